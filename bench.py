#!/usr/bin/env python
"""DPV hot-path benchmark (contract: one JSON line on stdout from rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): default_stereo, batch 8 per GPU, D=64 bins, image
256x384, features / cost volume 64x96, C=67 channels, V=1 source view, fp32.  One "step" pushes
one batch of 8 frames through the hot-path kernels: cost volume with the 1/4-res log-softmax in its
epilogue, then the full-res head (log-softmax, E[d], Var, arg-max, 1/4 hand-off) fused with the
uncertainty field -- three launches; the CNN blocks between them are cuDNN's and are not part of
the path (their outputs are synthetic inputs here).
`value`  : frames/s with every input already resident in HBM, device-timed, max over ranks.
`e2e`    : frames/s through the host-buffer C-ABI pipeline (dpv_pipeline_run): pinned host
           inputs are copied in, results copied back, inside the timed region.
`roofline`: the dominant kernel (fused full-res head + UF; --no-fuse-uf: the head alone) against the
           measured HBM peak; `kernels` / `sweep`: per-kernel times and the sweep's FP32 figures.
`cpu_baseline` / --impl reference: the CPU port of the reference's PyTorch path (oracle/) on the
           host cores, bounded sample.
"""
import argparse
import importlib
import json
import os
import statistics
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "DPV frames/sec (D=64, 256x384)"
UNIT = "frames/s"
WL = dict(B=8, V=1, C=67, D=64, h=64, w=96, H=256, W=384)
WORKLOAD = "default_stereo: batch 8/GPU, D=64, image 256x384, features 64x96, C=67, V=1"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def host_inputs(dpv, batch, seed=0):
    s = dpv.synth
    d = s.depth_candidates(5.0, 40.0, WL["D"], 1.0)
    cam = s.camera(WL["w"], WL["h"], batch)
    feats = s.randn(seed + 1, batch, WL["V"] + 1, WL["C"], WL["h"], WL["w"])
    poses = s.stereo_poses(batch)
    logits = s.ground_plane_logits(seed + 2, batch, WL["H"], WL["W"], d, cam["intrinsics_up"][0])
    return dict(d=d, feats=feats, poses=poses, K=cam["intrinsics"], rays=cam["unit_ray"],
                intr_up=cam["intrinsics_up"], logits=logits)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (NVML, ~10 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, nm in names.items():
                    if bits & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.01)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=1.0)
        return {"sm_mhz": int(statistics.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def cpu_port_fps(dpv, budget_s=15.0, frames_per_step=1, max_steps=50, warmup=1):
    """The reference's CPU PyTorch path (oracle port), frames/s on this host's cores."""
    from oracle import dpv_oracle as O
    T = torch.from_numpy
    torch.set_num_threads(os.cpu_count() or 1)
    hi = host_inputs(dpv, frames_per_step, seed=100)

    def one_step():
        for b in range(frames_per_step):
            O.frame_hot_path(T(hi["feats"][b:b + 1, -1]), T(hi["feats"][b:b + 1, :-1]), hi["d"],
                             T(hi["poses"][b, :-1, :3, :3]), T(hi["poses"][b, :-1, :3, 3]),
                             T(hi["K"][b]), T(hi["rays"][b]), 10.0, None, T(hi["logits"][b:b + 1]),
                             T(hi["intr_up"][b]))
    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    n = 0
    while n < max_steps:
        one_step()
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return n * frames_per_step / dt, n, dt


def run_reference(args, dpv):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import dpv_oracle as O
    T = torch.from_numpy
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    frames = 1                                  # bounded sample: 1 frame of the batch per step
    hi = host_inputs(dpv, frames, seed=100)

    def one_step():
        for b in range(frames):
            O.frame_hot_path(T(hi["feats"][b:b + 1, -1]), T(hi["feats"][b:b + 1, :-1]), hi["d"],
                             T(hi["poses"][b, :-1, :3, :3]), T(hi["poses"][b, :-1, :3, 3]),
                             T(hi["K"][b]), T(hi["rays"][b]), 10.0, None, T(hi["logits"][b:b + 1]),
                             T(hi["intr_up"][b]))
    steps = min(args.steps, 40)                 # keep the whole run within minutes
    for _ in range(min(args.warmup, 3)):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = time.perf_counter() - t0
    fps = steps * frames / dt
    sample = "%d step(s) x %d frame of the batch-8 workload, torch CPU ops, %d threads" % (steps, frames, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 3), "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs NVML reports as local to GPU `index`, so that the pinned host buffers
    of the end-to-end path are allocated (first touch) on the memory next to that GPU.  With one process per
    GPU on a two-socket host, unbound ranks otherwise push half of their H2D traffic across the socket link.
    Best effort: returns the number of CPUs bound to, or 0."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def alg_of(name, alg):
    """Algorithmic bytes of the launch(es) timed under `name` (FrameStep.algorithmic_bytes keys)."""
    return {"sweep": alg.get("sweep", 0), "head_quarter": alg.get("head_quarter", 0),
            "head_full": alg.get("head_full", 0), "ufield": alg.get("ufield", 0),
            "head_full_ufield": alg.get("head_full", 0), "bayes_fuse": alg.get("bayes_fuse", 0),
            "warp_feature": alg.get("warp_feature", 0), "feedback_fuse": alg.get("feedback_fuse", 0)}.get(name, 0)


def run_ours(args, dpv):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the DPV path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    bound_cpus = bind_to_gpu_numa_node(local) if world > 1 else 0
    out = sys.stdout
    if world > 1:
        # NCCL prints its version / debug lines on file descriptor 1 when NCCL_DEBUG is set in the
        # environment (NCCL_DEBUG_FILE did not catch the version line on the GPU boxes).  stdout carries
        # exactly one JSON line: keep a private copy of fd 1 for it and point fd 1 at stderr.
        sys.stdout.flush()
        out = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    frame_mod = importlib.import_module("probabilistic-depth_b200.frame")
    pipe_mod = importlib.import_module("probabilistic-depth_b200.pipeline")
    B = WL["B"]
    hi = host_inputs(dpv, B, seed=rank)
    step = frame_mod.FrameStep(B, WL["V"], WL["C"], WL["D"], WL["h"], WL["w"], WL["H"], WL["W"], hi["d"],
                               sigma=10.0, mode="default", device=dev, fuse_uf=not args.no_fuse_uf,
                               fuse_lsm=not args.no_fuse_lsm)
    # two input sets in HBM, alternated: 2 x 227 MB read + 227 MB written per step >> 126 MB L2
    nset = 2
    dsets = []
    for i in range(nset):
        h = hi if i == 0 else host_inputs(dpv, B, seed=rank + 1000 * i)
        dsets.append({k: torch.from_numpy(np.ascontiguousarray(h[k])).to(dev)
                      for k in ("feats", "poses", "K", "rays", "logits", "intr_up")})

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    cur = {"i": 0, "on": False}

    HOOK_EVERY = 8     # an event pair between kernels costs ~3 us of bubble: sample every 8th step

    def hook(which):
        if cur["on"] and cur["i"] % HOOK_EVERY == 0:
            ev[cur["i"]][which].record()

    def one(i, timed=False):
        s = dsets[i % nset]
        cur["i"], cur["on"] = i, timed
        step.run(s["feats"], s["poses"], s["K"], s["rays"], s["logits"], s["intr_up"], head_hook=hook)

    # One CUDA graph per input set: a replay re-runs the step's three launches with one host call.  Every
    # HOOK_EVERY-th step is launched call by call instead, with the events that time the dominant kernel.
    graphs = None
    if not args.no_graph:
        try:
            graphs = [step.capture(s_["feats"], s_["poses"], s_["K"], s_["rays"], s_["logits"], s_["intr_up"])
                      for s_ in dsets]
        except Exception as exc:     # capture is an optimisation of the launch path, not of the kernels
            print("bench.py: CUDA graph capture failed (%s); launching call by call" % exc, file=sys.stderr)
            graphs = None
            torch.cuda.synchronize()
    replays = 0

    def go(i, timed):
        if graphs is not None and not (timed and i % HOOK_EVERY == 0):
            graphs[i % nset].replay()
            return 1
        one(i, timed)
        return 0

    for i in range(max(args.warmup, 3)):
        go(i + 1, False)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = dpv._lib.launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(args.steps):
        replays += go(i, True)
    t_end.record()
    barrier()
    launches = dpv._lib.launch_count() - l0 + replays * step.launches_per_step()
    ms = t_start.elapsed_time(t_end)
    head_ms = [a.elapsed_time(b) for i, (a, b) in enumerate(ev) if i % HOOK_EVERY == 0]

    # ---- per-kernel breakdown (separate short loop: events between every launch) -----------
    kev = {}
    kcur = {"i": 0}
    nbk = 20

    def khook(name, which):
        lst = kev.setdefault(name, [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(nbk)])
        lst[kcur["i"]][which].record()

    for i in range(nbk):
        s_ = dsets[i % nset]
        kcur["i"] = i
        step.run(s_["feats"], s_["poses"], s_["K"], s_["rays"], s_["logits"], s_["intr_up"], kernel_hook=khook)
    torch.cuda.synchronize()
    kernel_ms = {n: statistics.median(a.elapsed_time(b) for a, b in lst) for n, lst in kev.items()}

    # ---- end to end through the host-buffer pipeline --------------------------------------
    pipe = pipe_mod.FramePipeline(B, WL["V"], WL["C"], WL["D"], WL["h"], WL["w"], WL["H"], WL["W"],
                                  device=local)
    pin = {k: torch.from_numpy(np.ascontiguousarray(hi[k])).pin_memory()
           for k in ("feats", "poses", "K", "rays", "logits", "intr_up")}
    outs = pipe.outputs(pinned=True)
    e2e_steps = max(3, min(args.steps, 30))

    def e2e_one():
        pipe.run(pin["feats"], pin["poses"], pin["K"], pin["rays"], hi["d"], pin["logits"],
                 pin["intr_up"], 10.0, outs)
    for _ in range(3):
        e2e_one()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_one()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    h2d, d2h = pipe.last_bytes()
    clocks = sampler.finish()
    pipe.close()

    # ---- max over ranks ---------------------------------------------------------------------
    stats = torch.tensor([ms, e2e_s, statistics.mean(head_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    ms, e2e_s, head_mean_ms = [float(v) for v in stats.cpu()]

    if rank == 0:
        peak, peak_kind = measured_peak()
        alg = step.algorithmic_bytes()
        head_name, head_bytes = step.dominant_kernel()
        achieved = head_bytes / (head_mean_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "head_uf_tile_traffic.json" if step.fused_uf
                                   else "head_full_traffic.json")) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": world * B * args.steps / (ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "l2": "inputs larger than L2 (2 alternating 227 MB sets)",
                       "kernels_per_step": step.launches_per_step(),
                       "algorithmic_bytes_per_step": alg,
                       "uf_fused_into_head": bool(step.fused_uf),
                       "cuda_graph_replay": graphs is not None,
                       "cpus_bound_per_rank": bound_cpus,
                       "quarter_log_softmax_in_sweep_epilogue": bool(step.fuse_lsm),
                       "frame_hbm_frac": sum(alg.values()) / (ms / args.steps * 1e-3) / 1e9 / peak,
                       "frame_hbm_frac_note": "SURVEY 8d bytes/frame (K5 counted as its own pass) / time / peak"},
            "roofline": {"kernel": head_name, "bound": "hbm",
                         "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "algorithmic_bytes_per_launch": head_bytes, "ms_per_launch": head_mean_ms,
                         "note": ("fused K3+K5: bytes of the fused operation itself (logits in, log-DPV and the "
                                  "per-pixel / per-column products out); SURVEY 8d counts K5 as a second pass "
                                  "over the DPV, see unfused_bytes_frac") if step.fused_uf else "K3 alone",
                         "unfused_bytes_frac": None},
            "e2e": {"value": world * B * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "api": "dpv_pipeline_run (host buffers, pinned)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        # where the step goes (rank 0, events around every launch, separate 20-step loop) and the
        # sweep kernel against the roof that binds it: it moves 6 % of the bytes but is FP32-bound
        fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12     # SMs x lanes x 2 x max clock (derived)
        sw_ms = kernel_ms.get("sweep")
        line["kernels"] = {
            n: {"ms": ms_k, "algorithmic_GBs": (alg_of(n, alg) / (ms_k * 1e-3) / 1e9) if alg_of(n, alg) else None}
            for n, ms_k in kernel_ms.items()}
        if sw_ms:
            line["sweep"] = {"bound": "fp32", "direct_form_flops_per_launch": step.sweep_flops(),
                             "achieved_TFLOPs_direct_form": step.sweep_flops() / (sw_ms * 1e-3) / 1e12,
                             "fp32_peak_TFLOPs": fp32_peak, "peak_kind": "derived",
                             "note": "flops of the direct (per-plane, 4-tap) form; the Gram form executes ~2.3x fewer",
                             "includes_quarter_log_softmax": bool(step.fuse_lsm)}
        if world == 1 and not args.no_cpu_baseline:
            import warnings
            warnings.filterwarnings("ignore")
            fps, n, dt = cpu_port_fps(dpv)
            line["cpu_baseline"] = {
                "value": fps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                "sample": "%d frame(s) of the batch-8 workload in %.1f s, torch CPU ops" % (n, dt)}
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fuse-uf", action="store_true",
                    help="run K3 and K5 as dpv_head + dpv_ufield (4 launches) instead of the fused tile kernel")
    ap.add_argument("--no-graph", action="store_true",
                    help="launch every step call by call instead of replaying per-input-set CUDA graphs")
    ap.add_argument("--no-fuse-lsm", action="store_true",
                    help="1/4-res log-softmax as its own dpv_head launch instead of the sweep kernel's epilogue")
    args = ap.parse_args()
    dpv = importlib.import_module("probabilistic-depth_b200")
    if args.impl == "reference":
        run_reference(args, dpv)
    else:
        run_ours(args, dpv)


if __name__ == "__main__":
    main()
