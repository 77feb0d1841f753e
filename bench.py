#!/usr/bin/env python
"""DPV hot-path benchmark (contract: one JSON line on stdout from rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads (BASELINE.json configs; D=64 bins, image 256x384, features / cost volume 64x96, C=67, V=1, fp32):
  stereo   (default, configs[1]) default_stereo, batch 8 per GPU: cost volume (+ 1/4-res log-softmax) ->
           full-res head (log-softmax, E[d], Var, arg-max, 1/4 hand-off) fused with the uncertainty field.
  stereo_refine  the same with the model's 1/4-res head between the cost volume and the soft-max: the three D -> D
           3x3 convolutions on the tcgen05 tensor cores, random-init weights.
  feedback (configs[2]) default_mono_feedback: + feedback warp of the previous DPV and log_softmax(BV + resi);
           8 sequences per GPU, a step = one frame of every sequence (16 steps = the 16-frame sequence).
  feedback_refine  the same with the model's Base3D inside the step (second half): the residual is
           computed from cat(BV, prev_output, warped features) by the 3-D convolution stack on the tcgen05 tensor
           cores instead of being a synthetic input; every arm (ours, incumbent, CPU reference) runs it.
  upsample (configs[3]) default_mono_upsample: + LiDAR prior and Bayesian fusion; GLOBAL batch 32 split over
           the ranks ("strong" scaling).
  stress   north_star's literal shape: cost volume + soft-max head directly on 256x384 features, C=67
           (SURVEY.md 8d "stress shape"), batch 4 per GPU; reports both the HBM and the FP32 roof.
  large_d  (configs[4]) D=256 at 384x1280: the head over a volume whose planes are sharded over the ranks
           (NCCL exchange of the soft-max statistics); N=1 runs the unsharded head.
The CNN blocks between the kernels are cuDNN's and are not part of the path (their outputs are synthetic inputs).

`value`    : frames/s with every input already resident in HBM, device-timed, max over ranks.
`e2e`      : frames/s through the host-buffer API: pinned host inputs are copied in, results copied back, inside
             the timed region (stereo: dpv_pipeline_submit / _wait, two batches in flight; other workloads:
             FrameStep.run double-buffered between pinned copies on side streams).  `e2e.ceiling` is a copy-only run of
             the same bytes.
`roofline` : the dominant kernel against the measured HBM peak; `kernels` / `sweep`: per-kernel times and the
             sweep's FP32 figures; `frame_hbm_frac` counts the bytes of the launched (fused) configuration,
             `survey_hbm_frac` the SURVEY 8d accounting (every reference pass on its own).
`cpu_baseline` / --impl reference: the UNMODIFIED reference's own functions (oracle/reference_loader.py:
             /root/reference or the copy build() ships to baseline/_ref) on the host cores, bounded sample;
             `incumbent`: the same functions through torch-CUDA on this GPU (what the reference does on a B200).
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
BASE = dict(V=1, C=67, D=64, h=64, w=96, H=256, W=384)
WORKLOADS = {
    "stereo": dict(BASE, B=8, mode="default", pose="stereo", scaling="weak",
                   desc="default_stereo: batch 8/GPU, D=64, image 256x384, features 64x96, C=67, V=1"),
    "stereo_refine": dict(BASE, B=8, mode="default", pose="stereo", scaling="weak", refine=True,
                          desc="default_stereo with the model's 1/4-res head: cost volume -> conv0 / conv0_1 / conv0_2 "
                               "(tcgen05, TF32x3) -> log-softmax -> full-res head + UF; batch 8/GPU, D=64, 256x384"),
    "feedback": dict(BASE, B=8, mode="feedback", pose="mono", scaling="weak",
                     desc="default_mono_feedback: 8 sequences/GPU, one frame of each per step (16 steps = the 16-frame "
                          "sequence), feedback warp + fusion, D=64, image 256x384, features 64x96, C=67, V=1"),
    "feedback_refine": dict(BASE, B=8, mode="feedback", pose="mono", scaling="weak", base3d=True,
                            desc="default_mono_feedback with the model's Base3D in the step: cost volume -> warp_feature -> "
                                 "Base3D(cat(BV, prev, warped)) (tcgen05, TF32x3; residual blocks on batch statistics as in "
                                 "the reference) -> log_softmax(BV + resi) -> full-res head + UF; 8 sequences/GPU, D=64, 256x384"),
    "upsample": dict(BASE, B=32, mode="upsample", pose="stereo", scaling="strong",
                     desc="default_mono_upsample: global batch 32 split over the GPUs, sparse depth prior + Bayesian "
                          "fusion, D=64, image 256x384, features 64x96, C=67, V=1"),
    "stress": dict(BASE, B=4, h=256, w=384, mode="stress", pose="stereo", scaling="weak",
                   desc="stress shape: cost volume + soft-max head directly on 256x384 features, C=67, D=64, "
                        "batch 4/GPU"),
    "large_d": dict(V=1, C=16, D=256, h=384, w=1280, H=384, W=1280, B=1, mode="large_d", pose="mono",
                    scaling="strong",
                    desc="large-D: head over a D=256, 384x1280 volume, depth planes sharded over the GPUs"),
}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def host_inputs(dpv, wl, batch, seed=0):
    """Seeded synthetic host inputs of one batch of `wl`."""
    s = dpv.synth
    d = s.depth_candidates(5.0, 40.0, wl["D"], 1.0)
    cam = s.camera(wl["w"], wl["h"], batch)
    out = dict(d=d, feats=s.randn(seed + 1, batch, wl["V"] + 1, wl["C"], wl["h"], wl["w"]),
               poses=s.stereo_poses(batch) if wl["pose"] == "stereo" else s.mono_poses(batch),
               K=cam["intrinsics"], rays=cam["unit_ray"], intr_up=cam["intrinsics_up"])
    if wl["mode"] in ("default", "feedback", "upsample"):
        one = s.ground_plane_logits(seed + 2, min(batch, 8), wl["H"], wl["W"], d, cam["intrinsics_up"][0])
        out["logits"] = np.concatenate([one] * ((batch + 7) // 8))[:batch] if batch > 8 else one
    if wl["mode"] == "upsample":
        out["dmaps"], out["masks"] = s.sparse_depth(seed + 3, batch, wl["h"], wl["w"])
    if wl["mode"] == "feedback":
        out["feat_raw"] = s.randn(seed + 4, batch, wl["V"] + 1, wl["D"], wl["h"], wl["w"])
        if wl.get("base3d"):       # the previous frame's 1/4-res hand-off (a log-DPV); the residual is computed in the step
            z = 2.0 * s.randn(seed + 5, batch, wl["D"], wl["h"], wl["w"])
            z = z - z.max(axis=1, keepdims=True)
            out["prev"] = (z - np.log(np.exp(z).sum(axis=1, keepdims=True))).astype(np.float32)
        else:
            out["bv_resi"] = 0.5 * s.randn(seed + 5, batch, wl["D"], wl["h"], wl["w"])
    return out


TENSOR_KEYS = ("feats", "poses", "K", "rays", "logits", "intr_up", "dmaps", "masks", "feat_raw", "bv_resi", "prev")


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


def base3d_module(dpv, wl, device, models=None):
    """Base3D as the feedback model builds it (models/models.py:464), random init, eval(): dres0 / classify on their
    running statistics, the unregistered residual blocks on batch statistics.  models: the module that defines Base3D
    (default: our mirror; the reference arms pass the reference's)."""
    import importlib
    M = models or importlib.import_module("probabilistic-depth_b200.models.models")
    torch.manual_seed(11)
    real_cuda = torch.nn.Module.cuda
    if str(device) == "cpu":
        torch.nn.Module.cuda = lambda self, device=None: self      # the reference's __init__ calls .cuda() on the blocks
    try:
        net = M.Base3D(wl["V"] + 3, dres_count=2, feature_dim=32, bn_running_avg=True, id=0)
    finally:
        torch.nn.Module.cuda = real_cuda
    net = net.to(device).eval()
    net.dres_modules = [b.to(device) for b in net.dres_modules]
    return net


# ------------------------------------------------------------------------------------ reference arms
def reference_step_fn(dpv, wl, frames, device):
    """A closure running `frames` frames of `wl` through the reference's own functions on `device`
    (oracle/reference_frame.py on the imported, unmodified reference), or through the oracle port when no
    reference tree is present.  Returns (fn, kind)."""
    T = torch.from_numpy
    hi = host_inputs(dpv, wl, frames, seed=100)
    mode = wl["mode"]
    try:
        from oracle import reference_loader, reference_frame
        ref = reference_loader.load() if reference_loader.available() else None
    except Exception:
        ref = None
    if ref is not None:
        if mode == "large_d":
            x = torch.randn((frames, wl["D"], wl["H"], wl["W"]), device=device)
            return (lambda: reference_frame.head_batch(ref, x, hi["d"], None, uf=False)), "reference"
        t = {k: T(np.ascontiguousarray(hi[k])).to(device) for k in TENSOR_KEYS if k in hi}
        if mode == "stress":
            return (lambda: reference_frame.stress_hot_path(ref, t["feats"], t["poses"], t["K"], t["rays"], hi["d"],
                                                            10.0, want=False)), "reference"
        kw = {}
        if mode == "upsample":
            kw = dict(mode="upsample", dmaps=t["dmaps"], masks=t["masks"])
        elif mode == "feedback":
            kw = dict(mode="feedback", feat_raw=t["feat_raw"], bv_resi=t.get("bv_resi"))
            if wl.get("base3d"):
                net = base3d_module(dpv, wl, device, models=ref.models)

                def base3d_ref(vol):
                    with torch.no_grad():
                        return net(vol, prob=False)
                kw.update(prev=t["prev"], base3d=base3d_ref)
        if wl.get("refine"):
            # the model's own 1/4-res head modules (models/models.py:456-460), random-init as the model builds them
            M = ref.models
            torch.manual_seed(7)
            mods = torch.nn.Sequential(M.conv2d_leakyRelu(wl["D"], wl["D"], 3, 1, 1), M.conv2d_leakyRelu(wl["D"], wl["D"], 3, 1, 1),
                                       torch.nn.Conv2d(wl["D"], wl["D"], 3, 1, 1)).to(device).eval()

            def refine_ref(c):
                with torch.no_grad():
                    return mods(c)
            kw["refine"] = refine_ref
        return (lambda: reference_frame.frame_hot_path(ref, t["feats"], t["poses"], t["K"], t["rays"], hi["d"], 10.0,
                                                       t["logits"], t["intr_up"], want=False, **kw)), "reference"
    # no reference tree on this machine: the restatement (CPU only, default frame)
    from oracle import dpv_oracle as O
    if mode not in ("default", "feedback", "upsample") or str(device) != "cpu":
        return None, "port"

    def port():
        for b in range(frames):
            O.frame_hot_path(T(hi["feats"][b:b + 1, -1]), T(hi["feats"][b:b + 1, :-1]), hi["d"],
                             T(hi["poses"][b, :-1, :3, :3]), T(hi["poses"][b, :-1, :3, 3]),
                             T(hi["K"][b]), T(hi["rays"][b]), 10.0, None, T(hi["logits"][b:b + 1]),
                             T(hi["intr_up"][b]))
    return port, "port"


def timed_cpu(fn, budget_s, max_steps, warmup=1):
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    n = 0
    while n < max_steps:
        fn()
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    return n, time.perf_counter() - t0


def run_reference(args, dpv, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import warnings
    warnings.filterwarnings("ignore")
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    frames = 1                                  # bounded sample: 1 frame of the workload's batch per step
    fn, kind = reference_step_fn(dpv, wl, frames, "cpu")
    if fn is None:
        print(json.dumps({"impl": "reference", "unavailable": "no reference tree and no CPU port for workload %s"
                          % args.workload}), flush=True)
        return
    steps = max(1, min(args.steps, 40))         # keep the whole run within minutes
    warm = max(1, min(args.warmup, 3))
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = time.perf_counter() - t0
    fps = steps * frames / dt
    sample = "%d step(s) x %d frame of the workload, %s, torch CPU ops, %d threads" % (
        steps, frames, "the reference's own functions" if kind == "reference" else "oracle port", cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": wl["desc"], "sample": sample},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs NVML reports as local to GPU `index`, so that the pinned host buffers
    of the end-to-end path are allocated (first touch) on the memory next to that GPU.  Best effort: returns
    the number of CPUs bound to, or 0."""
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


# ------------------------------------------------------------------------------------ steps
class StressStep:
    """K1-K3 directly on full-resolution features: cost volume, then the head on the cost volume itself
    (models/packnet.py:380-394 style).  Same interface subset as FrameStep."""

    def __init__(self, dpv, wl, B, d, dev):
        self.ops, self.lib, self._lib = dpv.ops, dpv._lib.load(), dpv._lib
        self.B, self.V, self.C, self.D, self.h, self.w = B, wl["V"], wl["C"], wl["D"], wl["h"], wl["w"]
        e = lambda *s, dt=torch.float32: torch.empty(s, device=dev, dtype=dt)
        self.d = dpv.ops.depth_bins(d, dev)
        self.cost, self.logp = e(B, self.D, self.h, self.w), e(B, self.D, self.h, self.w)
        self.depth, self.var = e(B, self.h, self.w), e(B, self.h, self.w)
        self.argmax = e(B, self.h, self.w, dt=torch.int64)
        self.ws = e(int(self.lib.dpv_sweep_workspace_floats(B, self.V, self.h, self.w)))
        self.mode, self.fused_uf, self.fuse_lsm = "stress", False, False

    def run(self, s, head_hook=None, kernel_hook=None):
        B, V, C, D, h, w = self.B, self.V, self.C, self.D, self.h, self.w
        st = torch.cuda.current_stream().cuda_stream
        hk = kernel_hook if kernel_hook is not None else (lambda n, i: None)
        chw, fp, p = C * h * w, s["feats"].data_ptr(), (lambda t: t.data_ptr())
        hk("sweep", 0)
        self._lib.check(self.lib.dpv_sweep_cost_volume_ws(
            fp + 4 * V * chw, fp, p(s["poses"]), p(s["K"]), p(s["rays"]), p(self.d), p(self.cost), None,
            B, V, C, D, h, w, (V + 1) * chw, (V + 1) * chw, chw, (V + 1) * 16, 9, 3 * h * w, 10.0, 0, 0, p(self.ws), st))
        hk("sweep", 1)
        if head_hook:
            head_hook(0)
        hk("head_full", 0)
        self._lib.check(self.lib.dpv_head(p(self.cost), None, p(self.d), p(self.logp), None, p(self.depth), p(self.var),
                                          p(self.argmax), None, B, D, h, w, self.ops.IN_LOGITS, st))
        hk("head_full", 1)
        if head_hook:
            head_hook(1)

    def algorithmic_bytes(self):
        hw = self.h * self.w
        return {"sweep": self.B * (4 * hw * (self.C * (1 + self.V) + self.D) + 12 * hw),
                "head_full": self.B * (8 * hw * self.D + 16 * hw)}

    survey_bytes = algorithmic_bytes

    def sweep_flops(self):
        return self.B * self.h * self.w * self.D * self.V * (11 * self.C + 40)

    def dominant_kernel(self):
        return "dpv::head_kernel<64,1,LOGITS,...> (soft-max head on the cost volume)", self.algorithmic_bytes()["head_full"]

    def results(self):
        return dict(depth=self.depth, variance=self.var, argmax=self.argmax)


class LargeDStep:
    """The head over a [1, D, H, W] volume whose depth planes are split over the ranks (sharding.PlaneShardedHead;
    world 1: the plain dpv_head)."""

    def __init__(self, dpv, wl, d, dev, rank, world):
        sh = importlib.import_module("probabilistic-depth_b200.sharding")
        self.ops, self.d, self.world = dpv.ops, d, world
        self.D, self.H, self.W = wl["D"], wl["H"], wl["W"]
        self.lo, self.hi = sh.plane_range(self.D, rank, world)
        self.head = sh.PlaneShardedHead(self.D) if world > 1 else None
        self.mode, self.fused_uf, self.fuse_lsm = "large_d", False, False
        self.out = None

    def run(self, s, head_hook=None, kernel_hook=None):
        if head_hook:
            head_hook(0)
        if self.head is not None:
            self.out = self.head(s["x"], self.d, variance=True, argmax=True, logp=True)
        else:
            self.out = self.ops.head(s["x"], self.d, logp=True, depth=True, variance=True, argmax=True)
        if head_hook:
            head_hook(1)

    def algorithmic_bytes(self):
        HW, Dl = self.H * self.W, self.hi - self.lo
        return {"head_sharded" if self.head is not None else "head_full": 8 * HW * Dl + 16 * HW}

    survey_bytes = algorithmic_bytes

    def dominant_kernel(self):
        n = list(self.algorithmic_bytes().items())[0]
        return ("plane-sharded head: local statistics pass + exchange + merge / finish pass (per rank)"
                if self.head is not None else "dpv::head_kernel<256,4,...>"), n[1]

    def results(self):
        return {k: v for k, v in self.out.items() if k != "logp"}


def frame_results(step):
    r = dict(bv=step.bv, depth=step.depth, variance=step.var, argmax=step.argmax, uf=step.uf, depth_zero=step.dz,
             quarter=step.quarter)
    if step.mode == "feedback":
        r["bv_upd"] = step.bv_upd
    if step.mode == "upsample":
        r["logfused"] = step.logfused
    return r


def run_ours(args, dpv, wl):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the DPV path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    bound_cpus = 0
    out = sys.stdout
    if world > 1:
        # NCCL prints its version / debug lines on file descriptor 1 when NCCL_DEBUG is set in the
        # environment.  stdout carries exactly one JSON line: keep a private copy of fd 1 for it and point
        # fd 1 at stderr.
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
    mode = wl["mode"]
    strong = wl["scaling"] == "strong"
    if strong and mode != "large_d" and wl["B"] % world:
        raise SystemExit("bench.py: workload %s needs the global batch %d divisible by the GPU count"
                         % (args.workload, wl["B"]))
    B = wl["B"] // world if (strong and mode != "large_d") else wl["B"]
    frames_per_step_global = wl["B"] if strong else wl["B"] * world
    hi = (host_inputs(dpv, wl, B, seed=rank) if mode != "large_d"
          else dict(d=dpv.synth.depth_candidates(5.0, 40.0, wl["D"])))
    to_dev = lambda h: {k: torch.from_numpy(np.ascontiguousarray(h[k])).to(dev) for k in TENSOR_KEYS if k in h}

    # ---- the step and its input sets (rotated: more bytes than the 126 MB L2 between two uses) -------------
    if mode in ("default", "feedback", "upsample"):
        refine = None
        if wl.get("refine"):
            gw = torch.Generator(device=dev).manual_seed(7)
            std = (2.0 / (9 * 64)) ** 0.5                       # models/models.py weight_init
            refine = dpv.ops.CostRefine([torch.randn((64, 64, 3, 3), device=dev, generator=gw) * std for _ in range(3)],
                                        [torch.randn((64,), device=dev, generator=gw) * 0.1 for _ in range(3)])
        base3d = None
        if wl.get("base3d"):
            base3d = dpv.ops.Base3DConvs.from_module(base3d_module(dpv, wl, dev))
        step = frame_mod.FrameStep(B, wl["V"], wl["C"], wl["D"], wl["h"], wl["w"], wl["H"], wl["W"], hi["d"],
                                   sigma=10.0, mode=mode, device=dev, fuse_uf=not args.no_fuse_uf,
                                   fuse_lsm=not args.no_fuse_lsm, refine=refine, base3d=base3d)
        nset = 2 if B <= 16 else 1
        dsets = [to_dev(hi if i == 0 else host_inputs(dpv, wl, B, seed=rank + 1000 * i)) for i in range(nset)]

        def run_step(s, **kw):
            step.run(s["feats"], s["poses"], s["K"], s["rays"], s["logits"], s["intr_up"], dmaps=s.get("dmaps"),
                     masks=s.get("masks"), feat_raw=s.get("feat_raw"), bv_resi=s.get("bv_resi"), prev=s.get("prev"), **kw)
        results = lambda: frame_results(step)
    elif mode == "stress":
        step = StressStep(dpv, wl, B, hi["d"], dev)
        nset = 2
        dsets = [to_dev(hi if i == 0 else host_inputs(dpv, wl, B, seed=rank + 1000 * i)) for i in range(nset)]
        run_step = lambda s, **kw: step.run(s, **kw)
        results = step.results
    else:
        step = LargeDStep(dpv, wl, hi["d"], dev, rank, world)
        nset = 2
        g = torch.Generator(device=dev).manual_seed(1234)       # the same global volume on every rank, sliced
        dsets = []
        for i in range(nset):
            full = torch.randn((1, wl["D"], wl["H"], wl["W"]), device=dev, generator=g) * 3
            dsets.append({"x": full[:, step.lo:step.hi].contiguous()})
            del full
        run_step = lambda s, **kw: step.run(s, **kw)
        results = step.results
    set_bytes = sum(t.numel() * t.element_size() for t in dsets[0].values())

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    cur = {"i": 0, "on": False}
    HOOK_EVERY = 8     # an event pair between kernels costs ~3 us of bubble: sample every 8th step

    def hook(which):
        if cur["on"] and cur["i"] % HOOK_EVERY == 0:
            ev[cur["i"]][which].record()

    def one(i, timed=False):
        cur["i"], cur["on"] = i, timed
        run_step(dsets[i % nset], head_hook=hook)

    # One CUDA graph per input set: a replay re-runs the step's launches with one host call.  Every
    # HOOK_EVERY-th step is launched call by call instead, with the events that time the dominant kernel.
    graphs = None
    launches_per_step = None
    if not args.no_graph and mode != "large_d":       # (the sharded head issues NCCL collectives through torch)
        try:
            graphs = []
            for s_ in dsets:
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    n0 = dpv._lib.launch_count()
                    run_step(s_)
                    launches_per_step = dpv._lib.launch_count() - n0
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_):
                    run_step(s_)
                graphs.append(g_)
        except Exception as exc:     # capture is an optimisation of the launch path, not of the kernels
            print("bench.py: CUDA graph capture failed (%s); launching call by call" % exc, file=sys.stderr)
            graphs = None
            torch.cuda.synchronize()
    if launches_per_step is None:
        n0 = dpv._lib.launch_count()
        run_step(dsets[0])
        launches_per_step = dpv._lib.launch_count() - n0
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
    launches = dpv._lib.launch_count() - l0 + replays * launches_per_step
    ms = t_start.elapsed_time(t_end)
    head_ms = [a.elapsed_time(b) for i, (a, b) in enumerate(ev) if i % HOOK_EVERY == 0]
    # The dominant kernel on its own: the full-res head (+ UF) of the step replayed back to back over the
    # rotating input sets (graph replays, no events between launches), one event pair around the lot.  The
    # in-step figure above brackets single launches with events, each of which costs ~3 us of bubble.
    head_alone_ms = None
    if mode in ("default", "feedback", "upsample"):
        try:
            hgraphs = []
            for s_ in dsets:
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    step.run_head(s_["logits"], s_["intr_up"])
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_):
                    step.run_head(s_["logits"], s_["intr_up"])
                hgraphs.append(g_)
            for i in range(4):
                hgraphs[i % nset].replay()
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            h0.record()
            for i in range(args.steps):
                hgraphs[i % nset].replay()
            h1.record()
            torch.cuda.synchronize()
            head_alone_ms = h0.elapsed_time(h1) / args.steps
            launches += args.steps * (2 if step.fused_uf else 1)
        except Exception as exc:
            print("bench.py: head-only graph timing failed (%s)" % exc, file=sys.stderr)
            torch.cuda.synchronize()

    # ---- per-kernel breakdown (separate short loop: events between every launch) -----------
    kev, kcur, nbk = {}, {"i": 0}, 20

    def khook(name, which):
        lst = kev.setdefault(name, [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(nbk)])
        lst[kcur["i"]][which].record()

    kernel_ms = {}
    if mode != "large_d":
        for i in range(nbk):
            kcur["i"] = i
            run_step(dsets[i % nset], kernel_hook=khook)
        torch.cuda.synchronize()
        kernel_ms = {n: statistics.median(a.elapsed_time(b) for a, b in lst) for n, lst in kev.items()}

    # ---- end to end through the host-buffer API ----------------------------------------------
    # (the pinned host buffers are first touched here: bind to the CPUs next to this GPU for this part only -- a
    # narrow affinity mask during the device-timed loops above would make the host threads of the ranks compete)
    affinity0 = os.sched_getaffinity(0)
    bound_cpus = bind_to_gpu_numa_node(local) if world > 1 else 0
    e2e_steps = max(4, min(args.steps, 30))
    pipe = None
    # (the C pipeline has no convolution stage: the refine workload goes through FrameStep.run between pinned copies)
    if mode == "default" and not wl.get("refine"):
        pipe = pipe_mod.FramePipeline(B, wl["V"], wl["C"], wl["D"], wl["h"], wl["w"], wl["H"], wl["W"], device=local)
        pins = [{k: torch.from_numpy(np.ascontiguousarray(h_[k])).pin_memory()
                 for k in ("feats", "poses", "K", "rays", "logits", "intr_up")}
                for h_ in (hi, host_inputs(dpv, wl, B, seed=rank + 77))]
        outs = [pipe.outputs(pinned=True) for _ in range(2)]

        def e2e_run(n):
            """n batches, two in flight: batch i + 1 is copied in while batch i computes (submit / wait)."""
            for i in range(n):
                pn = pins[i % 2]
                pipe.submit(pn["feats"], pn["poses"], pn["K"], pn["rays"], hi["d"], pn["logits"], pn["intr_up"], 10.0,
                            outs[i % 2])
                if i >= 1:
                    pipe.wait()
            pipe.wait()
        e2e_api = "dpv_pipeline_submit / dpv_pipeline_wait (host buffers, pinned, two batches in flight)"
        e2e_run(3)
        h2d, d2h = pipe.last_bytes()
        copy_in = [pins[0][k] for k in ("feats", "logits")]
    else:
        # Double-buffered through the public FrameStep API, as a streaming caller would: batch i + 1 is copied in on a
        # copy stream while batch i computes; the results of batch i go back on a third stream and must have left
        # the step's output buffers before batch i + 1 overwrites them.
        if len(dsets) < 2:
            dsets = dsets + [{k: torch.empty_like(v) for k, v in dsets[0].items()}]
        pin_in = [{k: v.cpu().pin_memory() for k, v in dsets[0].items()} for _ in range(2)]
        run_step(dsets[0])
        pin_out = [{k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in results().items()} for _ in range(2)]
        h2d = sum(t.numel() * t.element_size() for t in pin_in[0].values())
        d2h = sum(t.numel() * t.element_size() for t in pin_out[0].values())
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        mk_ev = lambda: [torch.cuda.Event() for _ in range(2)]
        ev_in, ev_done, ev_out = mk_ev(), mk_ev(), mk_ev()

        def e2e_run(n):
            main = torch.cuda.current_stream()
            for i in range(n):
                j = i % 2
                with torch.cuda.stream(s_in):
                    if i >= 2:
                        s_in.wait_event(ev_done[j])           # the step that last read input set j is done
                    for k, t in pin_in[j].items():
                        dsets[j][k].copy_(t, non_blocking=True)
                    ev_in[j].record(s_in)
                main.wait_event(ev_in[j])
                if i >= 1:
                    main.wait_event(ev_out[(i - 1) % 2])      # the previous results have left the output buffers
                run_step(dsets[j])
                ev_done[j].record(main)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_done[j])
                    if i >= 2:
                        ev_out[j].synchronize()               # (host) pin_out[j] of batch i - 2 has been consumed
                    for k, t in results().items():
                        pin_out[j][k].copy_(t, non_blocking=True)
                    ev_out[j].record(s_out)
            torch.cuda.synchronize()
        e2e_api = ("%s.run, double-buffered: pinned-host copies of every input (copy stream) and result (third stream) "
                   "overlap the step" % type(step).__name__)
        e2e_run(3)
        copy_in = list(pin_in[0].values())
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    # copy-only ceiling: the same input bytes host -> device, nothing else, all ranks at once
    dst = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in copy_in]
    for a_, b_ in zip(dst, copy_in):
        a_.copy_(b_, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        for a_, b_ in zip(dst, copy_in):
            a_.copy_(b_, non_blocking=True)
    torch.cuda.synchronize()
    copy_s = (time.perf_counter() - t0) / 5
    barrier()
    copy_bytes = sum(t.numel() * t.element_size() for t in copy_in)
    clocks = sampler.finish()
    if pipe is not None:
        pipe.close()
    del dst
    try:
        os.sched_setaffinity(0, affinity0)
    except Exception:
        pass

    # ---- max over ranks ---------------------------------------------------------------------
    stats = torch.tensor([ms, e2e_s, statistics.mean(head_ms), copy_s, head_alone_ms or 0.0], device=dev,
                         dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    ms, e2e_s, head_in_step_ms, copy_s, head_alone_ms = [float(v) for v in stats.cpu()]
    head_mean_ms = head_alone_ms if head_alone_ms > 0 else head_in_step_ms

    if rank == 0:
        peak, peak_kind = measured_peak()
        alg, surv = step.algorithmic_bytes(), step.survey_bytes()
        head_name, head_bytes = step.dominant_kernel()
        achieved = head_bytes / (head_mean_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        if mode in ("default", "feedback", "upsample"):
            tf = "head_uf_tile_traffic.json" if step.fused_uf else "head_full_traffic.json"
            try:
                with open(os.path.join(ROOT, "profiles", tf)) as f:
                    traffic = json.load(f).get("dram_bytes_per_launch")
                traffic_src = "static: offline ncu capture of the B=8 launch, profiles/%s (not measured in this run)" % tf
                if B != 8:
                    traffic = traffic * B / 8
            except Exception:
                pass
        step_s = ms / args.steps * 1e-3
        line = {
            "metric": METRIC, "value": frames_per_step_global * args.steps / (ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": wl["scaling"],
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "workload_name": args.workload, "batch_per_gpu": B,
                       "l2": "inputs larger than L2 (%d alternating %.0f MB sets + as much written)"
                             % (nset, set_bytes / 1e6),
                       "kernels_per_step": launches_per_step,
                       "algorithmic_bytes_per_step": alg,
                       "uf_fused_into_head": bool(step.fused_uf),
                       "cuda_graph_replay": graphs is not None,
                       "cpus_bound_per_rank": bound_cpus,
                       "quarter_log_softmax_in_sweep_epilogue": bool(step.fuse_lsm),
                       "frame_hbm_frac": sum(alg.values()) / step_s / 1e9 / peak,
                       "frame_hbm_frac_note": "bytes of the launched (fused) kernels: every input read once, every "
                                              "output written once / time / peak",
                       "survey_hbm_frac": sum(surv.values()) / step_s / 1e9 / peak,
                       "survey_hbm_frac_note": "SURVEY 8d accounting (K5 and the 1/4-res soft-max as passes of their "
                                               "own, as the reference runs them) / time / peak"},
            "roofline": {"kernel": head_name, "bound": "hbm",
                         "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": head_bytes, "ms_per_launch": head_mean_ms,
                         "ms_per_launch_how": ("back-to-back graph replays of the head over the rotating input sets, "
                                               "one CUDA-event pair around %d launches" % args.steps) if head_alone_ms > 0
                         else "event pairs around single launches inside the step (every 8th step)",
                         "ms_per_launch_in_step": head_in_step_ms,
                         "frac_in_step": head_bytes / (head_in_step_ms * 1e-3) / 1e9 / peak},
            "e2e": {"value": frames_per_step_global * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "api": e2e_api,
                    "ceiling": {"value": frames_per_step_global / (copy_s * h2d / copy_bytes), "unit": UNIT,
                                "h2d_GBs_per_gpu": copy_bytes / copy_s / 1e9,
                                "note": "copy-only: the step's input bytes host -> device from pinned memory, all "
                                        "ranks at once, nothing else running"}},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        line["e2e"]["frac_of_ceiling"] = line["e2e"]["value"] / line["e2e"]["ceiling"]["value"]
        # where the step goes (rank 0, events around every launch, separate 20-step loop) and the
        # sweep kernel against the roof that binds it: it moves few bytes and is FP32-bound
        fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12     # SMs x lanes x 2 x max clock (derived)
        line["kernels"] = {n: {"ms": ms_k,
                               "algorithmic_GBs": (alg.get(n, 0) / (ms_k * 1e-3) / 1e9) if alg.get(n) else None}
                           for n, ms_k in kernel_ms.items()}
        sw_ms = kernel_ms.get("sweep")
        if sw_ms:
            tf_ = step.sweep_flops() / (sw_ms * 1e-3) / 1e12
            line["sweep"] = {"bound": "fp32", "direct_form_flops_per_launch": step.sweep_flops(),
                             "achieved_TFLOPs_direct_form": tf_,
                             "fp32_peak_TFLOPs": fp32_peak, "peak_kind": "derived",
                             "fp32_frac_direct_form": tf_ / fp32_peak,
                             "hbm_frac": alg.get("sweep", 0) / (sw_ms * 1e-3) / 1e9 / peak,
                             "note": "flops of the direct (per-plane, 4-tap) form; the cross-correlation form "
                                     "executes several times fewer",
                             "includes_quarter_log_softmax": bool(step.fuse_lsm)}
        if world == 1 and not args.no_cpu_baseline:
            import warnings
            warnings.filterwarnings("ignore")
            # the incumbent: the reference's own functions through torch-CUDA on this GPU
            try:
                fn, kind = reference_step_fn(dpv, wl, B, dev)
                if fn is not None and kind == "reference":
                    for _ in range(2):
                        fn()
                    torch.cuda.synchronize()
                    a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    n_inc = 5
                    a_.record()
                    for _ in range(n_inc):
                        fn()
                    b_.record()
                    torch.cuda.synchronize()
                    inc_ms = a_.elapsed_time(b_) / n_inc
                    inc_fps = B / (inc_ms * 1e-3)
                    line["incumbent"] = {"value": inc_fps, "unit": UNIT, "ms_per_step": inc_ms,
                                         "what": "the unmodified reference's hot-path functions (est_swp_volume_v4, "
                                                 "log_softmax, dpv_to_depthmap, variance, gen_ufield ...) through "
                                                 "torch-CUDA on this GPU, same batch, device-timed",
                                         "speedup_device_timed": line["value"] / inc_fps}
                del fn
                torch.cuda.empty_cache()
            except Exception as exc:
                line["incumbent"] = {"unavailable": repr(exc)[:200]}
            torch.set_num_threads(os.cpu_count() or 1)
            fn, kind = reference_step_fn(dpv, wl, 1, "cpu")
            if fn is not None:
                n, dt = timed_cpu(fn, 15.0, 50)
                line["cpu_baseline"] = {
                    "value": n / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                    "sample": "%d frame(s) of the workload in %.1f s, %s, torch CPU ops" % (
                        n, dt, "the reference's own functions" if kind == "reference" else "oracle port")}
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="stereo", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fuse-uf", action="store_true",
                    help="run K3 and K5 as dpv_head + dpv_ufield (4 launches) instead of the fused tile kernel")
    ap.add_argument("--no-graph", action="store_true",
                    help="launch every step call by call instead of replaying per-input-set CUDA graphs")
    ap.add_argument("--no-fuse-lsm", action="store_true",
                    help="1/4-res log-softmax as its own dpv_head launch instead of the sweep kernel's epilogue")
    args = ap.parse_args()
    dpv = importlib.import_module("probabilistic-depth_b200")
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, dpv, wl)
    else:
        run_ours(args, dpv, wl)


if __name__ == "__main__":
    main()
