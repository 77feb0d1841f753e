#!/usr/bin/env python
"""Per-kernel timing (CUDA events, inputs rotated so every launch misses L2).  Dev tool."""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dpv = importlib.import_module("probabilistic-depth_b200")
ops = dpv.ops
PEAK = 6554.9


GRAPH = False


def timeit(fn, nrot, iters=20, warm=3):
    if GRAPH:
        return timeit_graph(fn, nrot, iters, warm)
    for i in range(warm):
        fn(i % nrot)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i, (a, b) in enumerate(evs):
        a.record()
        fn(i % nrot)
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def timeit_graph(fn, nrot, iters=20, warm=3):
    """One CUDA graph holding `nrot` back-to-back calls (rotating inputs): device time per call with
    no host launch overhead in the way."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(max(warm, nrot)):
            fn(i % nrot)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    keep = []
    with torch.cuda.graph(g):
        for i in range(nrot):
            keep.append(fn(i))
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / nrot)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--only", default="")
    ap.add_argument("--pose", default="stereo")
    args = ap.parse_args()
    global GRAPH
    GRAPH = args.graph
    B, V, C, D, h, w, H, W = args.B, 1, 67, 64, 64, 96, 256, 384
    s = dpv.synth
    d = s.depth_candidates(5, 40, D)
    cam = s.camera(w, h, B)
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    K, rays, Ku = cu(cam["intrinsics"]), cu(cam["unit_ray"]), cu(cam["intrinsics_up"])
    poses = cu(s.stereo_poses(B) if args.pose == "stereo" else s.mono_poses(B))
    res = {}

    # bring the clocks up before the first measurement (an idle B200 sits at 120 MHz and takes a few
    # hundred ms of load to reach its boost clock: the first section measured would otherwise look slow)
    _w = torch.empty((64, 1024, 1024), device="cuda")
    _t0 = torch.cuda.Event(enable_timing=True); _t1 = torch.cuda.Event(enable_timing=True)
    _t0.record()
    for _ in range(300):
        _w.mul_(1.0001)
    _t1.record()
    torch.cuda.synchronize()
    del _w

    def want(n):
        return not args.only or n in args.only.split(",")

    nrot = 3
    if want("sweep"):
        feats = [torch.randn((B, V + 1, C, h, w), device="cuda") for _ in range(nrot)]
        for algo in (4, 3, 2, 1):
            f = lambda i: ops.sweep_cost_volume(feats[i][:, -1], feats[i][:, :-1], poses[:, :-1], K, rays, d, 10.0, algo=algo)
            med, best = timeit(f, nrot, iters=10 if algo == 1 else 20)
            byt = B * (4 * h * w * (C * (1 + V) + D) + 12 * h * w)
            res["sweep_algo%d" % algo] = dict(ms=med, best_ms=best, GBs=byt / med / 1e6, frac=byt / med / 1e6 / PEAK)
    if want("head"):
        xs = [torch.randn((B, D, H, W), device="cuda") * 3 for _ in range(nrot)]
        f = lambda i: ops.head(xs[i], d, logp=True, depth=True, variance=True, argmax=True, quarter=True)
        med, best = timeit(f, nrot)
        byt = B * (8 * H * W * D + 16 * H * W)
        res["head_full"] = dict(ms=med, best_ms=best, GBs=byt / med / 1e6, frac=byt / med / 1e6 / PEAK)
        f = lambda i: ops.head(xs[i], d, logp=True)
        med, best = timeit(f, nrot)
        byt = B * (8 * H * W * D)
        res["head_full_logp_only"] = dict(ms=med, best_ms=best, GBs=byt / med / 1e6, frac=byt / med / 1e6 / PEAK)
        xq = [torch.randn((B, D, h, w), device="cuda") * 3 for _ in range(nrot)]
        f = lambda i: ops.head(xq[i], d, logp=True)
        med, best = timeit(f, nrot)
        byt = B * (8 * h * w * D)
        res["head_quarter"] = dict(ms=med, best_ms=best, GBs=byt / med / 1e6, frac=byt / med / 1e6 / PEAK)
        # plain copy of the same volume for reference
        ys = [torch.empty_like(xs[0]) for _ in range(nrot)]
        f = lambda i: ys[i].copy_(xs[i])
        med, best = timeit(f, nrot)
        byt = B * (8 * H * W * D)
        res["torch_copy_same_bytes"] = dict(ms=med, best_ms=best, GBs=byt / med / 1e6, frac=byt / med / 1e6 / PEAK)
    if want("ufield"):
        lg = cu(s.ground_plane_logits(2, B, H, W, d, cam["intrinsics_up"][0]))
        hd = ops.head(lg, d, logp=True, depth=True)
        lps = [hd["logp"].clone() for _ in range(nrot)]
        f = lambda i: ops.ufield(lps[i], d, Ku, depth=hd["depth"])
        med, best = timeit(f, nrot)
        byt = B * (4 * H * W * D + 4 * D * W + 8 * H * W)
        res["ufield_given_depth"] = dict(ms=med, best_ms=best, GBs=byt / med / 1e6, frac=byt / med / 1e6 / PEAK)
        f = lambda i: ops.ufield(lps[i], d, Ku)
        med, best = timeit(f, nrot)
        res["ufield_standalone"] = dict(ms=med, best_ms=best, GBs=byt / med / 1e6, frac=byt / med / 1e6 / PEAK)
        lgs = [lg.clone() for _ in range(nrot)]
        f = lambda i: ops.head_ufield(lgs[i], d, Ku, logp=True, depth=True, variance=True, argmax=True, quarter=True)
        med, best = timeit(f, nrot)
        byt = B * (8 * H * W * D + 16 * H * W + 4 * D * W + 4 * H * W)
        res["head_full_uf_fused"] = dict(ms=med, best_ms=best, GBs=byt / med / 1e6, frac=byt / med / 1e6 / PEAK)
        f = lambda i: ops.head(lgs[i], d, logp=True, depth=True, variance=True, argmax=True, quarter=True)
        med, best = timeit(f, nrot)
        byt = B * (8 * H * W * D + 16 * H * W)
        res["head_full_groundplane"] = dict(ms=med, best_ms=best, GBs=byt / med / 1e6, frac=byt / med / 1e6 / PEAK)
    if want("fuse"):
        bv = [torch.log_softmax(torch.randn((B, D, h, w), device="cuda"), 1) for _ in range(nrot)]
        dm, mk = s.sparse_depth(3, B, h, w)
        dm, mk = cu(dm), cu(mk)
        f = lambda i: ops.bayes_fuse(bv[i], d, dmaps=dm, masks=mk)
        med, best = timeit(f, nrot)
        byt = B * (12 * h * w * D + 8 * h * w)
        res["bayes_fuse"] = dict(ms=med, best_ms=best, GBs=byt / med / 1e6, frac=byt / med / 1e6 / PEAK)
        ad = [0.5 * torch.randn((B, D, h, w), device="cuda") for _ in range(nrot)]
        f = lambda i: ops.head(bv[i], d, addend=ad[i], logp=True)
        med, best = timeit(f, nrot)
        byt = B * (12 * h * w * D)
        res["feedback_fuse"] = dict(ms=med, best_ms=best, GBs=byt / med / 1e6, frac=byt / med / 1e6 / PEAK)
        fr = [torch.randn((B, 2, D, h, w), device="cuda") for _ in range(nrot)]
        f = lambda i: ops.warp_feature(fr[i], poses, K, rays, d)
        med, best = timeit(f, nrot)
        byt = B * (8 * h * w * D * 2 + 12 * h * w)
        res["warp_feature"] = dict(ms=med, best_ms=best, GBs=byt / med / 1e6, frac=byt / med / 1e6 / PEAK)
    if want("refine"):
        # conv0 -> LeakyReLU -> conv0_1 -> LeakyReLU -> conv0_2 -> log_softmax at [B,64,64,96]
        F = torch.nn.functional
        std = (2.0 / (9 * 64)) ** 0.5
        ws = [torch.randn((64, 64, 3, 3), device="cuda") * std for _ in range(3)]
        bs = [torch.randn((64,), device="cuda") * 0.1 for _ in range(3)]
        costs = [torch.randn((B, D, h, w), device="cuda") * 4 + 10 for _ in range(nrot)]
        refine = ops.CostRefine(ws, bs)
        flops = 3 * 2.0 * B * h * w * 64 * 64 * 9

        def torch_chain(i):
            y = F.leaky_relu(F.conv2d(costs[i], ws[0], bs[0], padding=1), 0.01)
            y = F.leaky_relu(F.conv2d(y, ws[1], bs[1], padding=1), 0.01)
            return torch.log_softmax(F.conv2d(y, ws[2], bs[2], padding=1), dim=1)
        med, best = timeit(lambda i: refine(costs[i]), nrot)
        res["refine_tcgen05_tf32x3"] = dict(ms=med, best_ms=best, GBs=0.0, frac=0.0, TFLOPs=flops / med / 1e9)
        for tf32 in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            for _ in range(3):
                torch_chain(0)
            med, best = timeit(torch_chain, nrot)
            res["refine_torch_cudnn_%s" % ("tf32" if tf32 else "fp32")] = dict(ms=med, best_ms=best, GBs=0.0, frac=0.0,
                                                                              TFLOPs=flops / med / 1e9)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        for k in ("refine_tcgen05_tf32x3", "refine_torch_cudnn_fp32", "refine_torch_cudnn_tf32"):
            print("%-26s %8.4f ms  %6.1f TFLOP/s (conv flops of the three layers)" % (k, res[k]["ms"], res[k]["TFLOPs"]))
    if want("base3d") and args.only:          # (only on request: 3 GB of buffers, ~100 ms per cuDNN pass)
        # SURVEY 8f rank 2, second half: Base3D(4, dres_count=2, feature_dim=32) on [Bv, 4, 64, 64, 96]
        import importlib as _il
        M = _il.import_module("probabilistic-depth_b200.models.models")
        Bv = int(os.environ.get("BASE3D_B", "2"))
        net = M.Base3D(4, dres_count=2, feature_dim=32, bn_running_avg=True).cuda().eval()
        net.dres_modules = [b.cuda() for b in net.dres_modules]
        vols = [torch.randn((Bv, 4, D, h, w), device="cuda") for _ in range(2)]
        flops = 2.0 * Bv * D * h * w * 27 * (4 * 32 + 6 * 32 * 32 + 32)
        with torch.no_grad():
            tc = ops.Base3DConvs.from_module(net)
            med, best = timeit(lambda i: tc(vols[i]), 2, iters=5)
            res["base3d_tcgen05_tf32x3"] = dict(ms=med, best_ms=best, GBs=0.0, frac=0.0, TFLOPs=flops / med / 1e9)
            os.environ["DPV_BASE3D_TC"] = "0"
            for tf32 in (False, True):
                torch.backends.cudnn.allow_tf32 = tf32
                for _ in range(2):
                    net(vols[0])
                med, best = timeit(lambda i: net(vols[i]), 2, iters=5)
                res["base3d_torch_cudnn_%s" % ("tf32" if tf32 else "fp32")] = dict(ms=med, best_ms=best, GBs=0.0, frac=0.0,
                                                                                  TFLOPs=flops / med / 1e9)
            torch.backends.cudnn.allow_tf32 = False
            os.environ.pop("DPV_BASE3D_TC")
        for k in ("base3d_tcgen05_tf32x3", "base3d_torch_cudnn_fp32", "base3d_torch_cudnn_tf32"):
            print("%-26s %8.4f ms  %6.1f TFLOP/s (conv flops of the eight layers, batch %d)" % (k, res[k]["ms"], res[k]["TFLOPs"], Bv))
    if want("corr"):
        x1 = [torch.randn((2, 32, 96, 208), device="cuda") for _ in range(nrot)]
        x2 = [torch.randn((2, 32, 96, 208), device="cuda") for _ in range(nrot)]
        f = lambda i: ops.correlation(x1[i], x2[i], 4)
        med, best = timeit(f, nrot)
        byt = 2 * 4 * 96 * 208 * (2 * 32 + 81)
        res["corr_2x32x96x208"] = dict(ms=med, best_ms=best, GBs=byt / med / 1e6, frac=byt / med / 1e6 / PEAK)
    for k, v in res.items():
        print("%-26s %8.4f ms (best %8.4f)  %8.1f GB/s  %5.1f%% of %.0f" % (k, v["ms"], v["best_ms"], v["GBs"], 100 * v["frac"], PEAK))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
