#!/bin/bash
# sweep kernel variants (dev tool)
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "sweep" 2>&1 | tail -15
for MB in 6 5; do echo "== DPV_XC_MINB=$MB"; DPV_XC_MINB=$MB python - <<'PY'
import importlib, numpy as np, torch, sys
sys.path.insert(0, '.')
dpv = importlib.import_module("probabilistic-depth_b200")
ops, s = dpv.ops, dpv.synth
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
B, V, C, D, h, w = 8, 1, 67, 64, 64, 96
d = s.depth_candidates(5, 40, D); cam = s.camera(w, h, B)
K, rays = cu(cam["intrinsics"]), cu(cam["unit_ray"])
_w = torch.empty((64, 1024, 1024), device="cuda")
for _ in range(300): _w.mul_(1.0001)
for pose_name in ("stereo", "mono"):
    poses = cu(s.stereo_poses(B) if pose_name == "stereo" else s.mono_poses(B))
    feats = [torch.randn((B, V + 1, C, h, w), device="cuda") for _ in range(3)]
    ref = ops.sweep_cost_volume(feats[0][:, -1], feats[0][:, :-1], poses[:, :-1], K, rays, d, 10.0, algo=1)
    for algo in (5, 4):
        got = ops.sweep_cost_volume(feats[0][:, -1], feats[0][:, :-1], poses[:, :-1], K, rays, d, 10.0, algo=algo)
        err = float(((got - ref).abs() / ref.abs().clamp_min(0.1)).max())
        f = lambda i: ops.sweep_cost_volume(feats[i % 3][:, -1], feats[i % 3][:, :-1], poses[:, :-1], K, rays, d, 10.0, algo=algo)
        for i in range(10): f(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(6): f(i)
        ts = []
        for rep in range(7):
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b_.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b_) / 6)
        ts.sort()
        print("%-6s algo %d: %.4f ms (best %.4f)  max rel err vs direct %.2e" % (pose_name, algo, ts[3], ts[0], err))
PY
done
