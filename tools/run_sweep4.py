#!/usr/bin/env python
"""One launch of the TMA sweep kernel on a small case (for compute-sanitizer / debugging)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dpv = importlib.import_module("probabilistic-depth_b200")
from oracle import dpv_oracle as O
s = dpv.synth
B, V, C, D, h, w = 2, 1, 67, 64, 16, 24
d = s.depth_candidates(5, 40, D)
cam = s.camera(w, h, B)
feats = s.randn(1, B, V + 1, C, h, w)
poses = s.stereo_poses(B) if "mono" not in sys.argv else s.mono_poses(B)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
f, p = cu(feats), cu(poses)
algo = int(os.environ.get("ALGO", "4"))
cost = dpv.ops.sweep_cost_volume(f[:, -1], f[:, :-1], p[:, :-1], cu(cam["intrinsics"]), cu(cam["unit_ray"]), d, 10.0, algo=algo)
torch.cuda.synchronize()
T = torch.from_numpy
for b in range(B):
    want = O.plane_sweep_cost(T(feats[b:b + 1, -1]), T(feats[b:b + 1, :-1]), d, T(poses[b, :-1, :3, :3]),
                              T(poses[b, :-1, :3, 3]), T(cam["intrinsics"][b]), T(cam["unit_ray"][b]), 10.0, "L2")
    err = ((cost[b:b + 1].cpu() - want).abs() / want.abs().clamp_min(1)).max()
    print("item", b, "max rel err", float(err))
