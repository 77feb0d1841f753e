#!/usr/bin/env python
"""Launch each hot-path kernel a few times at the bench shape (for ncu captures)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dpv = importlib.import_module("probabilistic-depth_b200")
frame = importlib.import_module("probabilistic-depth_b200.frame")
B, V, C, D, h, w, H, W = 8, 1, 67, 64, 64, 96, 256, 384
s = dpv.synth
d = s.depth_candidates(5, 40, D)
cam = s.camera(w, h, B)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
pose = s.stereo_poses(B) if "mono" not in sys.argv else s.mono_poses(B)
feats = torch.randn((B, V + 1, C, h, w), device="cuda")
logits = cu(s.ground_plane_logits(2, B, H, W, d, cam["intrinsics_up"][0]))
step = frame.FrameStep(B, V, C, D, h, w, H, W, d)
args = (feats, cu(pose), cu(cam["intrinsics"]), cu(cam["unit_ray"]), logits, cu(cam["intrinsics_up"]))
n = int(os.environ.get("N", "3"))
for _ in range(n):
    step.run(*args)
if "algo2" in sys.argv:
    p = cu(pose)
    for _ in range(n):
        dpv.ops.sweep_cost_volume(feats[:, -1], feats[:, :-1], p[:, :-1], args[2], args[3], d, 10.0, algo=2)
torch.cuda.synchronize()
print("done")
