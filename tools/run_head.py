#!/usr/bin/env python
"""Launch the full-resolution head a few times at the bench shape (for ncu captures)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dpv = importlib.import_module("probabilistic-depth_b200")
B, D, H, W = 8, 64, 256, 384
s = dpv.synth
d = s.depth_candidates(5, 40, D)
cam = s.camera(W // 4, H // 4, B)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
lg = cu(s.ground_plane_logits(2, B, H, W, d, cam["intrinsics_up"][0]))
Ku = cu(cam["intrinsics_up"])
n = int(os.environ.get("N", "3"))
for _ in range(n):
    dpv.ops.head(lg, d, logp=True, depth=True, variance=True, argmax=True, quarter=True)
    if "uf" in sys.argv:
        dpv.ops.head_ufield(lg, d, Ku, logp=True, depth=True, variance=True, argmax=True, quarter=True)
torch.cuda.synchronize()
print("done")
