#!/usr/bin/env python
"""Launch the stand-alone uncertainty-field kernels a few times at the bench shape (for ncu)."""
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
hd = dpv.ops.head(lg, d, logp=True, depth=True)
for _ in range(int(os.environ.get("N", "2"))):
    uf, dz = dpv.ops.ufield(hd["logp"], d, Ku, depth=hd["depth"])
torch.cuda.synchronize()
print("in-band fraction", float((dz != 0).float().mean()), "rows with band", int((dz != 0).any(2).sum()), "of", B * H)
