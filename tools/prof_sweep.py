#!/usr/bin/env python
"""A few launches of one sweep algorithm at the bench shape (for ncu).  usage: prof_sweep.py ALGO [mono]"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dpv = importlib.import_module("probabilistic-depth_b200")
ops, s = dpv.ops, dpv.synth
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
algo = int(sys.argv[1])
B, V, C, D, h, w = 8, 1, 67, 64, 64, 96
d = s.depth_candidates(5, 40, D); cam = s.camera(w, h, B)
K, rays = cu(cam["intrinsics"]), cu(cam["unit_ray"])
poses = cu(s.mono_poses(B) if "mono" in sys.argv else s.stereo_poses(B))
feats = torch.randn((B, V + 1, C, h, w), device="cuda")
for _ in range(3):
    ops.sweep_cost_volume(feats[:, -1], feats[:, :-1], poses[:, :-1], K, rays, d, 10.0, algo=algo)
torch.cuda.synchronize()
