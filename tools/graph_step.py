#!/usr/bin/env python
"""The frame step launched call by call vs replayed from CUDA graphs (is the step host-bound?)."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dpv = importlib.import_module("probabilistic-depth_b200")
frame = importlib.import_module("probabilistic-depth_b200.frame")
s = dpv.synth
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
B, V, C, D, h, w, H, W = 8, 1, 67, 64, 64, 96, 256, 384
d = s.depth_candidates(5, 40, D)
cam = s.camera(w, h, B)
step = frame.FrameStep(B, V, C, D, h, w, H, W, d, mode="default")
sets = []
for i in range(2):
    sets.append(dict(feats=cu(s.randn(1 + 1000 * i, B, V + 1, C, h, w)), poses=cu(s.stereo_poses(B)), K=cu(cam["intrinsics"]),
                     rays=cu(cam["unit_ray"]), logits=cu(s.ground_plane_logits(2 + 1000 * i, B, H, W, d, cam["intrinsics_up"][0])),
                     Ku=cu(cam["intrinsics_up"])))
def run(i):
    a = sets[i % 2]
    step.run(a["feats"], a["poses"], a["K"], a["rays"], a["logits"], a["Ku"])
for i in range(10): run(i)
torch.cuda.synchronize()
n = 200
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
c0 = time.perf_counter()
t0.record()
for i in range(n): run(i)
t1.record()
c1 = time.perf_counter()
torch.cuda.synchronize()
print("call by call: GPU %.4f ms/step, host enqueue %.4f ms/step" % (t0.elapsed_time(t1) / n, 1e3 * (c1 - c0) / n))
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for i in range(4): run(i)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    run(0); run(1)
g.replay(); torch.cuda.synchronize()
t0.record()
for i in range(n // 2): g.replay()
t1.record(); torch.cuda.synchronize()
print("graph replay : GPU %.4f ms/step" % (t0.elapsed_time(t1) / n))
