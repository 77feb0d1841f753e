#!/usr/bin/env python
"""Frame step in the three nmodes of the reference at BASELINE.json's batch sizes (dev tool; the contract
benchmark is bench.py).  configs[1] default_stereo B=8, configs[2] feedback B=8 (a 16-frame sequence is 16
such steps per item), configs[3] upsample B=32.  Prints frames/s per GPU, inputs resident and rotated."""
import importlib, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dpv = importlib.import_module("probabilistic-depth_b200")
frame = importlib.import_module("probabilistic-depth_b200.frame")
s = dpv.synth
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
V, C, D, h, w, H, W = 1, 67, 64, 64, 96, 256, 384
d = s.depth_candidates(5, 40, D)
# SURVEY 8d bytes per frame
BYTES = {"default": 86138880, "feedback": 98795520, "upsample": 90906624}
out = {}
_w = torch.empty((64, 1024, 1024), device="cuda")
for _ in range(200):
    _w.mul_(1.0001)
for mode, B, pose in (("default", 8, "stereo"), ("feedback", 8, "mono"), ("upsample", 32, "stereo")):
    cam = s.camera(w, h, B)
    step = frame.FrameStep(B, V, C, D, h, w, H, W, d, mode=mode)
    nset = 2 if B <= 8 else 1
    sets = []
    for i in range(nset):
        kw = {}
        if mode == "upsample":
            dm, mk = s.sparse_depth(7 + i, B, h, w)
            kw = dict(dmaps=cu(dm), masks=cu(mk))
        if mode == "feedback":
            kw = dict(feat_raw=torch.randn((B, V + 1, D, h, w), device="cuda"),
                      bv_resi=0.5 * torch.randn((B, D, h, w), device="cuda"))
        one = s.ground_plane_logits(2 + i, 8, H, W, d, cam["intrinsics_up"][0])
        logits = cu(one).repeat(B // 8, 1, 1, 1) if B > 8 else cu(one)
        sets.append((dict(feats=torch.randn((B, V + 1, C, h, w), device="cuda"),
                          poses=cu(s.stereo_poses(B) if pose == "stereo" else s.mono_poses(B)),
                          K=cu(cam["intrinsics"]), rays=cu(cam["unit_ray"]), logits=logits,
                          Ku=cu(cam["intrinsics_up"])), kw))

    def one_step(i):
        a, kw = sets[i % nset]
        step.run(a["feats"], a["poses"], a["K"], a["rays"], a["logits"], a["Ku"], **kw)
    for i in range(10):
        one_step(i)
    torch.cuda.synchronize()
    n = 100
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(n):
        one_step(i)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / n
    fps = B / (ms * 1e-3)
    out[mode] = dict(batch=B, ms_per_step=ms, frames_per_s=fps, frame_hbm_frac_of_6650=BYTES[mode] * fps / 6650e9,
                     launches_per_step=step.launches_per_step())
    print("%-9s B=%2d  %.4f ms/step  %8.0f frames/s  frame HBM fraction %.3f" % (mode, B, ms, fps, out[mode]["frame_hbm_frac_of_6650"]))
    del step, sets
    torch.cuda.empty_cache()
print(json.dumps(out))
