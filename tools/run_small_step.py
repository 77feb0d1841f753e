#!/usr/bin/env python
"""One small frame step in every mode (target for compute-sanitizer memcheck / racecheck)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
dpv = importlib.import_module("probabilistic-depth_b200")
frame = importlib.import_module("probabilistic-depth_b200.frame")
s = dpv.synth
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
B, V, C, D, h, w, H, W = 2, 1, 19, 64, 16, 24, 64, 96
d = s.depth_candidates(5, 40, D)
cam = s.camera(w, h, B)
for mode in ("default", "upsample", "feedback"):
    step = frame.FrameStep(B, V, C, D, h, w, H, W, d, mode=mode)
    kw = {}
    if mode == "upsample":
        dm, mk = s.sparse_depth(7, B, h, w)
        kw = dict(dmaps=cu(dm), masks=cu(mk))
    if mode == "feedback":
        kw = dict(feat_raw=cu(s.randn(3, B, V + 1, D, h, w)), bv_resi=cu(s.randn(4, B, D, h, w)))
    step.run(cu(s.randn(1, B, V + 1, C, h, w)), cu(s.mono_poses(B)), cu(cam["intrinsics"]), cu(cam["unit_ray"]),
             cu(s.ground_plane_logits(2, B, H, W, d, cam["intrinsics_up"][0])), cu(cam["intrinsics_up"]), **kw)
    torch.cuda.synchronize()
    print(mode, "ok", float(step.uf[torch.isfinite(step.uf)].sum()))
# wide / many-run sweep shapes (fallback + two passes) and the metrics / lidar kernels
f = cu(s.randn(5, 1, 2, 9, 8, 128))
p = cu(np.stack([np.stack([s.pose(None, (-2.2, 0, 0)), s.pose()])]).astype(np.float32))
c2 = s.camera(128, 8, 1)
dpv.ops.sweep_cost_volume(f[:, -1], f[:, :-1], p[:, :-1], cu(c2["intrinsics"]), cu(c2["unit_ray"]), s.depth_candidates(5, 40, 32), 10.0, algo=4)
f = cu(s.randn(6, 1, 2, 11, 8, 16))
p = cu(np.stack([np.stack([s.pose(None, (-7.0, 0, 0)), s.pose()])]).astype(np.float32))
c2 = s.camera(16, 8, 1)
dpv.ops.sweep_cost_volume(f[:, -1], f[:, :-1], p[:, :-1], cu(c2["intrinsics"]), cu(c2["unit_ray"]), d, 10.0, algo=4, log_softmax=False)
# the tcgen05 convolution chain on a small cost volume
ws = [cu(0.06 * s.randn(20 + i, 64, 64, 3, 3)) for i in range(3)]
bs = [cu(0.1 * s.randn(30 + i, 64)) for i in range(3)]
lp = dpv.ops.CostRefine(ws, bs)(cu(4 * s.randn(40, 2, 64, 16, 24) + 10))
torch.cuda.synchronize()
print("refine ok", float(torch.logsumexp(lp, 1).abs().max()))
# the 3-D convolution stack (Base3D): folded and batch-statistics BatchNorms, residual blocks, 4 -> 32 -> ... -> 1
import cases
g3 = np.load(os.path.join(ROOT, "tests", "golden", "base3d.npz"))
resi = dpv.ops.Base3DConvs(cases.base3d_layers(g3, cu))(cu(g3["volume"]))
torch.cuda.synchronize()
print("base3d ok", float(np.abs(resi.cpu().numpy() - g3["resi"]).max()))
a = cu(np.abs(s.randn(8, 2, 32, 48)) + 1)
dpv.ops.depth_errors(a, a * 1.1)
torch.cuda.synchronize()
print("extras ok")
