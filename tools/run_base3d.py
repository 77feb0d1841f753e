#!/usr/bin/env python
"""Dev tool: N passes of the Base3D stack at the model's shape (for ncu).  BASE3D_B = batch (default 2)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dpv = importlib.import_module("probabilistic-depth_b200")
M = importlib.import_module("probabilistic-depth_b200.models.models")
B = int(os.environ.get("BASE3D_B", "2"))
net = M.Base3D(4, dres_count=2, feature_dim=32, bn_running_avg=True).cuda().eval()
net.dres_modules = [b.cuda() for b in net.dres_modules]
vol = torch.randn((B, 4, 64, 64, 96), device="cuda")
with torch.no_grad():
    tc = dpv.ops.Base3DConvs.from_module(net)
    for _ in range(int(os.environ.get("N", "2"))):
        out = tc(vol)
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
