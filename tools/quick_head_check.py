#!/usr/bin/env python
"""Tiny correctness probe of the head kernels against torch ops on the GPU (dev tool)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dpv = importlib.import_module("probabilistic-depth_b200")
torch.manual_seed(0)
for (B, D, H, W) in [(1, 64, 8, 128), (2, 64, 16, 24), (2, 64, 64, 96), (3, 32, 12, 260), (8, 64, 256, 384)]:
    d = dpv.synth.depth_candidates(5, 40, D)
    x = torch.randn((B, D, H, W), device="cuda") * 3
    out = dpv.ops.head(x, d, logp=True, depth=True, variance=True, argmax=True, quarter=(H % 4 == 0))
    torch.cuda.synchronize()
    ref = torch.log_softmax(x, 1)
    dd = torch.tensor(d, dtype=torch.float32, device="cuda").view(1, D, 1, 1)
    e = (ref.exp() * dd).sum(1)
    print(B, D, H, W, "logp err", float((out["logp"] - ref).abs().max()), "depth err", float((out["depth"] - e).abs().max()),
          "argmax eq", bool(torch.equal(out["argmax"], ref.argmax(1))),
          "q eq", bool(torch.equal(out["quarter"], out["logp"][:, :, ::4, ::4])) if "quarter" in out else None, flush=True)
print("ok")
