import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dpv = importlib.import_module("probabilistic-depth_b200")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
d = dpv.synth.depth_candidates(5, 40, 64)
x = torch.randn((B, 64, 256, 384), device="cuda")
out = dpv.ops.head(x, d, logp=True, depth=True, variance=True, argmax=True, quarter=True)
torch.cuda.synchronize()
print("ok", float(out["depth"].mean()))
