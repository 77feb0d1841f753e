import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dpv = importlib.import_module("probabilistic-depth_b200")
sh = importlib.import_module("probabilistic-depth_b200.sharding")
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
s = dpv.synth
for (D, world, C, h, w) in [(128, 4, 16, 384, 1280), (64, 2, 16, 64, 96), (128, 4, 16, 96, 320)]:
    B = 1
    d = s.depth_candidates(5, 40, D)
    cam = s.camera(w, h, B)
    feats = cu(s.randn(800 + D, B, 2, C, h, w))
    poses = cu(s.mono_poses(B).astype(np.float32))
    K, rays = cu(cam["intrinsics"]), cu(cam["unit_ray"])
    res = {}
    for algo in (4, 3, 2, 1):
        res[algo] = dpv.ops.sweep_cost_volume(feats[:, -1], feats[:, :-1], poses[:, :-1], K, rays, d, 10.0, algo=algo)
    torch.cuda.synchronize()
    for algo in (4, 3, 2):
        e = ((res[algo] - res[1]).abs() / res[1].abs().clamp_min(1.0))
        print(D, h, w, "algo", algo, "vs 1: max", float(e.max()), "frac>1e-4", float((e > 1e-4).float().mean()),
              "argmax idx", np.unravel_index(int(e.argmax()), e.shape))
    parts = [sh.plane_sharded_sweep(feats[:, -1], feats[:, :-1], poses[:, :-1], K, rays, d, 10.0, r, world)[0] for r in range(world)]
    e = ((torch.cat(parts, 1) - res[4]).abs() / res[4].abs().clamp_min(1.0))
    print("   sharded vs full: max", float(e.max()), "frac>1e-5", float((e > 1e-5).float().mean()))
