#!/usr/bin/env python
"""Dev tool: feedback-mode forward on the GPU kernels vs the same model on CPU with oracle ops."""
import sys, importlib, warnings, os
import numpy as np, torch
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import model_cases as MC
from oracle import dpv_oracle as O
dpv = importlib.import_module("probabilistic-depth_b200")
OM = importlib.import_module("probabilistic-depth_b200.models.models")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
name = "feedback_mono"
torch.manual_seed(0); g = MC.batch_stat_norm(OM.BaseModel(MC.cfg(name), 0).cuda())
torch.manual_seed(0); c = MC.batch_stat_norm(OM.BaseModel(MC.cfg(name), 0))
mi = MC.frame_inputs(name, 0)
tg = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) and k != "d_candi" else v) for k, v in mi.items()}
tc = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) and k != "d_candi" else v) for k, v in mi.items()}
d = mi["d_candi"]
with torch.no_grad():
    lg, costg, fg, _, wg = g._encode(tg, True)
    # CPU side with oracle ops
    ops = dpv.ops
    real = (ops.sweep_cost_volume, ops.warp_feature)
    ops.sweep_cost_volume = lambda ref, src, poses, K, rays, d, sigma, dist="L2", **kw: torch.cat([O.plane_sweep_cost(ref[i:i+1], src[i:i+1], d, poses[i,:,:3,:3], poses[i,:,:3,3], K[i], rays[i], sigma, dist) for i in range(ref.shape[0])])
    ops.warp_feature = lambda feat, poses, K, rays, d: torch.cat([O.warp_feature_diag(feat[i:i+1], d, poses[i,:,:3,:3], poses[i,:,:3,3], K[i], rays[i]) for i in range(feat.shape[0])])
    lc, costc, fc, _, wc = c._encode(tc, True)
    ops.sweep_cost_volume, ops.warp_feature = real
    rel = lambda a, b: float(((a.cpu() - b).abs() / b.abs().clamp_min(1.0)).max())
    print("cost", rel(costg, costc), "logits", rel(lg, lc), "warped", rel(wg, wc), "feat", rel(fg[0], fc[0]))
    BVg = ops.head(lg, d, logp=True)["logp"]; BVc = torch.log_softmax(lc, 1)
    print("BV", rel(BVg, BVc))
    prevg = torch.zeros_like(BVg).unsqueeze(1) + 1.0 / 64; prevc = prevg.cpu()
    resig = g.based_3d(torch.cat([BVg.unsqueeze(1), prevg, wg], 1)); resic = c.based_3d(torch.cat([BVc.unsqueeze(1), prevc, wc], 1))
    print("resi", rel(resig, resic), float(resic.abs().max()))
    # same inputs to both 3-D nets: how much do cuDNN conv3d + BN differ from the CPU by themselves
    resig2 = g.based_3d(torch.cat([BVc.unsqueeze(1), prevc, wc], 1).cuda())
    print("resi (identical inputs)", rel(resig2, resic))
    updg = ops.head(BVg, d, addend=resig.contiguous(), logp=True)["logp"]; updc = torch.log_softmax(BVc + resic, 1)
    print("upd", rel(updg, updc))
    upd2 = ops.head(BVc.cuda(), d, addend=resic.cuda().contiguous(), logp=True)["logp"]
    print("upd (identical inputs)", rel(upd2, updc))
