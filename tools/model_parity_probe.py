"""Where does the model-level difference to the reference come from?  (dev tool, GPU box)

Runs the reference BaseModel (baseline/_ref or /root/reference) on cuda:0 and compares, stage by stage,
against the same model with the hot path swapped for our kernels, for several sweep-kernel variants, and
against the reference's own CPU-vs-CUDA spread of est_swp_volume_v4 (the noise floor of the contract).
    python tools/model_parity_probe.py [algo]      # env DPV_SWEEP_TMA_EXACT=0/1 selects the coordinate form
"""
import importlib
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import model_cases as MC  # noqa: E402
from oracle import reference_loader  # noqa: E402

warnings.filterwarnings("ignore")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dpv = importlib.import_module("probabilistic-depth_b200")
ref = reference_loader.load()
algo = int(sys.argv[1]) if len(sys.argv) > 1 else 0
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()


def sc(a, b):
    a, b = a.double(), b.double()
    return float(((a - b).abs() / b.abs().clamp_min(1.0)).max())


name, batch = "default_stereo", 2
torch.manual_seed(0)
model = MC.batch_stat_norm(ref.models.BaseModel(MC.cfg(name), 0)).cuda()
mi = MC.frame_inputs(name, 0, batch)
t = {k: (cu(v) if isinstance(v, np.ndarray) and k != "d_candi" else v) for k, v in mi.items()}
t["prev_output"] = None
ours_h = dpv.warping.homography
orig = ref.homography.est_swp_volume_v4


def est(*a, **k):     # our sweep with the requested algo
    feat_img_ref, feat_img_src, d_candi, R, tt, cam, sigma = a[:7]
    poses = ours_h._pack_poses(R, tt, feat_img_ref.device)
    return dpv.ops.sweep_cost_volume(feat_img_ref, feat_img_src, poses, cam['intrinsic_M_cuda'], cam['unit_ray_array_2D'],
                                     d_candi, sigma, dist="L2", algo=algo)


with torch.no_grad():
    BV_r, cost_r, _, _ = model.forward_encoder(t)
    ref.homography.est_swp_volume_v4 = est
    BV_o, cost_o, _, _ = model.forward_encoder(t)
    ref.homography.est_swp_volume_v4 = orig
    # noise floor: the reference's own function, CPU vs CUDA, on the same features
    rgb = t["rgb"].view(-1, 3, MC.H, MC.W)
    l1, raw, feat = model.base_encoder(rgb)
    fa = torch.cat((feat, torch.nn.functional.avg_pool2d(rgb, 4)), 1).view(batch, 2, 67, 64, 96)
    i = 0
    cam = {"intrinsic_M_cuda": t["intrinsics"][i], "intrinsic_M": t["intrinsics"][i].cpu().numpy(),
           "unit_ray_array_2D": t["unit_ray"][i]}
    camc = {"intrinsic_M_cuda": t["intrinsics"][i].cpu(), "intrinsic_M": t["intrinsics"][i].cpu().numpy(),
            "unit_ray_array_2D": t["unit_ray"][i].cpu()}
    R, tt = t["src_cam_poses"][i, :-1, :3, :3], t["src_cam_poses"][i, :-1, :3, 3]
    c_cuda = orig(fa[i, -1].unsqueeze(0), fa[i, :-1].unsqueeze(0), MC.D_CANDI, R, tt, cam, 10.0)
    c_cpu = orig(fa[i, -1].unsqueeze(0).cpu(), fa[i, :-1].unsqueeze(0).cpu(), MC.D_CANDI, R.cpu(), tt.cpu(), camc, 10.0)
    c_ours = est(fa[i, -1].unsqueeze(0), fa[i, :-1].unsqueeze(0), MC.D_CANDI, R, tt, cam, 10.0)
    # and the conv stack's sensitivity: perturb the reference's cost volume by 1 ulp-scale noise
    lg = lambda c: model.conv0_2(model.conv0_1(model.conv0(c)))
    noise = cost_r * (1 + 6e-8 * torch.randn_like(cost_r))
    BV_n = torch.log_softmax(lg(noise), 1)
rel = lambda a, b: float(((a - b).abs() / b.abs().clamp_min(1e-3)).max())
print("algo", algo, "exact env", os.environ.get("DPV_SWEEP_TMA_EXACT"))
print("cost volume range", float(cost_r.min()), float(cost_r.max()), "BV range", float(BV_r.min()), float(BV_r.max()))
print("cost: ours vs ref-cuda  max rel %.3g  max abs %.3g" % (rel(cost_o, cost_r), float((cost_o - cost_r).abs().max())))
print("cost: ref-cpu vs ref-cuda max rel %.3g  max abs %.3g" % (rel(c_cpu.cuda(), c_cuda), float((c_cpu.cuda() - c_cuda).abs().max())))
print("cost: ours vs ref-cpu   max rel %.3g  max abs %.3g" % (rel(c_ours, c_cpu.cuda()), float((c_ours - c_cpu.cuda()).abs().max())))
print("BV:   ours vs ref-cuda  scaled %.3g  max abs %.3g" % (sc(BV_o, BV_r), float((BV_o - BV_r).abs().max())))
print("BV:   ref with 1-ulp-noise cost (6e-8 rel) vs ref: scaled %.3g  max abs %.3g" % (sc(BV_n, BV_r), float((BV_n - BV_r).abs().max())))
lo, lr = lg(cost_o), lg(cost_r)
print("logits range", float(lr.min()), float(lr.max()), "logit abs diff", float((lo - lr).abs().max()))
