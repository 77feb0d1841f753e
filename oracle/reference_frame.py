"""TEST / BENCH INFRASTRUCTURE: one batch of frames through the UNMODIFIED reference's own hot-path functions.

`frame_hot_path(ref, ...)` calls the imported reference (oracle/reference_loader.py) in the order its eval
frame does (models/models.py:528-560, :351; trainer/default_trainer.py:221-244,333-336;
utils/img_utils.py:52-61,268-358), on whatever device the input tensors live on: CPU tensors give the
reference's CPU PyTorch path (bench.py --impl reference, cpu_baseline.kind = "reference"), CUDA tensors
give the reference through torch-CUDA on the same B200 -- the incumbent and
the second oracle of SURVEY.md 8c.  The CNN blocks between the functions are out of scope; their
outputs are the synthetic `logits_full` (and, for the 1/4-res soft-max, the cost volume itself, as
models/packnet.py:394 does).  Same step as probabilistic-depth_b200.frame.FrameStep(mode="default").
"""
import torch
import torch.nn.functional as F

from .reference_loader import KittiCfg


def frame_hot_path(ref, feats, poses, K, rays, d_candi, sigma, logits_full, intr_up, want=True):
    """feats [B,V+1,C,h,w] (reference view last), poses [B,V+1,4,4], K [B,3,3], rays [B,3,h*w],
    logits_full [B,D,H,W], intr_up [B,3,3]; d_candi numpy float64 [D].  Returns a dict of batched
    results (or nothing when want=False: timing only)."""
    hom, iu = ref.homography, ref.img_utils
    B = feats.shape[0]
    costs = []
    for i in range(B):                                              # models/models.py:528-550
        cam = {"intrinsic_M_cuda": K[i], "intrinsic_M": K[i].cpu().numpy(), "unit_ray_array_2D": rays[i]}
        costs.append(hom.est_swp_volume_v4(feats[i, -1].unsqueeze(0), feats[i, :-1].unsqueeze(0), d_candi,
                                           poses[i, :-1, :3, :3], poses[i, :-1, :3, 3], cam, sigma,
                                           feat_dist='L2'))
    cost = torch.cat(costs, dim=0)                                  # :552
    bv = F.log_softmax(cost, dim=1)                                 # :560 / packnet.py:394
    refined = F.log_softmax(logits_full, dim=1)                     # :351
    prev = F.interpolate(refined, scale_factor=0.25, mode='nearest')   # default_trainer.py:221-222
    depth, var, amax, uf, dz = [], [], [], [], []
    dd = torch.tensor(d_candi).unsqueeze(1).unsqueeze(1).to(refined.device)   # float64, as :333
    for i in range(B):                                              # the trainer works item by item
        r = refined[i:i + 1]
        depth.append(iu.dpv_to_depthmap(r, d_candi, BV_log=True))   # default_trainer.py:232-233
        z = torch.exp(r.squeeze(0))                                 # :333-336
        mean = torch.sum(dd * z, dim=0)
        var.append(torch.sum(((dd - mean) ** 2) * z, dim=0))
        amax.append(torch.argmax(r, dim=1))
        u, d0 = iu.gen_ufield(r, d_candi, intr_up[i], BV_log=True, cfg=KittiCfg)   # :243 -> img_utils.py:178-181
        uf.append(u)
        dz.append(d0)
    if not want:
        return None
    return dict(cost=cost, bv=bv, refined=refined, quarter=prev, depth=torch.cat(depth), var=torch.stack(var),
                argmax=torch.cat(amax), uf=torch.cat(uf), depth_zero=torch.cat(dz))
