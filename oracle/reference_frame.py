"""TEST / BENCH INFRASTRUCTURE: one batch of frames through the UNMODIFIED reference's own hot-path functions.

`frame_hot_path(ref, ...)` calls the imported reference (oracle/reference_loader.py) in the order its eval
frame does (models/models.py:528-560, :351, :666-672, :686-694; trainer/default_trainer.py:221-244,333-336;
utils/img_utils.py:52-61,268-358), on whatever device the input tensors live on: CPU tensors give the
reference's CPU PyTorch path (bench.py --impl reference, cpu_baseline.kind = "reference"), CUDA tensors
give the reference through torch-CUDA on the same B200 -- the incumbent and
the second oracle of SURVEY.md 8c.  The CNN blocks between the functions are out of scope; their
outputs are the synthetic `logits_full` / `bv_resi` (and, for the 1/4-res soft-max, the cost volume
itself, as models/packnet.py:394 does).  Same step as probabilistic-depth_b200.frame.FrameStep.
"""
import torch
import torch.nn.functional as F

from .reference_loader import KittiCfg


def sweep_batch(ref, feats, poses, K, rays, d_candi, sigma):
    """models/models.py:528-552: est_swp_volume_v4 item by item, concatenated."""
    hom = ref.homography
    costs = []
    for i in range(feats.shape[0]):
        cam = {"intrinsic_M_cuda": K[i], "intrinsic_M": K[i].cpu().numpy(), "unit_ray_array_2D": rays[i]}
        costs.append(hom.est_swp_volume_v4(feats[i, -1].unsqueeze(0), feats[i, :-1].unsqueeze(0), d_candi,
                                           poses[i, :-1, :3, :3], poses[i, :-1, :3, 3], cam, sigma,
                                           feat_dist='L2'))
    return torch.cat(costs, dim=0)


def head_batch(ref, logits_full, d_candi, intr_up, uf=True):
    """The decoder's final log-softmax (models/models.py:351) and what the eval loop derives from it item by
    item (trainer/default_trainer.py:221-244,333-336)."""
    iu = ref.img_utils
    refined = F.log_softmax(logits_full, dim=1)
    prev = F.interpolate(refined, scale_factor=0.25, mode='nearest')            # default_trainer.py:221-222
    depth, var, amax, ufs, dz = [], [], [], [], []
    dd = torch.tensor(d_candi).unsqueeze(1).unsqueeze(1).to(refined.device)     # float64, as :333
    for i in range(refined.shape[0]):
        r = refined[i:i + 1]
        depth.append(iu.dpv_to_depthmap(r, d_candi, BV_log=True))               # :232-233
        z = torch.exp(r.squeeze(0))                                             # :333-336
        mean = torch.sum(dd * z, dim=0)
        var.append(torch.sum(((dd - mean) ** 2) * z, dim=0))
        amax.append(torch.argmax(r, dim=1))
        if uf:
            u, d0 = iu.gen_ufield(r, d_candi, intr_up[i], BV_log=True, cfg=KittiCfg)   # :243 -> img_utils.py:178-181
            ufs.append(u)
            dz.append(d0)
    out = dict(refined=refined, quarter=prev, depth=torch.cat(depth), var=torch.stack(var), argmax=torch.cat(amax))
    if uf:
        out.update(uf=torch.cat(ufs), depth_zero=torch.cat(dz))
    return out


def frame_hot_path(ref, feats, poses, K, rays, d_candi, sigma, logits_full, intr_up, mode="default",
                   dmaps=None, masks=None, feat_raw=None, bv_resi=None, want=True, refine=None, prev=None,
                   base3d=None):
    """feats [B,V+1,C,h,w] (reference view last), poses [B,V+1,4,4], K [B,3,3], rays [B,3,h*w],
    logits_full [B,D,H,W], intr_up [B,3,3]; d_candi numpy float64 [D].  mode "upsample": dmaps [B,h,w],
    masks [B,1,h,w]; mode "feedback": feat_raw [B,V+1,D,h,w], bv_resi [B,D,h,w].  Returns a dict of batched
    results (or nothing when want=False: timing only).  refine: optional callable cost -> logits (the model's
    conv0 / conv0_1 / conv0_2 modules, models/models.py:555-557) applied before the 1/4-res soft-max.  base3d:
    optional callable volume -> residual (the model's Base3D, models/models.py:692-693) fed with
    cat(BV, prev, warped); bv_resi is then ignored."""
    hom, iu = ref.homography, ref.img_utils
    cost = sweep_batch(ref, feats, poses, K, rays, d_candi, sigma)
    bv = F.log_softmax(cost if refine is None else refine(cost), dim=1)         # models.py:560 / packnet.py:394
    extra = {}
    if mode == "upsample":                                                      # models.py:666-672
        prior = iu.gen_dpv_withmask(dmaps, masks, d_candi, 0.3)
        fused = torch.exp(bv + torch.log(prior))
        fused = fused / torch.sum(fused, dim=1).unsqueeze(1)
        fused = torch.clamp(fused, iu.epsilon, 1.)
        extra = dict(fused=fused, logfused=torch.log(fused))
    elif mode == "feedback":                                                    # models.py:614-627,694
        warped = []
        for i in range(feats.shape[0]):
            cam = {"intrinsic_M_cuda": K[i], "intrinsic_M": K[i].cpu().numpy(), "unit_ray_array_2D": rays[i]}
            warped.append(hom.warp_feature(feat_raw[i].unsqueeze(0), d_candi, poses[i, :, :3, :3],
                                           poses[i, :, :3, 3], cam))
        warped = torch.cat(warped, dim=0)
        if base3d is not None:                                                  # models.py:692-693
            bv_resi = base3d(torch.cat([bv.unsqueeze(1), prev.unsqueeze(1), warped], dim=1))
        extra = dict(warped=warped, bv_upd=F.log_softmax(bv + bv_resi, dim=1))
    out = head_batch(ref, logits_full, d_candi, intr_up)
    if not want:
        return None
    out.update(cost=cost, bv=bv, **extra)
    return out


def stress_hot_path(ref, feats, poses, K, rays, d_candi, sigma, want=True):
    """K1-K3 directly on full-resolution features (SURVEY.md 8d stress shape; the packnet-style pipeline,
    models/packnet.py:380-394): cost volume, log-softmax over the planes, E[d] / Var / arg-max."""
    cost = sweep_batch(ref, feats, poses, K, rays, d_candi, sigma)
    out = head_batch(ref, cost, d_candi, None, uf=False)
    if not want:
        return None
    out["cost"] = cost
    return out
