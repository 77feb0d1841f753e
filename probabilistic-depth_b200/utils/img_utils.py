"""Drop-in for the DPV helpers of the reference's utils/img_utils.py.

Mirrors dpv_to_depthmap (:52-61), powerf (:80-85), compute_unc_field (:178-181),
gen_ufield (:268-358) and gen_dpv_withmask (:360-375) of
/root/reference/utils/img_utils.py with the same names, arguments and error
behaviour; the arithmetic runs in libdpv_sm100a.so.
"""
import numpy as np
import torch

from .. import ops

epsilon = torch.finfo(float).eps


def powerf(d_min, d_max, nDepth, power):
    x = np.power(np.linspace(start=0, stop=1, num=nDepth), power)
    return np.array([d_min + (d_max - d_min) * v for v in x])


def dpv_to_depthmap(dpv, d_candi, BV_log=False):
    if dpv.shape[0] != 1:
        raise Exception('Unable to handle this case')
    out = ops.head(dpv, d_candi, mode="logprob" if BV_log else "prob", logp=False, depth=True)
    return out["depth"]


def depth_variance(dpv, d_candi, BV_log=True):
    """Var[d] per pixel (trainer/default_trainer.py:333-336 computes it inline, in float64)."""
    out = ops.head(dpv, d_candi, mode="logprob" if BV_log else "prob", logp=False, variance=True)
    return out["variance"]


def gen_dpv_withmask(dmaps, masks, d_candi, var=0.3):
    return ops.lidar_prior(dmaps, masks, d_candi, var)


def _ufield_params(cfg, cfgx):
    if cfgx is not None:
        zstart = cfgx["unc_shift"]
        return dict(pshift=cfgx["unc_ang"], zstart=zstart, zend=zstart + cfgx["unc_span"],
                    maxd=100., mind=3.), True
    if "kitti" in cfg.data.dataset_path:
        return dict(pshift=5, zstart=0.6, zend=0.6 + 0.3, maxd=100., mind=0.), False
    if "ilim" in cfg.data.dataset_path:
        return dict(pshift=0, zstart=1.0, zend=1.0 + 0.3, maxd=100., mind=3.), True
    raise UnboundLocalError("gen_ufield: dataset_path names neither kitti nor ilim")


def gen_ufield(dpv_predicted, d_candi, intr_up, visualizer=None, img=None, BV_log=True,
               normalize=False, mask=None, cfg=None, cfgx=None):
    params, quash_limit = _ufield_params(cfg, cfgx)
    if quash_limit:
        raise NotImplementedError("gen_ufield: the quash_limit branch (ILIM / cfgx) is outside the "
                                  "KITTI hot path and has no kernel yet")
    if normalize:
        raise NotImplementedError("gen_ufield: normalize=True (visualisation only) has no kernel")
    if dpv_predicted.shape[0] != 1:
        raise Exception('Unable to handle this case')
    uf, depth_zero = ops.ufield(dpv_predicted, d_candi, intr_up,
                                mode="logprob" if BV_log else "prob", mask=mask, params=params)
    return uf, depth_zero


def compute_unc_field(dpv_refined_predicted, dpv_refined_truth, d_candi, intr_refined,
                      mask_refined, cfg):
    unc_field_truth, _ = gen_ufield(dpv_refined_truth, d_candi, intr_refined.squeeze(0),
                                    BV_log=False, mask=mask_refined, cfg=cfg)
    unc_field_predicted, debugmap = gen_ufield(dpv_refined_predicted, d_candi,
                                               intr_refined.squeeze(0), BV_log=True, cfg=cfg)
    return unc_field_truth, unc_field_predicted, debugmap
