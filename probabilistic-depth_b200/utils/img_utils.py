"""Drop-in for the DPV helpers of the reference's utils/img_utils.py.

Mirrors eval_errors (:14-15), depth_error (:17-22), dpv_to_depthmap (:52-61), powerf (:80-85),
compute_unc_field (:178-181), compute_unc_rmse (:183-202), gen_ufield (:268-358) and
gen_dpv_withmask (:360-375) of
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


def minpool(tensor, scale, default=0):
    """utils/img_utils.py:87-95."""
    return ops.minpool(tensor, scale, default)


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


QUASH_RANGE = 1.0      # utils/img_utils.py:326


def _ufield_params(cfg, cfgx):
    """utils/img_utils.py:269-290: the `cfgx` dict of the ROS caller (ros/ros_net.py:279) or the
    dataset constants; quash_limit (:325-332) for cfgx and ILIM."""
    if cfgx is not None:
        zstart = cfgx["unc_shift"]
        return dict(pshift=cfgx["unc_ang"], zstart=zstart, zend=zstart + cfgx["unc_span"],
                    maxd=100., mind=3., quash_range=QUASH_RANGE)
    if "kitti" in cfg.data.dataset_path:
        return dict(pshift=5, zstart=0.6, zend=0.6 + 0.3, maxd=100., mind=0.)
    if "ilim" in cfg.data.dataset_path:
        return dict(pshift=0, zstart=1.0, zend=1.0 + 0.3, maxd=100., mind=3., quash_range=QUASH_RANGE)
    raise UnboundLocalError("gen_ufield: dataset_path names neither kitti nor ilim")


def gen_ufield(dpv_predicted, d_candi, intr_up, visualizer=None, img=None, BV_log=True,
               normalize=False, mask=None, cfg=None, cfgx=None):
    params = _ufield_params(cfg, cfgx)
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


# ---------------------------------------------------------------- eval metrics
def _as_cuda(a):
    if isinstance(a, torch.Tensor):
        if not a.is_cuda:
            raise RuntimeError("depth_error: tensors must live on a CUDA device (no CPU path)")
        return a.float()
    # the reference passes .cpu().numpy() arrays (trainer/default_trainer.py:255-256); take them back
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def depth_error(predicted, truth):
    """utils/img_utils.py:17-22.  [H,W] arrays or CUDA tensors -> list of the nine metrics, as the
    reference's pybind call returns it.  (ops.depth_errors keeps a whole batch on the device.)"""
    out, counts = ops.depth_errors(_as_cuda(predicted), _as_cuda(truth), want_counts=True)
    if int(counts[0]) == 0:
        # evaluate_depth.h:92-95 prints this and throws
        raise RuntimeError("ERROR: Ground truth defect => Please write me an email!")
    return [float(v) for v in out[0].cpu()]


def eval_errors(errors):
    """utils/img_utils.py:14-15 -> evaluateErrors (external/deval_lib/src/evaluate_depth.h:121-142):
    {metric: [mean, min, max]} over the per-item vectors.  Host-side bookkeeping in the reference too
    (a std::vector of 9-vectors); kept in float32 with utils.h's accumulators -- the sum runs in item
    order, the minimum starts at 1 and the maximum at 0 (utils.h:24-56)."""
    errs = np.asarray(errors, dtype=np.float32).reshape(-1, 9)
    out = {}
    for j, name in enumerate(ops.METRIC_NAMES):
        col = errs[:, j]
        mean = np.cumsum(col, dtype=np.float32)[-1] / np.float32(len(col))
        mn = np.float32(1)
        mx = np.float32(0)
        for v in col:
            if v < mn:
                mn = v
            if v > mx:
                mx = v
        out[name] = [float(mean), float(mn), float(mx)]
    return out


def compute_unc_rmse(unc_field_truth, unc_field_predicted, d_candi, plot=False):
    """utils/img_utils.py:183-202 (the plot branch needs matplotlib and is visualisation only)."""
    if plot:
        raise NotImplementedError("compute_unc_rmse: plot=True is visualisation only")
    return ops.unc_rmse(unc_field_truth, unc_field_predicted, d_candi)[0]
