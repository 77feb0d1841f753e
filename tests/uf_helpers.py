"""Shared by the GPU uncertainty-field tests: which columns hold a rounding-sensitive pixel.

gen_ufield (reference utils/img_utils.py:268-358) takes per-pixel yes/no decisions against fixed
thresholds; a pixel whose height / depth sits within fp32 rounding of a threshold may legitimately land
on the other side when E[d] is summed in a different order, and it then moves its whole column by a
discrete amount.  The tests compare every OTHER column at 1e-4 and require those to be > 90 % of all.
"""
import torch

from oracle import dpv_oracle as O


def near_threshold_columns(dpvv, d_candi, intr_up, log, params=None, rel=1e-4):
    """dpvv [1,D,H,W] torch CPU (log-probabilities when log); returns a bool numpy array [W]."""
    p = dict(O.KITTI_UF if params is None else params)
    H, W = dpvv.shape[2:]
    if p["pshift"] != 0:
        g_fwd, _ = O._shift_grids(H, W, p["pshift"])
        shifted = torch.nn.functional.grid_sample(dpvv, g_fwd, mode="nearest", align_corners=False)
    else:
        shifted = dpvv
    z = O.expected_depth(shifted, d_candi, log=log)
    pts = O.depth_to_points(z, intr_up)
    bad = torch.zeros((H, W), dtype=torch.bool)
    for val, thr in ((pts[1], p["zend"]), (pts[1], p["zstart"]), (pts[2], p["maxd"] - 1), (pts[2], p["mind"])):
        gap = (val - thr).abs()
        bad |= (gap > 0) & (gap <= rel * max(1.0, abs(thr)))   # exact hits (zero padding) are not rounding-sensitive
    if p.get("quash_limit", False) or p.get("quash_range", 0):
        zmask = (~((pts[1] > p["zend"]) | (pts[1] < p["zstart"]) | (pts[2] > p["maxd"] - 1) | (pts[2] < p["mind"]))).float()
        cleaned = (z * zmask).squeeze(0)
        cleaned[cleaned == 0] = 1000
        cmin, _ = torch.min(cleaned, dim=0)
        for edge in (cmin - 1.0, cmin + 1.0):
            gap = (cleaned - edge).abs()
            bad |= (gap <= rel * edge.abs().clamp_min(1.0)) & (cleaned < 999)
    # the mask is shifted back up and keeps its column: same column index
    return bad.any(0).numpy()
