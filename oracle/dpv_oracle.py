"""CPU oracle for the DPV hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A restatement, in plain PyTorch CPU ops, of what soulslicer/probabilistic-depth
computes on the depth-probability-volume path.  Only tests/, bench.py's
cpu_baseline / --impl reference leg and __graft_entry__.smoke() may import it;
the shipped package (probabilistic-depth_b200/) never does and has no CPU
fallback.

Parity status: the reference ships NO golden vectors, known-answer tests or
fixtures for this path.  This oracle is therefore
pinned against outputs of the reference itself: tests/golden/make_golden.py
imports the unmodified reference from /root/reference in the build container,
runs its functions on seeded synthetic inputs and commits the results under
tests/golden/*.npz; tests/test_oracle_golden.py checks every function below
against those files.

The bilinear gather and the soft-max arithmetic the reference relies on live in
PyTorch ATen (F.grid_sample, F.log_softmax), not in the reference tree, and the
reference pins no torch version; the oracle calls the same ATen ops of the torch
in this image (2.11), i.e. grid_sample with align_corners=False.  A second,
tap-by-tap numpy restatement of those ATen ops is in oracle/dpv_oracle_np.py.

Each function cites the reference file:line it follows (paths relative to
/root/reference).
"""
import numpy as np
import torch
import torch.nn.functional as F

# utils/img_utils.py:12 -- float64 machine epsilon used as an fp32 clamp floor.
EPSILON = float(np.finfo(np.float64).eps)


# --------------------------------------------------------------------------- K1
def sweep_terms(K, R, t, rays):
    """term1 = K t, term2 = (K R) rays   (warping/homography.py:119-121)."""
    term1 = K.matmul(t).reshape(3, 1)
    term2 = K.matmul(R).matmul(rays)
    return term1, term2


def sweep_grid(term1, term2, d, cx, cy, H, W):
    """Normalised sampling grid [D,H,W,2] (warping/homography.py:183-196).

    p_src ~ term1 + term2 * d, perspective divide with the +1e-10 guard, then
    (u - cx) / cx, (v - cy) / cy with cx, cy taken from the intrinsics as fp32.
    """
    nd = d.numel()
    P = term1.unsqueeze(0) + term2.unsqueeze(0).expand(nd, -1, -1) * d.reshape(nd, 1, 1)
    P = P / (P[:, 2, :].unsqueeze(1) + 1e-10)
    gx = (P[:, 0, :] - cx) / cx
    gy = (P[:, 1, :] - cy) / cy
    return torch.stack((gx, gy), dim=-1).reshape(nd, H, W, 2)


def back_warp_planes(src, d, term1, term2, cx, cy, H, W):
    """Bilinear plane-sweep warp (warping/homography.py:170-198).

    src [C,H,W] is sampled once per depth plane -> [D,C,H,W]; zeros padding,
    align_corners=False (the default the unmodified reference gets on torch>=1.3).
    The reference materialises src D times with .repeat (:123); so do we, the
    CPU baseline is meant to cost what the reference costs.
    """
    nd = d.numel()
    grid = sweep_grid(term1, term2, d, cx, cy, H, W)
    stack = src.unsqueeze(0).repeat(nd, 1, 1, 1)
    return F.grid_sample(stack, grid, mode="bilinear", padding_mode="zeros",
                         align_corners=False)


# -------------------------------------------------------------------------- K2a
def plane_sweep_cost(ref, src, d_candi, R, t, K, rays, sigma, dist="L2"):
    """Plane-sweep cost volume [1,D,H,W] (warping/homography.py:98-135, 80-86).

    ref [1,C,H,W]; src [1,V,C,H,W]; R [V,3,3]; t [V,3]; K [3,3]; rays [3,H*W];
    d_candi float64 numpy (cast to fp32 like :115).
    """
    H, W = ref.shape[2], ref.shape[3]
    d = torch.from_numpy(np.asarray(d_candi).astype(np.float32))
    cx = K[0, 2].item()
    cy = K[1, 2].item()
    cx, cy = np.float32(cx), np.float32(cy)
    cost = torch.zeros(1, d.numel(), H, W)
    for v in range(src.shape[1]):
        term1, term2 = sweep_terms(K, R[v], t[v], rays)
        warped = back_warp_planes(src[0, v], d, term1, term2, cx, cy, H, W)
        if dist == "L2":
            per_plane = torch.sum((warped - ref) ** 2, 1)
        elif dist == "L1":
            per_plane = torch.sum(torch.abs(warped - ref), 1)
        else:
            raise Exception("undefined metric for feature distance ...")
        cost[0] = cost[0] + per_plane / sigma
    return cost


# -------------------------------------------------------------------------- K4a
def warp_feature_diag(feat, d_candi, R, t, K, rays):
    """Per-plane warp keeping channel k on plane k (warping/homography.py:137-168).

    feat [1,V,C,H,W] with C == D  ->  [1,V,D,H,W].
    """
    if feat.shape[0] != 1:
        raise Exception("Warped Accum Error")
    H, W = feat.shape[3], feat.shape[4]
    d = torch.from_numpy(np.asarray(d_candi).astype(np.float32))
    cx, cy = np.float32(K[0, 2].item()), np.float32(K[1, 2].item())
    out = torch.zeros(feat.shape)
    for v in range(feat.shape[1]):
        term1, term2 = sweep_terms(K, R[v], t[v], rays)
        warped = back_warp_planes(feat[0, v], d, term1, term2, cx, cy, H, W)
        idx = torch.arange(d.numel())
        out[0, v] = warped[idx, idx]
    return out


# --------------------------------------------------------------------------- K3
def log_softmax_bins(x):
    """models/models.py:351,560,637,694 -- log-softmax over the depth-bin axis."""
    return F.log_softmax(x, dim=1)


def expected_depth(dpv, d_candi, log=False):
    """E[d] in fp32, batch must be 1 (utils/img_utils.py:52-61)."""
    if dpv.shape[0] != 1:
        raise Exception("Unable to handle this case")
    z = dpv.squeeze(0)
    if log:
        z = torch.exp(z)
    dd = torch.tensor(np.asarray(d_candi)).unsqueeze(1).unsqueeze(1).float()
    return torch.sum(dd * z, dim=0).unsqueeze(0)


def depth_variance(log_dpv, d_candi):
    """Var[d] in float64 (trainer/default_trainer.py:333-336); log_dpv [D,H,W]."""
    z = torch.exp(log_dpv)
    dd = torch.tensor(np.asarray(d_candi)).unsqueeze(1).unsqueeze(1)      # float64
    mean = torch.sum(dd * z, dim=0)
    return torch.sum(((dd - mean) ** 2) * z, dim=0)


def argmax_bin(dpv):
    """MAP bin.  Not in the reference: torch.argmax, first max wins."""
    return torch.argmax(dpv, dim=1)


def quarter_nearest(x):
    """Feedback hand-off (trainer/default_trainer.py:221-222)."""
    return F.interpolate(x, scale_factor=0.25, mode="nearest")


# -------------------------------------------------------------------------- K4b
def feedback_fuse(bv_cur, bv_resi):
    """models/models.py:694 -- multiply in probability space and renormalise."""
    return F.log_softmax(bv_cur + bv_resi, dim=1)


# -------------------------------------------------------------------------- K4c
def soft_label(d_candi, depthmap, variance, zero_invalid=False):
    """Gaussian soft label over the bins (utils/img_utils.py:24-47)."""
    dd = torch.tensor(np.asarray(d_candi)).float().unsqueeze(-1).unsqueeze(-1)
    dd = dd.repeat(1, depthmap.shape[0], depthmap.shape[1])
    sig = torch.sqrt(variance)
    dists = torch.exp(-torch.pow(torch.abs(dd - depthmap), 2.0) / (2 * torch.pow(sig, 2.0)))
    dists = dists / torch.sum(dists, dim=0)
    if zero_invalid:
        dists[dists != dists] = -1
    return dists


def lidar_prior(dmaps, masks, d_candi, var=0.3):
    """Prior DPV from a sparse depth map (utils/img_utils.py:360-375, 49-50)."""
    nb = len(d_candi)
    vv = torch.tensor(var)
    rows = []
    for b in range(dmaps.shape[0]):
        m = masks[b, 0].unsqueeze(0)
        g = soft_label(d_candi, dmaps[b], vv, zero_invalid=True)
        uni = torch.ones((nb, dmaps.shape[1], dmaps.shape[2])) / nb
        rows.append((g * m + uni * (1.0 - m)).unsqueeze(0))
    prior = torch.cat(rows)
    return torch.clamp(prior, EPSILON, 1.0)


def bayes_fuse(bv_cur, prior):
    """models/models.py:669-672 -> (fused_dpv, log fused_dpv)."""
    fused = torch.exp(bv_cur + torch.log(prior))
    fused = fused / torch.sum(fused, dim=1).unsqueeze(1)
    fused = torch.clamp(fused, EPSILON, 1.0)
    return fused, torch.log(fused)


# --------------------------------------------------------------------------- K5
KITTI_UF = dict(pshift=5, zstart=0.6, zend=0.6 + 0.3, maxd=100.0, mind=0.0)


def _shift_grids(H, W, pshift):
    """utils/img_utils.py:170-176 + 295-299: normalised grids for +/- pshift rows."""
    yv, xv = torch.meshgrid([torch.arange(0, H).float(), torch.arange(0, W).float()],
                            indexing="ij")
    ystep = 2.0 / float(H - 1)
    xstep = 2.0 / float(W - 1)
    grids = []
    for s in (float(pshift), -float(pshift)):
        flow = torch.zeros((1, H, W, 2)).float()
        flow[:, :, :, 1] = s
        flow[0, :, :, 0] = -1 + xv * xstep - flow[0, :, :, 0] * xstep
        flow[0, :, :, 1] = -1 + yv * ystep - flow[0, :, :, 1] * ystep
        grids.append(flow)
    return grids


def depth_to_points(depth, intr):
    """utils/img_utils.py:111-135; depth [1,H,W] -> [3,H,W]."""
    d = depth[0]
    fx, cx, fy, cy = intr[0, 0], intr[0, 2], intr[1, 1], intr[1, 2]
    yf, xf = torch.meshgrid([torch.arange(0, d.shape[0]).float(),
                             torch.arange(0, d.shape[1]).float()], indexing="ij")
    yf = (yf - cy) / fy
    xf = (xf - cx) / fx
    return torch.cat([(xf * d).unsqueeze(0), (yf * d).unsqueeze(0), d.unsqueeze(0)], 0)


def cfgx_params(cfgx):
    """utils/img_utils.py:269-275: parameters of a `cfgx` call (ros/ros_net.py:279)."""
    return dict(pshift=cfgx["unc_ang"], zstart=cfgx["unc_shift"], zend=cfgx["unc_shift"] + cfgx["unc_span"],
                maxd=100.0, mind=3.0, quash_limit=True)


def uncertainty_field(dpv, d_candi, intr_up, log=True, mask=None, params=None):
    """Road-surface uncertainty-field collapse (utils/img_utils.py:268-358).

    dpv [1,D,H,W] (log-probabilities when log=True); intr_up [3,3]; optional
    mask [1,H,W].  params defaults to the KITTI constants (:277-283); params["quash_limit"]
    (the `cfgx` caller ros/ros_net.py:279 and ILIM data, :269-275,289-290) adds the column gate of
    :325-332.  Returns (UF [1,D,W], depth * mask [1,H,W]).
    """
    p = dict(KITTI_UF if params is None else params)
    pshift, zstart, zend, maxd, mind = p["pshift"], p["zstart"], p["zend"], p["maxd"], p["mind"]
    H, W = dpv.shape[2], dpv.shape[3]
    if pshift != 0:
        g_fwd, g_inv = _shift_grids(H, W, pshift)
        shifted = F.grid_sample(dpv, g_fwd, mode="nearest", align_corners=False)
    else:
        shifted = dpv.clone()
    depth_s = expected_depth(shifted, d_candi, log=log)
    depth_p = expected_depth(dpv, d_candi, log=log)
    pts = depth_to_points(depth_s, intr_up)
    zmask = (~((pts[1] > zend) | (pts[1] < zstart) | (pts[2] > maxd - 1) | (pts[2] < mind))).float()
    if mask is not None:
        if pshift != 0:
            ms = F.grid_sample(mask.unsqueeze(1), g_fwd, mode="nearest",
                               align_corners=False).squeeze(1)
        else:
            ms = mask.clone()
        zmask = zmask * ms.squeeze(0)
    if p.get("quash_limit", False):                       # :325-332
        cleaned = (depth_s * zmask).squeeze(0)
        cleaned[cleaned == 0] = 1000
        col_min, _ = torch.min(cleaned, dim=0)
        zmask = zmask * ((cleaned > col_min - 1.0) & (cleaned < col_min + 1.0)).float()
    if pshift != 0:
        zmask_p = F.grid_sample(zmask.unsqueeze(0).unsqueeze(0), g_inv, mode="nearest",
                                align_corners=False).squeeze(0).squeeze(0)
    else:
        zmask_p = zmask.clone()
    depth_zero = depth_p * zmask_p
    zm = zmask_p.unsqueeze(0).unsqueeze(0)
    probs = torch.exp(dpv) if log else dpv
    plane = torch.sum(probs * zm, dim=2)
    plane = plane / torch.sum(zmask, dim=0)
    return plane, depth_zero


# -------------------------------------------------------------------------- K2b
def local_correlation(x1, x2, max_displacement=4):
    """(2r+1)^2-displacement correlation, channel mean (models/correlation_native.py:13-23)."""
    r = max_displacement
    n = 2 * r + 1
    B, C, H, W = x1.size()
    x2p = F.pad(x2, [r] * 4)
    planes = []
    for i in range(n):
        for j in range(n):
            planes.append(torch.mean(x1 * x2p[:, :, i:i + H, j:j + W], 1, keepdim=True))
    return torch.cat(planes, 1)


# --------------------------------------------------------------- 8f rank 2
def cost_refine(cost, weights, biases, slope=0.01):
    """conv0 -> LeakyReLU -> conv0_1 -> LeakyReLU -> conv0_2 -> log_softmax over the depth bins
    (models/models.py:456-460 definitions with conv2d_leakyRelu :38-46, applied at :555-560 and :632-637):
    Conv2d(D, D, 3, stride 1, padding 1, bias), nn.LeakyReLU() default slope 0.01.  Evaluated in float64 so that
    the oracle carries no summation-order noise of its own.  Returns (log-DPV, conv0_2 output) as float64."""
    x = cost.double()
    for i in range(3):
        x = F.conv2d(x, weights[i].double(), biases[i].double(), stride=1, padding=1)
        if i < 2:
            x = F.leaky_relu(x, slope)
    return F.log_softmax(x, dim=1), x


def base3d(volume, layers):
    """Base3D.forward(volume, prob=False) (models/models.py:376-438, called at :693): dres0 -> residual blocks ->
    classify, every convolution Conv3d(k=3, stride 1, padding 1, bias=False) (models/models.py:31-36, :403), float64.
    layers: list of dicts {weight, bn: None | dict(gamma, beta, mean, var, eps, batch_stats), relu, block: None | "in"
    | "out"}: "in" = first layer of a residual block (its input is the skip), "out" = the layer the skip is added to
    (:421-422: curr = dres(curr) + curr, no ReLU after the sum).  batch_stats: the BatchNorm normalises with the
    biased statistics of the batch (training mode / track_running_stats=False), else with its running statistics."""
    x = volume.double()
    skip = None
    for L in layers:
        if L.get("block") == "in":
            skip = x
        x = F.conv3d(x, L["weight"].double(), None, stride=1, padding=1)
        bn = L.get("bn")
        if bn is not None:
            g, b = bn["gamma"].double(), bn["beta"].double()
            if bn["batch_stats"]:
                x = F.batch_norm(x, None, None, g, b, True, 0.0, float(bn["eps"]))
            else:
                x = F.batch_norm(x, bn["mean"].double(), bn["var"].double(), g, b, False, 0.0, float(bn["eps"]))
        if L.get("block") == "out":
            x = x + skip
            skip = None
        if L.get("relu"):
            x = F.relu(x)
    return x.squeeze(1)


# ------------------------------------------------------------- whole-frame port
def frame_hot_path(ref, src, d_candi, R, t, K, rays, sigma, logits_quarter, logits_full,
                   intr_up):
    """One `default`/stereo frame through the hot-path functions, CPU port.

    Mirrors the order of models/models.py:504-565,644-656 and
    trainer/default_trainer.py:221-244,333-336 with the CNN blocks left out (their
    outputs are the synthetic `logits_*` inputs):
    cost volume -> log-softmax (1/4 res) -> log-softmax (full res) -> E[d], Var,
    argmax -> 1/4 nearest hand-off -> uncertainty field.
    `logits_quarter` may be None to soft-max the cost volume itself
    (models/packnet.py:394).
    """
    cost = plane_sweep_cost(ref, src, d_candi, R, t, K, rays, sigma, "L2")
    bv = log_softmax_bins(cost if logits_quarter is None else logits_quarter)
    refined = log_softmax_bins(logits_full)
    depth_q = expected_depth(bv, d_candi, log=True)
    depth = expected_depth(refined, d_candi, log=True)
    var = depth_variance(refined[0], d_candi)
    amax = argmax_bin(refined)
    prev = quarter_nearest(refined)
    uf, depth_zero = uncertainty_field(refined, d_candi, intr_up, log=True)
    return dict(cost=cost, bv=bv, refined=refined, depth_q=depth_q, depth=depth, var=var,
                argmax=amax, prev=prev, uf=uf, depth_zero=depth_zero)
