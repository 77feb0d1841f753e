"""Seeded input builders shared by make_golden.py (reference side) and the tests.

Every case is a function seed -> dict of numpy arrays, so the golden files only
need to hold the reference OUTPUTS; inputs are regenerated bit-identically from
the legacy numpy RandomState streams.
"""
import importlib
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
synth = importlib.import_module("probabilistic-depth_b200.synth")

D_CANDI = synth.depth_candidates(5.0, 40.0, 64, 1.0)


def _cam(w, h):
    return synth.intrinsics(w, h), synth.unit_rays(w, h)


def sweep_case(name):
    """Inputs for est_swp_volume_v4.  Returns dict(ref, src, R, t, K, rays, d_candi, sigma)."""
    spec = {
        # name:            (h,  w,  C,  V, D,  pose kind,             seed)
        "mono_small":      (16, 24, 67, 1, 64, "mono", 11),
        "mono_yaw_2view":  (20, 28, 19, 2, 32, "mono2", 12),
        "stereo_small":    (16, 24, 67, 1, 64, "stereo", 13),
        "oob_heavy":       (12, 20, 8, 1, 16, "oob", 14),
        "mono_ref_shape":  (64, 96, 67, 1, 64, "mono", 15),
        "stereo_ref_shape": (64, 96, 67, 1, 64, "stereo", 16),
        "odd_dims":        (13, 17, 5, 1, 7, "mono", 17),
    }[name]
    h, w, C, V, D, kind, seed = spec
    K, rays = _cam(w, h)
    d = synth.depth_candidates(5.0, 40.0, D, 1.0)
    ref = synth.randn(seed, 1, C, h, w)
    src = synth.randn(seed + 1000, 1, V, C, h, w)
    if kind == "mono":
        poses = [synth.pose(synth.yaw_matrix(0.7), (0.05, -0.02, 0.8))]
    elif kind == "mono2":
        poses = [synth.pose(synth.yaw_matrix(0.7), (0.05, -0.02, 0.8)),
                 synth.pose(synth.yaw_matrix(-1.0), (-0.1, 0.03, -0.6))]
    elif kind == "stereo":
        poses = [synth.pose(None, (-0.54, 0.0, 0.0))]
    elif kind == "oob":
        poses = [synth.pose(synth.yaw_matrix(12.0), (1.5, 0.4, 0.5))]
    P = np.stack(poses)
    return dict(ref=ref, src=src, R=np.ascontiguousarray(P[:, :3, :3]),
                t=np.ascontiguousarray(P[:, :3, 3]), K=K, rays=rays, d_candi=d,
                sigma=10.0, h=h, w=w)


def sweep_wide_case(name):
    """Round 2: images wider than 192 px, where dpv_sweep_cost_volume selects the exact-coordinate
    (IEEE division) variant of the TMA kernel, up to the north-star's literal shape (features at 256x384,
    C=67, D=64) and one 1280-wide large-D slab.  `sub` = (row start, row step, col start, col step) of the
    part of the volume the golden file stores."""
    spec = {
        # name:             (h,   w,    C,  V, D,   pose kind, seed, sub)
        "wide_48x256":      (48,  256,  19, 1, 32,  "mono",   71, (0, 1, 0, 1)),
        "wide_stereo_24x320": (24, 320, 12, 1, 32,  "stereo", 72, (0, 1, 0, 1)),
        "stress_256x384":   (256, 384,  67, 1, 64,  "mono",   73, (1, 4, 2, 4)),
        "stress_stereo_256x384": (256, 384, 67, 1, 64, "stereo", 74, (2, 4, 1, 4)),
        "large_d_96x1280":  (96,  1280, 8,  1, 128, "mono",   75, (0, 3, 5, 8)),
    }[name]
    h, w, C, V, D, kind, seed, sub = spec
    K, rays = _cam(w, h)
    d = synth.depth_candidates(5.0, 40.0, D, 1.0)
    ref = synth.randn(seed, 1, C, h, w)
    src = synth.randn(seed + 1000, 1, V, C, h, w)
    if kind == "mono":
        poses = [synth.pose(synth.yaw_matrix(0.7), (0.05, -0.02, 0.8))]
    else:
        poses = [synth.pose(None, (-0.54, 0.0, 0.0))]
    P = np.stack(poses)
    return dict(ref=ref, src=src, R=np.ascontiguousarray(P[:, :3, :3]),
                t=np.ascontiguousarray(P[:, :3, 3]), K=K, rays=rays, d_candi=d,
                sigma=10.0, h=h, w=w, sub=sub)


def sub_view(vol, sub):
    r0, rs, c0, cs = sub
    return vol[..., r0::rs, c0::cs]


SWEEP_WIDE_CASES = ["wide_48x256", "wide_stereo_24x320", "stress_256x384", "stress_stereo_256x384",
                    "large_d_96x1280"]

SWEEP_CASES = ["mono_small", "mono_yaw_2view", "stereo_small", "oob_heavy",
               "mono_ref_shape", "stereo_ref_shape", "odd_dims"]


def warp_feature_case(name):
    spec = {
        "small":     (16, 24, 16, 21),
        "ref_shape": (64, 96, 64, 22),
    }[name]
    h, w, D, seed = spec
    K, rays = _cam(w, h)
    d = synth.depth_candidates(5.0, 40.0, D, 1.0)
    feat = synth.randn(seed, 1, 2, D, h, w)
    P = np.stack([synth.pose(synth.yaw_matrix(0.7), (0.05, -0.02, 0.8)), synth.pose()])
    return dict(feat=feat, R=np.ascontiguousarray(P[:, :3, :3]),
                t=np.ascontiguousarray(P[:, :3, 3]), K=K, rays=rays, d_candi=d, h=h, w=w)


WARP_FEATURE_CASES = ["small", "ref_shape"]


def softmax_case(name):
    spec = {
        "small":    (2, 64, 16, 24, 31, 3.0),
        "wide":     (1, 64, 32, 48, 32, 12.0),
        "d32":      (3, 32, 9, 13, 33, 2.0),
        "d128":     (1, 128, 8, 12, 34, 5.0),
    }[name]
    B, D, h, w, seed, scale = spec
    x = synth.randn(seed, B, D, h, w) * np.float32(scale)
    if name == "wide":
        # exact ties and near-ties for the arg-max rule (first maximum wins)
        x[0, 5, 0, :] = x[0, 9, 0, :] = np.float32(40.0)
        x[0, 7, 1, :] = np.float32(41.0)
        x[0, 3, 1, :] = np.nextafter(np.float32(41.0), np.float32(0.0))
        x[0, :, 2, :] = np.float32(1.25)
    d = synth.depth_candidates(5.0, 40.0, D, 1.0)
    return dict(x=x, d_candi=d)


SOFTMAX_CASES = ["small", "wide", "d32", "d128"]


def fuse_case(name):
    spec = {"small": (2, 64, 16, 24, 41), "ref_shape": (1, 64, 64, 96, 42)}[name]
    B, D, h, w, seed = spec
    d = synth.depth_candidates(5.0, 40.0, D, 1.0)
    bv_logits = synth.randn(seed, B, D, h, w) * np.float32(2.0)
    resi = synth.randn(seed + 1, B, D, h, w)
    dm, mask = synth.sparse_depth(seed + 2, B, h, w, keep=0.3)
    # a few depths outside the bin range (NaN -> -1 -> clamp branch, img_utils.py:45)
    dm[0, 0, :4] = np.float32([0.0, 400.0, 4.9, 41.0])
    mask[0, 0, 0, :4] = 1.0
    return dict(bv_logits=bv_logits, resi=resi, dmaps=dm, masks=mask, d_candi=d)


FUSE_CASES = ["small", "ref_shape"]


def ufield_case(name):
    spec = {
        # name          (H,   W,   seed, with_mask, log)
        "small_log":    (64, 96, 51, False, True),
        "full_log":     (256, 384, 52, False, True),
        "full_mask_lin": (256, 384, 53, True, False),
        "odd_log":      (37, 51, 54, False, True),
    }[name]
    H, W, seed, with_mask, log = spec
    K = synth.intrinsics(W // 4 if W % 4 == 0 else W, H // 4 if H % 4 == 0 else H)
    Ku = synth.intrinsics_up(K) if (W % 4 == 0 and H % 4 == 0) else K
    logits = synth.ground_plane_logits(seed, 1, H, W, D_CANDI, Ku)
    mask = None
    if with_mask:
        mask = (synth.rng(seed + 1).uniform(size=(1, H, W)) < 0.7).astype(np.float32)
    return dict(logits=logits, d_candi=D_CANDI, intr_up=Ku, mask=mask, log=log)


UFIELD_CASES = ["small_log", "full_log", "full_mask_lin", "odd_log"]


def ufield_cfgx_case(name):
    """Round 2: gen_ufield called with `cfgx` as ros/ros_net.py:279 does (utils/img_utils.py:269-275:
    pshift = unc_ang, zstart = unc_shift, span = unc_span, mind = 3, quash_limit -> :325-332)."""
    spec = {
        # name            (H,   W,   seed, with_mask, log,  cfgx)
        "ros_shift5":     (64,  96,  55, False, True,  dict(unc_ang=5, unc_shift=0.6, unc_span=0.3)),
        "ros_noshift":    (128, 192, 56, False, True,  dict(unc_ang=0, unc_shift=0.55, unc_span=0.4)),
        "ros_full_mask":  (256, 384, 57, True,  False, dict(unc_ang=3, unc_shift=0.6, unc_span=0.3)),
        "ros_odd":        (37,  51,  58, False, True,  dict(unc_ang=2, unc_shift=0.5, unc_span=0.5)),
    }[name]
    H, W, seed, with_mask, log, cfgx = spec
    K = synth.intrinsics(W // 4 if W % 4 == 0 else W, H // 4 if H % 4 == 0 else H)
    Ku = synth.intrinsics_up(K) if (W % 4 == 0 and H % 4 == 0) else K
    logits = synth.ground_plane_logits(seed, 1, H, W, D_CANDI, Ku)
    mask = None
    if with_mask:
        mask = (synth.rng(seed + 1).uniform(size=(1, H, W)) < 0.7).astype(np.float32)
    return dict(logits=logits, d_candi=D_CANDI, intr_up=Ku, mask=mask, log=log, cfgx=cfgx)


UFIELD_CFGX_CASES = ["ros_shift5", "ros_noshift", "ros_full_mask", "ros_odd"]


def corr_case(name):
    spec = {
        "tiny":   (2, 8, 6, 13, 61),
        "lvl":    (2, 32, 24, 52, 62),
        "c96":    (1, 96, 12, 26, 63),
    }[name]
    B, C, H, W, seed = spec
    return dict(x1=synth.randn(seed, B, C, H, W), x2=synth.randn(seed + 1, B, C, H, W))


CORR_CASES = ["tiny", "lvl", "c96"]


# ------------------------------------------------------------------ eval metrics
def metrics_case(name):
    """(predicted, truth) depth maps as trainer/default_trainer.py:246-256 hands them to depth_error:
    truth sparse (LiDAR-like, zeros = no return), prediction dense times the truth's mask."""
    spec = {
        # name:        (H,   W,   keep, noise, seed)
        "small":       (16,  24,  0.5,  1.0,   41),
        "quarter":     (64,  96,  0.3,  1.5,   42),
        "full":        (256, 384, 0.2,  2.0,   43),
        "dense":       (37,  51,  1.0,  0.5,   44),
        "one_pixel":   (8,   8,   0.0,  1.0,   45),
    }[name]
    H, W, keep, noise, seed = spec
    r = synth.rng(seed)
    truth = r.uniform(5.0, 45.0, (H, W)).astype(np.float32)
    m = (r.uniform(0, 1, (H, W)) < keep)
    if name == "one_pixel":
        m[3, 5] = True
    truth = (truth * m).astype(np.float32)
    pred = np.clip(truth + noise * r.standard_normal((H, W)).astype(np.float32), 0.5, 60.0).astype(np.float32)
    pred = (pred * m).astype(np.float32)
    return dict(predicted=pred, truth=truth, mask=m.astype(np.float32), d_max=40.0)


METRICS_CASES = ["small", "quarter", "full", "dense", "one_pixel"]


def unc_rmse_case(name):
    """Two linear uncertainty fields [1, D, W] with NaN columns, as compute_unc_field produces them."""
    D, W, seed = {"small": (16, 24, 51), "full": (64, 384, 52)}[name]
    r = synth.rng(seed)
    d = synth.depth_candidates(5.0, 40.0, D, 1.0)

    def field(shift):
        x = r.standard_normal((1, D, W)).astype(np.float32) * 2 + shift
        e = np.exp(x - x.max(1, keepdims=True))
        f = (e / e.sum(1, keepdims=True)).astype(np.float32)
        nan_cols = r.uniform(0, 1, W) < 0.2
        f[:, :, nan_cols] = np.nan
        return f
    return dict(truth=field(0.0), pred=field(0.3), d_candi=d)


UNC_RMSE_CASES = ["small", "full"]


# ------------------------------------------------------------------ LiDAR depth maps
def lidar_case(name):
    """A synthetic Velodyne sweep in front of a KITTI-like camera: points on a ground plane, two walls
    and a few boxes (so that the z-buffer sees occlusions and the filter sees depth steps), some behind
    the camera, some outside the image."""
    n, width, height, seed = {"small": (4000, 96, 64, 61), "kitti": (60000, 384, 256, 62),
                              "sparse": (300, 52, 36, 63)}[name]
    r = synth.rng(seed)
    az = r.uniform(-np.pi, np.pi, n)
    el = r.uniform(-0.43, 0.04, n)
    rng_m = np.minimum(1.73 / np.maximum(np.tan(-el), 1e-3), r.choice([8.0, 15.0, 30.0, 60.0], n))
    rng_m = rng_m * r.uniform(0.98, 1.02, n)
    velo = np.stack([rng_m * np.cos(el) * np.cos(az), rng_m * np.cos(el) * np.sin(az), rng_m * np.sin(el),
                     np.ones(n)], 1).astype(np.float32)
    # velodyne (x fwd, y left, z up) -> camera (x right, y down, z fwd), small lever arm
    M = np.array([[0, -1, 0, 0.02], [0, 0, -1, -0.08], [1, 0, 0, -0.27], [0, 0, 0, 1]], dtype=np.float32)
    fx = 0.75 * width
    intr = np.array([[fx, 0, width / 2.0, 0], [0, fx, height / 2.0, 0], [0, 0, 1, 0]], dtype=np.float32)
    return dict(velo=velo, intr=intr, M=M, width=width, height=height, filtering=2, filterdiff=1.0)


LIDAR_CASES = ["small", "kitti", "sparse"]


def base3d_layers(g, T=None):
    """Layer specs of the Base3D golden (base3d.npz) in the form oracle.dpv_oracle.base3d and ops.Base3DConvs take.
    T: array -> tensor (default torch.from_numpy)."""
    import torch
    T = T or torch.from_numpy
    relu = [True, True, True, False, True, False, True, False]
    block = [None, None, "in", "out", "in", "out", None, None]
    layers = []
    for i in range(8):
        bn = None
        if ("gamma%d" % i) in g:
            bn = dict(gamma=T(g["gamma%d" % i]), beta=T(g["beta%d" % i]), mean=T(g["mean%d" % i]), var=T(g["var%d" % i]),
                      eps=float(g["eps%d" % i]), batch_stats=bool(int(g["batch%d" % i])))
        layers.append(dict(weight=T(g["w%d" % i]), bn=bn, relu=relu[i], block=block[i]))
    return layers
