"""Seeded synthetic inputs for the DPV hot path (tests, goldens, bench).

Everything here is host-side numpy so that the build container (no GPU), the
golden generator (which imports the reference) and the GPU box regenerate
bit-identical inputs from a seed.  The camera follows the recipe the
reference's loader applies to KITTI calibration
(/root/reference/kittiloader/kitti.py:270-320): the principal point is moved to
the image centre and the focal lengths are re-derived from the field of view;
the per-pixel unit rays follow /root/reference/warping/view.py:16-62
(pixel centres at +0.5, z = 1).
"""
import math

import numpy as np

# KITTI raw calibration the survey used.
_KITTI_FX_RAW = 2.0 * 721.5377   # crop_amt = 2 in x
_KITTI_FY_RAW = 721.5377
_KITTI_CX_RAW = 609.5593
_KITTI_CY_RAW = 172.854


def depth_candidates(d_min=5.0, d_max=40.0, n=64, power=1.0):
    """Depth-bin centres, float64 (reference utils/img_utils.py:80-85)."""
    x = np.power(np.linspace(0.0, 1.0, num=n), power)
    return d_min + (d_max - d_min) * x


def fov_deg():
    hf = math.degrees(math.atan(_KITTI_CX_RAW / _KITTI_FX_RAW) * 2.0)
    vf = math.degrees(math.atan(_KITTI_CY_RAW / _KITTI_FY_RAW) * 2.0)
    return hf, vf


def intrinsics(width, height):
    """3x3 float32 K for a `width` x `height` grid (kitti.py:284-293)."""
    hf, vf = fov_deg()
    K = np.zeros((3, 3), dtype=np.float64)
    K[0, 0] = (width / 2.0) / math.tan(math.radians(hf / 2.0))
    K[1, 1] = (height / 2.0) / math.tan(math.radians(vf / 2.0))
    K[0, 2] = width / 2.0
    K[1, 2] = height / 2.0
    K[2, 2] = 1.0
    return K.astype(np.float32)


def intrinsics_up(K, scale=4.0):
    """Full-resolution intrinsics (kittiloader/batch_scheduler.py:194-195)."""
    Ku = (K * np.float32(scale)).astype(np.float32)
    Ku[2, 2] = 1.0
    return Ku


def unit_rays(width, height):
    """[3, height*width] float32 rays (x, y, 1) (view.py:16-62, kitti.py:309-310)."""
    hf, vf = fov_deg()
    th = math.tan(math.radians(hf / 2.0))
    tv = math.tan(math.radians(vf / 2.0))
    xs = th * ((2.0 * ((np.arange(width) + 0.5) / width)) - 1.0)
    ys = tv * ((2.0 * ((np.arange(height) + 0.5) / height)) - 1.0)
    rays = np.empty((3, height, width), dtype=np.float64)
    rays[0] = xs[None, :]
    rays[1] = ys[:, None]
    rays[2] = 1.0
    return rays.reshape(3, -1).astype(np.float32)


def yaw_matrix(deg):
    a = math.radians(deg)
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]], dtype=np.float64)


def pose(R=None, t=(0.0, 0.0, 0.0)):
    T = np.eye(4, dtype=np.float64)
    if R is not None:
        T[:3, :3] = R
    T[:3, 3] = np.asarray(t, dtype=np.float64)
    return T.astype(np.float32)


def mono_poses(batch, yaw_deg=0.7, t=(0.05, -0.02, 0.8)):
    """[B, 2, 4, 4]: source (t-1) pose then identity (models/models.py:530-535)."""
    P = np.stack([pose(yaw_matrix(yaw_deg), t), pose()])
    return np.repeat(P[None], batch, axis=0).astype(np.float32)


def stereo_poses(batch, baseline=0.54, side="left"):
    """Other eye -> this eye, pure x translation (batch_scheduler.py:84-94)."""
    tx = -baseline if side == "left" else baseline
    P = np.stack([pose(None, (tx, 0.0, 0.0)), pose()])
    return np.repeat(P[None], batch, axis=0).astype(np.float32)


def rng(seed):
    # Legacy RandomState: its streams are frozen across numpy versions.
    return np.random.RandomState(seed)


def randn(seed, *shape):
    return rng(seed).standard_normal(shape).astype(np.float32)


def camera(width, height, batch=1):
    """Dict of batched camera tensors as the model input dict carries them."""
    K = intrinsics(width, height)
    return {
        "intrinsics": np.repeat(K[None], batch, axis=0),
        "intrinsics_up": np.repeat(intrinsics_up(K)[None], batch, axis=0),
        "unit_ray": np.repeat(unit_rays(width, height)[None], batch, axis=0),
    }


def sparse_depth(seed, batch, height, width, keep=0.1, d_min=5.0, d_max=40.0):
    """LiDAR-like sparse depth + mask for the upsample mode."""
    r = rng(seed)
    mask = (r.uniform(size=(batch, 1, height, width)) < keep).astype(np.float32)
    dm = r.uniform(d_min, d_max, size=(batch, height, width)).astype(np.float32)
    return dm * mask[:, 0], mask


def ground_plane_logits(seed, batch, height, width, d_candi, K_up, cam_height=0.75,
                        noise=1.0, sharp=0.3):
    """Pre-softmax logits whose softmax peaks on a flat road below the camera.

    Gives the uncertainty-field mask (utils/img_utils.py:316) something to
    select: depth = cam_height * fy / (y - cy) clamped to the bin range, turned
    into a Gaussian bump over the bins plus seeded noise.
    """
    fy, cy = float(K_up[1, 1]), float(K_up[1, 2])
    ys = np.arange(height, dtype=np.float64)
    with np.errstate(divide="ignore"):
        depth = cam_height * fy / np.maximum(ys - cy, 1e-6)
    depth = np.clip(depth, d_candi[0], d_candi[-1])
    bump = -((d_candi[:, None] - depth[None, :]) ** 2) / (2.0 * sharp)      # [D, H]
    logits = np.repeat(bump[None, :, :, None], batch, axis=0)
    logits = np.repeat(logits, width, axis=3)
    logits = logits + noise * rng(seed).standard_normal(logits.shape)
    return np.clip(logits, -60.0, 0.0 + 10.0).astype(np.float32)
