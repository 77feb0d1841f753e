"""Tap-by-tap numpy restatement of the two ATen ops the DPV path rests on  --  TEST
INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/dpv_oracle.py for the rules and the pinning).

oracle/dpv_oracle.py calls F.grid_sample and F.log_softmax of the torch in this image, exactly
as the reference does (warping/homography.py:197, models/models.py:351,560).  So that the oracle
does not rest on "torch agrees with torch", this file restates what those two library calls
compute, following ATen's CPU kernels:

  * bilinear grid_sample, padding_mode="zeros", align_corners=False
      torch/include/ATen/native/GridSampler.h:27-36 (unnormalise: ((g + 1) * size - 1) / 2),
      aten/src/ATen/native/GridSampler.cpp (floor, the four weights
      (ix_se - ix)(iy_se - iy) ..., out-of-bounds taps contribute 0, accumulation order nw, ne, sw, se)
  * log_softmax over one axis: x - max - log(sum(exp(x - max)))

and builds the plane-sweep cost volume and the diagonal feature warp from them in float32 with
the reference's operation order (warping/homography.py:98-198).  tests/test_oracle_golden.py
checks it against the reference-generated fixtures next to the torch-based oracle.
"""
import numpy as np

F32 = np.float32


def unnormalize(g, size):
    """ATen grid_sampler_unnormalize, align_corners=False, in float32."""
    g = g.astype(F32)
    return ((g + F32(1.0)) * F32(size) - F32(1.0)) / F32(2.0)


def grid_sample_bilinear_zeros(img, grid):
    """img [N,C,H,W] float32, grid [N,Ho,Wo,2] (x, y in [-1, 1]) -> [N,C,Ho,Wo]."""
    img = img.astype(F32)
    N, C, H, W = img.shape
    ix = unnormalize(grid[..., 0], W)
    iy = unnormalize(grid[..., 1], H)
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    x1 = x0 + F32(1.0)
    y1 = y0 + F32(1.0)
    # weights as ATen forms them
    w_nw = (x1 - ix) * (y1 - iy)
    w_ne = (ix - x0) * (y1 - iy)
    w_sw = (x1 - ix) * (iy - y0)
    w_se = (ix - x0) * (iy - y0)
    out = np.zeros((N, C) + ix.shape[1:], dtype=F32)
    n_idx = np.arange(N).reshape(N, 1, 1)
    for xs, ys, wt in ((x0, y0, w_nw), (x1, y0, w_ne), (x0, y1, w_sw), (x1, y1, w_se)):
        finite = np.isfinite(xs) & np.isfinite(ys)
        xi = np.where(finite, xs, -1).astype(np.int64)
        yi = np.where(finite, ys, -1).astype(np.int64)
        ok = finite & (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
        xi = np.clip(xi, 0, W - 1)
        yi = np.clip(yi, 0, H - 1)
        tap = img[n_idx, :, yi, xi]                        # [N,Ho,Wo,C]
        tap = np.where(ok[..., None], tap, F32(0.0))
        out = out + np.moveaxis(tap * wt[..., None].astype(F32), -1, 1)
    return out.astype(F32)


def log_softmax(x, axis=1):
    x = x.astype(F32)
    m = np.max(x, axis=axis, keepdims=True)
    s = np.sum(np.exp(x - m), axis=axis, keepdims=True, dtype=F32)
    return (x - m) - np.log(s)


def sweep_grid(K, R, t, rays, d, H, W):
    """warping/homography.py:119-121,183-196 in float32: [D,H,W,2] normalised grid."""
    K, R, t, rays, d = (np.asarray(a, dtype=F32) for a in (K, R, t, rays, d))
    term1 = (K @ t.reshape(3, 1)).astype(F32)
    term2 = ((K @ R) @ rays).astype(F32)
    P = term1[None] + term2[None] * d.reshape(-1, 1, 1)
    P = (P / (P[:, 2:3, :] + F32(1e-10))).astype(F32)
    cx, cy = K[0, 2], K[1, 2]
    gx = (P[:, 0, :] - cx) / cx
    gy = (P[:, 1, :] - cy) / cy
    return np.stack([gx, gy], -1).reshape(len(d), H, W, 2).astype(F32)


def plane_sweep_cost(ref, src, d_candi, R, t, K, rays, sigma, dist="L2"):
    """est_swp_volume_v4 (warping/homography.py:98-135): ref [1,C,H,W], src [1,V,C,H,W]."""
    ref = np.asarray(ref, dtype=F32)
    src = np.asarray(src, dtype=F32)
    H, W = ref.shape[2:]
    d = np.asarray(d_candi).astype(F32)
    cost = np.zeros((1, len(d), H, W), dtype=F32)
    for v in range(src.shape[1]):
        grid = sweep_grid(K, R[v], t[v], rays, d, H, W)
        stack = np.broadcast_to(src[0, v][None], (len(d),) + src.shape[2:])
        warped = grid_sample_bilinear_zeros(stack, grid)
        diff = warped - ref
        per_plane = np.sum(diff * diff if dist == "L2" else np.abs(diff), axis=1, dtype=F32)
        cost[0] = cost[0] + per_plane / F32(sigma)
    return cost


def warp_feature_diag(feat, d_candi, R, t, K, rays):
    """warp_feature (warping/homography.py:137-168): feat [1,V,D,H,W] -> [1,V,D,H,W]."""
    feat = np.asarray(feat, dtype=F32)
    H, W = feat.shape[3:]
    d = np.asarray(d_candi).astype(F32)
    out = np.zeros_like(feat)
    for v in range(feat.shape[1]):
        grid = sweep_grid(K, R[v], t[v], rays, d, H, W)
        # plane k only ever needs channel k
        img = feat[0, v][:, None]                         # [D,1,H,W]
        out[0, v] = grid_sample_bilinear_zeros(img, grid)[:, 0]
    return out
