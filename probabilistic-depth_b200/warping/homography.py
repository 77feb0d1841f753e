"""Drop-in for the reference's warping/homography.py hot-path functions.

Same names, argument meaning and error behaviour as
/root/reference/warping/homography.py (est_swp_volume_v4 :98-135, warp_feature
:137-168, _back_warp_homo_parallel :170-198, get_rel_extrinsicM :260-262); the work
is done by libdpv_sm100a.so.  These per-item signatures exist for compatibility
with callers written against the reference (models/models.py:541,625,
models/packnet.py:380); the batched entry points in ..ops avoid the per-item
Python loop, the D-fold feature copy and the per-call host syncs.
"""
import numpy as np
import torch

from .. import ops


def _pack_poses(R, t, device):
    """[V,3,3], [V,3] -> [1,V,4,4] row-major [R|t]."""
    V = R.shape[0]
    P = torch.zeros((1, V, 4, 4), device=device, dtype=torch.float32)
    P[0, :, :3, :3] = R.to(device=device, dtype=torch.float32)
    P[0, :, :3, 3] = t.to(device=device, dtype=torch.float32).reshape(V, 3)
    P[0, :, 3, 3] = 1.0
    return P


def est_swp_volume_v4(feat_img_ref, feat_img_src, d_candi, R, t, cam_intrinsic, costV_sigma,
                      feat_dist='L2', debug_ipdb=False):
    r'''
    feat_img_ref - NCHW tensor (N = 1)
    feat_img_src - NVCHW tensor.  V is for different views
    R, t - R[idx_view, :, :] - 3x3 rotation matrix
           t[idx_view, :] - 3x1 transition vector
    returns costV [1, D, H, W]
    '''
    if feat_dist not in ('L2', 'L1'):
        raise Exception('undefined metric for feature distance ...')
    device = feat_img_ref.device
    if feat_img_ref.shape[0] != 1 or feat_img_src.shape[0] != 1:
        raise Exception('est_swp_volume_v4 handles one item per call; use ops.sweep_cost_volume for a batch')
    K = cam_intrinsic['intrinsic_M_cuda'].to(device)
    rays = cam_intrinsic['unit_ray_array_2D'].to(device)
    poses = _pack_poses(R, t, device)
    return ops.sweep_cost_volume(feat_img_ref, feat_img_src, poses, K, rays, d_candi,
                                 costV_sigma, dist=feat_dist)


def warp_feature(feat_img_src, d_candi, R, t, cam_intrinsic):
    r'''
    feat_img_src - NVCHW tensor (N = 1, C = len(d_candi)).  Returns [1, V, D, H, W] where
    plane i of view v is channel i of that view warped with depth d_candi[i].
    '''
    if feat_img_src.shape[0] != 1:
        raise Exception("Warped Accum Error")
    device = feat_img_src.device
    K = cam_intrinsic['intrinsic_M_cuda'].to(device)
    rays = cam_intrinsic['unit_ray_array_2D'].to(device)
    poses = _pack_poses(R, t, device)
    return ops.warp_feature(feat_img_src, poses, K, rays, d_candi)


def _back_warp_homo_parallel(img_src, D, term1, term2, cam_intrinsics, H, W, debug_inputs=None):
    r'''
    p_src ~ term1 + term2 * d for every depth d in D; bilinear sampling of img_src
    [len(D), C, H, W] (zeros padding).  term1 [3,1], term2 [3, H*W].
    '''
    u_center = cam_intrinsics['intrinsic_M'][0, 2]
    v_center = cam_intrinsics['intrinsic_M'][1, 2]
    return ops.warp_planes(img_src, D, term1, term2, float(u_center), float(v_center), H, W)


def get_rel_extrinsicM(ext_ref, ext_src):
    ''' Get the extrinisc matrix from ref_view to src_view '''
    return np.asarray(ext_src).dot(np.linalg.inv(np.asarray(ext_ref)))
