"""LiDAR -> sparse depth maps.

generate_depth is PARITY UNPINNED against the reference (its C++ needs Eigen / OpenCV / pybind11, absent
here): the kernels are held bit-exact to the C restatement oracle/c/lidar_depthmap.c, which fixes the one
thing Eigen leaves open (summation order of the 4-term products).  minpool is pinned: the goldens are the
reference's own Python applied to the same maps.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases  # noqa: E402
from oracle import lidar_depthmap as L  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lidar.npz"))


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("name", cases.LIDAR_CASES)
def test_oracle_vs_golden(name):
    c = cases.lidar_case(name)
    dmap = L.generate_depth(c["velo"], c["intr"], c["M"], c["width"], c["height"], c["filtering"], c["filterdiff"])
    assert np.array_equal(dmap, GOLD[name + "_dmap_oracle"])
    assert np.array_equal(L.minpool(dmap, 4, 1000.0), GOLD[name + "_small_ref"])          # reference's minpool
    assert np.array_equal(L.minpool(dmap, 4), GOLD[name + "_small_ref_plain"])


def test_oracle_filter_and_zbuffer_semantics():
    """Hand-made cases of utils_lib.cpp:121-157: nearest return wins; a return more than `filterdiff`
    behind a neighbour is dropped; the border ring of `filtering` (+1 on the far side) stays empty."""
    M = np.eye(4, dtype=np.float32)
    intr = np.array([[10, 0, 8, 0], [0, 10, 8, 0], [0, 0, 1, 0]], dtype=np.float32)

    def pt(u, v, z):   # lands in pixel (u, v): (int)(x - 0.5)
        return [(u + 0.7 - 8) * z / 10.0, (v + 0.7 - 8) * z / 10.0, z, 1.0]
    velo = np.array([pt(5, 5, 9.0), pt(5, 5, 4.0), pt(6, 5, 7.0), pt(9, 9, 3.0), pt(1, 1, 2.0), pt(5, 5, -3.0)], np.float32)
    d = L.generate_depth(velo, intr, M, 16, 16, filtering=2, filterdiff=1.0)
    assert d[5, 5] == 4.0                  # z-buffer kept the nearer of the two returns
    assert d[5, 6] == 0.0                  # 7.0 is more than 1.0 behind its neighbour 4.0: filtered
    assert d[9, 9] == 3.0
    assert d[1, 1] == 0.0 and (d[:2] == 0).all() and (d[:, :2] == 0).all() and (d[13:] == 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", cases.LIDAR_CASES)
def test_kernels_bit_exact_vs_oracle_and_reference_minpool(dpv, name):
    c = cases.lidar_case(name)
    dmap, small, mask = dpv.ops.lidar_depthmap(cu(c["velo"]), cu(c["intr"]), cu(c["M"]), c["width"], c["height"],
                                               c["filtering"], c["filterdiff"])
    assert np.array_equal(dmap.cpu().numpy(), GOLD[name + "_dmap_oracle"])
    assert np.array_equal(small.cpu().numpy(), GOLD[name + "_small_ref"])
    assert np.array_equal(mask.cpu().numpy(), (GOLD[name + "_small_ref"] >= 0.01).astype(np.float32))
    # the stand-alone minpool and the img_utils mirror, both conventions
    t = dmap.reshape(1, 1, *dmap.shape)
    assert np.array_equal(dpv.utils.img_utils.minpool(t, 4, 1000)[0, 0].cpu().numpy(), GOLD[name + "_small_ref"])
    assert np.array_equal(dpv.utils.img_utils.minpool(t, 4)[0, 0].cpu().numpy(), GOLD[name + "_small_ref_plain"])
    # the pybind-shaped mirror: numpy in, numpy out
    import importlib
    ku = importlib.import_module("probabilistic-depth_b200.external.utils_lib.utils_lib")
    got = ku.generate_depth(c["velo"], c["intr"], c["M"], c["width"], c["height"], {"filtering": 2, "upsample": 0})
    assert isinstance(got, np.ndarray) and np.array_equal(got, GOLD[name + "_dmap_oracle"])
    with pytest.raises(NotImplementedError):
        ku.generate_depth(c["velo"], c["intr"], c["M"], c["width"], c["height"], {"filtering": 2, "upsample": 2})


@pytest.mark.gpu
def test_lidar_feeds_the_upsample_fusion(dpv):
    """dmap_small / mask_small are the `dmaps` / `masks` of the Bayesian fusion (models/models.py:666-672)."""
    c = cases.lidar_case("kitti")
    _, small, mask = dpv.ops.lidar_depthmap(cu(c["velo"]), cu(c["intr"]), cu(c["M"]), c["width"], c["height"])
    D, h, w = 64, small.shape[0], small.shape[1]
    d = dpv.synth.depth_candidates(5, 40, D)
    bv = torch.full((1, D, h, w), float(np.log(1.0 / D)), device="cuda")
    fused, _ = dpv.ops.bayes_fuse(bv, d, dmaps=small.unsqueeze(0), masks=mask.reshape(1, 1, h, w))
    hit = (mask > 0) & (small >= 5) & (small <= 40)
    near = torch.argmin((small.unsqueeze(0) - cu(d.astype(np.float32)).view(D, 1, 1)).abs(), dim=0)
    assert int(hit.sum()) > 100
    assert torch.equal(torch.argmax(fused[0], 0)[hit], near[hit])
    assert float((fused[0][:, mask == 0] - 1.0 / D).abs().max()) < 1e-6
