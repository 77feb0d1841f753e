"""LiDAR -> sparse depth maps.

generate_depth: the goldens (`*_dmap_ref`) come from the reference's OWN source file compiled against minimal
Eigen / OpenCV / pybind11 stand-ins (oracle/ref_utils_lib -> oracle/_ref/libutils_ref.so; those libraries are
absent here); the C restatement oracle/c/lidar_depthmap.c and the kernels are held bit-exact to them.  The
stand-in states the one thing Eigen leaves open (the association of the 4-term products).  minpool is pinned:
the goldens are the reference's own Python applied to the same maps.
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
    assert np.array_equal(dmap, GOLD[name + "_dmap_ref"])                                # the reference's own C++
    assert np.array_equal(L.minpool(dmap, 4, 1000.0), GOLD[name + "_small_ref"])          # reference's minpool
    assert np.array_equal(L.minpool(dmap, 4), GOLD[name + "_small_ref_plain"])


@pytest.mark.skipif(not L.reference_available(), reason="oracle/_ref/libutils_ref.so not built "
                    "(__graft_entry__.build() compiles it where /root/reference exists)")
def test_oracle_vs_compiled_reference_on_random_clouds():
    """The restatement against the reference's own generate_depth (compiled from /root/reference) on seeded random
    point clouds, several image sizes, filter radii 0-3: bit-identical maps."""
    rng = np.random.RandomState(7)
    for it in range(12):
        width, height = int(rng.choice([16, 48, 96, 200])), int(rng.choice([12, 40, 64]))
        n = int(rng.choice([50, 2000, 20000]))
        velo = np.concatenate([rng.uniform(-30, 30, (n, 2)), rng.uniform(-5, 60, (n, 1)), np.ones((n, 1))], 1).astype(np.float32)
        f = float(rng.uniform(0.5, 1.5) * width)
        intr = np.array([[f, 0, width / 2.0, 0], [0, f, height / 2.0, 0], [0, 0, 1, 0]], dtype=np.float32)
        ang = rng.uniform(-0.2, 0.2)
        M = np.eye(4, dtype=np.float32)
        M[:3, :3] = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]], dtype=np.float32)
        M[:3, 3] = rng.uniform(-0.5, 0.5, 3)
        filt, fd = int(rng.randint(0, 4)), float(rng.choice([0.5, 1.0, 2.0]))
        a = L.generate_depth(velo, intr, M, width, height, filt, fd)
        b = L.reference_generate_depth(velo, intr, M, width, height, filt, fd)
        assert np.array_equal(a, b), (it, width, height, n, filt, fd)
        assert (a > 0).any() or n == 50


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
