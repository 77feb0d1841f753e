"""Pin the CPU oracle against outputs of the reference itself (tests/golden/*.npz).

The reference has no tests or golden vectors of its own for this path
(SURVEY.md section 4); the fixtures were produced by tests/golden/make_golden.py,
which imports the unmodified reference.  The oracle calls the same ATen ops, so
agreement is expected to be bit-exact or within a few ulp.
"""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import dpv_oracle as O

T = torch.from_numpy


def _close(a, b, rtol=1e-6, atol=1e-6):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    np.testing.assert_array_equal(np.isnan(a), np.isnan(b))
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, equal_nan=True)


@pytest.mark.parametrize("name", cases.SWEEP_CASES)
@pytest.mark.parametrize("dist", ["L2", "L1"])
def test_sweep(golden, name, dist):
    g = golden("sweep")
    key = "%s_%s" % (name, dist)
    if key not in g.files:
        pytest.skip("no golden for this combination")
    c = cases.sweep_case(name)
    cv = O.plane_sweep_cost(T(c["ref"]), T(c["src"]), c["d_candi"], T(c["R"]), T(c["t"]),
                            T(c["K"]), T(c["rays"]), c["sigma"], dist)
    _close(cv.numpy(), g[key], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name", cases.SWEEP_WIDE_CASES)
def test_sweep_wide(golden, name):
    """Images wider than 192 px up to the north-star's literal 256x384 / C=67 / D=64 shape and a
    1280-wide large-D slab (round-2 goldens, tests/golden/make_golden_r2.py)."""
    c = cases.sweep_wide_case(name)
    cv = O.plane_sweep_cost(T(c["ref"]), T(c["src"]), c["d_candi"], T(c["R"]), T(c["t"]),
                            T(c["K"]), T(c["rays"]), c["sigma"], "L2")
    _close(cases.sub_view(cv.numpy(), c["sub"]), golden("sweep_wide")[name + "_L2"], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name", cases.WARP_FEATURE_CASES)
def test_warp_feature(golden, name):
    c = cases.warp_feature_case(name)
    wf = O.warp_feature_diag(T(c["feat"]), c["d_candi"], T(c["R"]), T(c["t"]), T(c["K"]),
                             T(c["rays"])).numpy()
    if name != "small":
        wf = wf[:, :, ::4]
    _close(wf, golden("warp_feature")[name])


@pytest.mark.parametrize("name", cases.SOFTMAX_CASES)
def test_softmax_moments(golden, name):
    g = golden("softmax")
    c = cases.softmax_case(name)
    x = T(c["x"])
    ls = O.log_softmax_bins(x)
    np.testing.assert_array_equal(ls.numpy(), g[name + "_logdpv"])
    for b in range(x.shape[0]):
        np.testing.assert_array_equal(
            O.expected_depth(ls[b:b + 1], c["d_candi"], log=True)[0].numpy(), g[name + "_depth"][b])
        np.testing.assert_array_equal(O.depth_variance(ls[b], c["d_candi"]).numpy(),
                                      g[name + "_var"][b])
    np.testing.assert_array_equal(O.argmax_bin(ls).numpy(), g[name + "_argmax"])
    if name + "_quarter" in g.files:
        np.testing.assert_array_equal(O.quarter_nearest(ls).numpy(), g[name + "_quarter"])


@pytest.mark.parametrize("name", cases.FUSE_CASES)
def test_fusion(golden, name):
    g = golden("fuse")
    c = cases.fuse_case(name)
    bv = O.log_softmax_bins(T(c["bv_logits"]))
    prior = O.lidar_prior(T(c["dmaps"]), T(c["masks"]), c["d_candi"], 0.3)
    fused, logf = O.bayes_fuse(bv, prior)
    np.testing.assert_array_equal(logf.numpy(), g[name + "_logfused"])
    if name == "small":
        np.testing.assert_array_equal(prior.numpy(), g[name + "_prior"])
        np.testing.assert_array_equal(fused.numpy(), g[name + "_fused"])
        np.testing.assert_array_equal(O.feedback_fuse(bv, T(c["resi"])).numpy(),
                                      g[name + "_feedback"])


@pytest.mark.parametrize("name", cases.UFIELD_CASES)
def test_ufield(golden, name):
    g = golden("ufield")
    c = cases.ufield_case(name)
    ls = O.log_softmax_bins(T(c["logits"]))
    dpv = ls if c["log"] else torch.exp(ls)
    mask = None if c["mask"] is None else T(c["mask"])
    uf, dz = O.uncertainty_field(dpv, c["d_candi"], T(c["intr_up"]), log=c["log"], mask=mask)
    np.testing.assert_array_equal(np.isnan(uf.numpy()), np.isnan(g[name + "_uf"]))
    np.testing.assert_array_equal(uf.numpy(), g[name + "_uf"])
    np.testing.assert_array_equal(dz.numpy(), g[name + "_depthzero"])
    assert np.isfinite(g[name + "_uf"]).any(), "mask selects nothing: vacuous case"


@pytest.mark.parametrize("name", cases.UFIELD_CFGX_CASES)
def test_ufield_cfgx_quash_limit(golden, name):
    """gen_ufield as the ROS caller invokes it (cfgx -> quash_limit, utils/img_utils.py:269-275,325-332)."""
    g = golden("ufield_cfgx")
    c = cases.ufield_cfgx_case(name)
    ls = O.log_softmax_bins(T(c["logits"]))
    dpv = ls if c["log"] else torch.exp(ls)
    mask = None if c["mask"] is None else T(c["mask"])
    uf, dz = O.uncertainty_field(dpv, c["d_candi"], T(c["intr_up"]), log=c["log"], mask=mask,
                                 params=O.cfgx_params(c["cfgx"]))
    np.testing.assert_array_equal(uf.numpy(), g[name + "_uf"])
    np.testing.assert_array_equal(dz.numpy(), g[name + "_depthzero"])
    assert np.isfinite(g[name + "_uf"]).any(), "mask selects nothing: vacuous case"
    # the gate does something: without it the field differs
    uf0, _ = O.uncertainty_field(dpv, c["d_candi"], T(c["intr_up"]), log=c["log"], mask=mask,
                                 params=dict(O.cfgx_params(c["cfgx"]), quash_limit=False))
    assert not np.array_equal(np.nan_to_num(uf0.numpy()), np.nan_to_num(uf.numpy()))


@pytest.mark.parametrize("name", cases.CORR_CASES)
def test_correlation(golden, name):
    c = cases.corr_case(name)
    y = O.local_correlation(T(c["x1"]), T(c["x2"]), 4)
    np.testing.assert_array_equal(y.numpy(), golden("correlation")[name])


# ------------------------------------------------- second restatement: ATen ops tap by tap
from oracle import dpv_oracle_np as ONP   # noqa: E402


@pytest.mark.parametrize("name", ["mono_small", "mono_yaw_2view", "stereo_small", "oob_heavy", "odd_dims"])
@pytest.mark.parametrize("dist", ["L2", "L1"])
def test_numpy_restatement_sweep(golden, name, dist):
    """grid_sample(bilinear, zeros, align_corners=False) restated in numpy reproduces the
    reference's est_swp_volume_v4 outputs (so the torch-based oracle is not self-referential)."""
    c = cases.sweep_case(name)
    want = golden("sweep")[name + "_" + dist]
    got = ONP.plane_sweep_cost(c["ref"], c["src"], c["d_candi"], c["R"], c["t"], c["K"], c["rays"],
                               c["sigma"], dist)
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-5)


def test_numpy_restatement_warp_feature_and_softmax(golden):
    name = cases.WARP_FEATURE_CASES[0]
    c = cases.warp_feature_case(name)
    got = ONP.warp_feature_diag(c["feat"], c["d_candi"], c["R"], c["t"], c["K"], c["rays"])
    np.testing.assert_allclose(got, golden("warp_feature")[name], rtol=2e-5, atol=2e-5)
    s = cases.softmax_case("small")
    np.testing.assert_allclose(ONP.log_softmax(s["x"], 1), golden("softmax")["small_logdpv"],
                               rtol=1e-5, atol=2e-6)


def test_cost_refine_oracle_vs_the_reference_models_modules():
    """the oracle's conv0 -> conv0_1 -> conv0_2 -> log_softmax (float64) against the reference
    BaseModel's own modules run in fp32 on the CPU (tests/golden/conv_refine.npz, make_golden_r2.py)."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "conv_refine.npz"))
    T = torch.from_numpy
    bv, logits = O.cost_refine(T(g["cost"]), [T(g["w%d" % i]) for i in range(3)], [T(g["b%d" % i]) for i in range(3)],
                               float(g["slope"]))
    scale = float(np.abs(g["logits"]).max())
    assert float((logits - T(g["logits"]).double()).abs().max()) <= 2e-6 * scale      # fp32 summation noise of the reference
    # (1.1e-5 measured: the fp32 convolutions of the reference's CPU build against float64)
    assert float(((bv - T(g["bv"]).double()).abs() / T(g["bv"]).double().abs().clamp_min(1.0)).max()) <= 3e-5


def test_base3d_oracle_vs_the_reference_module():
    """SURVEY 8f rank 2, second half: the oracle's Base3D (float64) against the reference's own Base3D module run in
    fp32 on the CPU (tests/golden/base3d.npz, make_golden_base3d.py) -- including the four BatchNorms of the
    unregistered residual blocks, which normalise with batch statistics even after eval()."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "base3d.npz"))
    layers = cases.base3d_layers(g)
    assert [bool(L["bn"] and L["bn"]["batch_stats"]) for L in layers] == [False, False, True, True, True, True, False, False]
    got = O.base3d(torch.from_numpy(g["volume"]), layers)
    scale = float(np.abs(g["resi"]).max())
    assert tuple(got.shape) == g["resi"].shape
    assert float((got - torch.from_numpy(g["resi"]).double()).abs().max()) <= 1e-5 * scale   # fp32 noise of the reference
