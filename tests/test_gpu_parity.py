"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the golden
fixtures produced by the reference itself.

Tolerances (BASELINE.json north_star: "argmax depth-bin indices bit-exact, and DPV, expected
depth and variance within 1e-4 relative in fp32"):
  cost volume / warps       rtol 1e-4 (+ atol 1e-5 for values near zero)
  log-DPV                   |y - y_ref| <= 1e-4 * max(1, |y_ref|)
  E[d], Var[d]              rtol 1e-4 (Var oracle is float64)
  argmax                    exact
  UF                        same NaN pattern, rtol 1e-4 on columns without a pixel within 1e-4 of a
                            mask threshold (a flipped pixel changes a column discretely)
  correlation               atol 2e-7 + rtol 1e-6 (the reference's own bar is atol 1e-7 between its
                            CUDA and torch versions on C>=128 inputs, i.e. |values| ~ 0.1,
                            models/correlation_native.py:64; summation order differs, 2 ulp)
"""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import dpv_oracle as O
from uf_helpers import near_threshold_columns

pytestmark = pytest.mark.gpu
T = torch.from_numpy


def cu(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def close(a, b, rtol=1e-4, atol=1e-5):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    np.testing.assert_array_equal(np.isnan(a), np.isnan(b))
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, equal_nan=True)


def logclose(a, b, tol=1e-4):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape
    err = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    assert err.max() <= tol, "max scaled error %g" % err.max()


def cam_dict(c):
    return {"intrinsic_M_cuda": cu(c["K"]), "intrinsic_M": c["K"], "unit_ray_array_2D": cu(c["rays"])}


# ------------------------------------------------------------------------- K1 + K2a
@pytest.mark.parametrize("name", cases.SWEEP_CASES)
@pytest.mark.parametrize("dist,algo", [("L2", 0), ("L2", 5), ("L2", 4), ("L2", 3), ("L2", 2), ("L2", 1), ("L1", 1)])
def test_sweep_vs_golden(dpv, golden, name, dist, algo):
    g = golden("sweep")
    if algo in (4, 5) and cases.sweep_case(name)["w"] % 4 != 0:
        pytest.skip("the TMA kernel needs 16-byte row strides; algo=0 falls back (covered above)")
    key = "%s_%s" % (name, dist)
    c = cases.sweep_case(name)
    if key in g.files:
        want = g[key]
    else:
        want = O.plane_sweep_cost(T(c["ref"]), T(c["src"]), c["d_candi"], T(c["R"]), T(c["t"]),
                                  T(c["K"]), T(c["rays"]), c["sigma"], dist).numpy()
    V = c["src"].shape[1]
    poses = np.zeros((1, V, 4, 4), np.float32)
    poses[0, :, :3, :3] = c["R"]
    poses[0, :, :3, 3] = c["t"]
    poses[0, :, 3, 3] = 1
    got = dpv.ops.sweep_cost_volume(cu(c["ref"]), cu(c["src"]), cu(poses), cu(c["K"]), cu(c["rays"]),
                                    c["d_candi"], c["sigma"], dist=dist, algo=algo)
    close(got, want)


@pytest.mark.parametrize("name", cases.SWEEP_WIDE_CASES)
@pytest.mark.parametrize("algo", [0, 5, 4, 1])
def test_sweep_wide_vs_golden(dpv, golden, name, algo):
    """Images wider than 192 px -- the exact-coordinate variant of the TMA kernel
    (sweep_gram_tma_kernel<.,EXACT>) -- against the reference's own outputs, up to the north-star's
    literal shape (features at 256x384, C=67, D=64) and a 1280-wide, D=128 slab."""
    c = cases.sweep_wide_case(name)
    V = c["src"].shape[1]
    poses = np.zeros((1, V, 4, 4), np.float32)
    poses[0, :, :3, :3] = c["R"]
    poses[0, :, :3, 3] = c["t"]
    poses[0, :, 3, 3] = 1
    got = dpv.ops.sweep_cost_volume(cu(c["ref"]), cu(c["src"]), cu(poses), cu(c["K"]), cu(c["rays"]),
                                    c["d_candi"], c["sigma"], dist="L2", algo=algo)
    close(cases.sub_view(got, c["sub"]), golden("sweep_wide")[name + "_L2"])


@pytest.mark.parametrize("name", ["mono_small", "mono_yaw_2view", "oob_heavy"])
def test_est_swp_volume_v4_mirror(dpv, golden, name):
    """Through the reference-shaped function (warping/homography.py:98)."""
    c = cases.sweep_case(name)
    got = dpv.warping.homography.est_swp_volume_v4(cu(c["ref"]), cu(c["src"]), c["d_candi"],
                                                   cu(c["R"]), cu(c["t"]), cam_dict(c), c["sigma"],
                                                   feat_dist="L2")
    close(got, golden("sweep")[name + "_L2"])
    with pytest.raises(Exception, match="undefined metric"):
        dpv.warping.homography.est_swp_volume_v4(cu(c["ref"]), cu(c["src"]), c["d_candi"], cu(c["R"]),
                                                 cu(c["t"]), cam_dict(c), c["sigma"], feat_dist="L3")


def test_sweep_batched_strided_and_fused_softmax(dpv):
    """One launch over B items laid out as the model holds them ([B, 2, C, h, w], reference view
    last, models/models.py:519-535), per-item poses, against the per-item oracle."""
    B, C, h, w, D = 3, 67, 16, 24, 64
    synth = dpv.synth
    d = synth.depth_candidates(5, 40, D)
    feats = synth.randn(71, B, 2, C, h, w)
    poses = np.stack([
        np.stack([synth.pose(synth.yaw_matrix(0.7), (0.05, -0.02, 0.8)), synth.pose()]),
        np.stack([synth.pose(None, (-0.54, 0, 0)), synth.pose()]),
        np.stack([synth.pose(synth.yaw_matrix(-2.0), (0.3, 0.1, -0.4)), synth.pose()])]).astype(np.float32)
    cam = synth.camera(w, h, B)
    f = cu(feats)
    p = cu(poses)
    cost, lsm = dpv.ops.sweep_cost_volume(f[:, -1], f[:, :-1], p[:, :-1], cu(cam["intrinsics"]),
                                          cu(cam["unit_ray"]), d, 10.0, log_softmax=True)
    for b in range(B):
        want = O.plane_sweep_cost(T(feats[b:b + 1, -1]), T(feats[b:b + 1, :-1]), d,
                                  T(poses[b, :-1, :3, :3]), T(poses[b, :-1, :3, 3]),
                                  T(cam["intrinsics"][b]), T(cam["unit_ray"][b]), 10.0, "L2")
        close(cost[b:b + 1], want.numpy())
        logclose(lsm[b:b + 1], O.log_softmax_bins(want).numpy())


@pytest.mark.parametrize("algo", [5, 4])
@pytest.mark.parametrize("kind", ["wide_baseline", "many_runs", "tall_motion", "two_views_strided"])
def test_sweep_tma_window_and_pass_edges(dpv, kind, algo):
    """The TMA kernels' (5: cross-correlation form, 4: Gram form) edge paths against the oracle: a source window wider than its 48-column
    capacity (per-thread gather fallback), more runs per pixel than one pass holds (16), a window
    taller than 8 rows, and two views read through a strided [B, V+1, C, h, w] allocation."""
    synth = dpv.synth
    if kind == "wide_baseline":      # disparity 6..50 px: window > 48 columns
        C, h, w, D, B = 9, 8, 128, 32, 2
        pose = [synth.pose(None, (-2.2, 0.0, 0.0))]
    elif kind == "many_runs":        # ~24 cells per pixel, window fits
        C, h, w, D, B = 11, 8, 16, 64, 1
        pose = [synth.pose(None, (-7.0, 0.0, 0.0))]
    elif kind == "tall_motion":      # vertical parallax: > 8 window rows
        C, h, w, D, B = 7, 48, 32, 32, 1
        pose = [synth.pose(None, (0.0, -1.5, 0.0))]
    else:
        C, h, w, D, B = 67, 16, 24, 64, 3
        pose = [synth.pose(synth.yaw_matrix(0.7), (0.05, -0.02, 0.8)),
                synth.pose(synth.yaw_matrix(-1.0), (-0.1, 0.03, -0.6))]
    V = len(pose)
    d = synth.depth_candidates(5, 40, D)
    feats = synth.randn(91, B, V + 1, C, h, w)
    poses = np.stack([np.stack(pose + [synth.pose()])] * B).astype(np.float32)
    cam = synth.camera(w, h, B)
    f, p = cu(feats), cu(poses)
    cost = dpv.ops.sweep_cost_volume(f[:, -1], f[:, :-1], p[:, :-1], cu(cam["intrinsics"]),
                                     cu(cam["unit_ray"]), d, 10.0, algo=algo)
    for b in range(B):
        want = O.plane_sweep_cost(T(feats[b:b + 1, -1]), T(feats[b:b + 1, :-1]), d,
                                  T(poses[b, :-1, :3, :3]), T(poses[b, :-1, :3, 3]),
                                  T(cam["intrinsics"][b]), T(cam["unit_ray"][b]), 10.0, "L2")
        close(cost[b:b + 1], want.numpy())


def test_sweep_identity_pose_is_near_zero(dpv):
    """Size-independent property at the model's shape: source == reference under the identity
    pose gives a cost volume that is ~0 (only coordinate rounding, SURVEY.md 7 'bit-level')."""
    C, h, w, D = 67, 64, 96, 64
    synth = dpv.synth
    ref = synth.randn(5, 1, C, h, w)
    poses = synth.pose()[None, None]
    cam = synth.camera(w, h, 1)
    for algo in (1, 2, 3, 4, 5):
        cost = dpv.ops.sweep_cost_volume(cu(ref), cu(ref[:, None]), cu(poses), cu(cam["intrinsics"]),
                                         cu(cam["unit_ray"]), synth.depth_candidates(5, 40, D), 10.0,
                                         algo=algo)
        assert float(cost.abs().max()) < 1e-5 * C


# ------------------------------------------------------------------------- K1 alone, K4a
def test_warp_planes_vs_oracle(dpv):
    c = cases.sweep_case("mono_small")
    K, R, t, rays = T(c["K"]), T(c["R"][0]), T(c["t"][0]), T(c["rays"])
    term1, term2 = O.sweep_terms(K, R, t, rays)
    d = torch.from_numpy(c["d_candi"].astype(np.float32))
    cx, cy = np.float32(c["K"][0, 2]), np.float32(c["K"][1, 2])
    want = O.back_warp_planes(T(c["src"][0, 0]), d, term1, term2, cx, cy, c["h"], c["w"])
    src = cu(c["src"][0, 0])
    stack = src.unsqueeze(0).repeat(d.numel(), 1, 1, 1)
    got = dpv.warping.homography._back_warp_homo_parallel(
        stack, d.cuda(), term1.cuda(), term2.cuda(), {"intrinsic_M": c["K"]}, c["h"], c["w"])
    close(got, want.numpy(), rtol=1e-4, atol=2e-5)
    got2 = dpv.ops.warp_planes(src.unsqueeze(0), d.cuda(), term1.cuda(), term2.cuda(), cx, cy,
                               c["h"], c["w"])
    assert torch.equal(got, got2)


@pytest.mark.parametrize("name", cases.WARP_FEATURE_CASES)
def test_warp_feature(dpv, golden, name):
    c = cases.warp_feature_case(name)
    got = dpv.warping.homography.warp_feature(cu(c["feat"]), c["d_candi"], cu(c["R"]), cu(c["t"]),
                                              cam_dict(c))
    got = got if name == "small" else got[:, :, ::4]
    close(got, golden("warp_feature")[name], rtol=1e-4, atol=2e-5)
    with pytest.raises(Exception, match="Warped Accum Error"):
        dpv.warping.homography.warp_feature(cu(np.concatenate([c["feat"]] * 2)), c["d_candi"],
                                            cu(c["R"]), cu(c["t"]), cam_dict(c))


# ------------------------------------------------------------------------- K3
@pytest.mark.parametrize("name", cases.SOFTMAX_CASES)
def test_head_vs_golden(dpv, golden, name):
    g = golden("softmax")
    c = cases.softmax_case(name)
    x = cu(c["x"])
    out = dpv.ops.head(x, c["d_candi"], logp=True, prob=True, depth=True, variance=True, argmax=True,
                       quarter=(name + "_quarter") in g.files)
    logclose(out["logp"], g[name + "_logdpv"])
    close(out["prob"], np.exp(g[name + "_logdpv"]), rtol=1e-4, atol=1e-9)
    close(out["depth"], g[name + "_depth"], rtol=1e-4, atol=0)
    close(out["variance"], g[name + "_var"].astype(np.float32), rtol=1e-4, atol=1e-6)
    np.testing.assert_array_equal(out["argmax"].cpu().numpy(), g[name + "_argmax"])
    if "quarter" in out:
        logclose(out["quarter"], g[name + "_quarter"])
        assert torch.equal(out["quarter"], out["logp"][:, :, ::4, ::4])


def test_head_modes_and_mirrors(dpv, golden):
    """dpv_to_depthmap on log / linear DPVs (utils/img_utils.py:52-61) and its B != 1 error."""
    g = golden("softmax")
    c = cases.softmax_case("wide")
    logdpv = cu(g["wide_logdpv"])
    iu = dpv.utils.img_utils
    close(iu.dpv_to_depthmap(logdpv, c["d_candi"], BV_log=True), g["wide_depth"], rtol=1e-4, atol=0)
    close(iu.dpv_to_depthmap(torch.exp(logdpv), c["d_candi"], BV_log=False), g["wide_depth"],
          rtol=1e-4, atol=0)
    close(iu.depth_variance(logdpv, c["d_candi"]), g["wide_var"].astype(np.float32), rtol=1e-4, atol=1e-6)
    with pytest.raises(Exception, match="Unable to handle this case"):
        iu.dpv_to_depthmap(torch.cat([logdpv, logdpv]), c["d_candi"], BV_log=True)


def test_head_full_size_properties(dpv):
    """BASELINE size (B=8, D=64, 256x384): invariants that need no oracle."""
    B, D, H, W = 8, 64, 256, 384
    d = dpv.synth.depth_candidates(5, 40, D)
    gen = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn((B, D, H, W), device="cuda", generator=gen) * 4.0
    out = dpv.ops.head(x, d, logp=True, prob=True, depth=True, variance=True, argmax=True, quarter=True)
    total = out["prob"].sum(1)
    assert float((total - 1).abs().max()) < 1e-5
    assert torch.equal(out["argmax"], torch.argmax(out["logp"], dim=1))
    picked = torch.gather(out["logp"], 1, out["argmax"].unsqueeze(1)).squeeze(1)
    assert torch.equal(picked, out["logp"].max(1).values)
    assert float(out["depth"].min()) >= 5.0 - 1e-4 and float(out["depth"].max()) <= 40.0 + 1e-4
    assert float(out["variance"].min()) >= 0.0
    assert torch.equal(out["quarter"], out["logp"][:, :, ::4, ::4])
    # idempotence: the head applied to its own log-DPV returns it unchanged (to rounding)
    again = dpv.ops.head(out["logp"], d, logp=True)["logp"]
    assert float((again - out["logp"]).abs().max()) < 2e-6
    # shift invariance of soft-max
    shifted = dpv.ops.head(x + 3.0, d, logp=True)["logp"]
    assert float((shifted - out["logp"]).abs().max()) < 1e-5


# ------------------------------------------------------------------------- K4b, K4c
@pytest.mark.parametrize("name", cases.FUSE_CASES)
def test_fusion(dpv, golden, name):
    g = golden("fuse")
    c = cases.fuse_case(name)
    bv = dpv.ops.head(cu(c["bv_logits"]), c["d_candi"], logp=True)["logp"]
    fused, logf = dpv.ops.bayes_fuse(bv, c["d_candi"], dmaps=cu(c["dmaps"]), masks=cu(c["masks"]), var=0.3)
    logclose(logf, g[name + "_logfused"])
    if name == "small":
        prior = dpv.utils.img_utils.gen_dpv_withmask(cu(c["dmaps"]), cu(c["masks"]), c["d_candi"], 0.3)
        close(prior, g[name + "_prior"], rtol=1e-4, atol=1e-12)
        close(fused, g[name + "_fused"], rtol=1e-4, atol=1e-12)
        fused2, logf2 = dpv.ops.bayes_fuse(bv, c["d_candi"], prior=prior)
        assert torch.equal(fused2, fused) and torch.equal(logf2, logf)
        upd = dpv.ops.head(bv, c["d_candi"], addend=cu(c["resi"]), logp=True)["logp"]
        logclose(upd, g[name + "_feedback"])


# ------------------------------------------------------------------------- K5
def _near_threshold_columns(c, log, params=None):
    """Columns holding a pixel whose height/depth is within 1e-4 (relative) of a mask threshold."""
    ls = O.log_softmax_bins(T(c["logits"]))
    return near_threshold_columns(ls if log else torch.exp(ls), c["d_candi"], T(c["intr_up"]), log, params)


@pytest.mark.parametrize("name", cases.UFIELD_CASES)
def test_ufield(dpv, golden, name):
    g = golden("ufield")
    c = cases.ufield_case(name)
    ls = dpv.ops.head(cu(c["logits"]), c["d_candi"], logp=True)["logp"]
    vol = ls if c["log"] else torch.exp(ls)

    class cfg:
        class data:
            dataset_path = "./kitti/"
    uf, dz = dpv.utils.img_utils.gen_ufield(vol, c["d_candi"], cu(c["intr_up"]), BV_log=c["log"],
                                            mask=cu(c["mask"]), cfg=cfg)
    want_uf, want_dz = g[name + "_uf"], g[name + "_depthzero"]
    keep = ~_near_threshold_columns(c, c["log"])
    assert keep.mean() > 0.9
    uf = uf.cpu().numpy()
    np.testing.assert_array_equal(np.isnan(uf[:, :, keep]), np.isnan(want_uf[:, :, keep]))
    np.testing.assert_allclose(uf[:, :, keep], want_uf[:, :, keep], rtol=1e-4, atol=1e-7, equal_nan=True)
    np.testing.assert_allclose(dz.cpu().numpy()[:, :, keep], want_dz[:, :, keep], rtol=1e-4, atol=0)


@pytest.mark.parametrize("name", cases.UFIELD_CFGX_CASES)
def test_ufield_cfgx_quash_limit(dpv, golden, name):
    """gen_ufield(..., cfgx=...) as ros/ros_net.py:279 calls it: quash_limit branch
    (utils/img_utils.py:269-275,325-332), with and without a row shift / GT mask / linear input."""
    g = golden("ufield_cfgx")
    c = cases.ufield_cfgx_case(name)
    ls = dpv.ops.head(cu(c["logits"]), c["d_candi"], logp=True)["logp"]
    vol = ls if c["log"] else torch.exp(ls)
    uf, dz = dpv.utils.img_utils.gen_ufield(vol, c["d_candi"], cu(c["intr_up"]), BV_log=c["log"],
                                            mask=cu(c["mask"]), cfgx=c["cfgx"])
    want_uf, want_dz = g[name + "_uf"], g[name + "_depthzero"]
    keep = ~_near_threshold_columns(c, c["log"], O.cfgx_params(c["cfgx"]))
    assert keep.mean() > 0.9
    uf = uf.cpu().numpy()
    np.testing.assert_array_equal(np.isnan(uf[:, :, keep]), np.isnan(want_uf[:, :, keep]))
    np.testing.assert_allclose(uf[:, :, keep], want_uf[:, :, keep], rtol=1e-4, atol=1e-7, equal_nan=True)
    np.testing.assert_allclose(dz.cpu().numpy()[:, :, keep], want_dz[:, :, keep], rtol=1e-4, atol=0)
    # the fused entry point takes the same branch through its two-kernel form
    if c["log"] and c["mask"] is None:
        p = dict(dpv.utils.img_utils._ufield_params(None, c["cfgx"]))
        out = dpv.ops.head_ufield(cu(c["logits"]), c["d_candi"], cu(c["intr_up"]), params=p)
        np.testing.assert_allclose(out["uf"].cpu().numpy()[:, :, keep], want_uf[:, :, keep], rtol=1e-4,
                                   atol=1e-7, equal_nan=True)


@pytest.mark.parametrize("name", ["small_log", "full_log"])
def test_head_ufield_fused_vs_golden(dpv, golden, name):
    """K3 + K5 in one pass (dpv_head_ufield) against the reference's gen_ufield output."""
    g = golden("ufield")
    c = cases.ufield_case(name)
    x = cu(c["logits"])
    out = dpv.ops.head_ufield(x, c["d_candi"], cu(c["intr_up"]), mode="logits", logp=True, depth=True,
                              variance=True, argmax=True, quarter=True)
    plain = dpv.ops.head(x, c["d_candi"], logp=True, depth=True, variance=True, argmax=True, quarter=True)
    # dpv_head may take a different kernel (bins split over lanes for small volumes): same values up
    # to the order of the per-pixel sums
    assert torch.equal(out["argmax"], plain["argmax"])
    assert torch.equal(out["quarter"], out["logp"][:, :, ::4, ::4])
    for k in ("logp", "depth", "variance"):
        assert float(((out[k] - plain[k]).abs() / plain[k].abs().clamp_min(1.0)).max()) < 1e-5, k
    want_uf, want_dz = g[name + "_uf"], g[name + "_depthzero"]
    keep = ~_near_threshold_columns(c, True)
    assert keep.mean() > 0.9
    uf = out["uf"].cpu().numpy()
    np.testing.assert_array_equal(np.isnan(uf[:, :, keep]), np.isnan(want_uf[:, :, keep]))
    np.testing.assert_allclose(uf[:, :, keep], want_uf[:, :, keep], rtol=1e-4, atol=1e-7, equal_nan=True)
    np.testing.assert_allclose(out["depth_zero"].cpu().numpy()[:, :, keep], want_dz[:, :, keep], rtol=1e-4, atol=0)
    # the two-kernel path (dpv_head + dpv_ufield) takes the same per-pixel decisions: same NaN
    # pattern and depth_zero everywhere, UF equal up to the summation order
    uf2, dz2 = dpv.ops.ufield(out["logp"], c["d_candi"], cu(c["intr_up"]), depth=out["depth"])
    assert torch.equal(out["depth_zero"], dz2)
    assert torch.equal(torch.isnan(out["uf"]), torch.isnan(uf2))
    ok = ~torch.isnan(uf2)
    assert float(((out["uf"][ok] - uf2[ok]).abs() / uf2[ok].abs().clamp_min(1e-6)).max()) < 1e-5
    # log-probability input gives the same field as logits input
    again = dpv.ops.head_ufield(plain["logp"], c["d_candi"], cu(c["intr_up"]), mode="logprob", logp=False)
    assert torch.equal(torch.isnan(again["uf"]), torch.isnan(out["uf"]))


def test_head_ufield_batched_full_size(dpv):
    """BASELINE size (B=8, 256x384): fused == two-kernel path per item, bit-reproducible run to run."""
    B, D, H, W = 8, 64, 256, 384
    s = dpv.synth
    d = s.depth_candidates(5, 40, D)
    cam = s.camera(W // 4, H // 4, B)
    x = cu(s.ground_plane_logits(7, B, H, W, d, cam["intrinsics_up"][0]))
    Ku = cu(cam["intrinsics_up"])
    a = dpv.ops.head_ufield(x, d, Ku, logp=True, depth=True)
    b = dpv.ops.head_ufield(x, d, Ku, logp=True, depth=True)
    assert torch.equal(a["uf"].view(torch.int32), b["uf"].view(torch.int32))
    uf2, dz2 = dpv.ops.ufield(a["logp"], d, Ku, depth=a["depth"])
    assert torch.equal(a["depth_zero"], dz2)
    assert torch.equal(torch.isnan(a["uf"]), torch.isnan(uf2))
    ok = ~torch.isnan(uf2)
    assert ok.float().mean() > 0.5
    assert float(((a["uf"][ok] - uf2[ok]).abs() / uf2[ok].abs().clamp_min(1e-6)).max()) < 1e-5


# ------------------------------------------------------------------------- K2b
@pytest.mark.parametrize("name", cases.CORR_CASES)
def test_correlation(dpv, golden, name):
    c = cases.corr_case(name)
    want = golden("correlation")[name]
    native = dpv.models.correlation_native.Correlation(max_displacement=4)
    got = native(cu(c["x1"]), cu(c["x2"]))
    close(got, want, rtol=1e-6, atol=2e-7)
    ext = dpv.models.correlation_package.correlation.Correlation(
        pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1, corr_multiply=1)
    assert torch.equal(ext(cu(c["x1"]), cu(c["x2"])), got)
    # generic-radius kernel against the oracle
    got3 = dpv.ops.correlation(cu(c["x1"]), cu(c["x2"]), 2)
    close(got3, O.local_correlation(T(c["x1"]), T(c["x2"]), 2).numpy(), rtol=1e-6, atol=2e-7)


def test_correlation_linearity_full_size(dpv):
    """PWC-Lite's largest level (models/pwclite.py:177-186): corr is linear in its first input."""
    gen = torch.Generator(device="cuda").manual_seed(9)
    x1 = torch.randn((2, 32, 96, 208), device="cuda", generator=gen)
    x1b = torch.randn((2, 32, 96, 208), device="cuda", generator=gen)
    x2 = torch.randn((2, 32, 96, 208), device="cuda", generator=gen)
    f = dpv.ops.correlation
    lhs = f(2.0 * x1 + x1b, x2)
    rhs = 2.0 * f(x1, x2) + f(x1b, x2)
    assert float((lhs - rhs).abs().max()) < 5e-6


# ------------------------------------------------------------------------- no CPU path
def test_cpu_tensors_are_refused(dpv):
    c = cases.softmax_case("small")
    with pytest.raises(dpv.DpvError):
        dpv.ops.head(T(c["x"]), c["d_candi"])


# ------------------------------------------------------------------------- host pipeline
def test_host_pipeline_matches_ops(dpv):
    B, V, C, D, h, w, H, W = 2, 1, 67, 64, 16, 24, 64, 96
    synth = dpv.synth
    d = synth.depth_candidates(5, 40, D)
    feats = synth.randn(81, B, V + 1, C, h, w)
    poses = synth.stereo_poses(B)
    cam = synth.camera(w, h, B)
    logits = synth.ground_plane_logits(82, B, H, W, d, cam["intrinsics_up"][0])
    from importlib import import_module
    pl = import_module("probabilistic-depth_b200.pipeline")
    pipe = pl.FramePipeline(B, V, C, D, h, w, H, W)
    out = pipe.run(T(feats).pin_memory(), T(poses).pin_memory(), T(cam["intrinsics"]).pin_memory(),
                   T(cam["unit_ray"]).pin_memory(), d, T(logits).pin_memory(),
                   T(cam["intrinsics_up"]).pin_memory(), 10.0, pipe.outputs())
    h2d, d2h = pipe.last_bytes()
    assert h2d > feats.nbytes + logits.nbytes and d2h > 0
    f, p = cu(feats), cu(poses)
    cost = dpv.ops.sweep_cost_volume(f[:, -1], f[:, :-1], p[:, :-1], cu(cam["intrinsics"]),
                                     cu(cam["unit_ray"]), d, 10.0)
    bv = dpv.ops.head(cost, d, logp=True)["logp"]
    hd = dpv.ops.head(cu(logits), d, logp=True, depth=True, variance=True, argmax=True, quarter=True)
    uf, dz = dpv.ops.ufield(hd["logp"], d, cu(cam["intrinsics_up"]), depth=hd["depth"])
    assert torch.equal(out["bv"], bv.cpu())
    # the pipeline runs head + UF fused (another kernel than dpv_head): same values up to the order
    # of the per-pixel sums, same arg-max
    for k in ("depth", "variance", "quarter"):
        np.testing.assert_allclose(out[k].numpy(), hd[k].cpu().numpy(), rtol=1e-5, atol=1e-6)
    assert torch.equal(out["argmax"], hd["argmax"].cpu())
    # the pipeline runs head + UF fused, item by item: same per-pixel decisions, and a UF equal up
    # to the order in which the rows of a column are added
    assert torch.equal(torch.isnan(out["uf"]), torch.isnan(uf.cpu()))
    np.testing.assert_allclose(out["uf"].numpy(), uf.cpu().numpy(), rtol=1e-5, atol=1e-30, equal_nan=True)
    np.testing.assert_allclose(out["depth_zero"].numpy(), dz.cpu().numpy(), rtol=1e-5, atol=0)
    pipe.close()


def test_host_pipeline_submit_wait_overlaps_batches_with_identical_results(dpv):
    """dpv_pipeline_submit / dpv_pipeline_wait: four batches through the two staging slots, two in flight at a
    time, each bit-identical to the same batch through the synchronous dpv_pipeline_run."""
    B, V, C, D, h, w, H, W = 2, 1, 67, 64, 16, 24, 64, 96
    synth = dpv.synth
    d = synth.depth_candidates(5, 40, D)
    cam = synth.camera(w, h, B)
    from importlib import import_module
    pl = import_module("probabilistic-depth_b200.pipeline")
    pipe = pl.FramePipeline(B, V, C, D, h, w, H, W)
    pin = lambda a: T(a).pin_memory()
    const = [pin(synth.stereo_poses(B)), pin(cam["intrinsics"]), pin(cam["unit_ray"]), pin(cam["intrinsics_up"])]
    batches = [(pin(synth.randn(300 + i, B, V + 1, C, h, w)),
                pin(synth.ground_plane_logits(400 + i, B, H, W, d, cam["intrinsics_up"][0]))) for i in range(4)]
    want = []
    for f, lg in batches:
        o = pipe.run(f, const[0], const[1], const[2], d, lg, const[3], 10.0, pipe.outputs())
        want.append({k: v.clone() for k, v in o.items()})
    with pytest.raises(dpv.DpvError):
        pipe.wait()                                   # nothing outstanding
    outs = [pipe.outputs() for _ in batches]
    for i, (f, lg) in enumerate(batches):             # the third submit waits for the first internally
        pipe.submit(f, const[0], const[1], const[2], d, lg, const[3], 10.0, outs[i])
    pipe.wait()
    pipe.wait()                                       # two were still outstanding
    with pytest.raises(dpv.DpvError):
        pipe.wait()
    for o, wnt in zip(outs, want):
        for k in wnt:
            assert torch.equal(torch.nan_to_num(o[k].float(), nan=-7.0), torch.nan_to_num(wnt[k].float(), nan=-7.0)), k
    pipe.close()


# ------------------------------------------------------------------------- 8f rank 2: D -> D 3x3 convolutions
@pytest.mark.parametrize("B,h,w", [(2, 64, 96), (1, 16, 24), (3, 7, 13)])
def test_cost_refine_convs_tensor_core_vs_torch_fp32(dpv, B, h, w):
    """conv0 -> LeakyReLU -> conv0_1 -> LeakyReLU -> conv0_2 -> log_softmax (models/models.py:456-460,555-560) on the
    tcgen05 tensor cores with TF32 x 3 split precision, against torch's fp32 convolutions on the CPU (no TF32, no
    cuDNN): logits within 1e-4 relative of their scale, log-DPV within 1e-4, arg-max identical except inside that noise."""
    g = torch.Generator().manual_seed(1234 + h)
    std = (2.0 / (9 * 64)) ** 0.5                                  # models/models.py weight_init
    ws = [torch.randn((64, 64, 3, 3), generator=g) * std for _ in range(3)]
    bs = [torch.randn((64,), generator=g) * 0.1 for _ in range(3)]
    cost = torch.randn((B, 64, h, w), generator=g) * 4.0 + 10.0    # a cost volume: positive, O(10)
    F = torch.nn.functional
    x = F.leaky_relu(F.conv2d(cost.double(), ws[0].double(), bs[0].double(), padding=1), 0.01)
    x = F.leaky_relu(F.conv2d(x, ws[1].double(), bs[1].double(), padding=1), 0.01)
    want_logits = F.conv2d(x, ws[2].double(), bs[2].double(), padding=1)
    want = torch.log_softmax(want_logits, dim=1)
    refine = dpv.ops.CostRefine([t.cuda() for t in ws], [t.cuda() for t in bs])
    got, got_logits = refine(cost.cuda(), want_logits=True)
    scale = float(want_logits.abs().max())
    e_logits = float((got_logits.cpu().double() - want_logits).abs().max()) / scale
    e_logp = float(((got.cpu().double() - want).abs() / want.abs().clamp_min(1.0)).max())
    print("cost refine %dx%dx%d: logits scale %.1f, max error %.2e of it; log-DPV scaled error %.2e" % (B, h, w, scale, e_logits, e_logp))
    assert e_logits <= 1e-4
    assert e_logp <= 1e-4
    flips = torch.argmax(got.cpu(), 1) != torch.argmax(want, 1)
    top2 = torch.topk(want, 2, dim=1).values
    assert not flips.any() or float((top2[:, 0] - top2[:, 1])[flips].max()) <= 2e-4
    # fp32 reference on the same device (what the model runs: cuDNN, TF32 off) for the record
    torch.backends.cudnn.allow_tf32 = False
    y = F.leaky_relu(F.conv2d(cost.cuda(), ws[0].cuda(), bs[0].cuda(), padding=1), 0.01)
    y = F.leaky_relu(F.conv2d(y, ws[1].cuda(), bs[1].cuda(), padding=1), 0.01)
    y = F.conv2d(y, ws[2].cuda(), bs[2].cuda(), padding=1)
    assert float((got_logits - y).abs().max()) <= 1e-4 * scale


def test_cost_refine_vs_reference_golden(dpv):
    """The tcgen05 convolution chain against the reference BaseModel's own modules (fixture conv_refine.npz)."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "conv_refine.npz"))
    refine = dpv.ops.CostRefine([cu(g["w%d" % i]) for i in range(3)], [cu(g["b%d" % i]) for i in range(3)],
                                slope=float(g["slope"]))
    bv, logits = refine(cu(g["cost"]), want_logits=True)
    scale = float(np.abs(g["logits"]).max())
    assert float(np.abs(logits.cpu().numpy() - g["logits"]).max()) <= 1e-5 * scale
    logclose(bv, g["bv"])
    want, _ = O.cost_refine(T(g["cost"]), [T(g["w%d" % i]) for i in range(3)], [T(g["b%d" % i]) for i in range(3)], float(g["slope"]))
    logclose(bv, want.float().numpy())


# ------------------------------------------------------------------ 8f rank 2, second half: Base3D
def test_base3d_vs_reference_golden(dpv):
    """The tcgen05 3-D convolution stack against the reference's own Base3D module (fixture base3d.npz: fp32 CPU,
    eval(): dres0 / classify on running statistics, the unregistered residual blocks on batch statistics)."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "base3d.npz"))
    net = dpv.ops.Base3DConvs(cases.base3d_layers(g, cu))
    got = net(cu(g["volume"]))
    scale = float(np.abs(g["resi"]).max())
    err = float(np.abs(got.cpu().numpy() - g["resi"]).max()) / scale
    print("Base3D vs the reference module: max error %.2e of the residual's scale %.1f" % (err, scale))
    assert err <= 2e-5
    want = O.base3d(T(g["volume"]), cases.base3d_layers(g))
    assert float((got.cpu().double() - want).abs().max()) <= 2e-5 * scale
    # the residual's use (models/models.py:693-694): log_softmax(BV_cur + BV_resi) within the DPV tolerance
    bv = torch.log_softmax(T(g["volume"])[:, 0], dim=1)
    upd_want = torch.log_softmax(bv.double() + want, dim=1)
    upd_got = torch.log_softmax(bv.double() + got.cpu().double(), dim=1)
    assert float(((upd_got - upd_want).abs() / upd_want.abs().clamp_min(1.0)).max()) <= 1e-4


@pytest.mark.parametrize("shape", [(1, 4, 8, 9, 11), (2, 4, 16, 12, 20), (1, 4, 3, 5, 130)])
def test_base3d_tensor_core_vs_oracle(dpv, shape):
    """Other volume shapes (odd sizes, rows longer than a tile, two items sharing the batch statistics) against the
    float64 oracle, with the golden's parameters."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "base3d.npz"))
    net = dpv.ops.Base3DConvs(cases.base3d_layers(g, cu))
    gen = torch.Generator().manual_seed(sum(shape))
    vol = torch.randn(shape, generator=gen)
    want = O.base3d(vol, cases.base3d_layers(g))
    got = net(vol.cuda()).cpu().double()
    scale = float(want.abs().max())
    err = float((got - want).abs().max()) / scale
    print("Base3D %s: max error %.2e of the residual's scale %.1f" % (shape, err, scale))
    assert err <= 2e-5
    again = net(vol.cuda()).cpu().double()          # buffers and statistics are reused call after call
    assert float((again - want).abs().max()) <= 2e-5 * scale


def test_base3d_all_running_statistics(dpv):
    """Every BatchNorm folded into its convolution (a registered, eval-mode stack; bn_avg: true)."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "base3d.npz"))
    spec, spec_cu = cases.base3d_layers(g), cases.base3d_layers(g, cu)
    for L in spec + spec_cu:
        if L["bn"] is not None:
            L["bn"]["batch_stats"] = False
    vol = torch.randn((1, 4, 6, 7, 9), generator=torch.Generator().manual_seed(9))
    want = O.base3d(vol, spec)
    got = dpv.ops.Base3DConvs(spec_cu)(vol.cuda()).cpu().double()
    assert float((got - want).abs().max()) <= 2e-5 * float(want.abs().max())
