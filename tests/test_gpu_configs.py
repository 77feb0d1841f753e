"""Config-level GPU tests: BASELINE.json configs[2..4] as parity cases.

Small shapes are compared with the CPU oracle; the full configured sizes are checked through
size-independent properties (normalisation, hand-off identity, uniform-prior invariance, agreement of
the production sweep kernel with the operation-by-operation one, shard-merge == unsharded).
"""
import importlib
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
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def scaled_err(a, b):
    a = a.detach().cpu().double()
    b = b.detach().cpu().double() if isinstance(b, torch.Tensor) else torch.from_numpy(np.asarray(b)).double()
    return float(((a - b).abs() / b.abs().clamp_min(1.0)).max())


def frame_mod():
    return importlib.import_module("probabilistic-depth_b200.frame")


def _inputs(dpv, B, V, C, D, h, w, H, W, seed, mono=True):
    s = dpv.synth
    d = s.depth_candidates(5, 40, D)
    cam = s.camera(w, h, B)
    feats = s.randn(seed, B, V + 1, C, h, w)
    poses = (s.mono_poses(B) if mono else s.stereo_poses(B)).astype(np.float32)
    logits = (3.0 * s.randn(seed + 1, B, D, H, W)).astype(np.float32)
    return d, cam, feats, poses, logits


def _oracle_quarter(feats, poses, cam, d, b):
    cost = O.plane_sweep_cost(T(feats[b:b + 1, -1]), T(feats[b:b + 1, :-1]), d, T(poses[b, :-1, :3, :3]),
                              T(poses[b, :-1, :3, 3]), T(cam["intrinsics"][b]), T(cam["unit_ray"][b]), 10.0, "L2")
    return cost, O.log_softmax_bins(cost)


# ----------------------------------------------------------------- FrameStep, three modes, small
@pytest.mark.parametrize("mode", ["default", "upsample", "feedback"])
def test_frame_step_modes_vs_oracle(dpv, mode):
    B, V, C, D, h, w, H, W = 2, 1, 67, 64, 16, 24, 64, 96
    d, cam, feats, poses, logits = _inputs(dpv, B, V, C, D, h, w, H, W, seed=300)
    step = frame_mod().FrameStep(B, V, C, D, h, w, H, W, d, mode=mode)
    kw = {}
    if mode == "upsample":
        dm, mk = dpv.synth.sparse_depth(7, B, h, w)
        kw = dict(dmaps=cu(dm), masks=cu(mk))
    if mode == "feedback":
        feat_raw = dpv.synth.randn(301, B, V + 1, D, h, w)
        resi = (0.5 * dpv.synth.randn(302, B, D, h, w)).astype(np.float32)
        kw = dict(feat_raw=cu(feat_raw), bv_resi=cu(resi))
    step.run(cu(feats), cu(poses), cu(cam["intrinsics"]), cu(cam["unit_ray"]), cu(logits),
             cu(cam["intrinsics_up"]), **kw)
    torch.cuda.synchronize()
    for b in range(B):
        cost, bv = _oracle_quarter(feats, poses, cam, d, b)
        assert scaled_err(step.cost[b:b + 1], cost) < 1e-4
        assert scaled_err(step.bv[b:b + 1], bv) < 1e-4
        refined = O.log_softmax_bins(T(logits[b:b + 1]))
        assert scaled_err(step.refined[b:b + 1], refined) < 1e-4
        assert scaled_err(step.depth[b:b + 1], O.expected_depth(refined, d, log=True)) < 1e-4
        assert scaled_err(step.var[b], O.depth_variance(refined[0], d)) < 1e-4
        # arg-max: exact on the kernel's own log-DPV, and equal to the oracle's wherever the oracle's
        # top-2 margin exceeds the fp32 noise of a log-softmax
        assert torch.equal(step.argmax[b].cpu(), torch.argmax(step.refined[b].cpu(), dim=0))
        top2 = torch.topk(refined[0], 2, dim=0).values
        clear = (top2[0] - top2[1]) > 1e-5
        assert torch.equal(step.argmax[b].cpu()[clear], O.argmax_bin(refined)[0][clear])
        assert torch.equal(step.quarter[b], step.refined[b, :, ::4, ::4])   # hand-off is a pure copy
        uf, dz = O.uncertainty_field(refined, d, T(cam["intrinsics_up"][b]), log=True)
        # same rule as tests/test_gpu_parity.py::test_ufield: columns holding a pixel within 1e-4 of a
        # band threshold are excluded explicitly (a flipped pixel moves its column discretely); every
        # other column must have the oracle's NaN pattern and agree to 1e-4
        keep = torch.from_numpy(~near_threshold_columns(refined, d, T(cam["intrinsics_up"][b]), True))
        assert float(keep.float().mean()) > 0.9
        got = step.uf[b:b + 1].cpu()[:, :, keep]
        want_uf = uf[:, :, keep]
        assert torch.equal(torch.isnan(got), torch.isnan(want_uf))
        fin = torch.isfinite(want_uf)
        assert float(((got[fin] - want_uf[fin]).abs() / want_uf[fin].abs().clamp_min(1e-3)).max()) < 1e-4
        assert float(((step.dz[b:b + 1].cpu() - dz).abs()[:, :, keep] / dz.abs().clamp_min(1.0)[:, :, keep]).max()) < 1e-4
        if mode == "upsample":
            prior = O.lidar_prior(T(dm[b:b + 1]), T(mk[b:b + 1]), d, 0.3)
            fused, logf = O.bayes_fuse(bv, prior)
            assert float(((step.fused[b:b + 1].cpu() - fused).abs() / fused).max()) < 2e-4
            assert scaled_err(step.logfused[b:b + 1], logf) < 1e-4
        if mode == "feedback":
            want = O.warp_feature_diag(T(feat_raw[b:b + 1]), d, T(poses[b, :, :3, :3]), T(poses[b, :, :3, 3]),
                                       T(cam["intrinsics"][b]), T(cam["unit_ray"][b]))
            assert float((step.warped[b:b + 1].cpu() - want).abs().max()) < 1e-4
            assert scaled_err(step.bv_upd[b:b + 1], O.feedback_fuse(bv, T(resi[b:b + 1]))) < 1e-4


# ----------------------------------------------------------------- configs[2]: feedback sequence
def test_feedback_sequence_16_frames(dpv):
    """16 consecutive frames through the feedback step; frame i's 1/4-res hand-off (what frame i+1
    receives as prev_output, trainer/default_trainer.py:221-222) is checked against the oracle chain
    on every frame, and the per-frame outputs against the oracle on frames 0, 7 and 15."""
    B, V, C, D, h, w, H, W = 1, 1, 67, 64, 16, 24, 64, 96
    s = dpv.synth
    step = frame_mod().FrameStep(B, V, C, D, h, w, H, W, s.depth_candidates(5, 40, D), mode="feedback")
    prev = torch.full((B, D, h, w), float(np.log(1.0 / D)))          # first frame: uniform log-DPV
    for i in range(16):
        d, cam, feats, poses, logits = _inputs(dpv, B, V, C, D, h, w, H, W, seed=400 + 10 * i)
        feat_raw = s.randn(401 + 10 * i, B, V + 1, D, h, w)
        # stand-in for the 3-D conv residual: any function of (BV_cur, prev) will do for the kernels
        resi = (0.25 * s.randn(402 + 10 * i, B, D, h, w) + 0.1 * prev.numpy()).astype(np.float32)
        step.run(cu(feats), cu(poses), cu(cam["intrinsics"]), cu(cam["unit_ray"]), cu(logits),
                 cu(cam["intrinsics_up"]), feat_raw=cu(feat_raw), bv_resi=cu(resi))
        torch.cuda.synchronize()
        refined = O.log_softmax_bins(T(logits))
        want_prev = O.quarter_nearest(refined)
        assert scaled_err(step.quarter, want_prev) < 1e-4
        if i in (0, 7, 15):
            _, bv = _oracle_quarter(feats, poses, cam, d, 0)
            assert scaled_err(step.bv_upd, O.feedback_fuse(bv, T(resi))) < 1e-4
            lse = torch.logsumexp(step.bv_upd, dim=1)
            assert float(lse.abs().max()) < 1e-5
        prev = step.quarter.cpu().clone()


def test_feedback_full_size_properties(dpv):
    B, V, C, D, h, w, H, W = 2, 1, 67, 64, 64, 96, 256, 384
    d, cam, feats, poses, logits = _inputs(dpv, B, V, C, D, h, w, H, W, seed=500)
    step = frame_mod().FrameStep(B, V, C, D, h, w, H, W, d, mode="feedback")
    feat_raw = cu(dpv.synth.randn(501, B, V + 1, D, h, w))
    resi = cu((0.5 * dpv.synth.randn(502, B, D, h, w)).astype(np.float32))
    step.run(cu(feats), cu(poses), cu(cam["intrinsics"]), cu(cam["unit_ray"]), cu(logits),
             cu(cam["intrinsics_up"]), feat_raw=feat_raw, bv_resi=resi)
    torch.cuda.synchronize()
    assert float(torch.logsumexp(step.bv_upd, 1).abs().max()) < 1e-5
    assert float(torch.logsumexp(step.refined, 1).abs().max()) < 1e-5
    # the identity (reference) view of warp_feature samples pixel centres: it returns its input
    assert float((step.warped[:, -1] - feat_raw[:, -1]).abs().max()) < 1e-4
    # adding a per-pixel constant to the residual does not change the fused DPV
    shifted = dpv.ops.head(step.bv, d, addend=resi + 3.0, logp=True)["logp"]
    assert float((shifted - step.bv_upd).abs().max()) < 1e-5


# ----------------------------------------------------------------- configs[3]: LiDAR upsample
def test_upsample_batch32_full_size_properties(dpv):
    B, D, h, w = 32, 64, 64, 96
    s = dpv.synth
    d = s.depth_candidates(5, 40, D)
    bv = torch.log_softmax(cu((2.0 * s.randn(600, B, D, h, w)).astype(np.float32)), 1)
    dm, mk = s.sparse_depth(601, B, h, w)
    fused, logf = dpv.ops.bayes_fuse(bv, d, dmaps=cu(dm), masks=cu(mk))
    assert float((fused.sum(1) - 1).abs().max()) < 1e-5
    assert float((torch.log(fused) - logf).abs().max()) < 1e-5
    # where no LiDAR return exists the prior is uniform: multiply-and-renormalise leaves the DPV alone
    no_hit = (cu(mk)[:, 0] == 0).unsqueeze(1).expand_as(fused)
    p = torch.exp(bv).clamp(2.220446049250313e-16, 1.0)
    assert float(((fused - p).abs() / p)[no_hit].max()) < 1e-4
    # where one exists the fused mode moves to (or stays at) the bin nearest the return when the
    # DPV is flat there: use a flat DPV
    flat = torch.full_like(bv, float(np.log(1.0 / D)))
    f2, _ = dpv.ops.bayes_fuse(flat, d, dmaps=cu(dm), masks=cu(mk))
    hit = cu(mk)[:, 0] == 1
    near = torch.argmin((cu(dm).unsqueeze(1) - cu(d.astype(np.float32)).view(1, D, 1, 1)).abs(), dim=1)
    assert torch.equal(torch.argmax(f2, 1)[hit], near[hit])
    # one launch over the batch == per-item launches (the reference loops over items, img_utils.py:365)
    one = dpv.ops.bayes_fuse(bv[5:6].contiguous(), d, dmaps=cu(dm)[5:6].contiguous(), masks=cu(mk)[5:6].contiguous())[0]
    assert torch.equal(one, fused[5:6])
    # sharded over 2/4/8 ranks by items: every rank's slice is the same computation
    sh = importlib.import_module("probabilistic-depth_b200.sharding")
    for world in (2, 4, 8):
        parts = []
        for r in range(world):
            idx = sh.shard_units(B, r, world)
            parts.append(dpv.ops.bayes_fuse(bv[idx].contiguous(), d, dmaps=cu(dm)[idx].contiguous(),
                                            masks=cu(mk)[idx].contiguous())[0])
        assert torch.equal(torch.stack(sh.merge_units(parts, B)), fused)


# ----------------------------------------------------------------- configs[4]: large D, plane shards
def _merge_plane_shards(sh, x, d, world):
    """PlaneShardedHead's two kernels with the all-gather done by hand on one device."""
    k = sh.CudaShardKernels()
    B, D, H, W = x.shape
    xs, stats = [], []
    for r in range(world):
        lo, hi = sh.plane_range(D, r, world)
        xs.append(x[:, lo:hi].contiguous().reshape(B, hi - lo, H * W))
        stats.append(k.local_stats(xs[r], cu(np.asarray(d, np.float64).astype(np.float32)[lo:hi]), lo))
    gathered = torch.stack(stats).contiguous()                                           # all-gather
    outs = [k.merge_finish(xs[r], gathered, True, True, True) for r in range(world)]
    for o in outs[1:]:      # the per-pixel products come out replicated, bit for bit
        assert torch.equal(o[1], outs[0][1]) and torch.equal(o[2], outs[0][2]) and torch.equal(o[3], outs[0][3])
    logp = torch.cat([o[0].reshape(B, -1, H, W) for o in outs], 1)
    # the reduce-scatter form (all-to-all + all-gather, what PlaneShardedHead uses from 3 ranks on): same results
    n = B * H * W
    sl = (n + world - 1) // world
    sends = []
    for r in range(world):
        lo, hi = sh.plane_range(D, r, world)
        buf = torch.full((world, 5, sl), float("nan"), device=x.device)
        sends.append(k.local_stats(xs[r], cu(np.asarray(d, np.float64).astype(np.float32)[lo:hi]), lo, out=buf, slice_len=sl))
    merged_all = torch.empty((world, 5, sl), device=x.device)
    for r in range(world):                                                               # all-to-all, then merge
        recv = torch.stack([sends[g][r] for g in range(world)]).contiguous()
        k.merge_slice(recv, merged_all[r])                                               # all-gather = the stack
    outs2 = [k.finish(xs[r], merged_all, True, True, True) for r in range(world)]
    logp2 = torch.cat([o[0].reshape(B, -1, H, W) for o in outs2], 1)
    assert torch.equal(outs2[0][3], outs[0][3])
    assert float((logp2 - logp).abs().max()) < 1e-5
    assert float((outs2[0][1] - outs[0][1]).abs().max()) < 1e-5
    assert float(((outs2[0][2] - outs[0][2]).abs() / outs[0][2].abs().clamp_min(1e-3)).max()) < 1e-5
    return logp, outs[0][1].reshape(B, H, W), outs[0][2].reshape(B, H, W), outs[0][3].reshape(B, H, W)


@pytest.mark.parametrize("D,world", [(128, 2), (128, 8), (256, 4), (256, 2), (256, 8), (100, 3)])
def test_large_d_plane_sharded_head_equals_unsharded(dpv, D, world):
    sh = importlib.import_module("probabilistic-depth_b200.sharding")
    B, H, W = (1, 48, 160) if D != 100 else (1, 47, 161)     # (D = 100, world 3: 7567 pixels, ragged slices)
    d = dpv.synth.depth_candidates(5, 40, D)
    x = cu((3.0 * dpv.synth.randn(700 + D, B, D, H, W)).astype(np.float32))
    logp, depth, var, amax = _merge_plane_shards(sh, x, d, world)
    ref = torch.log_softmax(x.cpu(), 1)
    assert scaled_err(logp, ref) < 1e-4
    assert scaled_err(depth, O.expected_depth(ref, d, log=True)) < 1e-4
    assert scaled_err(var[0], O.depth_variance(ref[0], d)) < 1e-4
    assert torch.equal(amax.cpu(), torch.argmax(x.cpu(), 1))
    one = dpv.ops.head(x, d, logp=True, depth=True, variance=True, argmax=True)
    assert scaled_err(logp, one["logp"]) < 1e-5 and scaled_err(depth, one["depth"]) < 1e-5
    assert torch.equal(amax, one["argmax"])


@pytest.mark.parametrize("D,world,h,w", [(128, 4, 384, 1280), (256, 8, 384, 1280), (128, 4, 96, 320)])
def test_large_d_plane_sharded_sweep_full_resolution(dpv, D, world, h, w):
    """384 x 1280 features, D = 128 / 256 (configs[4]): the plane shards of the cost volume,
    concatenated, equal the unsharded volume to fp32 rounding, and the production kernel agrees with the
    operation-by-operation kernel (algo 1, pinned to the reference goldens at small sizes)."""
    sh = importlib.import_module("probabilistic-depth_b200.sharding")
    B, C = 1, 16   # (96 x 320: tiles that mix pixels with <= 16 and > 16 runs, i.e. one and two passes)
    s = dpv.synth
    d = s.depth_candidates(5, 40, D)
    cam = s.camera(w, h, B)
    feats = cu(s.randn(800 + D, B, 2, C, h, w))
    poses = cu(s.mono_poses(B).astype(np.float32))
    K, rays = cu(cam["intrinsics"]), cu(cam["unit_ray"])
    full = dpv.ops.sweep_cost_volume(feats[:, -1], feats[:, :-1], poses[:, :-1], K, rays, d, 10.0)
    parts = [sh.plane_sharded_sweep(feats[:, -1], feats[:, :-1], poses[:, :-1], K, rays, d, 10.0, r, world)[0]
             for r in range(world)]
    # (not bit-identical: a shard walks its own plane range, so a run of planes may be anchored at a
    # different source cell than in the unsharded launch; the results differ by fp32 rounding only)
    assert scaled_err(torch.cat(parts, 1), full) < 1e-4
    direct = dpv.ops.sweep_cost_volume(feats[:, -1], feats[:, :-1], poses[:, :-1], K, rays, d, 10.0, algo=1)
    err = ((full - direct).abs() / direct.abs().clamp_min(1.0)).max()
    assert float(err) < 1e-4
    logp, depth, var, amax = _merge_plane_shards(sh, full, d, world)
    assert float(torch.logsumexp(logp, 1).abs().max()) < 1e-4
    assert torch.equal(amax, torch.argmax(full, 1))


# ----------------------------------------------------------------- CUDA-graph replay of the step
@pytest.mark.parametrize("mode", ["default", "feedback"])
def test_frame_step_graph_replay_is_bit_identical(dpv, mode):
    """FrameStep.capture(): a replayed step writes exactly what the call-by-call step writes, also after
    the inputs changed in place (the graph holds pointers, not values)."""
    B, V, C, D, h, w, H, W = 2, 1, 67, 64, 16, 24, 64, 96
    d, cam, feats, poses, logits = _inputs(dpv, B, V, C, D, h, w, H, W, seed=900)
    step = frame_mod().FrameStep(B, V, C, D, h, w, H, W, d, mode=mode)
    args = [cu(feats), cu(poses), cu(cam["intrinsics"]), cu(cam["unit_ray"]), cu(logits), cu(cam["intrinsics_up"])]
    kw = {}
    if mode == "feedback":
        kw = dict(feat_raw=cu(dpv.synth.randn(901, B, V + 1, D, h, w)),
                  bv_resi=cu((0.5 * dpv.synth.randn(902, B, D, h, w)).astype(np.float32)))
    g = step.capture(*args, **kw)
    names = ["cost", "bv", "refined", "depth", "var", "argmax", "quarter", "uf", "dz"] + \
        (["warped", "bv_upd"] if mode == "feedback" else [])
    for trial in range(2):
        if trial == 1:       # new values in the same buffers
            args[0].copy_(cu(dpv.synth.randn(903, B, V + 1, C, h, w)))
            args[4].mul_(0.5)
        step.run(*args, **kw)
        torch.cuda.synchronize()
        want = {n: getattr(step, n).clone() for n in names}
        for n in names:
            getattr(step, n).fill_(0)
        g.replay()
        torch.cuda.synchronize()
        for n in names:
            a, b = getattr(step, n), want[n]
            same = torch.equal(a, b) if a.dtype != torch.float32 else \
                torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0))
            assert same, (mode, trial, n)


# ----------------------------------------------------------------- fused head + UF: other shapes
@pytest.mark.parametrize("D,H,W", [(64, 60, 100), (32, 64, 96), (32, 44, 68), (16, 64, 96), (64, 256, 384)])
def test_fused_head_ufield_equals_two_kernel_path(dpv, D, H, W):
    """dpv_head_ufield (tile kernel for D = 32 / 64, the persistent kernel for D = 16) against dpv_head +
    dpv_ufield on ragged shapes: widths that are not a multiple of the 32-column tile, heights that are not a
    multiple of the 8-row tile.  Same per-pixel decisions (depth_zero, NaN pattern), same sums up to order."""
    B = 2
    s = dpv.synth
    d = s.depth_candidates(5, 40, D)
    cam = s.camera(max(W // 4, 1), max(H // 4, 1), B)
    x = cu(s.ground_plane_logits(40 + D, B, H, W, d, cam["intrinsics_up"][0]))
    Ku = cu(cam["intrinsics_up"])
    fused = dpv.ops.head_ufield(x, d, Ku, logp=True, depth=True, variance=True, argmax=True, quarter=True)
    plain = dpv.ops.head(x, d, logp=True, depth=True, variance=True, argmax=True, quarter=True)
    assert torch.equal(fused["argmax"], plain["argmax"])
    assert torch.equal(fused["quarter"], fused["logp"][:, :, ::4, ::4][:, :, :H // 4, :W // 4])
    for k in ("logp", "depth", "variance"):
        assert scaled_err(fused[k], plain[k]) < 1e-5, k
    uf2, dz2 = dpv.ops.ufield(fused["logp"], d, Ku, depth=fused["depth"])
    assert torch.equal(fused["depth_zero"], dz2)
    assert torch.equal(torch.isnan(fused["uf"]), torch.isnan(uf2))
    ok = ~torch.isnan(uf2)
    assert int(ok.sum()) > 0
    assert float(((fused["uf"][ok] - uf2[ok]).abs() / uf2[ok].abs().clamp_min(1e-6)).max()) < 1e-5


def test_frame_step_with_the_models_quarter_res_head(dpv):
    """FrameStep(refine=CostRefine): the 1/4-res BV is log_softmax(conv0_2(conv0_1(conv0(cost)))) (models/models.py:555-560)
    instead of the soft-max of the cost volume; everything else is unchanged; the step replays from a CUDA graph."""
    frame = importlib.import_module("probabilistic-depth_b200.frame")
    s = dpv.synth
    B, V, C, D, h, w, H, W = 2, 1, 19, 64, 16, 24, 64, 96
    d = s.depth_candidates(5, 40, D)
    cam = s.camera(w, h, B)
    g = torch.Generator().manual_seed(3)
    ws = [(torch.randn((64, 64, 3, 3), generator=g) * 0.06).cuda() for _ in range(3)]
    bs = [(torch.randn((64,), generator=g) * 0.1).cuda() for _ in range(3)]
    args = (cu(s.randn(1, B, V + 1, C, h, w)), cu(s.stereo_poses(B)), cu(cam["intrinsics"]), cu(cam["unit_ray"]),
            cu(s.ground_plane_logits(2, B, H, W, d, cam["intrinsics_up"][0])), cu(cam["intrinsics_up"]))
    plain = frame.FrameStep(B, V, C, D, h, w, H, W, d)
    plain.run(*args)
    step = frame.FrameStep(B, V, C, D, h, w, H, W, d, refine=dpv.ops.CostRefine(ws, bs))
    step.run(*args)
    torch.cuda.synchronize()
    assert torch.equal(step.cost, plain.cost) and torch.equal(step.refined, plain.refined) and torch.equal(step.argmax, plain.argmax)
    F = torch.nn.functional
    x = F.leaky_relu(F.conv2d(step.cost.double(), ws[0].double(), bs[0].double(), padding=1), 0.01)
    x = F.leaky_relu(F.conv2d(x, ws[1].double(), bs[1].double(), padding=1), 0.01)
    want = torch.log_softmax(F.conv2d(x, ws[2].double(), bs[2].double(), padding=1), dim=1)
    assert scaled_err(step.bv, want.float()) < 1e-4
    first = step.bv.clone()
    graph = step.capture(*args)
    step.bv.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(step.bv, first)


def test_feedback_step_with_base3d_in_the_step(dpv):
    """FrameStep(mode="feedback", base3d=Base3DConvs): the residual is computed in the step from
    cat(BV, prev_output, warped features) (models/models.py:692-694) by the tensor-core 3-D convolution stack; checked
    against the float64 oracle evaluated on the step's own BV and warped features; the step replays from a CUDA graph."""
    frame = importlib.import_module("probabilistic-depth_b200.frame")
    s = dpv.synth
    B, V, C, D, h, w, H, W = 2, 1, 19, 64, 16, 24, 64, 96
    d = s.depth_candidates(5, 40, D)
    cam = s.camera(w, h, B)
    g3 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "base3d.npz"))
    net = dpv.ops.Base3DConvs(cases.base3d_layers(g3, cu))
    prev = torch.log_softmax(cu(2.0 * s.randn(9, B, D, h, w)), dim=1)
    args = (cu(s.randn(1, B, V + 1, C, h, w)), cu(s.mono_poses(B)), cu(cam["intrinsics"]), cu(cam["unit_ray"]),
            cu(s.ground_plane_logits(2, B, H, W, d, cam["intrinsics_up"][0])), cu(cam["intrinsics_up"]))
    kw = dict(feat_raw=cu(s.randn(4, B, V + 1, D, h, w)), prev=prev)
    step = frame.FrameStep(B, V, C, D, h, w, H, W, d, mode="feedback", base3d=net)
    step.run(*args, **kw)
    torch.cuda.synchronize()
    vol = torch.cat([step.bv.unsqueeze(1), prev.unsqueeze(1), step.warped], dim=1).cpu()
    resi = O.base3d(vol, cases.base3d_layers(g3))
    assert float((step.resi.cpu().double() - resi).abs().max()) <= 2e-5 * float(resi.abs().max())
    want = torch.log_softmax(step.bv.cpu().double() + resi, dim=1)
    assert scaled_err(step.bv_upd, want.float().cuda()) < 1e-4
    first = step.bv_upd.clone()
    graph = step.capture(*args, **kw)
    step.bv_upd.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(step.bv_upd, first)
