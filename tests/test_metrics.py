"""Eval metrics: oracle restatement vs the reference (goldens + oracle/_ref when
built), and the device kernels vs the goldens through the C ABI.

Tolerances: the reference accumulates in float32 pixel by pixel, the kernels in double in a fixed tree
order, numpy's float32 log differs from glibc's logf in the last bit -- 1e-6 relative between the
restatement and the reference, 1e-4 relative (the north_star floating-point bar) for the kernels.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases  # noqa: E402
from oracle import depth_metrics as M  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics.npz"))


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-12)))


def prepared(c):
    t = c["truth"].copy()
    t[t >= c["d_max"]] = c["d_max"]
    return c["predicted"] * c["mask"], t


# ------------------------------------------------------------------------------- CPU: the oracle
@pytest.mark.parametrize("name", cases.METRICS_CASES)
def test_oracle_depth_error_vs_golden(name):
    c = cases.metrics_case(name)
    assert rel(M.depth_error(c["predicted"], c["truth"]), GOLD[name + "_errors"]) < 1e-6
    assert rel(M.depth_error(*prepared(c)), GOLD[name + "_errors_prepared"]) < 1e-6


@pytest.mark.skipif(not M.reference_available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name", cases.METRICS_CASES)
def test_oracle_depth_error_vs_compiled_reference(name):
    c = cases.metrics_case(name)
    want = M.reference_depth_error(c["predicted"], c["truth"])
    assert np.array_equal(want, GOLD[name + "_errors"])            # the goldens are the reference's output
    assert rel(M.depth_error(c["predicted"], c["truth"]), want) < 1e-6


def test_oracle_eval_errors_and_unc_rmse_vs_golden():
    per_item = [GOLD[n + "_errors"] for n in cases.METRICS_CASES]
    got = M.eval_errors(per_item)
    assert rel([got[n] for n in M.METRICS], GOLD["eval_errors"] + 0.0) < 1e-6 or \
        np.allclose([got[n] for n in M.METRICS], GOLD["eval_errors"], rtol=1e-6, atol=1e-9)
    for name in cases.UNC_RMSE_CASES:
        c = cases.unc_rmse_case(name)
        assert rel(M.compute_unc_rmse(c["truth"], c["pred"], c["d_candi"]), GOLD["unc_rmse_" + name]) < 1e-5


def test_oracle_raises_without_valid_pixels():
    z = np.zeros((4, 4), np.float32)
    with pytest.raises(RuntimeError):
        M.depth_error(z, z)


def test_eval_errors_mirror_matches_oracle(dpv):
    per_item = [GOLD[n + "_errors"] for n in cases.METRICS_CASES]
    assert dpv.utils.img_utils.eval_errors(per_item) == M.eval_errors(per_item)


# ------------------------------------------------------------------------------- GPU: the kernels
def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("name", cases.METRICS_CASES)
def test_depth_errors_kernel_vs_golden(dpv, name):
    c = cases.metrics_case(name)
    got = dpv.utils.img_utils.depth_error(c["predicted"], c["truth"])          # numpy in, list out
    assert isinstance(got, list) and len(got) == 9
    assert np.allclose(got, GOLD[name + "_errors"], rtol=1e-4, atol=1e-7)
    # the trainer's preparation fused into the launch == done beforehand
    out, counts = dpv.ops.depth_errors(cu(c["predicted"]), cu(c["truth"]), mask=cu(c["mask"]),
                                       clamp_max=c["d_max"], want_counts=True)
    assert np.allclose(out[0].cpu().numpy(), GOLD[name + "_errors_prepared"], rtol=1e-4, atol=1e-7)
    assert int(counts[0]) == int((c["predicted"] * c["mask"] != 0).sum())


@pytest.mark.gpu
def test_depth_errors_batched_and_edge_cases(dpv):
    names = ["quarter", "quarter", "quarter"]
    cs = [cases.metrics_case(n) for n in names]
    cs[1]["predicted"] = cs[1]["predicted"] * 1.1
    cs[2]["predicted"] = np.zeros_like(cs[2]["predicted"])                        # no valid pixel at all
    out, counts = dpv.ops.depth_errors(cu(np.stack([c["predicted"] for c in cs])),
                                       cu(np.stack([c["truth"] for c in cs])), want_counts=True)
    out = out.cpu().numpy()
    assert np.allclose(out[0], GOLD["quarter_errors"], rtol=1e-4, atol=1e-7)
    assert np.allclose(out[1], M.depth_error(cs[1]["predicted"], cs[1]["truth"]), rtol=1e-4, atol=1e-7)
    assert int(counts[2]) == 0 and np.isnan(out[2]).all()
    with pytest.raises(RuntimeError, match="Ground truth defect"):
        dpv.utils.img_utils.depth_error(cs[2]["predicted"], cs[2]["truth"])
    # identical maps: every metric is exactly zero
    same = dpv.ops.depth_errors(cu(cs[0]["truth"]), cu(cs[0]["truth"]))[0].cpu().numpy()
    assert np.array_equal(same, np.zeros(9, np.float32))
    # a constant scale leaves the scale-invariant log error at ~0 and sets the log mae to log(scale)
    sc = dpv.ops.depth_errors(cu(cs[0]["truth"] * 2), cu(cs[0]["truth"]))[0].cpu().numpy()
    # (the radicand of evaluate_depth.h:113 is then 0 up to rounding: a tiny value or sqrt(-0.0...) = NaN,
    # in the reference too)
    assert abs(sc[4] - np.log(2)) < 1e-5 and (np.isnan(sc[6]) or sc[6] < 2e-3)
    with pytest.raises(RuntimeError):
        dpv.utils.img_utils.depth_error(torch.zeros(4, 4), torch.zeros(4, 4))    # CPU tensors are refused


@pytest.mark.gpu
@pytest.mark.parametrize("name", cases.UNC_RMSE_CASES)
def test_unc_rmse_kernel_vs_golden(dpv, name):
    c = cases.unc_rmse_case(name)
    got = dpv.utils.img_utils.compute_unc_rmse(cu(c["truth"]), cu(c["pred"]), c["d_candi"])
    assert abs(float(got) - float(GOLD["unc_rmse_" + name])) <= 1e-4 * float(GOLD["unc_rmse_" + name])
    both = dpv.ops.unc_rmse(cu(np.concatenate([c["truth"], c["pred"]])), cu(np.concatenate([c["pred"], c["pred"]])),
                            c["d_candi"]).cpu().numpy()
    assert abs(both[0] - float(GOLD["unc_rmse_" + name])) <= 1e-4 * both[0]
    # a field against itself is not zero: the first and last predicted columns are zeroed (:187-188)
    want_self = M.compute_unc_rmse(c["pred"], c["pred"], c["d_candi"])
    assert abs(both[1] - want_self) <= 1e-4 * max(want_self, 1e-6)
