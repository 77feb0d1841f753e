"""CPU restatement of the reference's eval metrics -- TEST INFRASTRUCTURE, never imported by the product.

  depth_error            utils/img_utils.py:17-22  ->  depthError  external/deval_lib/src/evaluate_depth.h:19-119
  eval_errors            utils/img_utils.py:14-15  ->  evaluateErrors  evaluate_depth.h:121-142,
                                                       statMean / statMin / statMax  external/deval_lib/src/utils.h:24-56
  compute_unc_rmse       utils/img_utils.py:183-202

Pinned: oracle/ref_deval/ compiles the reference's OWN evaluate_depth.h from /root/reference into
oracle/_ref/libdeval_ref.so (png++ stubbed, only the reference's file IO uses it);
tests/test_oracle_golden.py checks this restatement against it when it is present, and against
tests/golden/metrics.npz (generated from it by tests/golden/make_metrics_golden.py) everywhere.

Arithmetic follows the C++ statement by statement: per-pixel terms in float32 except the inverse
error (`1.0 / x` is double, evaluate_depth.h:61), accumulation in float32 in the reference's loop
order -- columns outer, rows inner (:50-51) -- which np.cumsum reproduces (sequential adds).
"""
import ctypes
import os

import numpy as np

METRICS = ["mae", "rmse", "inverse mae", "inverse rmse", "log mae", "log rmse", "scale invariant log",
           "abs relative", "squared relative"]
_F = np.float32


def _seq_sum(x):
    """Left-to-right float32 accumulation starting from 0.f."""
    x = np.asarray(x, dtype=_F)
    return _F(0) if x.size == 0 else np.cumsum(x, dtype=_F)[-1]


def depth_error_raw(first, second):
    """depthError(D_gt=first, D_ipol=second): pixels with first >= 0 count (io_depth.h:99-101)."""
    a = np.asarray(first, dtype=_F).T.reshape(-1)      # column-major traversal: u outer, v inner
    b = np.asarray(second, dtype=_F).T.reshape(-1)
    ok = a >= 0
    gt, ip = a[ok], b[ok]
    n = int(ok.sum())
    if n == 0:
        raise RuntimeError("ERROR: Ground truth defect")          # evaluate_depth.h:92-95 (throw 1)
    with np.errstate(all="ignore"):
        d_err = np.abs(gt - ip)
        d_sq = d_err * d_err
        d_inv = np.abs(1.0 / gt.astype(np.float64) - 1.0 / ip.astype(np.float64)).astype(_F)
        d_inv_sq = d_inv * d_inv
        lg, li = np.log(gt), np.log(ip)                            # std::log(float)
        d_log = np.abs(lg - li)
        d_log_sq = d_log * d_log
        e = np.zeros(9, dtype=_F)
        e[0] = _seq_sum(d_err)
        e[1] = _seq_sum(d_sq)
        e[2] = _seq_sum(d_inv)
        e[3] = _seq_sum(d_inv_sq)
        e[4] = _seq_sum(d_log)
        e[5] = _seq_sum(d_log_sq)
        log_sum = _seq_sum(lg - li)
        e[7] = _seq_sum(d_err / gt)
        e[8] = _seq_sum(d_sq / (gt * gt))
        nf = _F(n)
        e[0] = e[0] / nf
        e[1] = np.sqrt(e[1] / nf)
        e[2] = e[2] / nf
        e[3] = np.sqrt(e[3] / nf)
        e[4] = e[4] / nf
        nsl = e[5] / nf
        e[5] = np.sqrt(nsl)
        e[6] = np.sqrt(nsl - (log_sum * log_sum / (nf * nf)))
        e[7] = e[7] / nf
        e[8] = e[8] / nf
    return e


def depth_error(predicted, truth):
    """utils/img_utils.py:17-22: zeros are invalid (-1); `+ epsilon` (2.2e-16) is a float32 no-op."""
    p = np.array(predicted, dtype=_F, copy=True)
    t = np.array(truth, dtype=_F, copy=True)
    p[p == 0] = -1
    t[t == 0] = -1
    return depth_error_raw(p, t)


def eval_errors(errors):
    """evaluateErrors: {metric: [mean, min, max]} with utils.h's accumulators (float sum / n; the
    minimum starts at 1 and the maximum at 0, utils.h:42-55)."""
    errs = np.asarray(errors, dtype=_F).reshape(-1, 9)
    out = {}
    for j, name in enumerate(METRICS):
        col = errs[:, j]
        mean = _seq_sum(col) / _F(len(col))
        mn, mx = _F(1), _F(0)
        for v in col:
            if v < mn:
                mn = v
            if v > mx:
                mx = v
        out[name] = [float(mean), float(mn), float(mx)]
    return out


def compute_unc_rmse(uf_truth, uf_pred, d_candi):
    """utils/img_utils.py:183-194: mean |E_truth - E_pred| over the columns where both expected
    depths are finite, with the first and last predicted column zeroed.  uf_* [1, D, W] linear."""
    d = np.asarray(d_candi, dtype=np.float64).astype(_F)[:, None]
    with np.errstate(all="ignore"):
        et = (d * np.asarray(uf_truth, dtype=_F)[0]).sum(0, dtype=_F)
        ep = (d * np.asarray(uf_pred, dtype=_F)[0]).sum(0, dtype=_F)
    ep[0] = 0
    ep[-1] = 0
    m = ~np.isnan(et) & ~np.isnan(ep)
    et = np.where(m, et, _F(0))
    ep = np.where(m, ep, _F(0))
    with np.errstate(all="ignore"):
        return _F(np.abs(et - ep).sum(dtype=_F)) / _F(m.sum())


# ---- the reference itself, when oracle/_ref has been built (build container only) -----------------
_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libdeval_ref.so")


def reference_available():
    return os.path.exists(_REF)


def _ref():
    lib = ctypes.CDLL(_REF)
    fp = ctypes.POINTER(ctypes.c_float)
    lib.ref_depth_error.restype = ctypes.c_int
    lib.ref_depth_error.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int, fp]
    lib.ref_evaluate_errors.restype = None
    lib.ref_evaluate_errors.argtypes = [fp, ctypes.c_int, fp]
    return lib


def reference_depth_error(predicted, truth):
    """The reference's depthError through utils/img_utils.py:17-22's preparation."""
    p = np.array(predicted, dtype=_F, copy=True)
    t = np.array(truth, dtype=_F, copy=True)
    p[p == 0] = -1
    t[t == 0] = -1
    p, t = np.ascontiguousarray(p), np.ascontiguousarray(t)
    out = np.zeros(9, dtype=_F)
    fp = ctypes.POINTER(ctypes.c_float)
    rc = _ref().ref_depth_error(p.ctypes.data_as(fp), t.ctypes.data_as(fp), p.shape[1], p.shape[0],
                                out.ctypes.data_as(fp))
    if rc != 0:
        raise RuntimeError("reference depthError threw")
    return out


def reference_eval_errors(errors):
    errs = np.ascontiguousarray(np.asarray(errors, dtype=_F).reshape(-1, 9))
    out = np.zeros(27, dtype=_F)
    fp = ctypes.POINTER(ctypes.c_float)
    _ref().ref_evaluate_errors(errs.ctypes.data_as(fp), errs.shape[0], out.ctypes.data_as(fp))
    return {name: [float(v) for v in out[3 * j:3 * j + 3]] for j, name in enumerate(METRICS)}
