#!/usr/bin/env python
"""Golden outputs of the reference's eval metrics -> tests/golden/metrics.npz.

depth_error / eval_errors: the reference's OWN C++ (external/deval_lib/src/evaluate_depth.h) compiled by
oracle/ref_deval/Makefile into oracle/_ref/libdeval_ref.so, called through the preparation of
utils/img_utils.py:17-22.  compute_unc_rmse: the reference's Python (utils/img_utils.py:183-194) imported
from /root/reference with the shims of SURVEY.md 9.2.  Run in the build container only:
    make -C oracle/ref_deval && python tests/golden/make_metrics_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import cases  # noqa: E402
from oracle import depth_metrics as M  # noqa: E402

REF = "/root/reference"
sys.path.insert(0, REF)
import external.deval_lib as _dl  # noqa: E402
stub = types.ModuleType("external.deval_lib.pyevaluatedepth_lib")
stub.evaluateErrors = lambda e: {}
stub.depthError = lambda a, b: [0.] * 9
sys.modules[stub.__name__] = stub
_dl.pyevaluatedepth_lib = stub
import utils.img_utils as ref_u  # noqa: E402

assert M.reference_available(), "build oracle/_ref first: make -C oracle/ref_deval"
out = {}
per_item = []
for name in cases.METRICS_CASES:
    c = cases.metrics_case(name)
    e = M.reference_depth_error(c["predicted"], c["truth"])
    out[name + "_errors"] = e
    per_item.append(e)
    # with the trainer's preparation (default_trainer.py:247-254) done on the host as the reference does
    t = c["truth"].copy()
    t[t >= c["d_max"]] = c["d_max"]
    out[name + "_errors_prepared"] = M.reference_depth_error(c["predicted"] * c["mask"], t)
stats = M.reference_eval_errors(per_item)
out["eval_errors"] = np.array([stats[n] for n in M.METRICS], dtype=np.float32)
for name in cases.UNC_RMSE_CASES:
    c = cases.unc_rmse_case(name)
    r = ref_u.compute_unc_rmse(torch.from_numpy(c["truth"].copy()), torch.from_numpy(c["pred"].copy()), c["d_candi"])
    out["unc_rmse_" + name] = np.float32(r.item())
np.savez(os.path.join(HERE, "metrics.npz"), **out)
for k, v in out.items():
    print(k, v)
