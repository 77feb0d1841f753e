#!/usr/bin/env python
"""tests/golden/lidar.npz: depth maps from the reference's own generate_depth (utils_lib.cpp:86-160 compiled by
oracle/ref_utils_lib against minimal Eigen / OpenCV / pybind11 stand-ins: `*_dmap_ref`), the oracle's
(oracle/c/lidar_depthmap.c: `*_dmap_oracle`, identical), and the reference's minpool (utils/img_utils.py:87-95,
imported from /root/reference) applied to them as kittiloader/kitti.py:706 does."""
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
from oracle import lidar_depthmap as L  # noqa: E402

sys.path.insert(0, "/root/reference")
import external.deval_lib as _dl  # noqa: E402
stub = types.ModuleType("external.deval_lib.pyevaluatedepth_lib")
sys.modules[stub.__name__] = stub
_dl.pyevaluatedepth_lib = stub
import utils.img_utils as ref_u  # noqa: E402

out = {}
for name in cases.LIDAR_CASES:
    c = cases.lidar_case(name)
    dmap = L.generate_depth(c["velo"], c["intr"], c["M"], c["width"], c["height"], c["filtering"], c["filterdiff"])
    out[name + "_dmap_oracle"] = dmap
    ref_map = L.reference_generate_depth(c["velo"], c["intr"], c["M"], c["width"], c["height"], c["filtering"],
                                         c["filterdiff"])
    assert np.array_equal(ref_map, dmap), name
    out[name + "_dmap_ref"] = ref_map
    t = torch.Tensor(dmap).unsqueeze(0).unsqueeze(0)
    out[name + "_small_ref"] = ref_u.minpool(t, 4, 1000).squeeze(0).squeeze(0).numpy()      # kitti.py:706
    out[name + "_small_ref_plain"] = ref_u.minpool(t, 4).squeeze(0).squeeze(0).numpy()
    print(name, "returns", int((dmap > 0).sum()), "of", dmap.size, "pooled", int((out[name + "_small_ref"] > 0).sum()))
np.savez_compressed(os.path.join(HERE, "lidar.npz"), **out)
