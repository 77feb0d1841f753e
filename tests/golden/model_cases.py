"""Seeded model-level inputs (the input dict of BaseModel.forward, reference
models/models.py:504-539 and kittiloader/batch_scheduler.py:250-263) shared by
make_model_golden.py (reference side) and tests/test_model_mirror.py."""
import importlib
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
synth = importlib.import_module("probabilistic-depth_b200.synth")

H, W, D = 256, 384, 64
D_CANDI = synth.depth_candidates(5.0, 40.0, D, 1.0)
MODES = {
    # name: (nmode, bn_avg, pose kind, frames)
    "default_stereo": ("default", True, "stereo", 1),
    "upsample_mono": ("default_upsample", False, "mono", 1),
    "feedback_mono": ("default_feedback", True, "mono", 2),
}


class AttrDict(dict):
    __getattr__ = dict.__getitem__


def cfg(name):
    nmode, bn_avg, _, _ = MODES[name]
    return AttrDict(var=AttrDict(sigma_soft_max=10.0, feature_dim=64, nmode=nmode, ndepth=D, bn_avg=bn_avg))


def frame_inputs(name, frame, batch=1):
    """numpy input dict of frame `frame` (without prev_output, which the caller chains)."""
    _, _, kind, _ = MODES[name]
    seed = 7000 + 97 * sorted(MODES).index(name) + frame
    cam = synth.camera(W // 4, H // 4, batch)
    poses = synth.stereo_poses(batch) if kind == "stereo" else synth.mono_poses(batch)
    dm, mk = synth.sparse_depth(seed + 1, batch, H // 4, W // 4)
    return dict(rgb=synth.randn(seed, batch, 2, 3, H, W), intrinsics=cam["intrinsics"],
                unit_ray=cam["unit_ray"], src_cam_poses=poses, d_candi=D_CANDI, dmaps=dm, masks=mk)


def batch_stat_norm(model):
    """Put the BatchNorm layers (only) in batch-statistics mode.

    With random-init weights and identity running statistics the activations of the ~60-layer
    network grow to ~1e12 and the soft-max input is quantised at 65536 per ulp: nothing can be
    compared there.  Normalising with the batch statistics (deterministic for a given input) keeps
    every stage O(1-100); the reference and the mirror are driven the same way.  This is also what
    the reference itself does in eval with bn_avg=false (models/models.py:25-30)."""
    import torch.nn as nn
    model.eval()
    for m in model.modules():
        if isinstance(m, nn.modules.batchnorm._BatchNorm):
            m.train()
    for blk in getattr(getattr(model, "based_3d", None), "dres_modules", []):
        for m in blk.modules():
            if isinstance(m, nn.modules.batchnorm._BatchNorm):
                m.train()
    return model
