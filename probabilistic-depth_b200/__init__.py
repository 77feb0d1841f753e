"""B200-native (sm_100a) depth-probability-volume hot path of soulslicer/probabilistic-depth.

The directory name carries a hyphen, so import it with
    dpv = importlib.import_module("probabilistic-depth_b200")
Sub-modules mirror the reference's layout for the path only:
    dpv.warping.homography, dpv.utils.img_utils, dpv.models.correlation_native,
    dpv.models.correlation_package.correlation
`dpv.ops` holds the batched tensor-level entry points, `dpv.synth` the seeded synthetic
inputs, `dpv.patch_reference()` swaps the hot-path functions of an imported reference tree
for ours.
"""
import importlib as _importlib

from . import _lib, build, ops, synth   # noqa: F401
from ._lib import DpvError              # noqa: F401

_SUB = ("warping.homography", "utils.img_utils", "models.correlation_native",
        "models.correlation_package.correlation", "sharding", "pipeline", "models.models")


def __getattr__(name):
    if name in ("warping", "utils", "models", "sharding", "pipeline"):
        return _importlib.import_module("." + name, __name__)
    raise AttributeError(name)


class _FunctionalProxy:
    """Stands in for the `F` (torch.nn.functional) name of the reference's models/models.py: every
    attribute resolves to torch.nn.functional, except that the depth-bin log-softmax sites of the hot
    path (models/models.py:351,560,637,694 -- 4-D fp32 CUDA volumes, dim=1, no autograd) run on
    dpv_head.  Anything else (the 5-D soft-max of Base3D's prob branch, CPU tensors, other dtypes) is
    the reference's own call to torch, untouched."""

    def __init__(self, functional):
        self._functional = functional

    def __getattr__(self, name):
        return getattr(self._functional, name)

    def log_softmax(self, x, dim=None, **kw):
        import torch
        if (dim == 1 and not kw and isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 4
                and x.dtype == torch.float32 and not (torch.is_grad_enabled() and x.requires_grad)):
            return ops.log_softmax(x)
        return self._functional.log_softmax(x, dim=dim, **kw)


def patch_reference(homography=None, img_utils=None, models=None):
    """Replace the hot-path functions of already-imported reference modules by ours.

    The reference resolves them at call time through the module object
    (models/models.py:6,541,625; trainer/default_trainer.py:232-243), so patching the
    attributes is enough for `BaseModel.forward` and the eval loop to run on the kernels.
    `models` (the reference's models.models module) additionally gets its depth-bin
    `F.log_softmax` sites routed to dpv_head through a proxy for the module-level name `F`.

    FORWARD ONLY: the kernels record no autograd graph.  Every op raises DpvError when called
    with grad enabled on a tensor that requires grad (the reference's training loss calls
    dpv_to_depthmap under grad, losses/losses.py:82-88) -- run the patched reference under
    torch.no_grad(), as its eval loop does (trainer/default_trainer.py:171).
    tests/test_reference_gpu.py runs the reference BaseModel unpatched and patched on cuda:0.
    """
    from .warping import homography as ours_h
    from .utils import img_utils as ours_u
    if homography is not None:
        for n in ("est_swp_volume_v4", "warp_feature", "_back_warp_homo_parallel"):
            setattr(homography, n, getattr(ours_h, n))
    if img_utils is not None:
        for n in ("dpv_to_depthmap", "gen_dpv_withmask", "gen_ufield", "compute_unc_field",
                  "depth_error", "eval_errors", "compute_unc_rmse", "minpool"):
            setattr(img_utils, n, getattr(ours_u, n))
    if models is not None and not isinstance(models.F, _FunctionalProxy):
        models.F = _FunctionalProxy(models.F)
