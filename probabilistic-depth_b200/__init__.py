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


def patch_reference(homography=None, img_utils=None):
    """Replace the hot-path functions of already-imported reference modules by ours.

    The reference resolves them at call time through the module object
    (models/models.py:6,541,625; trainer/default_trainer.py:232-243), so patching the
    attributes is enough for `BaseModel.forward` and the eval loop to run on the kernels.
    """
    from .warping import homography as ours_h
    from .utils import img_utils as ours_u
    if homography is not None:
        for n in ("est_swp_volume_v4", "warp_feature", "_back_warp_homo_parallel"):
            setattr(homography, n, getattr(ours_h, n))
    if img_utils is not None:
        for n in ("dpv_to_depthmap", "gen_dpv_withmask", "gen_ufield", "compute_unc_field",
                  "depth_error", "eval_errors", "compute_unc_rmse"):
            setattr(img_utils, n, getattr(ours_u, n))
