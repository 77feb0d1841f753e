"""TEST INFRASTRUCTURE: import the UNMODIFIED reference (soulslicer/probabilistic-depth) as the second oracle.

Only tests/, tests/golden/make_*.py, __graft_entry__.smoke() and bench.py's reference / incumbent /
cpu_baseline legs may import this; nothing under probabilistic-depth_b200/ does.

Where the tree comes from, in this order: $DPV_REFERENCE, /root/reference (the build container),
baseline/_ref (a git-ignored copy of the reference's own .py files made by `ship()`, called from
__graft_entry__.build() -- it travels to the GPU box with the repo snapshot, where /root/reference
does not exist).  The files are imported as they are; three shims make them importable on this image
(SURVEY.md 8c): a stub for the C++ metric library utils/img_utils.py:11 imports at module scope, a
Tensor.repeat wrapper for the stray positional arguments at utils/img_utils.py:342 that torch >= 2
rejects, and -- on a CUDA-less host only -- a no-op nn.Module.cuda for models/models.py:399.
"""
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIPPED = os.path.join(ROOT, "baseline", "_ref")
# what the hot path and its callers need: the Python modules only (no C++ / CUDA extension, no data)
SHIP = ("warping/__init__.py", "warping/homography.py", "warping/view.py",
        "models/__init__.py", "models/models.py", "models/correlation_native.py", "models/packnet.py",
        "utils/__init__.py", "utils/img_utils.py",
        "external/__init__.py", "external/deval_lib/__init__.py",
        "configs/default_mono.json", "configs/default_stereo.json", "configs/default_mono_feedback.json",
        "configs/default_mono_upsample.json")


def ship(src="/root/reference", dst=SHIPPED):
    """Copy the reference's own files, unmodified, into the git-ignored baseline/_ref (the one place the
    base contract allows a reference install).  Returns the number of files copied (0 = no source)."""
    if not os.path.isdir(os.path.join(src, "warping")):
        return 0
    n = 0
    for rel in SHIP:
        s, d = os.path.join(src, rel), os.path.join(dst, rel)
        if os.path.isfile(s):
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
            n += 1
    return n


def find():
    for cand in (os.environ.get("DPV_REFERENCE"), "/root/reference", SHIPPED):
        if cand and os.path.isfile(os.path.join(cand, "warping", "homography.py")):
            return cand
    return None


def available():
    return find() is not None


_loaded = {}


def load():
    """Import the reference modules of the path; returns a namespace with `root`, `homography`,
    `img_utils`, `models` (models/models.py), `correlation_native`.  Raises if no tree is present."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    import torch
    root = find()
    if root is None:
        raise RuntimeError("reference tree not found (neither /root/reference nor baseline/_ref)")
    if root not in sys.path:
        sys.path.insert(0, root)
    import external.deval_lib as _dl
    stub = types.ModuleType("external.deval_lib.pyevaluatedepth_lib")
    stub.evaluateErrors = lambda e: {}
    stub.depthError = lambda a, b: [0.0] * 9
    sys.modules[stub.__name__] = stub
    _dl.pyevaluatedepth_lib = stub
    if not getattr(torch.Tensor.repeat, "_dpv_shim", False):
        _rep = torch.Tensor.repeat

        def _repeat(self, *a):
            if a and isinstance(a[0], (list, tuple)):
                a = (a[0],)
            return _rep(self, *a)
        _repeat._dpv_shim = True
        torch.Tensor.repeat = _repeat
    if not torch.cuda.is_available():
        torch.nn.Module.cuda = lambda s, *a, **k: s
    import warping.homography as homography
    import utils.img_utils as img_utils
    import models.models as models
    import models.correlation_native as correlation_native
    # pristine copies of everything patch_reference() may overwrite, so a test can always get back
    orig = {(m.__name__, n): getattr(m, n) for m, names in (
        (homography, ("est_swp_volume_v4", "warp_feature", "_back_warp_homo_parallel")),
        (img_utils, ("dpv_to_depthmap", "gen_dpv_withmask", "gen_ufield", "compute_unc_field", "depth_error",
                     "eval_errors", "compute_unc_rmse", "minpool")),
        (models, ("F",))) for n in names if hasattr(m, n)}
    _loaded.update(root=root, homography=homography, img_utils=img_utils, models=models,
                   correlation_native=correlation_native, _orig=orig)
    return types.SimpleNamespace(**_loaded)


def restore():
    """Undo patch_reference(): put the reference's own functions back."""
    if not _loaded:
        return
    for (mod, name), fn in _loaded["_orig"].items():
        setattr(sys.modules[mod], name, fn)


class KittiCfg:
    """gen_ufield only reads cfg.data.dataset_path (utils/img_utils.py:277)."""
    class data:
        dataset_path = "./kitti/"
