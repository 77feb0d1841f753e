"""models/models.py:BaseModel mirror against the reference BaseModel.

CPU part: same state_dict keys, shapes and (under the same torch seed) the same random-init
weights as the reference -- checked against the list make_model_golden.py stored, and against the
reference itself when /root/reference is present.  The reference loads checkpoints by POSITION
(trainer/base_trainer.py:83-90), so order and shapes are the contract.
GPU part: forward of all three nmodes on the sm_100a kernels vs the reference's CPU outputs
(tests/golden/model.npz); cuDNN convolutions run in full fp32 (TF32 off).
"""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import model_cases as MC  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "model.npz"))


def _ours(name):
    OM = importlib.import_module("probabilistic-depth_b200.models.models")
    torch.manual_seed(0)
    return OM.BaseModel(MC.cfg(name), 0)


@pytest.mark.parametrize("name", sorted(MC.MODES))
def test_state_dict_matches_reference_order_shapes_and_init(name):
    sd = _ours(name).state_dict()
    assert list(sd.keys()) == list(GOLD[name + "_keys"])
    assert [str(tuple(v.shape)) for v in sd.values()] == list(GOLD[name + "_shapes"])
    sums = np.array([float(v.double().sum()) for v in sd.values()])
    np.testing.assert_allclose(sums, GOLD[name + "_sums"], rtol=1e-12, atol=1e-12)
    assert len(sd) == {"default": 385, "default_upsample": 205, "default_feedback": 404}[MC.MODES[name][0]]


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference tree not present")
def test_weights_bit_identical_to_reference_under_same_seed():
    import warnings
    warnings.filterwarnings("ignore")
    import make_golden
    make_golden.import_reference()
    import models.models as RM
    for name in sorted(MC.MODES):
        torch.manual_seed(0)
        ref = RM.BaseModel(MC.cfg(name), 0).state_dict()
        ours = _ours(name).state_dict()
        assert all(torch.equal(ref[k], ours[k]) for k in ref), name


def _run(model, name, device):
    outs, prev = [], None
    for f in range(MC.MODES[name][3]):
        mi = MC.frame_inputs(name, f)
        t = {k: (torch.from_numpy(v).to(device) if isinstance(v, np.ndarray) and k != "d_candi" else v)
             for k, v in mi.items()}
        t["prev_output"] = prev
        with torch.no_grad():
            res = model([t])[0]
        outs.append(res)
        key = "%s_f%d_prev" % (name, f + 1)
        if key in GOLD.files:      # the reference's own hand-off, so that errors do not compound
            prev = torch.from_numpy(GOLD[key]).to(device)
            mine = torch.nn.functional.interpolate(res["output_refined"][-1], scale_factor=0.25, mode="nearest")
            assert float(((mine - prev).abs() / prev.abs().clamp_min(1.0)).max()) < 2e-2
    return outs


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(MC.MODES))
def test_forward_matches_reference_outputs(name):
    dpv = importlib.import_module("probabilistic-depth_b200")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = MC.batch_stat_norm(_ours(name).cuda())
    nmode = MC.MODES[name][0]
    launches0 = dpv._lib.launch_count()
    outs = _run(model, name, "cuda")
    assert dpv._lib.launch_count() - launches0 >= 3 * len(outs)       # the hot path ran on our kernels
    for f, res in enumerate(outs):
        assert set(res) == {"output", "output_refined", "flow", "flow_refined"}
        assert len(res["output"]) == (1 if nmode == "default" else 2)
        bv = res["output"][0] if nmode == "default_upsample" else res["output"][-1]
        refined = res["output_refined"][-1]
        assert tuple(bv.shape) == (1, MC.D, MC.H // 4, MC.W // 4) and tuple(refined.shape) == (1, MC.D, MC.H, MC.W)
        # ~60 fp32 conv layers (no normalisation in the decoder / 3-D net output) on cuDNN vs the CPU:
        # 1e-3-class agreement on log-probabilities; the kernels themselves agree to 1e-6 with the
        # oracle on identical inputs (test_gpu_parity.py)
        tol = 1e-2 if nmode == "default_feedback" else 2e-3
        want_bv, want_rf = GOLD["%s_f%d_bv" % (name, f)], GOLD["%s_f%d_refined" % (name, f)]
        got_bv, got_rf = bv[:, :, ::2, ::2].cpu().numpy(), refined[:, :, ::8, ::8].cpu().numpy()
        assert np.max(np.abs(got_bv - want_bv) / np.maximum(1.0, np.abs(want_bv))) < tol
        assert np.max(np.abs(got_rf - want_rf) / np.maximum(1.0, np.abs(want_rf))) < tol
        iu = dpv.utils.img_utils
        dq = iu.dpv_to_depthmap(bv, MC.D_CANDI, BV_log=True).cpu().numpy()
        dr = iu.dpv_to_depthmap(refined, MC.D_CANDI, BV_log=True).cpu().numpy()
        np.testing.assert_allclose(dq, GOLD["%s_f%d_depth_q" % (name, f)], rtol=tol, atol=tol)
        np.testing.assert_allclose(dr, GOLD["%s_f%d_depth" % (name, f)], rtol=tol, atol=tol)
        # log-DPVs are normalised
        assert float((torch.logsumexp(refined, 1)).abs().max()) < 1e-4


def _randomise_bn(m, seed):
    g = torch.Generator().manual_seed(seed)
    for mod in list(m.modules()) + [x for b in m.dres_modules for x in b.modules()]:
        if isinstance(mod, torch.nn.BatchNorm3d):
            n = mod.num_features
            mod.weight.data = 0.5 + torch.rand(n, generator=g)
            mod.bias.data = 0.2 * torch.randn(n, generator=g)
            if mod.running_mean is not None:
                mod.running_mean.data = 0.1 * torch.randn(n, generator=g)
                mod.running_var.data = 0.5 + torch.rand(n, generator=g)


@pytest.mark.parametrize("bn_avg", [True, False])
def test_base3d_layer_spec_describes_the_module(bn_avg):
    """`Base3DConvs.spec_from_module` (what the tensor-core path is built from: weights, which BatchNorms fold and
    which use batch statistics, ReLUs, residual blocks) evaluated by the float64 oracle reproduces the module's own
    forward -- for the mirror Base3D in eval(), with running statistics (bn_avg true: dres0 / classify fold, the
    unregistered residual blocks stay on batch statistics) and without (everything on batch statistics)."""
    from oracle import dpv_oracle as O
    OM = importlib.import_module("probabilistic-depth_b200.models.models")
    ops = importlib.import_module("probabilistic-depth_b200.ops")
    torch.manual_seed(3)
    m = OM.Base3D(4, dres_count=2, feature_dim=32, bn_running_avg=bn_avg).eval()
    _randomise_bn(m, 5)
    spec = ops.Base3DConvs.spec_from_module(m)
    assert [L.get("block") for L in spec] == [None, None, "in", "out", "in", "out", None, None]
    assert [bool(L["bn"] and L["bn"]["batch_stats"]) for L in spec] == \
        ([False, False, True, True, True, True, False, False] if bn_avg else [True] * 7 + [False])
    vol = torch.randn((2, 4, 5, 6, 7), generator=torch.Generator().manual_seed(8))
    with torch.no_grad():
        want = m(vol)
        got = O.base3d(vol, spec)
    assert float((got - want.double()).abs().max()) <= 1e-5 * float(want.abs().max())


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference tree not present")
def test_base3d_layer_spec_describes_the_reference_module():
    """The same for the reference's own Base3D (models/models.py:376-438), built on this GPU-less container with
    nn.Module.cuda replaced by the identity for the duration of its constructor."""
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import dpv_oracle as O, reference_loader
    ops = importlib.import_module("probabilistic-depth_b200.ops")
    ref = reference_loader.load()
    try:
        real_cuda = torch.nn.Module.cuda
        torch.nn.Module.cuda = lambda self, device=None: self
        try:
            torch.manual_seed(3)
            m = ref.models.Base3D(4, dres_count=2, feature_dim=32, bn_running_avg=True, id=0).eval()
        finally:
            torch.nn.Module.cuda = real_cuda
        _randomise_bn(m, 5)
        spec = ops.Base3DConvs.spec_from_module(m)
        vol = torch.randn((2, 4, 5, 6, 7), generator=torch.Generator().manual_seed(8))
        with torch.no_grad():
            want = m(vol, prob=False)
            got = O.base3d(vol, spec)
        assert float((got - want.double()).abs().max()) <= 1e-5 * float(want.abs().max())
    finally:
        reference_loader.restore()
