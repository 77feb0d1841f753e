"""CPU-side checks: the C-ABI library builds, loads and exports what include/dpv_b200.h
declares; host helpers (shift tables, constants, synthetic inputs) behave; the product refuses
to run without a GPU instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import dpv_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "dpv_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dpv_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_loads_and_exports_every_declared_symbol(dpv):
    path = dpv.build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libdpv_sm100a.so does not export %s" % n
    assert sorted(dpv._lib.PROTOTYPES) == names, "ctypes prototypes out of sync with the header"
    assert dpv._lib.load().dpv_abi_version() == 2
    assert dpv._lib.load().dpv_error_string(-1).decode().startswith("dpv: bad argument")


def test_sass_is_sm100a_only(dpv):
    out = os.popen("cuobjdump -lelf %s 2>/dev/null" % dpv.build.build()).read()
    if not out:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback(dpv):
    x = torch.zeros((1, 64, 4, 4))
    with pytest.raises(dpv.DpvError):
        dpv.ops.head(x, dpv.synth.depth_candidates())
    with pytest.raises(dpv.DpvError):
        dpv.ops.correlation(torch.zeros(1, 4, 8, 8), torch.zeros(1, 4, 8, 8))


def test_bad_arguments_are_rejected_without_a_device(dpv):
    lib = dpv._lib.load()
    assert lib.dpv_correlation(None, None, None, 1, 1, 1, 1, 4, None) == -1
    assert lib.dpv_head(None, None, None, None, None, None, None, None, None, 1, 64, 4, 4, 0, None) == -1
    with pytest.raises(dpv.DpvError):
        dpv._lib.check(-2)


@pytest.mark.parametrize("H,W", [(256, 384), (64, 96), (37, 51), (384, 1280)])
def test_shift_tables_reproduce_the_reference_shift(dpv, H, W):
    """ops.shift_luts against the oracle's restatement of utils/img_utils.py:291-302,336."""
    rf, ri, cf, ci = [t.numpy() for t in dpv.ops.shift_luts(H, W, 5, "cpu")]
    img = torch.arange(1, H * W + 1, dtype=torch.float32).reshape(1, 1, H, W)
    g_fwd, g_inv = O._shift_grids(H, W, 5)
    for grid, rows, cols in ((g_fwd, rf, cf), (g_inv, ri, ci)):
        want = torch.nn.functional.grid_sample(img, grid, mode="nearest", align_corners=False)[0, 0]
        src = np.where((rows[:, None] >= 0) & (cols[None, :] >= 0),
                       rows[:, None].astype(np.int64) * W + cols[None, :] + 1, 0)
        np.testing.assert_array_equal(want.numpy().astype(np.int64), src)
    # what SURVEY.md 7 records for even sizes: +5 drops rows 0-4, -5 drops the last six
    if H % 2 == 0:
        assert list(rf[:6]) == [-1] * 5 + [0] and rf[-1] == H - 6
        assert ri[0] == 5 and list(ri[-6:]) == [-1] * 6
    if W % 2 == 0:
        assert cf[-1] == -1 and cf[0] == 0


def test_two_sigma_sq_matches_reference_arithmetic(dpv):
    v = torch.tensor(0.3)
    want = float(2 * torch.pow(torch.sqrt(v), 2.0))
    assert dpv.ops.two_sigma_sq(0.3) == want


def test_synthetic_camera_matches_survey(dpv):
    K = dpv.synth.intrinsics(96, 64)
    np.testing.assert_allclose(K, [[113.636, 0, 48], [0, 133.576, 32], [0, 0, 1]], rtol=1e-5)
    rays = dpv.synth.unit_rays(96, 64)
    assert rays.shape == (3, 6144) and rays.dtype == np.float32
    # pixel centres project back onto themselves: K ray = (x + 0.5, y + 0.5, 1)
    uv = K @ rays
    np.testing.assert_allclose(uv[0].reshape(64, 96)[0], np.arange(96) + 0.5, atol=1e-4)
    np.testing.assert_allclose(uv[1].reshape(64, 96)[:, 0], np.arange(64) + 0.5, atol=1e-4)
    d = dpv.synth.depth_candidates()
    assert d.dtype == np.float64 and d[0] == 5.0 and d[-1] == 40.0 and len(d) == 64


def test_mirror_modules_keep_the_reference_names(dpv):
    h = dpv.warping.homography
    for n in ("est_swp_volume_v4", "warp_feature", "_back_warp_homo_parallel", "get_rel_extrinsicM"):
        assert callable(getattr(h, n))
    u = dpv.utils.img_utils
    for n in ("dpv_to_depthmap", "gen_dpv_withmask", "gen_ufield", "compute_unc_field", "powerf"):
        assert callable(getattr(u, n))
    assert u.epsilon == O.EPSILON
    np.testing.assert_array_equal(u.powerf(5.0, 40.0, 64, 1.0), dpv.synth.depth_candidates())
    c = dpv.models.correlation_native.Correlation(max_displacement=4)
    assert c.output_dim == 9 and c.pad_size == 4


@pytest.mark.parametrize("H,W", [(256, 384), (64, 96), (384, 1280)])
def test_fused_uf_tables_match_the_closed_form(dpv, H, W):
    """dpv_uf_fused_tables (host helper, no device work) on the reference's +/-5-row shifts gives
    the closed form SURVEY.md 8c verified against gen_ufield: numerator rows y <= H-7 test shifted
    row y+5, denominator source rows y <= H-6, column W-1 excluded everywhere."""
    import ctypes
    luts = [t.numpy().astype(np.int32).copy() for t in dpv.ops.shift_luts(H, W, 5, "cpu")]
    row_tab = np.zeros((H, 4), dtype=np.int32)
    col_tab = np.zeros((W,), dtype=np.int32)
    rc = dpv._lib.load().dpv_uf_fused_tables(*[a.ctypes.data for a in luts], H, W,
                                             row_tab.ctypes.data, col_tab.ctypes.data)
    assert rc == 0
    y = np.arange(H)
    np.testing.assert_array_equal(row_tab[:, 0], np.where(y <= H - 7, y + 5, -1))
    np.testing.assert_array_equal(row_tab[:, 1], 0)
    np.testing.assert_array_equal(row_tab[:, 2], np.where(y <= H - 6, y + 5, -1))
    np.testing.assert_array_equal(row_tab[:, 3], (y < 5).astype(np.int32))
    np.testing.assert_array_equal(col_tab[:-1], 1 | 4)
    assert col_tab[-1] == 8
    # shifts that do not compose to "same pixel or padding" are refused, not mis-handled
    bad = [a.copy() for a in luts]
    bad[1][:] = np.clip(np.arange(H) + 4, 0, H - 1)
    rc = dpv._lib.load().dpv_uf_fused_tables(*[a.ctypes.data for a in bad], H, W,
                                             row_tab.ctypes.data, col_tab.ctypes.data)
    assert rc == -2


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` needs no GPU: exactly one JSON line on stdout with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0
    # "reference" = the imported, unmodified reference (here: /root/reference or baseline/_ref); "port" only where
    # no reference tree exists
    from oracle import reference_loader
    assert d["cpu_baseline"]["kind"] == ("reference" if reference_loader.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]
