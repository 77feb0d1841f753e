#!/usr/bin/env python
"""Dev tool: the 3-D convolution kernels step by step against torch float64 (one layer raw output, batch statistics,
BatchNorm apply, the whole Base3D stack against the oracle)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
dpv = importlib.import_module("probabilistic-depth_b200")
from oracle import dpv_oracle as O
import cases
ops, _lib = dpv.ops, importlib.import_module("probabilistic-depth_b200._lib")
lib = _lib.load()
F = torch.nn.functional
p = lambda t: None if t is None else t.data_ptr()
st = torch.cuda.current_stream().cuda_stream
g = torch.Generator().manual_seed(3)
for (B, C, D, H, W) in ((1, 32, 5, 6, 7), (2, 4, 9, 10, 13), (1, 32, 16, 24, 40)):
    x = torch.randn((B, C, D, H, W), generator=g)
    w = torch.randn((32, C, 3, 3, 3), generator=g) * (2.0 / (27 * 32)) ** 0.5
    n = int(lib.dpv_conv3d_packed_floats(B, D, H, W))
    mk = lambda: torch.empty((n,), device="cuda")
    a_hi, a_lo, raw = mk(), mk(), mk()
    w_hi, w_lo = torch.empty((27, 32, 32), device="cuda"), torch.empty((27, 32, 32), device="cuda")
    stats = torch.zeros(64, device="cuda", dtype=torch.float64)
    _lib.check(lib.dpv_conv3d_pack(p(x.cuda()), p(a_hi), p(a_lo), B, C, D, H, W, st))
    _lib.check(lib.dpv_conv3d_pack_weights(p(w.cuda()), None, p(w_hi), p(w_lo), 32, C, st))
    _lib.check(lib.dpv_conv3d_c32(p(a_hi), p(a_lo), p(w_hi), p(w_lo), None, None, None, None, None, None, p(raw), p(stats),
                                  B, D, H, W, 0, C, st))
    torch.cuda.synchronize()
    want = F.conv3d(x.double(), w.double(), padding=1)                           # [B,32,D,H,W]
    got = raw.view(B, D + 2, H + 2, W + 2, 32)[:, 1:-1, 1:-1, 1:-1].permute(0, 4, 1, 2, 3).cpu().double()
    border = raw.view(B, D + 2, H + 2, W + 2, 32).clone()
    border[:, 1:-1, 1:-1, 1:-1] = 0
    scale = float(want.abs().max())
    print("conv %s: max err %.2e of scale %.2f; border max %.1e" % ((B, C, D, H, W), float((got - want).abs().max()) / scale, scale,
                                                                   float(border.abs().max())))
    s1, s2 = want.sum((0, 2, 3, 4)), (want * want).sum((0, 2, 3, 4))
    print("   stats: sum err %.2e  sumsq err %.2e (relative)" % (float(((stats[:32].cpu() - s1).abs() / s1.abs().clamp_min(1)).max()),
                                                                 float(((stats[32:].cpu() - s2).abs() / s2).max())))
    gamma, beta = torch.rand(32, generator=g) + 0.5, torch.randn(32, generator=g) * 0.2
    o_hi, o_lo = mk(), mk()
    g_cu, b_cu = gamma.cuda(), beta.cuda()          # held: a temporary's block would be reused by the next one
    _lib.check(lib.dpv_conv3d_bn_apply(p(raw), p(stats), p(g_cu), p(b_cu), 1e-5, p(a_hi), p(a_lo), p(o_hi), p(o_lo),
                                       B, D, H, W, 1, st))
    torch.cuda.synchronize()
    wantb = F.relu(F.batch_norm(want, None, None, gamma.double(), beta.double(), True, 0.0, 1e-5) +
                   (x.double() if C == 32 else F.pad(x.double(), (0, 0, 0, 0, 0, 0, 0, 32 - C))))
    gotb = (o_hi + o_lo).view(B, D + 2, H + 2, W + 2, 32)[:, 1:-1, 1:-1, 1:-1].permute(0, 4, 1, 2, 3).cpu().double()
    print("   bn_apply (+res, relu): max err %.2e of scale %.2f" % (float((gotb - wantb).abs().max()) / float(wantb.abs().max()),
                                                                   float(wantb.abs().max())))
# the whole stack: golden, then a random net against the oracle
gd = np.load(os.path.join(ROOT, "tests", "golden", "base3d.npz"))
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
net = ops.Base3DConvs(cases.base3d_layers(gd, cu))
got = net(cu(gd["volume"])).cpu().numpy()
print("Base3D vs reference golden: max err %.2e of scale %.1f" % (float(np.abs(got - gd["resi"]).max()) / float(np.abs(gd["resi"]).max()),
                                                                 float(np.abs(gd["resi"]).max())))
layers = cases.base3d_layers(gd)
vol = torch.randn((2, 4, 12, 10, 14), generator=g)
want = O.base3d(vol, layers)
got = net(vol.cuda()).cpu().double()
print("Base3D %s vs oracle: max err %.2e of scale %.1f" % (tuple(vol.shape), float((got - want).abs().max()) / float(want.abs().max()),
                                                         float(want.abs().max())))
