"""Golden for SURVEY 8f rank 2, second half: the UNMODIFIED reference's Base3D (models/models.py:376-438) on a seeded
volume, fp32 CPU (build container only):

    python tests/golden/make_golden_base3d.py

base3d.npz: volume [2,4,6,7,9], the residual Base3D(volume, prob=False) returns, and every layer's parameters
(convolution weights; BatchNorm gamma / beta / running mean / running var / eps and whether the layer normalises
with BATCH statistics).  The module is built as the feedback model builds it (Base3D(4, dres_count=2,
feature_dim=32, bn_running_avg=True), models/models.py:464) and put in eval(): dres0 and classify then use their
running statistics, while the UNREGISTERED dres_modules (a plain list, :394-399) never leave training mode and use
batch statistics -- the golden pins exactly that.  BatchNorm parameters and running statistics are randomised so
that none of the affine maps is the identity.  Base3D.__init__ calls .cuda() on the residual blocks; on this
GPU-less container nn.Module.cuda is replaced by the identity for the duration of the construction.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def layer_list(m):
    """(conv, bn or None) in forward order."""
    L = [(m.dres0[0][0], m.dres0[0][1]), (m.dres0[2][0], m.dres0[2][1])]
    for blk in m.dres_modules:
        L += [(blk[0][0], blk[0][1]), (blk[2][0], blk[2][1])]
    L += [(m.classify[0][0], m.classify[0][1]), (m.classify[2], None)]
    return L


def main():
    warnings.filterwarnings("ignore")
    from oracle import reference_loader
    ref = reference_loader.load()
    torch.manual_seed(5)
    torch.set_num_threads(8)
    real_cuda = torch.nn.Module.cuda
    torch.nn.Module.cuda = lambda self, device=None: self
    try:
        m = ref.models.Base3D(4, dres_count=2, feature_dim=32, bn_running_avg=True, id=0)
    finally:
        torch.nn.Module.cuda = real_cuda
    g = torch.Generator().manual_seed(11)
    for conv, bn in layer_list(m):
        if bn is None:
            continue
        n = bn.num_features
        bn.weight.data = 0.5 + torch.rand(n, generator=g)
        bn.bias.data = 0.2 * torch.randn(n, generator=g)
        bn.running_mean.data = 0.1 * torch.randn(n, generator=g)
        bn.running_var.data = 0.5 + torch.rand(n, generator=g)
    m.eval()
    out = {}
    for i, (conv, bn) in enumerate(layer_list(m)):
        out["w%d" % i] = conv.weight.detach().numpy().copy()
        if bn is not None:
            out["gamma%d" % i] = bn.weight.detach().numpy().copy()
            out["beta%d" % i] = bn.bias.detach().numpy().copy()
            out["mean%d" % i] = bn.running_mean.detach().numpy().copy()       # before the forward pass touches them
            out["var%d" % i] = bn.running_var.detach().numpy().copy()
            out["eps%d" % i] = np.float64(bn.eps)
            out["batch%d" % i] = np.int32(1 if (bn.training or not bn.track_running_stats) else 0)
    rs = np.random.RandomState(21)
    vol = rs.standard_normal((2, 4, 6, 7, 9)).astype(np.float32)
    vol[:, 0] = -4.0 + vol[:, 0]                       # channel 0 / 1: log-probability-like planes
    vol[:, 1] = -4.0 + 0.5 * vol[:, 1]
    with torch.no_grad():
        resi = m(torch.from_numpy(vol), prob=False)
    out["volume"] = vol
    out["resi"] = resi.numpy()
    print("layers", len(layer_list(m)), "batch-stat layers", [int(out.get("batch%d" % i, -1)) for i in range(8)])
    print("resi", tuple(resi.shape), float(resi.min()), float(resi.max()))
    np.savez_compressed(os.path.join(HERE, "base3d.npz"), **out)
    print("base3d.npz", os.path.getsize(os.path.join(HERE, "base3d.npz")))


if __name__ == "__main__":
    main()
