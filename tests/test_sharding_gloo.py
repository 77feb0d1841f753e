"""Multi-process (gloo, world_size 2, CPU) checks of the N > 1 host logic in sharding.py:

* unit round-robin (items / sequences) covers every unit exactly once and merges back in order;
* the depth-plane-sharded soft-max exchange protocol (PlaneShardedHead) reproduces the single-rank
  log-softmax / E[d] / Var / arg-max when the local passes are done by a stand-in that restates the
  dpv_shard_* kernels with torch CPU ops.  The stand-in is TEST code: the product's local passes
  are the CUDA kernels (sharding.CudaShardKernels) and there is no CPU implementation in the
  package; on the GPU box the same protocol runs over NCCL (tests/test_gpu_parity.py).
"""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class TorchShardStandIn:
    """What csrc/shard.cu computes, in torch CPU ops (test stand-in for CudaShardKernels)."""

    def local_stats(self, x, d_local, lo, out=None, slice_len=0):
        B, Dl, HW = x.shape
        m = x.max(dim=1).values
        first = (x == m.unsqueeze(1)).float().argmax(dim=1)          # first maximum wins inside the shard
        e = torch.exp(x - m.unsqueeze(1))
        s0 = e.sum(1)
        mu = (e * d_local.view(1, -1, 1)).sum(1) / s0
        m2 = (e * (d_local.view(1, -1, 1) - mu.unsqueeze(1)) ** 2).sum(1)
        st = torch.stack([m, s0, mu, m2, (first + lo).float()]).reshape(5, B * HW).contiguous()
        if slice_len:                                                # slice-major [G, 5, slice] (all-to-all layout)
            G = out.shape[0]
            pad = torch.zeros((5, G * slice_len))
            pad[:, :B * HW] = st
            st = pad.reshape(5, G, slice_len).permute(1, 0, 2).contiguous()
        if out is not None:
            out.copy_(st)
            return out
        return st

    @staticmethod
    def _merge(g):
        """g [G, 5, n] -> M, S, mean, var, arg-max (first rank attaining the maximum)."""
        m, s0, mu, m2, am = g[:, 0], g[:, 1], g[:, 2], g[:, 3], g[:, 4]
        M = m.max(dim=0).values
        sc = torch.exp(m - M.unsqueeze(0))
        w = s0 * sc
        S = w.sum(0)
        mean = (w * mu).sum(0) / S
        var = (m2 * sc + w * (mu - mean.unsqueeze(0)) ** 2).sum(0) / S
        first_rank = (m == M.unsqueeze(0)).float().argmax(dim=0)
        return M, S, mean, var, torch.gather(am, 0, first_rank.unsqueeze(0))[0]

    def merge_slice(self, recv, out):
        M, S, mean, var, am = self._merge(torch.nan_to_num(recv))
        out.copy_(torch.stack([M, torch.log(S), mean, var, am]))
        return out

    def finish(self, x, merged_all, want_logp, want_var, want_argmax):
        B, Dl, HW = x.shape
        G, _, sl = merged_all.shape
        flat = merged_all.permute(1, 0, 2).reshape(5, G * sl)[:, :B * HW].reshape(5, B, HW)
        logp = (x - flat[0].unsqueeze(1)) - flat[1].unsqueeze(1) if want_logp else None
        return logp, flat[2], (flat[3] if want_var else None), (flat[4].long() if want_argmax else None)

    def merge_finish(self, x, gathered, want_logp, want_var, want_argmax):
        B, Dl, HW = x.shape
        g = gathered.reshape(gathered.shape[0], 5, B, HW)
        m, s0, mu, m2, am = g[:, 0], g[:, 1], g[:, 2], g[:, 3], g[:, 4]
        M = m.max(dim=0).values
        sc = torch.exp(m - M.unsqueeze(0))
        w = s0 * sc
        S = w.sum(0)
        mean = (w * mu).sum(0) / S
        var = ((m2 * sc + w * (mu - mean.unsqueeze(0)) ** 2).sum(0) / S) if want_var else None
        first_rank = (m == M.unsqueeze(0)).float().argmax(dim=0)     # first rank attaining the maximum
        amax = torch.gather(am, 0, first_rank.unsqueeze(0))[0].long() if want_argmax else None
        logp = (x - M.unsqueeze(1)) - torch.log(S).unsqueeze(1) if want_logp else None
        return logp, mean, var, amax


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sharding = importlib.import_module("probabilistic-depth_b200.sharding")
        synth = importlib.import_module("probabilistic-depth_b200.synth")
        # ---- units: every rank takes r, r+G, ...; gather and merge -----------------------------
        n_units = 7
        mine = sharding.shard_units(n_units, rank, world)
        results = [u * 10 + 1 for u in mine]                      # "process" the unit
        gathered = [None] * world
        dist.all_gather_object(gathered, results)
        merged = sharding.merge_units(gathered, n_units)
        assert merged == [u * 10 + 1 for u in range(n_units)]
        t = torch.arange(n_units * 3).reshape(n_units, 3)
        assert torch.equal(sharding.shard_slice(t, rank, world), t[rank::world])
        assert sharding.max_over_ranks(float(rank + 1)) == float(world)

        # ---- planes: sharded soft-max == single-rank soft-max ----------------------------------
        B, D, H, W = 2, 10, 5, 7                                   # D not divisible by 4 ranks either
        g = torch.Generator().manual_seed(5)
        x = torch.randn((B, D, H, W), generator=g) * 4.0
        x[0, 3, 0, :] = x[0, 8, 0, :] = 50.0                       # tie across the two shards: first wins
        d = synth.depth_candidates(5.0, 40.0, D, 1.0)
        lo, hi = sharding.plane_range(D, rank, world)
        ref = torch.log_softmax(x, 1)
        dd = torch.from_numpy(d.astype(np.float32)).view(1, D, 1, 1)
        p = ref.exp()
        mean = (p * dd).sum(1)
        var = (p * (dd - mean.unsqueeze(1)) ** 2).sum(1)
        for exchange in ("gather", "scatter"):      # one all-gather / all-to-all + all-gather (B*H*W = 70 pixels:
            head = sharding.PlaneShardedHead(D, local=TorchShardStandIn(), exchange=exchange)   # ragged slices)
            assert (head.lo, head.hi) == (lo, hi)
            out = head(x[:, lo:hi].contiguous(), d)
            assert float((out["logp"] - ref[:, lo:hi]).abs().max()) < 1e-5, exchange
            assert float((out["depth"] - mean).abs().max()) < 1e-4, exchange
            assert float(((out["variance"] - var).abs() / var.clamp_min(1e-3)).max()) < 1e-4, exchange
            assert torch.equal(out["argmax"], torch.argmax(ref, 1)), exchange
            # every rank ends up with the same replicated per-pixel products
            rep = [None] * world
            dist.all_gather_object(rep, (out["depth"].numpy().tobytes(), out["argmax"].numpy().tobytes()))
            assert all(r == rep[0] for r in rep)
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_unit_and_plane_sharding_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


def test_plane_range_partitions_every_plane_once():
    sharding = importlib.import_module("probabilistic-depth_b200.sharding")
    for D in (64, 128, 10, 7):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.plane_range(D, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == D
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
