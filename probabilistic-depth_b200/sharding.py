"""Multi-GPU sharding of the DPV path: one process per GPU, `torch.distributed` for plumbing.

Two partitionings:

* items / sequences -- every batch item is independent (the reference loops over items,
  models/models.py:528-550,614-627) and in feedback mode the unit is the sequence (frames are
  chained through prev_output, trainer/default_trainer.py:136,221-222).  Rank r takes units
  r, r+G, r+2G, ... exactly like the reference's scene round-robin
  (kittiloader/batch_scheduler.py:345-352).  There is no data-path collective.

* depth planes (large D) -- rank r owns planes [lo, hi) of a cost volume whose D planes do not
  fit or are too slow on one GPU.  Not in the reference.  The soft-max over D then needs one
  exchange: per-pixel max (all-reduce MAX), per-pixel sums (all-reduce SUM) and, for the variance,
  the central second moment (all-reduce SUM); the MAP bin is merged first-maximum-wins from the
  gathered (value, index) candidates.  Payload: 4-8 bytes per pixel per collective, i.e. latency
  bound on NVLink; the collectives are NCCL calls enqueued on the kernels' stream.

The local passes are the dpv_shard_* kernels of libdpv_sm100a.so.  `PlaneShardedHead` takes the
object that provides them as `local` so that the exchange protocol can be exercised by the CPU
test-suite over gloo with a stand-in (tests/test_sharding_gloo.py); the default is the CUDA
library and there is no other implementation in the package.
"""
import numpy as np
import torch

from . import _lib, ops


# ------------------------------------------------------------------ items / sequences
def shard_units(n_units, rank, world):
    """Indices of the units (items, sequences) rank `rank` of `world` processes owns."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return list(range(rank, n_units, world))


def shard_slice(t, rank, world, dim=0):
    """The r::G slice of a tensor / array along `dim` (a view, no copy)."""
    idx = [slice(None)] * t.ndim
    idx[dim] = slice(rank, None, world)
    return t[tuple(idx)]


def merge_units(parts, n_units):
    """Inverse of shard_units for per-rank result lists: parts[r][i] is unit r + i*G."""
    world = len(parts)
    out = [None] * n_units
    for r, p in enumerate(parts):
        for i, v in enumerate(p):
            out[r + i * world] = v
    if any(v is None for v in out):
        raise ValueError("merge_units: missing units")
    return out


def plane_range(D, rank, world):
    """[lo, hi) of the depth planes rank owns: contiguous, sizes differ by at most one."""
    base, extra = divmod(D, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value, group=None, device=None):
    """Max of a host scalar over the ranks (device timings are reported as the slowest rank)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


# ------------------------------------------------------------------ depth planes
class CudaShardKernels:
    """The dpv_shard_* entry points (include/dpv_b200.h) on the current torch stream."""

    def __init__(self):
        self.lib = _lib.load()

    @staticmethod
    def _st():
        return torch.cuda.current_stream().cuda_stream

    def local_max(self, x, lo, want_argmax):
        ops._need(x, "x")
        B, Dl, HW = x.shape
        m = torch.empty((B, HW), device=x.device, dtype=torch.float32)
        am = torch.empty((B, HW), device=x.device, dtype=torch.float32) if want_argmax else None
        _lib.check(self.lib.dpv_shard_max(x.data_ptr(), m.data_ptr(), ops._p(am), B, Dl, HW, int(lo),
                                          self._st()))
        return m, am

    def local_sums(self, x, d_local, gmax):
        B, Dl, HW = x.shape
        s = torch.empty((2, B, HW), device=x.device, dtype=torch.float32)
        _lib.check(self.lib.dpv_shard_sums(x.data_ptr(), d_local.data_ptr(), gmax.data_ptr(),
                                           s.data_ptr(), B, Dl, HW, self._st()))
        return s

    def local_central(self, x, d_local, gmax, gsums):
        B, Dl, HW = x.shape
        c = torch.empty((B, HW), device=x.device, dtype=torch.float32)
        _lib.check(self.lib.dpv_shard_central(x.data_ptr(), d_local.data_ptr(), gmax.data_ptr(),
                                              gsums.data_ptr(), c.data_ptr(), B, Dl, HW, self._st()))
        return c

    def finish(self, x, gmax, gsums, gcentral, want_logp, want_depth):
        B, Dl, HW = x.shape
        logp = torch.empty_like(x) if want_logp else None
        depth = torch.empty((B, HW), device=x.device, dtype=torch.float32) if want_depth else None
        var = torch.empty((B, HW), device=x.device, dtype=torch.float32) if gcentral is not None else None
        _lib.check(self.lib.dpv_shard_finish(x.data_ptr(), gmax.data_ptr(), gsums.data_ptr(),
                                             ops._p(gcentral), ops._p(logp), ops._p(depth), ops._p(var),
                                             B, Dl, HW, self._st()))
        return logp, depth, var

    def argmax_merge(self, vals, idx):
        G, n = vals.shape[0], vals[0].numel()
        out = torch.empty((n,), device=vals.device, dtype=torch.int64)
        _lib.check(self.lib.dpv_shard_argmax_merge(vals.data_ptr(), idx.data_ptr(), out.data_ptr(),
                                                   G, n, self._st()))
        return out


class PlaneShardedHead:
    """log-softmax / E[d] / Var / arg-max over depth planes that live on several ranks.

    x_local [B, D_local, H, W] holds planes [lo, hi) of the global volume (plane_range);
    d_candi is the GLOBAL bin vector.  Returns a dict with `logp` for the local planes and the
    per-pixel `depth`, `variance`, `argmax` replicated on every rank.
    """

    def __init__(self, D, rank=None, world=None, group=None, local=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        on = dist.is_available() and dist.is_initialized()
        self.world = world if world is not None else (dist.get_world_size(group) if on else 1)
        self.rank = rank if rank is not None else (dist.get_rank(group) if on else 0)
        self.D = int(D)
        self.lo, self.hi = plane_range(self.D, self.rank, self.world)
        self.local = local if local is not None else CudaShardKernels()

    def _local_bins(self, d_candi, device):
        """This rank's slice of the (global) bin depths on the device; uploaded once per bin vector."""
        key = (id(d_candi) if not isinstance(d_candi, np.ndarray) else d_candi.ctypes.data, len(d_candi), str(device))
        hit = getattr(self, "_bins_cache", None)
        d_all = np.asarray(d_candi, dtype=np.float64).astype(np.float32)
        if hit is not None and hit[0] == key and np.array_equal(hit[1], d_all):
            return hit[2]
        t = torch.from_numpy(np.ascontiguousarray(d_all[self.lo:self.hi])).to(device)
        self._bins_cache = (key, d_all, t)
        return t

    def _all_reduce(self, t, op):
        if self.world > 1:
            self.dist.all_reduce(t, op=op, group=self.group)
        return t

    def __call__(self, x_local, d_candi, variance=True, argmax=True, logp=True):
        dist = self.dist
        B, Dl, H, W = x_local.shape
        if Dl != self.hi - self.lo:
            raise ValueError("rank %d owns %d planes, got %d" % (self.rank, self.hi - self.lo, Dl))
        x = x_local.contiguous().reshape(B, Dl, H * W)
        d_local = self._local_bins(d_candi, x.device)
        m, am = self.local.local_max(x, self.lo, argmax)
        out = {}
        if argmax:
            if self.world > 1:
                # one all-gather of (max, arg-max) per rank serves both the arg-max merge and the global
                # maximum (no separate MAX all-reduce)
                cand = torch.stack([m, am]).contiguous()
                gathered = [torch.empty_like(cand) for _ in range(self.world)]
                dist.all_gather(gathered, cand, group=self.group)
                vals = torch.stack([g[0] for g in gathered]).contiguous()
                idx = torch.stack([g[1] for g in gathered]).contiguous()
                gmax = vals.max(dim=0).values.contiguous()
            else:
                vals, idx = m.unsqueeze(0).contiguous(), am.unsqueeze(0).contiguous()
                gmax = m
            out["argmax"] = self.local.argmax_merge(vals, idx).reshape(B, H, W)
        else:
            gmax = self._all_reduce(m.clone(), dist.ReduceOp.MAX)
        gsums = self._all_reduce(self.local.local_sums(x, d_local, gmax), dist.ReduceOp.SUM)
        gcentral = None
        if variance:
            gcentral = self._all_reduce(self.local.local_central(x, d_local, gmax, gsums),
                                        dist.ReduceOp.SUM)
        lp, depth, var = self.local.finish(x, gmax, gsums, gcentral, logp, True)
        if logp:
            out["logp"] = lp.reshape(B, Dl, H, W)
        out["depth"] = depth.reshape(B, H, W)
        if variance:
            out["variance"] = var.reshape(B, H, W)
        return out


def plane_sharded_sweep(ref, src, poses, K, rays, d_candi, sigma, rank, world, dist="L2"):
    """Cost volume of the planes this rank owns: features replicated, planes split.

    Returns (cost_local [B, D_local, H, W], (lo, hi)).  Feed cost_local to PlaneShardedHead when
    the soft-max follows the cost volume directly (models/packnet.py:394).
    """
    d_all = np.asarray(d_candi, dtype=np.float64)
    lo, hi = plane_range(len(d_all), rank, world)
    cost = ops.sweep_cost_volume(ref, src, poses, K, rays, d_all[lo:hi], sigma, dist=dist)
    return cost, (lo, hi)
