"""Multi-GPU sharding of the DPV path: one process per GPU, `torch.distributed` for plumbing.

Two partitionings:

* items / sequences -- every batch item is independent (the reference loops over items,
  models/models.py:528-550,614-627) and in feedback mode the unit is the sequence (frames are
  chained through prev_output, trainer/default_trainer.py:136,221-222).  Rank r takes units
  r, r+G, r+2G, ... exactly like the reference's scene round-robin
  (kittiloader/batch_scheduler.py:345-352).  There is no data-path collective.

* depth planes (large D) -- rank r owns planes [lo, hi) of a cost volume whose D planes do not
  fit or are too slow on one GPU.  Not in the reference.  The soft-max over D then needs ONE
  exchange: every rank reduces its planes to five numbers per pixel (local max, sum of exponentials,
  local mean, local central second moment, first local arg-max: dpv_shard_stats, one pass with the
  planes in registers), the ranks all-gather those 20 bytes per pixel (one NCCL collective over
  NVLink, enqueued on the kernels' stream), and every rank merges the G records with the
  online-soft-max / pairwise-variance rules while it writes its planes of the log-softmax
  (dpv_shard_merge_finish).  Two reads and one write of the local slice, one collective.  From 3 ranks
  on, the all-gather (20 B x pixels x G per rank) would outweigh the local slice, so the exchange takes the
  reduce-scatter shape instead: an all-to-all hands rank g the G records of ITS 1/G of the pixels, it
  merges them (dpv_shard_merge_slice), an all-gather replicates the merged records, dpv_shard_finish
  writes the planes -- 20 B x pixels per rank in each collective, independent of G (world 8, D=256,
  384x1280: 0.15 ms instead of 0.23 ms; one GPU unsharded: 0.35 ms).

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

    NSTAT = 5

    def __init__(self):
        self.lib = _lib.load()

    @staticmethod
    def _st():
        return torch.cuda.current_stream().cuda_stream

    def local_stats(self, x, d_local, lo, out=None, slice_len=0):
        """x [B, D_local, HW] -> local max, sum exp, local mean, local M2, first arg-max per pixel: five planes
        [5, B*HW] (slice_len = 0) or slice-major [G, 5, slice_len] (the all-to-all layout)."""
        ops._need(x, "x")
        B, Dl, HW = x.shape
        st = out if out is not None else torch.empty((self.NSTAT, B * HW), device=x.device, dtype=torch.float32)
        _lib.check(self.lib.dpv_shard_stats(x.data_ptr(), d_local.data_ptr(), st.data_ptr(), B, Dl, HW, int(lo),
                                            int(slice_len), self._st()))
        return st

    def merge_slice(self, recv, out):
        """recv [G, 5, slice] (rank g's statistics of my pixels) -> out [5, slice]: M, log S, mean, Var, arg-max."""
        G, _, sl = recv.shape
        _lib.check(self.lib.dpv_shard_merge_slice(recv.data_ptr(), out.data_ptr(), G, sl, self._st()))
        return out

    def finish(self, x, merged_all, want_logp, want_var, want_argmax):
        """merged_all [G, 5, slice]: the merged record of every pixel -> (logp of the local planes, depth, variance,
        argmax)."""
        B, Dl, HW = x.shape
        sl = merged_all.shape[2]
        e = lambda dt=torch.float32: torch.empty((B, HW), device=x.device, dtype=dt)
        logp = torch.empty_like(x) if want_logp else None
        depth, var, am = e(), (e() if want_var else None), (e(torch.int64) if want_argmax else None)
        _lib.check(self.lib.dpv_shard_finish(x.data_ptr(), merged_all.data_ptr(), ops._p(logp), depth.data_ptr(),
                                             ops._p(var), ops._p(am), B, Dl, HW, sl, self._st()))
        return logp, depth, var, am

    def merge_finish(self, x, gathered, want_logp, want_var, want_argmax):
        """gathered [G, 5, B*HW] (ranks in plane order) -> (logp of the local planes, depth, variance, argmax)."""
        B, Dl, HW = x.shape
        G = gathered.shape[0]
        e = lambda dt=torch.float32: torch.empty((B, HW), device=x.device, dtype=dt)
        logp = torch.empty_like(x) if want_logp else None
        depth, var, am = e(), (e() if want_var else None), (e(torch.int64) if want_argmax else None)
        _lib.check(self.lib.dpv_shard_merge_finish(x.data_ptr(), gathered.data_ptr(), ops._p(logp), depth.data_ptr(),
                                                   ops._p(var), ops._p(am), G, B, Dl, HW, self._st()))
        return logp, depth, var, am


class PlaneShardedHead:
    """log-softmax / E[d] / Var / arg-max over depth planes that live on several ranks.

    x_local [B, D_local, H, W] holds planes [lo, hi) of the global volume (plane_range);
    d_candi is the GLOBAL bin vector.  Returns a dict with `logp` for the local planes and the
    per-pixel `depth`, `variance`, `argmax` replicated on every rank.
    """

    def __init__(self, D, rank=None, world=None, group=None, local=None, exchange="auto"):
        """exchange: "gather" (one all-gather), "scatter" (all-to-all + all-gather, the reduce-scatter shape) or
        "auto" (gather up to 2 ranks, scatter above)."""
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.exchange = exchange
        on = dist.is_available() and dist.is_initialized()
        self.world = world if world is not None else (dist.get_world_size(group) if on else 1)
        self.rank = rank if rank is not None else (dist.get_rank(group) if on else 0)
        self.D = int(D)
        self.lo, self.hi = plane_range(self.D, self.rank, self.world)
        self.local = local if local is not None else CudaShardKernels()
        self._buf = None        # (key, stats, gathered): reused call after call (no allocation on the path)

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

    def __call__(self, x_local, d_candi, variance=True, argmax=True, logp=True):
        B, Dl, H, W = x_local.shape
        if Dl != self.hi - self.lo:
            raise ValueError("rank %d owns %d planes, got %d" % (self.rank, self.hi - self.lo, Dl))
        x = x_local.contiguous().reshape(B, Dl, H * W)
        d_local = self._local_bins(d_candi, x.device)
        n = B * H * W
        G = self.world
        scatter = self.exchange == "scatter" or (self.exchange == "auto" and G > 2)
        sl = (n + G - 1) // G if scatter else 0
        key = (n, str(x.device), scatter)
        if self._buf is None or self._buf[0] != key:
            mk = lambda *shape: torch.empty(shape, device=x.device, dtype=torch.float32)
            self._buf = ((key, mk(G, 5, sl), mk(G, 5, sl), mk(5, sl), mk(G, 5, sl)) if scatter
                         else (key, mk(5, n), mk(G, 5, n)))
        if scatter:
            _, send, recv, merged, merged_all = self._buf
            self.local.local_stats(x, d_local, self.lo, out=send, slice_len=sl)
            if G > 1:
                self.dist.all_to_all_single(recv.view(G * 5 * sl), send.view(G * 5 * sl), group=self.group)
            else:
                recv = send
            self.local.merge_slice(recv, merged)
            if G > 1:
                self.dist.all_gather_into_tensor(merged_all.view(G * 5, sl), merged, group=self.group)
            else:
                merged_all = merged.unsqueeze(0)
            lp, depth, var, am = self.local.finish(x, merged_all, logp, variance, argmax)
        else:
            _, stats, gathered = self._buf
            stats = self.local.local_stats(x, d_local, self.lo, out=stats)
            if G > 1:
                # the one exchange (output = the ranks' records concatenated along dim 0, in rank = plane order)
                self.dist.all_gather_into_tensor(gathered.view(G * 5, n), stats, group=self.group)
            else:
                gathered = stats.unsqueeze(0)
            lp, depth, var, am = self.local.merge_finish(x, gathered, logp, variance, argmax)
        out = {"depth": depth.reshape(B, H, W)}
        if logp:
            out["logp"] = lp.reshape(B, Dl, H, W)
        if variance:
            out["variance"] = var.reshape(B, H, W)
        if argmax:
            out["argmax"] = am.reshape(B, H, W)
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
