"""Tensor-level entry points over the C ABI (include/dpv_b200.h).

PyTorch is used for device memory and streams only: every function checks its
arguments, allocates the outputs, and enqueues one (or two) of our kernels on
the current torch stream through ctypes.  Nothing here computes on the CPU and
nothing falls back to ATen ops; a missing library or a non-CUDA tensor raises.
"""
import numpy as np
import torch

from . import _lib

DIST = {"L2": 0, "L1": 1}
IN_LOGITS, IN_LOGPROB, IN_PROB = 0, 1, 2
_MODE = {"logits": IN_LOGITS, "logprob": IN_LOGPROB, "prob": IN_PROB}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _need(t, name, dtype=torch.float32):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise _lib.DpvError("%s must live on a CUDA device: the DPV kernels have no CPU path" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if t.requires_grad and torch.is_grad_enabled():
        # forward-only kernels: returning a detached result here would silently zero the gradients
        # of whatever loss sits downstream (the reference's losses/losses.py:82-88 calls
        # dpv_to_depthmap under grad)
        raise _lib.DpvError("%s requires grad: the DPV kernels are forward-only (eval path); call them "
                            "under torch.no_grad() or detach the input" % name)
    return t


def _inner_contiguous(t, n_inner):
    """True when the last n_inner dims are laid out contiguously."""
    stride = 1
    for size, st in zip(reversed(t.shape[-n_inner:]), reversed(t.stride()[-n_inner:])):
        if size != 1 and st != stride:
            return False
        stride *= size
    return True


def _per_item(t, B, shape, name):
    """Per-item constant given as [B,*shape] or broadcast as [*shape] / [1,*shape]."""
    t = t.reshape((-1,) + tuple(shape)).contiguous()
    if t.shape[0] == 1:
        return t, 0
    if t.shape[0] != B:
        raise ValueError("%s must have 1 or %d items, got %d" % (name, B, t.shape[0]))
    return t, int(np.prod(shape))


_dcache = {}      # (bin bytes, device) -> fp32 device tensor, owned here, never modified
_dsum = {}        # id(tensor owned by _dcache) -> fp32 sum of its bins
_DCACHE_MAX = 64


def _bin_sum(d):
    """sum_k d_k in fp32: E[d] of an all-zero log-DPV column (exp(0) = 1 per bin).

    Cached only for the tensors depth_bins() owns (alive and immutable for the life of the cache, so
    id() identifies them).  A caller-supplied tensor may be freed, re-allocated at the same address or
    changed in place, so its sum is taken from its content on every call (one 4*D-byte read-back)."""
    v = _dsum.get(id(d))
    if v is not None and _dcache.get(getattr(d, "_dpv_key", None)) is d:
        return v
    return float(np.sum(d.detach().cpu().numpy(), dtype=np.float32))


def depth_bins(d_candi, device):
    """fp32 device copy of the depth bins (the reference re-uploads them every call,
    warping/homography.py:115); cached per content and device."""
    if isinstance(d_candi, torch.Tensor):
        return _need(d_candi, "d_candi").contiguous()
    arr = np.ascontiguousarray(np.asarray(d_candi, dtype=np.float64))
    key = (arr.tobytes(), str(device))
    t = _dcache.get(key)
    if t is None:
        if len(_dcache) >= _DCACHE_MAX:      # bounded: bin sets are a handful per process in practice
            _dcache.clear()
            _dsum.clear()
        t = torch.from_numpy(arr.astype(np.float32)).to(device)
        t._dpv_key = key
        _dcache[key] = t
        _dsum[id(t)] = float(np.sum(arr.astype(np.float32), dtype=np.float32))
    return t


# ----------------------------------------------------------------------------- K1 + K2a
def sweep_cost_volume(ref, src, poses, K, rays, d_candi, sigma, dist="L2", algo=0,
                      log_softmax=False):
    """Plane-sweep cost volume, batched (reference warping/homography.py:98-135).

    ref [B,C,H,W]; src [B,V,C,H,W]; poses [B,V,4,4] ([R|t] of each source view);
    K [B,3,3] or [3,3]; rays [B,3,H*W] or [3,H*W].  The batch (and view) dimension
    may be strided views of larger tensors; the inner [C,H,W] must be contiguous.
    Returns cost [B,D,H,W] (and log_softmax(cost) when log_softmax=True).
    """
    _need(ref, "ref"), _need(src, "src"), _need(poses, "poses"), _need(K, "K"), _need(rays, "rays")
    B, C, H, W = ref.shape
    V = src.shape[1]
    if src.shape != (B, V, C, H, W):
        raise ValueError("src must be [B,V,C,H,W] matching ref")
    if not _inner_contiguous(ref, 3):
        ref = ref.contiguous()
    if not _inner_contiguous(src, 3):
        src = src.contiguous()
    if poses.shape != (B, V, 4, 4):
        raise ValueError("poses must be [B,V,4,4]")
    if not _inner_contiguous(poses, 2) or (V > 1 and poses.stride(1) != 16):
        poses = poses.contiguous()
    K, k_bs = _per_item(K, B, (3, 3), "K")
    rays, r_bs = _per_item(rays, B, (3, H * W), "rays")
    d = depth_bins(d_candi, ref.device)
    D = d.numel()
    cost = torch.empty((B, D, H, W), device=ref.device, dtype=torch.float32)
    lsm = torch.empty_like(cost) if log_softmax else None
    lib = _lib.load()
    # scratch for the cross-correlation form (algo 5, what "choose" takes for L2): source-only product maps
    ws = None
    if algo in (0, 5) and dist == "L2":
        ws = torch.empty((int(lib.dpv_sweep_workspace_floats(B, V, H, W)),), device=ref.device, dtype=torch.float32)
    _lib.check(lib.dpv_sweep_cost_volume_ws(
        _p(ref), _p(src), _p(poses), _p(K), _p(rays), _p(d), _p(cost), _p(lsm),
        B, V, C, D, H, W,
        ref.stride(0) if B > 1 else 0, src.stride(0) if B > 1 else 0, src.stride(1) if V > 1 else 0,
        poses.stride(0) if B > 1 else 0, k_bs, r_bs,
        float(sigma), DIST[dist], int(algo), _p(ws), _stream()))
    return (cost, lsm) if log_softmax else cost


def warp_planes(img, d, term1, term2, cx, cy, H, W):
    """Per-plane bilinear warp (reference warping/homography.py:170-198).

    img [N,C,H,W] (or [1,C,H,W] broadcast over the planes); d [N]; term1 [3,1];
    term2 [3,H*W] -> [N,C,H,W].
    """
    _need(img, "img"), _need(d, "d"), _need(term1, "term1"), _need(term2, "term2")
    N = d.numel()
    C = img.shape[1]
    if img.shape[0] not in (1, N) or img.shape[2:] != (H, W):
        raise ValueError("img must be [N or 1, C, H, W]")
    if not _inner_contiguous(img, 3):
        img = img.contiguous()
    nstride = 0 if img.shape[0] == 1 else img.stride(0)
    term1 = term1.contiguous()
    term2 = term2.contiguous()
    out = torch.empty((N, C, H, W), device=img.device, dtype=torch.float32)
    _lib.check(_lib.load().dpv_warp_planes(_p(img), _p(d.contiguous()), _p(term1), _p(term2), _p(out),
                                           N, C, H, W, nstride, float(cx), float(cy), _stream()))
    return out


def warp_feature(feat, poses, K, rays, d_candi):
    """Diagonal feature warp (reference warping/homography.py:137-168), batched.

    feat [B,V,D,H,W]; poses [B,V,4,4] -> [B,V,D,H,W].
    """
    _need(feat, "feat"), _need(poses, "poses"), _need(K, "K"), _need(rays, "rays")
    B, V, D, H, W = feat.shape
    d = depth_bins(d_candi, feat.device)
    if d.numel() != D:
        raise ValueError("warp_feature needs as many channels as depth planes (%d vs %d)" % (D, d.numel()))
    feat = feat.contiguous()
    poses = poses.contiguous()
    K, k_bs = _per_item(K, B, (3, 3), "K")
    rays, r_bs = _per_item(rays, B, (3, H * W), "rays")
    out = torch.empty_like(feat)
    _lib.check(_lib.load().dpv_warp_feature(_p(feat), _p(poses), _p(K), _p(rays), _p(d), _p(out),
                                            B, V, D, H, W, V * 16, k_bs, r_bs, _stream()))
    return out


# ----------------------------------------------------------------------------- K3 / K4b
def head(x, d_candi, addend=None, mode="logits", logp=True, prob=False, depth=False,
         variance=False, argmax=False, quarter=False):
    """Depth-bin head over x [B,D,H,W]; returns a dict with the requested outputs."""
    _need(x, "x")
    if x.dim() != 4:
        raise ValueError("x must be [B,D,H,W]")
    x = x.contiguous()
    B, D, H, W = x.shape
    if addend is not None:
        _need(addend, "addend")
        if mode != "logits":
            raise ValueError("addend is only defined for mode='logits' (log_softmax(x + addend), "
                             "models/models.py:694)")
        if addend.shape != x.shape:
            raise ValueError("addend must match x")
        addend = addend.contiguous()
    d = depth_bins(d_candi, x.device)
    if d.numel() != D:
        raise ValueError("d_candi has %d bins, x has %d" % (d.numel(), D))
    dev = x.device
    out = {}
    if logp:
        out["logp"] = torch.empty_like(x)
    if prob:
        out["prob"] = torch.empty_like(x)
    if depth:
        out["depth"] = torch.empty((B, H, W), device=dev, dtype=torch.float32)
    if variance:
        out["variance"] = torch.empty((B, H, W), device=dev, dtype=torch.float32)
    if argmax:
        out["argmax"] = torch.empty((B, H, W), device=dev, dtype=torch.int64)
    if quarter:
        out["quarter"] = torch.empty((B, D, H // 4, W // 4), device=dev, dtype=torch.float32)
    _lib.check(_lib.load().dpv_head(
        _p(x), _p(addend), _p(d), _p(out.get("logp")), _p(out.get("prob")), _p(out.get("depth")),
        _p(out.get("variance")), _p(out.get("argmax")), _p(out.get("quarter")),
        B, D, H, W, _MODE[mode], _stream()))
    return out


def log_softmax(x):
    """F.log_softmax(x, dim=1) over the bins of x [B,D,H,W] (reference models/models.py:351,560,637)."""
    _need(x, "x")
    D = x.shape[1]
    return head(x, np.arange(D, dtype=np.float64), logp=True)["logp"]     # bin depths unused for logp alone


# ----------------------------------------------------------------------------- K4c
def two_sigma_sq(var):
    """2 * pow(sqrt(var), 2) in fp32, as utils/img_utils.py:24-25,40 evaluates it."""
    s = np.sqrt(np.float32(var)).astype(np.float32)
    return float(np.float32(2.0) * (s * s).astype(np.float32))


def lidar_prior(dmaps, masks, d_candi, var=0.3):
    """Prior DPV from sparse depth (reference utils/img_utils.py:360-375).
    dmaps [B,H,W]; masks [B,1,H,W] -> [B,D,H,W]."""
    _need(dmaps, "dmaps"), _need(masks, "masks")
    B, H, W = dmaps.shape
    d = depth_bins(d_candi, dmaps.device)
    D = d.numel()
    dmaps = dmaps.contiguous()
    masks = masks.reshape(B, H, W).contiguous()
    prior = torch.empty((B, D, H, W), device=dmaps.device, dtype=torch.float32)
    _lib.check(_lib.load().dpv_lidar_prior(_p(dmaps), _p(masks), _p(d), _p(prior), B, D, H, W,
                                           two_sigma_sq(var), _stream()))
    return prior


def bayes_fuse(bv, d_candi, prior=None, dmaps=None, masks=None, var=0.3, want_fused=True,
               want_log=True):
    """Multiply-and-renormalise fusion (reference models/models.py:669-672).
    Either `prior` [B,D,H,W] or (`dmaps`, `masks`) must be given.  Returns (fused, log_fused)."""
    _need(bv, "bv")
    bv = bv.contiguous()
    B, D, H, W = bv.shape
    d = depth_bins(d_candi, bv.device)
    if prior is not None:
        prior = _need(prior, "prior").contiguous()
    else:
        dmaps = _need(dmaps, "dmaps").contiguous()
        masks = _need(masks, "masks").reshape(B, H, W).contiguous()
    fused = torch.empty_like(bv) if want_fused else None
    logf = torch.empty_like(bv) if want_log else None
    _lib.check(_lib.load().dpv_bayes_fuse(_p(bv), _p(prior), _p(dmaps), _p(masks), _p(d), _p(fused),
                                          _p(logf), B, D, H, W, two_sigma_sq(var), _stream()))
    return fused, logf


# ----------------------------------------------------------------------------- K5
_lut_cache = {}


def shift_luts(H, W, pshift, device):
    """Index maps of the reference's +/-pshift nearest-neighbour shifts.

    utils/img_utils.py:170-176,291-302,336 build a normalised grid with step
    2/(size-1) and sample it with grid_sample(nearest, align_corners=False); the
    resulting source indices differ from `y - pshift` at the borders (half-pixel
    ties).  Rather than re-deriving that rounding, run the same grid construction
    once per shape on an index image (host side, cached) and keep the integer maps.
    Returns int32 device tensors (row_fwd [H], row_inv [H], col_fwd [W], col_inv [W]),
    -1 meaning "samples the zero padding".
    """
    key = (H, W, int(pshift), str(device))
    hit = _lut_cache.get(key)
    if hit is not None:
        return hit
    if pshift == 0:
        rows = torch.arange(H, dtype=torch.int32)
        cols = torch.arange(W, dtype=torch.int32)
        luts = (rows, rows.clone(), cols, cols.clone())
    else:
        import torch.nn.functional as F
        yv, xv = torch.meshgrid([torch.arange(0, H).float(), torch.arange(0, W).float()], indexing="ij")
        ystep, xstep = 2.0 / float(H - 1), 2.0 / float(W - 1)
        index_img = (torch.arange(H * W, dtype=torch.float32) + 1.0).reshape(1, 1, H, W)
        luts = []
        maps = []
        for s in (float(pshift), -float(pshift)):
            g = torch.zeros((1, H, W, 2), dtype=torch.float32)
            g[:, :, :, 1] = s
            g[0, :, :, 0] = -1 + xv * xstep - g[0, :, :, 0] * xstep
            g[0, :, :, 1] = -1 + yv * ystep - g[0, :, :, 1] * ystep
            m = F.grid_sample(index_img, g, mode="nearest", align_corners=False)[0, 0].long() - 1
            maps.append(m)
        rl, cl = [], []
        for m in maps:
            valid = m >= 0
            sy = torch.where(valid, m // W, torch.full_like(m, -1)).max(dim=1).values
            sx = torch.where(valid, m % W, torch.full_like(m, -1)).max(dim=0).values
            rebuilt = torch.where((sy[:, None] >= 0) & (sx[None, :] >= 0), sy[:, None] * W + sx[None, :],
                                  torch.full_like(m, -1))
            if not torch.equal(rebuilt, m):
                raise _lib.DpvError("shift map is not separable; cannot build row/column tables")
            rl.append(sy.int())
            cl.append(sx.int())
        luts = (rl[0], rl[1], cl[0], cl[1])
    luts = tuple(t.contiguous().to(device) for t in luts)
    _lut_cache[key] = luts
    return luts


KITTI_UF = dict(pshift=5, zstart=0.6, zend=0.6 + 0.3, maxd=100.0, mind=0.0)


def ufield(dpv, d_candi, intr_up, mode="logprob", mask=None, depth=None, params=None):
    """Uncertainty-field collapse, batched (reference utils/img_utils.py:268-358).

    dpv [B,D,H,W] in `mode` ("logprob" or "prob"); intr_up [B,3,3] or [3,3]; mask
    [B,H,W] optional; depth [B,H,W] optional precomputed E[d].  params: pshift, zstart, zend,
    maxd, mind (default: the KITTI constants) and optionally quash_range > 0 for the
    quash_limit branch (:325-332; the reference's range is 1.0).
    Returns (uf [B,D,W], depth_zero [B,H,W]).
    """
    _need(dpv, "dpv")
    dpv = dpv.contiguous()
    B, D, H, W = dpv.shape
    p = dict(KITTI_UF if params is None else params)
    d = depth_bins(d_candi, dpv.device)
    if depth is None:
        depth = head(dpv, d, mode=mode, logp=False, depth=True)["depth"]
    depth = _need(depth, "depth").contiguous()
    intr_up, i_bs = _per_item(_need(intr_up, "intr_up"), B, (3, 3), "intr_up")
    if mask is not None:
        mask = _need(mask, "mask").reshape(B, H, W).contiguous()
    rf, ri, cf, ci = shift_luts(H, W, p["pshift"], dpv.device)
    lib = _lib.load()
    ws = torch.empty((int(lib.dpv_ufield_workspace_floats(B, D, H, W)),), device=dpv.device,
                     dtype=torch.float32)
    uf = torch.empty((B, D, W), device=dpv.device, dtype=torch.float32)
    dz = torch.empty((B, H, W), device=dpv.device, dtype=torch.float32)
    pad_depth = _bin_sum(d) if mode == "logprob" else 0.0
    f32 = lambda v: float(np.float32(v))
    _lib.check(lib.dpv_ufield(_p(dpv), _p(depth), _p(d), _p(intr_up), _p(mask), _p(rf), _p(ri),
                              _p(cf), _p(ci), _p(uf), _p(dz), _p(ws), B, D, H, W, i_bs, _MODE[mode],
                              f32(p["zstart"]), f32(p["zend"]), f32(p["maxd"]), f32(p["mind"]),
                              pad_depth, f32(p.get("quash_range", 0.0)), _stream()))
    return uf, dz


_tab_cache = {}


def uf_fused_tables(H, W, pshift, device):
    """Device tables of the fused head + UF kernel (dpv_uf_fused_tables) for one image shape, or
    None when the reference's shifts do not compose to "same pixel or padding" for it."""
    key = (H, W, int(pshift), str(device))
    if key in _tab_cache:
        return _tab_cache[key]
    luts = [t.cpu().numpy().astype(np.int32).copy() for t in shift_luts(H, W, pshift, "cpu")]
    row_tab = np.zeros((H, 4), dtype=np.int32)
    col_tab = np.zeros((W,), dtype=np.int32)
    rc = _lib.load().dpv_uf_fused_tables(luts[0].ctypes.data, luts[1].ctypes.data, luts[2].ctypes.data,
                                         luts[3].ctypes.data, H, W, row_tab.ctypes.data,
                                         col_tab.ctypes.data)
    if rc == -2:
        tabs = None
    else:
        _lib.check(rc)
        tabs = (torch.from_numpy(row_tab).to(device), torch.from_numpy(col_tab).to(device))
    _tab_cache[key] = tabs
    return tabs


def head_ufield(x, d_candi, intr_up, mode="logits", logp=True, depth=True, variance=False,
                argmax=False, quarter=False, params=None):
    """Depth-bin head and uncertainty field in ONE pass over x [B,D,H,W] (dpv_head_ufield).

    What the reference does back to back on the refined DPV (models/models.py:351 ->
    trainer/default_trainer.py:221-244,333-336 -> utils/img_utils.py:268-358).  `mode` is
    "logits" or "logprob"; no ground-truth mask (use `ufield` for that).  Falls back to
    head() + ufield() -- still our kernels -- for shapes the fused kernel does not take.
    Returns the dict of head() plus "uf" [B,D,W] and "depth_zero" [B,H,W].
    """
    _need(x, "x")
    x = x.contiguous()
    B, D, H, W = x.shape
    p = dict(KITTI_UF if params is None else params)
    d = depth_bins(d_candi, x.device)
    if d.numel() != D:
        raise ValueError("d_candi has %d bins, x has %d" % (d.numel(), D))
    lib = _lib.load()
    nws = int(lib.dpv_head_ufield_workspace_floats(B, D, H, W))
    quash = float(p.get("quash_range", 0.0)) > 0      # the column gate needs E[d] of the whole column first
    tabs = uf_fused_tables(H, W, p["pshift"], x.device) if (nws > 0 and W % 4 == 0 and not quash) else None
    if tabs is None or mode == "prob":
        out = head(x, d, mode=mode, logp=True, depth=True, variance=variance, argmax=argmax,
                   quarter=quarter)
        out["uf"], out["depth_zero"] = ufield(out["logp"] if mode != "prob" else x, d, intr_up,
                                              mode="prob" if mode == "prob" else "logprob",
                                              depth=out["depth"], params=p)
        return out
    intr_up, i_bs = _per_item(_need(intr_up, "intr_up"), B, (3, 3), "intr_up")
    dev = x.device
    e = lambda shape, dt=torch.float32: torch.empty(shape, device=dev, dtype=dt)
    out = {}
    if logp:
        out["logp"] = torch.empty_like(x)
    if depth:
        out["depth"] = e((B, H, W))
    if variance:
        out["variance"] = e((B, H, W))
    if argmax:
        out["argmax"] = e((B, H, W), torch.int64)
    if quarter:
        out["quarter"] = e((B, D, H // 4, W // 4))
    out["uf"] = e((B, D, W))
    out["depth_zero"] = e((B, H, W))
    ws = e((nws,))
    pad_depth = _bin_sum(d) if mode != "prob" else 0.0
    f32 = lambda v: float(np.float32(v))
    _lib.check(lib.dpv_head_ufield(
        _p(x), _p(d), _p(out.get("logp")), _p(out.get("depth")), _p(out.get("variance")),
        _p(out.get("argmax")), _p(out.get("quarter")), _p(intr_up), _p(tabs[0]), _p(tabs[1]),
        _p(out["uf"]), _p(out["depth_zero"]), _p(ws), B, D, H, W, i_bs, _MODE[mode],
        f32(p["zstart"]), f32(p["zend"]), f32(p["maxd"]), f32(p["mind"]), pad_depth, _stream()))
    return out


# ----------------------------------------------------------------------------- 8f rank 2: D -> D convolutions
class CostRefine:
    """conv0 -> LeakyReLU -> conv0_1 -> LeakyReLU -> conv0_2 -> log_softmax over the depth bins, the block between
    the cost volume and the 1/4-res BV (reference models/models.py:456-460,555-560), on the tcgen05 tensor cores
    at fp32 parity (TF32 x 3).  Built once per model from the three (weight [64,64,3,3], bias [64]) pairs: the
    weights are packed on the device here; `__call__(cost)` takes the cost volume [B,64,h,w] and returns the
    log-DPV (and, with `want_logits`, conv0_2's output before the soft-max)."""

    def __init__(self, weights, biases, slope=0.01):
        lib = _lib.load()
        if len(weights) != 3 or len(biases) != 3:
            raise ValueError("three (weight, bias) pairs: conv0, conv0_1, conv0_2")
        self.slope = float(slope)
        self.w, self.b = [], []
        for w, b in zip(weights, biases):
            w, b = w.detach(), b.detach()      # parameters: packed copies, no graph (the kernels are forward-only)
            _need(w, "weight"), _need(b, "bias")
            if tuple(w.shape) != (64, 64, 3, 3) or tuple(b.shape) != (64,):
                raise ValueError("the tensor-core path is built for D = 64 channels and 3x3 filters")
            hi = torch.empty((9, 64, 64), device=w.device, dtype=torch.float32)
            lo = torch.empty_like(hi)
            _lib.check(lib.dpv_conv3x3_pack_weights(_p(w.detach().contiguous()), _p(hi), _p(lo), 64, 64, _stream()))
            self.w.append((hi, lo))
            self.b.append(b.detach().contiguous().float())
        self._buf = None

    def __call__(self, cost, want_logits=False, want_logp=True, out=None):
        """`out`: optional pre-allocated [B,64,h,w] tensor for the log-DPV (no allocation on the path)."""
        _need(cost, "cost")
        cost = cost.contiguous()
        B, C, H, W = cost.shape
        if C != 64:
            raise ValueError("cost volume must have 64 planes")
        lib = _lib.load()
        n = int(lib.dpv_conv3x3_packed_floats(B, H, W))
        key = (B, H, W, str(cost.device))
        if self._buf is None or self._buf[0] != key:
            self._buf = (key, [torch.empty((n,), device=cost.device, dtype=torch.float32) for _ in range(4)])
        a_hi, a_lo, b_hi, b_lo = self._buf[1]
        st = _stream()
        _lib.check(lib.dpv_conv3x3_pack(_p(cost), _p(a_hi), _p(a_lo), B, C, H, W, st))
        _lib.check(lib.dpv_conv3x3_d64(_p(a_hi), _p(a_lo), _p(self.w[0][0]), _p(self.w[0][1]), _p(self.b[0]),
                                       _p(b_hi), _p(b_lo), None, B, H, W, 1, self.slope, st))
        _lib.check(lib.dpv_conv3x3_d64(_p(b_hi), _p(b_lo), _p(self.w[1][0]), _p(self.w[1][1]), _p(self.b[1]),
                                       _p(a_hi), _p(a_lo), None, B, H, W, 1, self.slope, st))
        logits = None
        if not want_logp:
            out = None
        if want_logits:
            logits = torch.empty_like(cost)
            _lib.check(lib.dpv_conv3x3_d64(_p(a_hi), _p(a_lo), _p(self.w[2][0]), _p(self.w[2][1]), _p(self.b[2]),
                                           None, None, _p(logits), B, H, W, 0, 0.0, st))
        if want_logp:
            if out is None:
                out = torch.empty_like(cost)
            else:
                _need(out, "out")
                if tuple(out.shape) != tuple(cost.shape) or not out.is_contiguous() or out.device != cost.device:
                    raise ValueError("out must be a contiguous tensor of the cost volume's shape on its device")
            _lib.check(lib.dpv_conv3x3_d64(_p(a_hi), _p(a_lo), _p(self.w[2][0]), _p(self.w[2][1]), _p(self.b[2]),
                                           None, None, _p(out), B, H, W, 2, 0.0, st))
        if want_logits and want_logp:
            return out, logits
        return logits if want_logits else out


class Base3DConvs:
    """Base3D, the 3-D convolution stack of the feedback mode (reference models/models.py:376-438, applied at :693):
    dres0 (two Conv3d + BatchNorm3d + ReLU), `dres_count` residual blocks (Conv3d + BN + ReLU, Conv3d + BN, + input),
    classify (Conv3d + BN + ReLU, Conv3d(32 -> 1)) over [B, C<=32, D, h, w], feature_dim = 32, on the tcgen05 tensor
    cores at fp32 parity (TF32 x 3).  A BatchNorm in eval with running statistics is folded into its convolution; one
    that uses batch statistics (training mode -- the reference's unregistered dres_modules never leave it -- or
    track_running_stats = False) is computed from sums the convolution's epilogue accumulates (running statistics
    are NOT updated by this path).  Built once per model with `from_module`; `__call__(volume)` returns the
    residual [B, D, h, w] (`prob=False` in the reference).  The packed activations and the statistics live in buffers
    the object owns (re-used call after call, sized by the first call's shape): one call in flight per object."""

    def __init__(self, layers):
        """layers: list of dicts {weight [Co,Ci,3,3,3], bn: None | dict(gamma, beta, mean, var, eps, batch_stats),
        relu: bool, block: None | "in" | "out"} -- "in" marks the first layer of a residual block (its input is the
        skip), "out" the layer whose result the skip is added to."""
        lib = _lib.load()
        import os
        self.layers = []
        for li, L in enumerate(layers):
            w = L["weight"].detach()
            _need(w, "weight")
            co, ci = int(w.shape[0]), int(w.shape[1])
            if tuple(w.shape[2:]) != (3, 3, 3) or co > 32 or ci > 32:
                raise ValueError("the tensor-core path is built for 3x3x3 filters and at most 32 channels")
            bn = L.get("bn")
            scale = shift = gamma = beta = None
            eps = 0.0
            batch_stats = False
            if bn is not None:
                batch_stats = bool(bn["batch_stats"])
                eps = float(bn["eps"])
                g = bn["gamma"].detach().float() if bn.get("gamma") is not None else torch.ones(co, device=w.device)
                b = bn["beta"].detach().float() if bn.get("beta") is not None else torch.zeros(co, device=w.device)
                if batch_stats:
                    gamma, beta = _pad32(g, 1.0), _pad32(b, 0.0)
                else:
                    sc = g.double() / torch.sqrt(bn["var"].detach().double() + eps)
                    scale = sc.float().contiguous()
                    shift = _pad32((b.double() - bn["mean"].detach().double() * sc).float(), 0.0)
            hi = torch.empty((27, 32, 32), device=w.device, dtype=torch.float32)
            lo = torch.empty_like(hi)
            # a first layer with few input channels is z-folded: the three z-planes become channels (a third of the copies)
            zfold = li == 0 and 3 * ci <= 32 and os.environ.get("DPV_BASE3D_ZFOLD", "1") != "0"
            pack_w = lib.dpv_conv3d_pack_weights_zfold if zfold else lib.dpv_conv3d_pack_weights
            _lib.check(pack_w(_p(w.contiguous().float()), _p(scale), _p(hi), _p(lo), co, ci, _stream()))
            # the 32 -> 1 classifier runs on the FP32 pipe with its weights as launch parameters (host copy, once)
            w_host = None
            if co == 1 and bn is None and not L.get("relu") and L.get("block") is None:
                w_host = np.ascontiguousarray(w.float().cpu().numpy().reshape(-1))
            self.layers.append(dict(w=(hi, lo), shift=shift, gamma=gamma, beta=beta, eps=eps, batch_stats=batch_stats,
                                    relu=bool(L.get("relu", False)), block=L.get("block"), co=co, ci=ci, w_host=w_host,
                                    zfold=zfold))
        if self.layers[-1]["co"] != 1:
            raise ValueError("the last layer must be the 32 -> 1 classifier")
        self.last_on_fp32_pipe = os.environ.get("DPV_BASE3D_LAST_TC", "0") != "1"    # (1: the tensor-core kernel, for timing)
        self._buf = None

    @staticmethod
    def spec_from_module(m):
        """The layer list of a Base3D (the reference's or the mirror's: dres0, dres_modules (a plain list), classify)
        in the form __init__ and oracle.dpv_oracle.base3d take.  Pure Python: no device, no library."""
        def bn_of(bn):
            use_batch = bn.training or not bn.track_running_stats
            return dict(gamma=bn.weight, beta=bn.bias, mean=bn.running_mean, var=bn.running_var, eps=bn.eps,
                        batch_stats=use_batch)
        L = [dict(weight=m.dres0[0][0].weight, bn=bn_of(m.dres0[0][1]), relu=True),
             dict(weight=m.dres0[2][0].weight, bn=bn_of(m.dres0[2][1]), relu=True)]
        for blk in m.dres_modules:
            L.append(dict(weight=blk[0][0].weight, bn=bn_of(blk[0][1]), relu=True, block="in"))
            L.append(dict(weight=blk[2][0].weight, bn=bn_of(blk[2][1]), relu=False, block="out"))
        L.append(dict(weight=m.classify[0][0].weight, bn=bn_of(m.classify[0][1]), relu=True))
        L.append(dict(weight=m.classify[2].weight, bn=None, relu=False))
        return L

    @classmethod
    def from_module(cls, m):
        """m: a Base3D (the reference's or the mirror's)."""
        return cls(cls.spec_from_module(m))

    def __call__(self, volume, out=None):
        _need(volume, "volume")
        volume = volume.contiguous()
        B, C, D, H, W = volume.shape
        if C != self.layers[0]["ci"]:
            raise ValueError("volume has %d channels, the first layer takes %d" % (C, self.layers[0]["ci"]))
        lib = _lib.load()
        n = int(lib.dpv_conv3d_packed_floats(B, D, H, W))
        key = (B, D, H, W, str(volume.device))
        if self._buf is None or self._buf[0] != key:
            mk = lambda: torch.empty((n,), device=volume.device, dtype=torch.float32)
            need_raw = any(L["batch_stats"] for L in self.layers)
            self._buf = (key, [(mk(), mk()) for _ in range(3)], mk() if need_raw else None,
                         torch.zeros((len(self.layers), 64), device=volume.device, dtype=torch.float64))
        _, bufs, raw, stats = self._buf
        st = _stream()
        if raw is not None:
            stats.zero_()
        cur = bufs[0]
        pack = lib.dpv_conv3d_pack_zfold if self.layers[0]["zfold"] else lib.dpv_conv3d_pack
        _lib.check(pack(_p(volume), _p(cur[0]), _p(cur[1]), B, C, D, H, W, st))
        if out is None:
            out = torch.empty((B, D, H, W), device=volume.device, dtype=torch.float32)
        else:
            _need(out, "out")
            if tuple(out.shape) != (B, D, H, W) or not out.is_contiguous() or out.device != volume.device:
                raise ValueError("out must be a contiguous [B, D, H, W] tensor on the volume's device")
        skip = None
        for i, L in enumerate(self.layers):
            last = i == len(self.layers) - 1
            if L["block"] == "in":
                skip = cur                         # the block's input: kept until the block's last layer adds it
            res = skip if L["block"] == "out" else None
            dst = (None, None) if last else next(b for b in bufs if b is not cur and b is not skip)
            hi, lo = L["w"]
            zf, cin = (4, 3 * L["ci"]) if L["zfold"] else (0, L["ci"])
            if last and L["w_host"] is not None and self.last_on_fp32_pipe:
                _lib.check(lib.dpv_conv3d_c32_to1(_p(cur[0]), _p(cur[1]), L["w_host"].ctypes.data, _p(out), B, D, H, W,
                                                  L["ci"], st))
            elif L["batch_stats"]:
                _lib.check(lib.dpv_conv3d_c32(_p(cur[0]), _p(cur[1]), _p(hi), _p(lo), None, None, None, None, None, None,
                                              _p(raw), stats[i].data_ptr(), B, D, H, W, zf, cin, st))
                _lib.check(lib.dpv_conv3d_bn_apply(_p(raw), stats[i].data_ptr(), _p(L["gamma"]), _p(L["beta"]), L["eps"],
                                                   _p(res[0]) if res else None, _p(res[1]) if res else None,
                                                   _p(dst[0]), _p(dst[1]), B, D, H, W, 1 if L["relu"] else 0, st))
            else:
                _lib.check(lib.dpv_conv3d_c32(_p(cur[0]), _p(cur[1]), _p(hi), _p(lo), _p(L["shift"]),
                                              _p(res[0]) if res else None, _p(res[1]) if res else None,
                                              _p(dst[0]), _p(dst[1]), _p(out) if last else None, None, None,
                                              B, D, H, W, (1 if L["relu"] else 0) | zf, cin, st))
            if L["block"] == "out":
                skip = None
            cur = dst
        return out


def _pad32(t, fill):
    """A per-channel vector padded to the kernels' 32 channels."""
    out = torch.full((32,), float(fill), device=t.device, dtype=torch.float32)
    out[:t.numel()] = t.float()
    return out


# ----------------------------------------------------------------------------- K2b
def correlation(x1, x2, max_displacement=4):
    """Local correlation [B,(2r+1)^2,H,W] (reference models/correlation_native.py:13-23)."""
    _need(x1, "x1"), _need(x2, "x2")
    if x1.shape != x2.shape or x1.dim() != 4:
        raise ValueError("x1, x2 must be [B,C,H,W] of equal shape")
    x1 = x1.contiguous()
    x2 = x2.contiguous()
    B, C, H, W = x1.shape
    n = 2 * int(max_displacement) + 1
    out = torch.empty((B, n * n, H, W), device=x1.device, dtype=torch.float32)
    _lib.check(_lib.load().dpv_correlation(_p(x1), _p(x2), _p(out), B, C, H, W,
                                           int(max_displacement), _stream()))
    return out


# ----------------------------------------------------------------------------- eval metrics
METRIC_NAMES = ["mae", "rmse", "inverse mae", "inverse rmse", "log mae", "log rmse",
                "scale invariant log", "abs relative", "squared relative"]


def depth_errors(predicted, truth, mask=None, clamp_max=None, zero_invalid=True, want_counts=False):
    """The nine KITTI depth metrics per item, on the device (reference utils/img_utils.py:17-22 around
    external/deval_lib/src/evaluate_depth.h:19-119).  predicted, truth [B,H,W] (or [H,W]); optional
    fused preparation of trainer/default_trainer.py:247-254: `mask` multiplies the prediction,
    `clamp_max` clamps the truth from above.  Returns [B,9] (METRIC_NAMES order) and, on request, the
    number of valid pixels per item."""
    _need(predicted, "predicted"), _need(truth, "truth")
    if predicted.dim() == 2:
        predicted, truth = predicted.unsqueeze(0), truth.unsqueeze(0)
        mask = None if mask is None else mask.reshape(1, *mask.shape[-2:])
    if predicted.shape != truth.shape or predicted.dim() != 3:
        raise ValueError("predicted and truth must both be [B,H,W]")
    predicted, truth = predicted.contiguous(), truth.contiguous()
    B, H, W = predicted.shape
    if mask is not None:
        mask = _need(mask, "mask").reshape(B, H, W).contiguous()
    lib = _lib.load()
    out = torch.empty((B, 9), device=predicted.device, dtype=torch.float32)
    counts = torch.empty((B,), device=predicted.device, dtype=torch.int32)
    ws = torch.empty((int(lib.dpv_depth_errors_workspace_doubles(B, H, W)),), device=predicted.device,
                     dtype=torch.float64)
    _lib.check(lib.dpv_depth_errors(_p(predicted), _p(truth), _p(mask),
                                    float(clamp_max) if clamp_max is not None else 0.0,
                                    1 if zero_invalid else 0, _p(out), _p(counts), _p(ws), B, H, W, _stream()))
    return (out, counts) if want_counts else out


def unc_rmse(uf_truth, uf_pred, d_candi):
    """compute_unc_rmse (reference utils/img_utils.py:183-194), batched: uf_* [B,D,W] -> [B]."""
    _need(uf_truth, "uf_truth"), _need(uf_pred, "uf_pred")
    if uf_truth.shape != uf_pred.shape or uf_truth.dim() != 3:
        raise ValueError("uncertainty fields must both be [B,D,W]")
    uf_truth, uf_pred = uf_truth.contiguous(), uf_pred.contiguous()
    B, D, W = uf_truth.shape
    d = depth_bins(d_candi, uf_truth.device)
    if d.numel() != D:
        raise ValueError("d_candi has %d bins, the fields %d" % (d.numel(), D))
    out = torch.empty((B,), device=uf_truth.device, dtype=torch.float32)
    _lib.check(_lib.load().dpv_unc_rmse(_p(uf_truth), _p(uf_pred), _p(d), _p(out), B, D, W, _stream()))
    return out


# ----------------------------------------------------------------------------- LiDAR depth maps
def lidar_depthmap(velo, intr, m_velo2cam, width, height, filtering=2, filterdiff=1.0, pool_scale=4,
                   pool_default=1000.0, want_large=True, want_small=True, want_mask=True):
    """generate_depth (reference external/utils_lib/python/utils_lib.cpp:86-160, upsample = 0) followed
    by minpool (utils/img_utils.py:87-95 via kittiloader/kitti.py:706).  velo [n,4] (x, y, z, 1), intr
    [3,4], m_velo2cam [4,4] on the device.  Returns (dmap [height,width], dmap_small, mask_small), None
    for the ones not requested."""
    _need(velo, "velo"), _need(intr, "intr"), _need(m_velo2cam, "m_velo2cam")
    if velo.dim() != 2 or velo.shape[1] != 4 or intr.shape != (3, 4) or m_velo2cam.shape != (4, 4):
        raise ValueError("velo [n,4], intr [3,4], m_velo2cam [4,4] expected")
    velo, intr, m_velo2cam = velo.contiguous(), intr.contiguous(), m_velo2cam.contiguous()
    dev = velo.device
    zbuf = torch.empty((height, width), device=dev, dtype=torch.int32)
    h2, w2 = height // pool_scale, width // pool_scale
    dmap = torch.empty((height, width), device=dev, dtype=torch.float32) if want_large else None
    small = torch.empty((h2, w2), device=dev, dtype=torch.float32) if want_small else None
    mask = torch.empty((h2, w2), device=dev, dtype=torch.float32) if want_mask else None
    _lib.check(_lib.load().dpv_lidar_depthmap(_p(velo), int(velo.shape[0]), _p(intr), _p(m_velo2cam),
                                              int(width), int(height), int(filtering), float(filterdiff),
                                              int(pool_scale), float(pool_default), _p(zbuf), _p(dmap),
                                              _p(small), _p(mask), _stream()))
    return dmap, small, mask


def minpool(x, scale, default=0):
    """minpool (reference utils/img_utils.py:87-95): [..., H, W] -> [..., H // scale, W // scale]."""
    _need(x, "x")
    x = x.contiguous()
    H, W = x.shape[-2:]
    n = x.numel() // (H * W)
    out = torch.empty(tuple(x.shape[:-2]) + (H // scale, W // scale), device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().dpv_minpool(_p(x), _p(out), n, H, W, int(scale), float(default), _stream()))
    return out
