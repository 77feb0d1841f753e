"""Device-resident frame step: the hot-path kernels of one batch of frames, back to back.

`FrameStep` pre-allocates every output once and then only enqueues kernels (no allocation, no
host sync, no Python loop over items), which also makes the step capturable in a CUDA graph.
Order follows the reference's eval frame (models/models.py:504-565,644-656 and
trainer/default_trainer.py:221-244,333-336) with the CNN blocks left to cuDNN:

  cost volume  -> log-softmax at 1/4 res  -> [decoder]  -> full-res head  -> uncertainty field
  (K1+K2a)        (K3)                                     (K3: log-DPV, E[d], Var, argmax,
                                                            1/4 hand-off)     (K5)
optionally with the LiDAR fusion (K4c, `upsample`) or the feedback warp + fusion (K4a + K4b,
`feedback`) between the 1/4-res soft-max and the decoder.
"""
import numpy as np
import torch

from . import _lib, ops


class FrameStep:
    def __init__(self, B, V, C, D, h, w, H, W, d_candi, sigma=10.0, mode="default", device=None,
                 fuse_uf=True, fuse_lsm=True, refine=None, base3d=None):
        """refine: optional ops.CostRefine (conv0 -> conv0_1 -> conv0_2 -> log-softmax, models/models.py:555-560):
        the 1/4-res BV is then the refined cost volume instead of the soft-max of the cost volume itself.
        base3d: optional ops.Base3DConvs (feedback mode): the residual is then computed in the step from
        cat(BV, prev_output, warped features) as the model does (models/models.py:692-694) instead of being an input."""
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dev = dev
        self.B, self.V, self.C, self.D, self.h, self.w, self.H, self.W = B, V, C, D, h, w, H, W
        self.mode = mode
        self.sigma = float(sigma)
        self.d = ops.depth_bins(d_candi, dev)
        self.pad_depth = ops._bin_sum(self.d)
        e = lambda *s, dt=torch.float32: torch.empty(s, device=dev, dtype=dt)
        self.cost = e(B, D, h, w)
        self.bv = e(B, D, h, w)
        self.refined = e(B, D, H, W)
        self.depth = e(B, H, W)
        self.var = e(B, H, W)
        self.argmax = e(B, H, W, dt=torch.int64)
        self.quarter = e(B, D, H // 4, W // 4)
        self.uf = e(B, D, W)
        self.dz = e(B, H, W)
        self.lib = _lib.load()
        # 1/4-res log-softmax as an epilogue of the sweep kernel (the cost tile is still in shared
        # memory; models/packnet.py:394 places them back to back) instead of a dpv_head launch.  The
        # TMA-fed kernel that has this epilogue needs 16-byte row strides.
        # The epilogue needs all planes of a pixel in one CTA, i.e. no plane split: only when the batch
        # alone fills the machine (the launcher splits planes below 148 x 8 warps of pixels).
        self.refine = refine
        self.fuse_lsm = bool(fuse_lsm) and refine is None and (w % 4 == 0) and B * h * ((w + 31) // 32) >= 148 * 8
        # K3 + K5 in one pass (fuse_uf, default: the tile kernel of dpv_head_uftile.cu takes gen_ufield's
        # column sums while the probabilities are in registers: 0.094 ms per batch of 8 x 256 x 384) or as
        # dpv_head followed by dpv_ufield's three launches (0.075 + 0.040 ms; also what shapes the fused
        # kernel does not take fall back to)
        self.tabs = ops.uf_fused_tables(H, W, ops.KITTI_UF["pshift"], dev) if fuse_uf else None
        nfused = int(self.lib.dpv_head_ufield_workspace_floats(B, D, H, W)) if self.tabs else 0
        self.fused_uf = nfused > 0
        self.ws = e(nfused if self.fused_uf else int(self.lib.dpv_ufield_workspace_floats(B, D, H, W)))
        self.luts = ops.shift_luts(H, W, ops.KITTI_UF["pshift"], dev)
        if mode == "upsample":
            self.fused = e(B, D, h, w)
            self.logfused = e(B, D, h, w)
        self.base3d = base3d if mode == "feedback" else None
        if mode == "feedback":
            self.warped = e(B, V + 1, D, h, w)
            self.bv_upd = e(B, D, h, w)
            if self.base3d is not None:
                self.comb = e(B, V + 3, D, h, w)
                self.resi = e(B, D, h, w)
        # scratch of the sweep's cross-correlation form (source-only product maps, one pre-pass per call)
        self.sweep_ws = e(int(self.lib.dpv_sweep_workspace_floats(B, V, h, w)))
        self.sweep_algo = 0
        self.two_sig = ops.two_sigma_sq(0.3)
        u = ops.KITTI_UF
        self.uf_params = [float(np.float32(u[k])) for k in ("zstart", "zend", "maxd", "mind")]

    # -- one batch of frames ---------------------------------------------------------------
    def run(self, feats, poses, K, rays, logits_full, intr_up, dmaps=None, masks=None,
            feat_raw=None, bv_resi=None, prev=None, head_hook=None, kernel_hook=None):
        """feats [B,V+1,C,h,w] (reference view last); poses [B,V+1,4,4]; K [B,3,3]; rays
        [B,3,h*w]; logits_full [B,D,H,W] (the decoder's pre-softmax output); intr_up [B,3,3].
        upsample: dmaps [B,h,w], masks [B,1,h,w].  feedback: feat_raw [B,V+1,D,h,w], bv_resi
        [B,D,h,w] (the 3-D conv residual) -- or, with base3d, prev [B,D,h,w] (the previous frame's 1/4-res
        hand-off) from which the residual is computed here.  All contiguous fp32 on this device."""
        B, V, C, D, h, w, H, W = self.B, self.V, self.C, self.D, self.h, self.w, self.H, self.W
        lib, st = self.lib, torch.cuda.current_stream().cuda_stream
        p = lambda t: None if t is None else t.data_ptr()
        hk = kernel_hook if kernel_hook is not None else (lambda name, which: None)
        chw = C * h * w
        fp = feats.data_ptr()
        hk("sweep", 0)
        _lib.check(lib.dpv_sweep_cost_volume_ws(
            fp + 4 * V * chw, fp, p(poses), p(K), p(rays), p(self.d), p(self.cost),
            p(self.bv) if self.fuse_lsm else None,
            B, V, C, D, h, w, (V + 1) * chw, (V + 1) * chw, chw, (V + 1) * 16, 9, 3 * h * w,
            self.sigma, 0, self.sweep_algo, p(self.sweep_ws), st))
        hk("sweep", 1)
        if self.refine is not None:
            hk("cost_refine", 0)
            self.refine(self.cost, out=self.bv)
            hk("cost_refine", 1)
        elif not self.fuse_lsm:
            hk("head_quarter", 0)
            _lib.check(lib.dpv_head(p(self.cost), None, p(self.d), p(self.bv), None, None, None, None,
                                    None, B, D, h, w, ops.IN_LOGITS, st))
            hk("head_quarter", 1)
        if self.mode == "upsample":
            hk("bayes_fuse", 0)
            _lib.check(lib.dpv_bayes_fuse(p(self.bv), None, p(dmaps), p(masks), p(self.d),
                                          p(self.fused), p(self.logfused), B, D, h, w, self.two_sig, st))
            hk("bayes_fuse", 1)
        elif self.mode == "feedback":
            hk("warp_feature", 0)
            _lib.check(lib.dpv_warp_feature(p(feat_raw), p(poses), p(K), p(rays), p(self.d),
                                            p(self.warped), B, V + 1, D, h, w, (V + 1) * 16, 9,
                                            3 * h * w, st))
            hk("warp_feature", 1)
            if self.base3d is not None:
                if prev is None or tuple(prev.shape) != (B, D, h, w):
                    raise ValueError("feedback step with base3d: prev [B,D,h,w] (the previous frame's 1/4-res hand-off) is required")
                hk("base3d", 0)
                self.comb[:, 0].copy_(self.bv)                 # models/models.py:692: cat(BV_cur, prev_output, warped)
                self.comb[:, 1].copy_(prev)
                self.comb[:, 2:].copy_(self.warped)
                bv_resi = self.base3d(self.comb, out=self.resi)
                hk("base3d", 1)
            hk("feedback_fuse", 0)
            _lib.check(lib.dpv_head(p(self.bv), p(bv_resi), p(self.d), p(self.bv_upd), None, None,
                                    None, None, None, B, D, h, w, ops.IN_LOGITS, st))
            hk("feedback_fuse", 1)
        self._full_res(lib, st, p, logits_full, intr_up, head_hook, hk)

    def run_head(self, logits_full, intr_up):
        """Only the full-resolution head (+ UF) of the step: what bench.py replays back to back to time the
        dominant kernel without the event bubbles of an in-step measurement."""
        p = lambda t: None if t is None else t.data_ptr()
        self._full_res(self.lib, torch.cuda.current_stream().cuda_stream, p, logits_full, intr_up, None,
                       lambda name, which: None)

    def _full_res(self, lib, st, p, logits_full, intr_up, head_hook, hk):
        B, D, H, W = self.B, self.D, self.H, self.W
        if head_hook is not None:
            head_hook(0)
        hk("head_full_ufield" if self.fused_uf else "head_full", 0)
        if self.fused_uf:
            _lib.check(lib.dpv_head_ufield(p(logits_full), p(self.d), p(self.refined), p(self.depth),
                                           p(self.var), p(self.argmax), p(self.quarter), p(intr_up),
                                           p(self.tabs[0]), p(self.tabs[1]), p(self.uf), p(self.dz),
                                           p(self.ws), B, D, H, W, 9, ops.IN_LOGITS, *self.uf_params,
                                           self.pad_depth, st))
            hk("head_full_ufield", 1)
            if head_hook is not None:
                head_hook(1)
            return
        _lib.check(lib.dpv_head(p(logits_full), None, p(self.d), p(self.refined), None, p(self.depth),
                                p(self.var), p(self.argmax), p(self.quarter), B, D, H, W,
                                ops.IN_LOGITS, st))
        hk("head_full", 1)
        if head_hook is not None:
            head_hook(1)
        rf, ri, cf, ci = self.luts
        hk("ufield", 0)
        _lib.check(lib.dpv_ufield(p(self.refined), p(self.depth), p(self.d), p(intr_up), None,
                                  p(rf), p(ri), p(cf), p(ci), p(self.uf), p(self.dz), p(self.ws),
                                  B, D, H, W, 9, ops.IN_LOGPROB, *self.uf_params, self.pad_depth, 0.0, st))
        hk("ufield", 1)

    def capture(self, *args, **kw):
        """Record one step on these (fixed) input tensors into a CUDA graph; `graph.replay()` then
        re-runs it with a single launch from the host.  Every buffer the step touches is owned by the
        caller or pre-allocated here, so the graph holds no allocations."""
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):      # warm-up outside the capture (module loading, attributes)
            n0 = _lib.launch_count()
            self.run(*args, **kw)
            self._launches = _lib.launch_count() - n0      # what a replay re-runs
        cur.wait_stream(side)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.run(*args, **kw)
        return g

    def launches_per_step(self):
        """Kernel launches of one step: counted by the library during capture()'s warm-up run when there was
        one, else by the formula below (sweep [+ 1/4-res soft-max], head, UF: weights + partial sums + finish;
        fused head + UF: tile kernel + finish)."""
        if getattr(self, "_launches", None):
            return int(self._launches)
        n = {"default": 6, "upsample": 7, "feedback": 8}[self.mode]
        if self.fuse_lsm:
            n -= 1
        return n - 2 if self.fused_uf else n

    # -- bytes per step --------------------------------------------------------------------
    def survey_bytes(self):
        """SURVEY.md section 8d accounting: every kernel of the reference's decomposition as its own pass
        (K5 re-reads the DPV, the 1/4-res soft-max is a separate read + write).  This is the numerator the
        survey's 60 % target is defined on; it is NOT what this step moves when passes are fused."""
        B, V, C, D = self.B, self.V, self.C, self.D
        hw, HW = self.h * self.w, self.H * self.W
        k = {
            "sweep": 4 * hw * (C * (1 + V) + D) + 12 * hw,
            "head_quarter": 8 * hw * D,
            "head_full": 8 * HW * D + 16 * HW,
            "ufield": 4 * HW * D + 4 * D * self.W + 8 * HW,
        }
        if self.mode == "upsample":
            k["bayes_fuse"] = 12 * hw * D + 8 * hw
        if self.mode == "feedback":
            k["warp_feature"] = 8 * hw * D * (V + 1) + 12 * hw
            k["feedback_fuse"] = 12 * hw * D
        return {n: v * B for n, v in k.items()}

    def algorithmic_bytes(self):
        """Bytes of the LAUNCHED configuration: each kernel's inputs read once and outputs written once.
        A pass that was fused away is not counted: with the 1/4-res log-softmax in the sweep epilogue the
        sweep additionally writes the log-DPV (4 hw D) and `head_quarter` disappears; with the fused head +
        UF the DPV is not re-read, the fused kernel additionally writes UF [D,W] and depth_zero [H,W]."""
        B, V, C, D = self.B, self.V, self.C, self.D
        hw, HW = self.h * self.w, self.H * self.W
        k = {"sweep": 4 * hw * (C * (1 + V) + D) + 12 * hw}
        if self.refine is not None:
            k["cost_refine"] = 8 * hw * D          # cost volume in, log-DPV out (the packed intermediates stay in L2)
        elif self.fuse_lsm:
            k["sweep"] += 4 * hw * D
        else:
            k["head_quarter"] = 8 * hw * D
        if self.fused_uf:
            k["head_full_ufield"] = 8 * HW * D + 16 * HW + 4 * D * self.W + 4 * HW
        else:
            k["head_full"] = 8 * HW * D + 16 * HW
            k["ufield"] = 4 * HW * D + 4 * D * self.W + 8 * HW
        if self.mode == "upsample":
            k["bayes_fuse"] = 12 * hw * D + 8 * hw
        if self.mode == "feedback":
            k["warp_feature"] = 8 * hw * D * (V + 1) + 12 * hw
            k["feedback_fuse"] = 12 * hw * D
            if self.base3d is not None:
                k["base3d"] = 4 * hw * D * (V + 3) + 4 * hw * D     # the volume in, the residual out (a tensor-pipe stage)
        return {n: v * B for n, v in k.items()}

    def sweep_flops(self):
        """Direct-form flops of K1+K2a per step: hw * D * V * (11 C + 40) per frame."""
        return self.B * self.h * self.w * self.D * self.V * (11 * self.C + 40)

    def dominant_kernel(self):
        """(name, algorithmic bytes per launch) of the kernel the roofline is quoted on: the
        full-resolution head.  Fused with the UF it reads the logits once and additionally writes
        UF [B,D,W] and depth_zero [B,H,W]; it does not re-read the DPV, so K5's 4*HW*D is not
        counted for it."""
        B, D, HW = self.B, self.D, self.H * self.W
        head = 8 * HW * D + 16 * HW
        if self.fused_uf:
            return "dpv::head_uf_tile_kernel<64,LOGITS,LOGP> + finish (full-res head + UF, fused)", \
                B * (head + 4 * D * self.W + 4 * HW)
        return "dpv::head_kernel<64,1,LOGITS,...> (full-res head)", B * head
