"""Host-buffer frame pipeline: thin handle over dpv_pipeline_* (include/dpv_b200.h).

This is the end-to-end entry a host application calls with HOST memory (numpy arrays or
pinned torch CPU tensors): copies, kernels and read-back are overlapped inside the library.
"""
import ctypes

import numpy as np
import torch

from . import _lib, ops


def _host_ptr(a):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        if a.is_cuda or not a.is_contiguous():
            raise ValueError("pipeline buffers must be contiguous HOST tensors")
        return a.data_ptr()
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("pipeline buffers must be C-contiguous")
    return a.ctypes.data


class FramePipeline:
    """B frames per call: cost volume -> 1/4-res log-softmax -> full-res head -> UF."""

    def __init__(self, B, V, C, D, h, w, H, W, device=0):
        self.shape = dict(B=B, V=V, C=C, D=D, h=h, w=w, H=H, W=W)
        self.device = int(device)
        self._h = ctypes.c_void_p()
        _lib.check(_lib.load().dpv_pipeline_create(ctypes.byref(self._h), self.device, B, V, C, D,
                                                   h, w, H, W))
        luts = ops.shift_luts(H, W, ops.KITTI_UF["pshift"], "cpu")
        self.luts = [t.numpy().copy() for t in luts]

    def outputs(self, pinned=True):
        """Allocate host result buffers (pinned so the D2H copies are asynchronous)."""
        s = self.shape
        B, D, h, w, H, W = s["B"], s["D"], s["h"], s["w"], s["H"], s["W"]
        mk = lambda shape, dt=torch.float32: torch.empty(shape, dtype=dt, pin_memory=pinned)
        return dict(bv=mk((B, D, h, w)), depth=mk((B, H, W)), variance=mk((B, H, W)),
                    argmax=mk((B, H, W), torch.int64), uf=mk((B, D, W)), depth_zero=mk((B, H, W)),
                    quarter=mk((B, D, H // 4, W // 4)))

    def run(self, feats, poses, K, rays, d_candi, logits_full, intr_up, sigma, out):
        """One batch, synchronously: returns when `out` holds the results (dpv_pipeline_run)."""
        return self._call("dpv_pipeline_run", feats, poses, K, rays, d_candi, logits_full, intr_up, sigma, out)

    def submit(self, feats, poses, K, rays, d_candi, logits_full, intr_up, sigma, out):
        """Enqueue one batch and return (dpv_pipeline_submit).  Up to two batches are in flight: the next one is
        copied in while this one computes.  Inputs and `out` must stay untouched until the matching wait()."""
        return self._call("dpv_pipeline_submit", feats, poses, K, rays, d_candi, logits_full, intr_up, sigma, out)

    def wait(self):
        """Block until the oldest outstanding submission has its results in host memory."""
        _lib.check(_lib.load().dpv_pipeline_wait(self._h))

    def _call(self, entry, feats, poses, K, rays, d_candi, logits_full, intr_up, sigma, out):
        d32 = np.ascontiguousarray(np.asarray(d_candi, dtype=np.float32))
        rf, ri, cf, ci = self.luts
        _lib.check(getattr(_lib.load(), entry)(
            self._h, _host_ptr(feats), _host_ptr(poses), _host_ptr(K), _host_ptr(rays),
            d32.ctypes.data, _host_ptr(logits_full), _host_ptr(intr_up),
            rf.ctypes.data, ri.ctypes.data, cf.ctypes.data, ci.ctypes.data, float(sigma),
            _host_ptr(out.get("bv")), _host_ptr(out.get("depth")), _host_ptr(out.get("variance")),
            _host_ptr(out.get("argmax")), _host_ptr(out.get("uf")), _host_ptr(out.get("depth_zero")),
            _host_ptr(out.get("quarter"))))
        return out

    def last_bytes(self):
        a, b = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(_lib.load().dpv_pipeline_last_bytes(self._h, ctypes.byref(a), ctypes.byref(b)))
        return int(a.value), int(b.value)

    def close(self):
        if self._h:
            _lib.load().dpv_pipeline_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
