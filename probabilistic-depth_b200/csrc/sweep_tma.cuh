// Device helpers shared by the TMA-fed plane-sweep kernels (sweep_tma.cu: per-cell Gram form,
// sweep_xcorr.cu: cross-correlation form): mbarrier / bulk-tensor copy wrappers, the coordinate map of
// warping/homography.py:187-196 in its fast and reference-order forms, the 10-term quadratic form of a
// plane in the Gram matrix of its cell, and the global-memory gather path for tiles whose source window
// does not fit in shared memory.
#pragma once
#include <cuda.h>

#include "sweep_common.cuh"

namespace dpv {

constexpr int TM_PX = 32;                        // reference pixels per tile (one row segment)
constexpr int TM_OS = TM_PX;                     // row stride of the result tile (lane = pixel = bank)
constexpr int kTmOutside = -1;                   // cell id of "no tap inside the image"
constexpr int kTmNone = -2;

__device__ __forceinline__ unsigned tm_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tm_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tm_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tm_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tm_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tm_mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n.reg .pred p;\n"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                     "selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(tm_smem(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tm_load_5d(void* dst, const CUtensorMap* map, int x, int y, int c, int v,
                                           int b, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
                 "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(tm_smem(dst)), "l"(map), "r"(x), "r"(y), "r"(c), "r"(v), "r"(b), "r"(tm_smem(bar))
                 : "memory");
}
__device__ __forceinline__ void tm_load_4d(void* dst, const CUtensorMap* map, int x, int y, int c, int b,
                                           unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
                 "[%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(tm_smem(dst)), "l"(map), "r"(x), "r"(y), "r"(c), "r"(b), "r"(tm_smem(bar))
                 : "memory");
}

// in-image cells have x0 in [-1, W-1], y0 in [-1, H-1]: packed ids are non-negative (W, H < 32767)
__device__ __forceinline__ int tm_pack(int x0, int y0) { return ((y0 + 1) << 16) | (x0 + 1); }
__device__ __forceinline__ int tm_cell_x(int p) { return (p & 0xffff) - 1; }
__device__ __forceinline__ int tm_cell_y(int p) { return (p >> 16) - 1; }

// Packed fp32x2 arithmetic (FFMA2 / FADD2 on sm_100a): two runs of a lane share every
// instruction of the Gram update, halving the issue slots and fma-pipe cycles of the hot loop.
typedef unsigned long long tm_f2;
__device__ __forceinline__ tm_f2 tm_pk(float lo, float hi) {
    tm_f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void tm_upk(tm_f2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ tm_f2 tm_sub2(tm_f2 a, tm_f2 b) {
    tm_f2 r;
    asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ void tm_fma2(tm_f2& acc, tm_f2 a, tm_f2 b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

struct TmShared {
    int bbox[4];       // x0 min, x0 max, y0 min, y0 max over the recorded in-image cells
    int max_runs;
    int overflow;
};

struct TmGeom {
    float t1x, t1y, t1z, cx, cy, inv_cx, inv_cy, half_w, half_h;
};

// sweep_coord_fast with the reciprocal on the SFU (MUFU.RCP, ~1 ulp; the IEEE __frcp_rn expands
// to ~10 instructions and was 6 % of the kernel).  Every phase uses this one function, so a plane
// gets the same coordinate wherever it is recomputed.
// EXACT: the reference's operation order with IEEE divisions (sweep_coord), for images wider than
// ~200 px, where one ulp of a coordinate (6e-5 px at 1000 px) times a unit feature gradient is
// already the size of the parity budget and the 2-3 ulp of the fast form would exceed it.
template <bool EXACT>
__device__ __forceinline__ void tm_coord(const TmGeom& g, const PixelTerm& p, float d, float& ix, float& iy) {
    if (EXACT) {
        sweep_coord(g.t1x, g.t1y, g.t1z, p, d, g.cx, g.cy, g.half_w, g.half_h, ix, iy);
        return;
    }
    const float px = fmaf(p.x, d, g.t1x);
    const float py = fmaf(p.y, d, g.t1y);
    const float pz = fmaf(p.z, d, g.t1z);
    float inv;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(pz + 1e-10f));
    const float u = px * inv, v = py * inv;
    ix = fmaf(fmaf(u - g.cx, g.inv_cx, 1.0f), g.half_w, -0.5f);
    iy = fmaf(fmaf(v - g.cy, g.inv_cy, 1.0f), g.half_h, -0.5f);
}
template <bool EXACT>
__device__ __forceinline__ Tap tm_tap(const TmGeom& g, const PixelTerm& pt, float d) {
    float ix, iy;
    tm_coord<EXACT>(g, pt, d, ix, iy);
    return make_tap(ix, iy);
}

// 10-term quadratic form of one plane in the Gram matrix of its cell
__device__ __forceinline__ float tm_quad(const float* G, float fx, float fy) {
    Tap tap;
    tap.x0 = 0; tap.y0 = 0; tap.fx = fx; tap.fy = fy;
    float nw, ne, sw, se;
    bilinear_weights(tap, nw, ne, sw, se);
    const float diag = nw * nw * G[0] + ne * ne * G[4] + sw * sw * G[7] + se * se * G[9];
    const float off = nw * (ne * G[1] + sw * G[2] + se * G[3]) + ne * (sw * G[5] + se * G[6]) + sw * se * G[8];
    return fmaf(2.0f, off, diag);
}

// Window does not fit: planes [ka, kb) of one pixel with per-thread gathers from global memory.
template <bool EXACT>
__device__ __noinline__ void tm_gather_planes(int C, int H, int W, const float* __restrict__ src,
                                              const float* __restrict__ refp, const TmGeom& g,
                                              const PixelTerm& pt, const float* d_s, int ka, int kb,
                                              float inv_sigma, float* out_col, bool first_view) {
    const int HW = H * W;
    int k = ka;
    while (k < kb) {
        float ix, iy;
        tm_coord<EXACT>(g, pt, d_s[k], ix, iy);
        const Tap tap = make_tap(ix, iy);
        const CellTaps cell = cell_taps(tap, H, W);
        float q[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) q[i] = 0.f;
        for (int c = 0; c < C; ++c) {
            const float r = __ldg(refp + (long long)c * HW);
            const float* sc = src + (long long)c * HW;
            const float e0 = (cell.v00 ? __ldg(sc + cell.o00) : 0.f) - r;
            const float e1 = (cell.v01 ? __ldg(sc + cell.o01) : 0.f) - r;
            const float e2 = (cell.v10 ? __ldg(sc + cell.o10) : 0.f) - r;
            const float e3 = (cell.v11 ? __ldg(sc + cell.o11) : 0.f) - r;
            q[0] = fmaf(e0, e0, q[0]); q[1] = fmaf(e0, e1, q[1]); q[2] = fmaf(e0, e2, q[2]);
            q[3] = fmaf(e0, e3, q[3]); q[4] = fmaf(e1, e1, q[4]); q[5] = fmaf(e1, e2, q[5]);
            q[6] = fmaf(e1, e3, q[6]); q[7] = fmaf(e2, e2, q[7]); q[8] = fmaf(e2, e3, q[8]);
            q[9] = fmaf(e3, e3, q[9]);
        }
        const float cx0 = (float)tap.x0, cy0 = (float)tap.y0;
        const bool outside = cell.id < 0;
        float fx = tap.fx, fy = tap.fy;
        for (;;) {
            const float val = (outside ? q[0] : tm_quad(q, fx, fy)) * inv_sigma;
            float* o = out_col + k * TM_OS;
            *o = first_view ? val : (*o + val);
            if (++k >= kb) break;
            tm_coord<EXACT>(g, pt, d_s[k], ix, iy);
            const Tap nt = make_tap(ix, iy);
            if (outside) {
                if (cell_taps(nt, H, W).id >= 0) break;
            } else {
                fx = ix - cx0; fy = iy - cy0;
                const bool same = (nt.x0 == tap.x0 && nt.y0 == tap.y0) ||
                                  (fx >= -kCellSlack && fx <= 1.0f + kCellSlack && fy >= -kCellSlack &&
                                   fy <= 1.0f + kCellSlack && nt.x0 > -1000000);
                if (!same) break;
            }
        }
    }
}


typedef CUresult (*tm_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                 CUtensorMapFloatOOBfill);
tm_encode_fn tm_encoder();      // cuTensorMapEncodeTiled through the runtime's driver entry point (sweep_tma.cu)

}  // namespace dpv
