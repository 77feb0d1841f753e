// Shared device helpers for the DPV kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <stdlib.h>
#include <atomic>

#include "../../include/dpv_b200.h"

namespace dpv {

extern std::atomic<long long> g_launch_count;   // defined in capi.cu; bumped by every launcher

#define DPV_CHECK_ARG(cond) do { if (!(cond)) return DPV_E_BADARG; } while (0)
#define DPV_LAUNCH_END() do { ::dpv::g_launch_count.fetch_add(1, std::memory_order_relaxed); cudaError_t e__ = cudaGetLastError(); \
                              if (e__ != cudaSuccess) return (int)e__; } while (0)

// ------------------------------------------------------------------ plane-sweep geometry
// Per (item, view) constants: term1 = K t and M = K R, so that term2(pixel) = M ray.
// Follows warping/homography.py:119-121; the per-plane part follows :187-196 with every
// fp32 rounding the reference's tensor ops perform kept as a separate _rn operation.
struct ViewGeom {
    float m[9];    // K R, row-major
    float t1[3];   // K t
    float cx, cy;  // K[0,2], K[1,2]
};

__device__ __forceinline__ float dot3(float a0, float a1, float a2, float b0, float b1, float b2) {
    return fmaf(a2, b2, fmaf(a1, b1, a0 * b0));
}

__device__ __forceinline__ ViewGeom load_view_geom(const float* __restrict__ K,
                                                   const float* __restrict__ pose) {
    ViewGeom g;
    float k[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) k[i] = __ldg(K + i);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j)
            g.m[i * 3 + j] = dot3(k[i * 3], k[i * 3 + 1], k[i * 3 + 2],
                                  __ldg(pose + j), __ldg(pose + 4 + j), __ldg(pose + 8 + j));
        g.t1[i] = dot3(k[i * 3], k[i * 3 + 1], k[i * 3 + 2],
                       __ldg(pose + 3), __ldg(pose + 7), __ldg(pose + 11));
    }
    g.cx = k[2];
    g.cy = k[5];
    return g;
}

struct PixelTerm { float x, y, z; };   // term2 for one reference pixel

__device__ __forceinline__ PixelTerm pixel_term(const ViewGeom& g, float rx, float ry, float rz) {
    PixelTerm p;
    p.x = dot3(g.m[0], g.m[1], g.m[2], rx, ry, rz);
    p.y = dot3(g.m[3], g.m[4], g.m[5], rx, ry, rz);
    p.z = dot3(g.m[6], g.m[7], g.m[8], rx, ry, rz);
    return p;
}

// A sampling position in source-pixel units, decomposed the way ATen's bilinear
// grid_sample does (align_corners=False): top-left tap (x0,y0) and the fractions.
struct Tap {
    int x0, y0;     // floor of the un-normalised coordinate (clamped sentinel when not finite)
    float fx, fy;   // coordinate - floor, in [0,1)
};

// term1 + term2*d -> perspective divide -> normalise by the principal point -> un-normalise.
__device__ __forceinline__ void sweep_coord(float t1x, float t1y, float t1z, const PixelTerm& p,
                                            float d, float cx, float cy, float half_w, float half_h,
                                            float& ix, float& iy) {
    float px = __fadd_rn(t1x, __fmul_rn(p.x, d));
    float py = __fadd_rn(t1y, __fmul_rn(p.y, d));
    float pz = __fadd_rn(t1z, __fmul_rn(p.z, d));
    float den = __fadd_rn(pz, 1e-10f);
    float u = __fdiv_rn(px, den);
    float v = __fdiv_rn(py, den);
    float gx = __fdiv_rn(__fsub_rn(u, cx), cx);
    float gy = __fdiv_rn(__fsub_rn(v, cy), cy);
    // ATen CPU kernel: (g + 1) * (size / 2) - 0.5
    ix = __fsub_rn(__fmul_rn(__fadd_rn(gx, 1.0f), half_w), 0.5f);
    iy = __fsub_rn(__fmul_rn(__fadd_rn(gy, 1.0f), half_h), 0.5f);
}

// Same mapping with one reciprocal instead of four IEEE divisions (used by the Gram kernels, whose
// channel contraction already differs from the reference's operation order at the 1e-6 level):
// u = Px * rcp(Pz + 1e-10); ix = ((u - cx) / cx + 1) * W/2 - 0.5 with 1/cx precomputed.  Agrees
// with sweep_coord to ~2 ulp of the coordinate (< 1e-5 px at the model's sizes).
__device__ __forceinline__ void sweep_coord_fast(float t1x, float t1y, float t1z, const PixelTerm& p,
                                                 float d, float cx, float cy, float inv_cx,
                                                 float inv_cy, float half_w, float half_h,
                                                 float& ix, float& iy) {
    const float px = fmaf(p.x, d, t1x);
    const float py = fmaf(p.y, d, t1y);
    const float pz = fmaf(p.z, d, t1z);
    const float inv = __frcp_rn(pz + 1e-10f);
    const float u = px * inv, v = py * inv;
    ix = fmaf(fmaf(u - cx, inv_cx, 1.0f), half_w, -0.5f);
    iy = fmaf(fmaf(v - cy, inv_cy, 1.0f), half_h, -0.5f);
}

// Non-finite or absurdly large coordinates are sent far outside the image so that every tap
// is out of bounds and samples zero (the behaviour of ATen's CUDA grid sampler).
__device__ __forceinline__ Tap make_tap(float ix, float iy) {
    Tap t;
    const float lim = 1.0e6f;
    bool ok = (ix > -lim) && (ix < lim) && (iy > -lim) && (iy < lim);   // false for NaN
    float fxf = floorf(ix), fyf = floorf(iy);
    t.x0 = ok ? (int)fxf : -1000000;
    t.y0 = ok ? (int)fyf : -1000000;
    t.fx = ok ? __fsub_rn(ix, fxf) : 0.0f;
    t.fy = ok ? __fsub_rn(iy, fyf) : 0.0f;
    return t;
}

// Bilinear weights as ATen forms them: w = x - floor(x), e = 1 - w; nw = s*e, ne = s*w, ...
__device__ __forceinline__ void bilinear_weights(const Tap& t, float& nw, float& ne, float& sw,
                                                 float& se) {
    float w = t.fx, e = __fsub_rn(1.0f, t.fx);
    float n = t.fy, s = __fsub_rn(1.0f, t.fy);
    nw = __fmul_rn(s, e);
    ne = __fmul_rn(s, w);
    sw = __fmul_rn(n, e);
    se = __fmul_rn(n, w);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Programmatic dependent launch (PDL).  A kernel launched with dpv_launch_pdl may become resident while
// the previous kernel of the stream is still draining its last CTAs; pdl_wait() at its top blocks until
// that kernel has completed and its writes are visible, so stream-order semantics are unchanged -- what
// disappears is the launch latency / drain bubble.  Used for ONE boundary only, the small finish kernel
// behind the fused head kernel (-1.6 us).  Measured the hard way: with all three kernels of the step
// chained this way the step went from 0.167 to 0.189 ms -- a big dependent grid that is already
// resident starts all its CTAs in the same instant, and the head kernel's load / compute phases then
// stay aligned across the SM instead of overlapping.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t dpv_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                         cudaStream_t st, Args&&... args) {
    static const bool off = [] { const char* e = getenv("DPV_NO_PDL"); return e && atoi(e) != 0; }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = off ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Streaming global accesses: the big volumes are touched once, keep them out of L1.
__device__ __forceinline__ float ld_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 ld_stream2(const float* p) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(float* p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void st_stream2(float* p, float2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

}  // namespace dpv
