// K1 stand-alone (per-plane warp) and K4a (diagonal feature warp).
//
// dpv_warp_planes replaces _back_warp_homo_parallel (reference warping/homography.py:170-198)
// for callers that need the warped [N,C,H,W] stack itself; dpv_warp_feature replaces
// warp_feature (:137-168), which in the reference builds that stack per view (100 MB at the
// model's shapes) only to keep its diagonal out[k] = warped[k, k].
#include "dpv_common.cuh"

namespace dpv {

// One thread per (plane n, pixel p), looping over channels: every load/store instruction of a
// warp touches 32 consecutive pixels of one channel plane.
__global__ void __launch_bounds__(128) warp_planes_kernel(
    const float* __restrict__ img, const float* __restrict__ d, const float* __restrict__ term1,
    const float* __restrict__ term2, float* __restrict__ out, int N, int C, int H, int W,
    long long img_ns, float cx, float cy) {
    const int HW = H * W;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = blockIdx.y;
    if (p >= HW) return;
    PixelTerm pt;
    pt.x = __ldg(term2 + p); pt.y = __ldg(term2 + HW + p); pt.z = __ldg(term2 + 2 * HW + p);
    float ix, iy;
    sweep_coord(__ldg(term1), __ldg(term1 + 1), __ldg(term1 + 2), pt, __ldg(d + n), cx, cy,
                (float)W * 0.5f, (float)H * 0.5f, ix, iy);
    const Tap tap = make_tap(ix, iy);
    float nw, ne, sw, se;
    bilinear_weights(tap, nw, ne, sw, se);
    const bool xl = (tap.x0 >= 0) & (tap.x0 < W), xr = (tap.x0 + 1 >= 0) & (tap.x0 + 1 < W);
    const bool yt = (tap.y0 >= 0) & (tap.y0 < H), yb = (tap.y0 + 1 >= 0) & (tap.y0 + 1 < H);
    const bool v00 = xl & yt, v01 = xr & yt, v10 = xl & yb, v11 = xr & yb;
    const int base = tap.y0 * W + tap.x0;
    const float* s = img + (long long)n * img_ns;
    const float* s00 = s + (v00 ? base : 0);
    const float* s01 = s + (v01 ? base + 1 : 0);
    const float* s10 = s + (v10 ? base + W : 0);
    const float* s11 = s + (v11 ? base + W + 1 : 0);
    float* o = out + (long long)n * C * HW + p;
#pragma unroll 4
    for (int c = 0; c < C; ++c) {
        float w = __fmul_rn(v00 ? __ldg(s00) : 0.f, nw);
        w = __fadd_rn(w, __fmul_rn(v01 ? __ldg(s01) : 0.f, ne));
        w = __fadd_rn(w, __fmul_rn(v10 ? __ldg(s10) : 0.f, sw));
        w = __fadd_rn(w, __fmul_rn(v11 ? __ldg(s11) : 0.f, se));
        st_stream(o, w);
        o += HW; s00 += HW; s01 += HW; s10 += HW; s11 += HW;
    }
}

// One thread per (item, view, pixel, group of WF_KPT planes): the view geometry (K R, K t) is
// computed once per CTA into shared memory, the pixel term once per thread, and each plane then
// costs one projection, 4 taps of channel k and a blend.  (One thread per output element
// recomputed the whole geometry -- 21 loads and ~70 flops -- for 4 taps.)  The projection keeps the
// reference's operation order (sweep_coord: a 1-ulp change of a 96-px coordinate is 8e-6 px, the
// size of the parity budget on a unit-gradient feature) and the blend keeps ATen's association.
constexpr int WF_KPT = 8;
__global__ void __launch_bounds__(128) warp_feature_kernel(
    const float* __restrict__ feat, const float* __restrict__ pose, const float* __restrict__ K,
    const float* __restrict__ rays, const float* __restrict__ d, float* __restrict__ out,
    int V, int D, int H, int W, long long pose_bs, long long k_bs, long long rays_bs) {
    __shared__ float geo_s[16];
    __shared__ float d_s[WF_KPT];
    const int HW = H * W;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int k0 = blockIdx.y * WF_KPT;
    const int b = blockIdx.z / V, v = blockIdx.z % V;
    const int tid = threadIdx.x;
    if (tid < 12) {   // same dot3 order as load_view_geom
        const float* Kp = K + (long long)b * k_bs + (tid < 9 ? tid / 3 : tid - 9) * 3;
        const float* pp = pose + (long long)b * pose_bs + (long long)v * 16 + (tid < 9 ? tid % 3 : 3);
        geo_s[tid] = dot3(__ldg(Kp), __ldg(Kp + 1), __ldg(Kp + 2), __ldg(pp), __ldg(pp + 4), __ldg(pp + 8));
    } else if (tid < 14) {
        geo_s[tid] = __ldg(K + (long long)b * k_bs + (tid == 12 ? 2 : 5));
    } else if (tid >= 32 && tid < 32 + WF_KPT) {
        d_s[tid - 32] = (k0 + tid - 32 < D) ? __ldg(d + k0 + tid - 32) : 1.0f;
    }
    __syncthreads();
    if (p >= HW) return;
    const float* r = rays + (long long)b * rays_bs;
    const float rx = __ldg(r + p), ry = __ldg(r + HW + p), rz = __ldg(r + 2 * HW + p);
    PixelTerm pt;
    pt.x = dot3(geo_s[0], geo_s[1], geo_s[2], rx, ry, rz);
    pt.y = dot3(geo_s[3], geo_s[4], geo_s[5], rx, ry, rz);
    pt.z = dot3(geo_s[6], geo_s[7], geo_s[8], rx, ry, rz);
    const float t1x = geo_s[9], t1y = geo_s[10], t1z = geo_s[11], cx = geo_s[12], cy = geo_s[13];
    const float half_w = (float)W * 0.5f, half_h = (float)H * 0.5f;
    const long long plane0 = (((long long)b * V + v) * D + k0) * HW;
    const int nk = min(WF_KPT, D - k0);
#pragma unroll
    for (int kk = 0; kk < WF_KPT; ++kk) {
        if (kk < nk) {
            float ix, iy;
            sweep_coord(t1x, t1y, t1z, pt, d_s[kk], cx, cy, half_w, half_h, ix, iy);
            const Tap tap = make_tap(ix, iy);
            float nw, ne, sw, se;
            bilinear_weights(tap, nw, ne, sw, se);
            const bool xl = (tap.x0 >= 0) & (tap.x0 < W), xr = (tap.x0 + 1 >= 0) & (tap.x0 + 1 < W);
            const bool yt = (tap.y0 >= 0) & (tap.y0 < H), yb = (tap.y0 + 1 >= 0) & (tap.y0 + 1 < H);
            const float* s = feat + plane0 + (long long)kk * HW;
            const int base = tap.y0 * W + tap.x0;
            float w = __fmul_rn((xl & yt) ? __ldg(s + base) : 0.f, nw);
            w = __fadd_rn(w, __fmul_rn((xr & yt) ? __ldg(s + base + 1) : 0.f, ne));
            w = __fadd_rn(w, __fmul_rn((xl & yb) ? __ldg(s + base + W) : 0.f, sw));
            w = __fadd_rn(w, __fmul_rn((xr & yb) ? __ldg(s + base + W + 1) : 0.f, se));
            st_stream(out + plane0 + (long long)kk * HW + p, w);
        }
    }
}

}  // namespace dpv

extern "C" int dpv_warp_planes(const float* img, const float* d, const float* term1,
                               const float* term2, float* out, int N, int C, int H, int W,
                               int64_t img_nstride, float cx, float cy, void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(img && d && term1 && term2 && out);
    DPV_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0);
    if (N > 65535) return DPV_E_UNSUPP;
    const int HW = H * W;
    dim3 grid((HW + 127) / 128, N), block(128);
    warp_planes_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(img, d, term1, term2, out, N, C, H,
                                                                 W, img_nstride, cx, cy);
    DPV_LAUNCH_END();
    return 0;
}

extern "C" int dpv_warp_feature(const float* feat, const float* pose, const float* K,
                                const float* rays, const float* d_candi, float* out, int B, int V,
                                int D, int H, int W, int64_t pose_bstride, int64_t k_bstride,
                                int64_t rays_bstride, void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(feat && pose && K && rays && d_candi && out);
    DPV_CHECK_ARG(B > 0 && V > 0 && D > 0 && H > 0 && W > 0);
    if (D > 65535 || (long long)B * V > 65535) return DPV_E_UNSUPP;
    const int HW = H * W;
    dim3 grid((HW + 127) / 128, (D + WF_KPT - 1) / WF_KPT, B * V), block(128);
    warp_feature_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(
        feat, pose, K, rays, d_candi, out, V, D, H, W, pose_bstride, k_bstride, rays_bstride);
    DPV_LAUNCH_END();
    return 0;
}
