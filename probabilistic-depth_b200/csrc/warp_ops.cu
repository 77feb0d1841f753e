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

// One thread per (item, view, plane k, pixel): 4 taps of channel k only.
__global__ void __launch_bounds__(128) warp_feature_kernel(
    const float* __restrict__ feat, const float* __restrict__ pose, const float* __restrict__ K,
    const float* __restrict__ rays, const float* __restrict__ d, float* __restrict__ out,
    int V, int D, int H, int W, long long pose_bs, long long k_bs, long long rays_bs) {
    const int HW = H * W;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    const int b = blockIdx.z / V, v = blockIdx.z % V;
    if (p >= HW) return;
    const ViewGeom g = load_view_geom(K + (long long)b * k_bs,
                                      pose + (long long)b * pose_bs + (long long)v * 16);
    const float* r = rays + (long long)b * rays_bs;
    const PixelTerm pt = pixel_term(g, __ldg(r + p), __ldg(r + HW + p), __ldg(r + 2 * HW + p));
    float ix, iy;
    sweep_coord(g.t1[0], g.t1[1], g.t1[2], pt, __ldg(d + k), g.cx, g.cy, (float)W * 0.5f,
                (float)H * 0.5f, ix, iy);
    const Tap tap = make_tap(ix, iy);
    float nw, ne, sw, se;
    bilinear_weights(tap, nw, ne, sw, se);
    const bool xl = (tap.x0 >= 0) & (tap.x0 < W), xr = (tap.x0 + 1 >= 0) & (tap.x0 + 1 < W);
    const bool yt = (tap.y0 >= 0) & (tap.y0 < H), yb = (tap.y0 + 1 >= 0) & (tap.y0 + 1 < H);
    const long long plane = (((long long)b * V + v) * D + k) * HW;
    const float* s = feat + plane;
    const int base = tap.y0 * W + tap.x0;
    float w = __fmul_rn((xl & yt) ? __ldg(s + base) : 0.f, nw);
    w = __fadd_rn(w, __fmul_rn((xr & yt) ? __ldg(s + base + 1) : 0.f, ne));
    w = __fadd_rn(w, __fmul_rn((xl & yb) ? __ldg(s + base + W) : 0.f, sw));
    w = __fadd_rn(w, __fmul_rn((xr & yb) ? __ldg(s + base + W + 1) : 0.f, se));
    out[plane + p] = w;
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
    dim3 grid((HW + 127) / 128, D, B * V), block(128);
    warp_feature_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(
        feat, pose, K, rays, d_candi, out, V, D, H, W, pose_bstride, k_bstride, rays_bstride);
    DPV_LAUNCH_END();
    return 0;
}
