// K1 + K2a: plane-sweep homography warp fused with the feature-distance reduction.
//
// Replaces est_swp_volume_v4 / _back_warp_homo_parallel / img_dis_L{1,2}_pard
// (reference warping/homography.py:80-86, 98-135, 170-198).  The reference materialises the
// source features D times, a [D,H,W,2] grid, the [D,C,H,W] warped stack and three more
// temporaries of that size per item and view; here one launch covers all items and views,
// reads each feature map once from HBM and writes only the [B,D,H,W] cost volume.
//
// Two formulations:
//  * direct : for every plane gather 4 taps per channel, blend, subtract, reduce.  Follows the
//             reference operation by operation; used for L1 and as the cross-check.
//  * gram   : (L2 only) consecutive planes of one reference pixel fall into the same source
//             2x2 cell (the sampling point moves < 1 px per plane at 1/4 resolution).  With
//             bilinear weights w_i (sum 1) and e_i = s_i - r  (s_i the 4 taps, 0 outside the
//             image; r the reference feature),
//                 sum_c (sum_i w_i s_i - r)^2 = sum_ij w_i w_j <e_i, e_j>,
//             so the channel contraction is done once per (pixel, cell) as a 4x4 Gram matrix of
//             differences (10 sums, no cancellation: every term is built from differences),
//             and every further plane in that cell costs a 10-term quadratic form.
#include <algorithm>
#include <cstdlib>

#include "sweep_common.cuh"

namespace dpv {

template <int NT>
__global__ void __launch_bounds__(NT) sweep_gram_kernel(const SweepArgs a) {
    extern __shared__ float smem[];
    const int HW = a.H * a.W;
    const int kper = (a.D + a.PS - 1) / a.PS;
    const int k0 = blockIdx.y * kper;
    const int k1 = min(a.D, k0 + kper);
    const int nk = k1 - k0;
    float* cost_s = smem;                 // [nk][NT]
    float* d_s = smem + kper * NT;        // [kper]
    const int tid = threadIdx.x;
    for (int k = tid; k < nk; k += NT) d_s[k] = __ldg(a.d + k0 + k);
    __syncthreads();
    if (nk <= 0) return;
    // Persistent over (item, pixel block) tiles: the number of resident warps per SM is chosen by
    // the launcher so that their combined footprint (C channel planes x a few 128 B lines each)
    // stays inside L1; consecutive tiles of a warp are neighbours in the image.
    const int blocks_per_item = (HW + NT - 1) / NT;
    const int ntiles = blocks_per_item * a.B;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = tile / blocks_per_item;
    const int p = (tile - b * blocks_per_item) * NT + tid;
    if (p >= HW) continue;

    const float* rays = a.rays + (long long)b * a.rays_bs;
    const float rx = __ldg(rays + p), ry = __ldg(rays + HW + p), rz = __ldg(rays + 2 * HW + p);
    const float* refp = a.ref + (long long)b * a.ref_bs + p;
    const float half_w = (float)a.W * 0.5f, half_h = (float)a.H * 0.5f;

    for (int v = 0; v < a.V; ++v) {
        const float* src = a.src + (long long)b * a.src_bs + (long long)v * a.src_vs;
        ViewGeom g = load_view_geom(a.K + (long long)b * a.k_bs,
                                    a.pose + (long long)b * a.pose_bs + (long long)v * 16);
        const PixelTerm pt = pixel_term(g, rx, ry, rz);
        const float t1x = g.t1[0], t1y = g.t1[1], t1z = g.t1[2], cx = g.cx, cy = g.cy;

        int k = 0;
        float ix, iy;
        sweep_coord(t1x, t1y, t1z, pt, d_s[0], cx, cy, half_w, half_h, ix, iy);
        Tap tap = make_tap(ix, iy);
        CellTaps cell = cell_taps(tap, a.H, a.W);
        while (k < nk) {
            // ---- Gram matrix of (tap - ref) differences for this cell
            float g00 = 0.f, g01 = 0.f, g02 = 0.f, g03 = 0.f, g11 = 0.f, g12 = 0.f, g13 = 0.f,
                  g22 = 0.f, g23 = 0.f, g33 = 0.f;
            {
                const float* s00 = src + cell.o00;
                const float* s01 = src + cell.o01;
                const float* s10 = src + cell.o10;
                const float* s11 = src + cell.o11;
                const bool v00 = cell.v00, v01 = cell.v01, v10 = cell.v10, v11 = cell.v11;
                const float* rp = refp;
#pragma unroll 4
                for (int c = 0; c < a.C; ++c) {
                    const float r = __ldg(rp);
                    const float e0 = (v00 ? __ldg(s00) : 0.f) - r;
                    const float e1 = (v01 ? __ldg(s01) : 0.f) - r;
                    const float e2 = (v10 ? __ldg(s10) : 0.f) - r;
                    const float e3 = (v11 ? __ldg(s11) : 0.f) - r;
                    g00 = fmaf(e0, e0, g00); g01 = fmaf(e0, e1, g01); g02 = fmaf(e0, e2, g02);
                    g03 = fmaf(e0, e3, g03); g11 = fmaf(e1, e1, g11); g12 = fmaf(e1, e2, g12);
                    g13 = fmaf(e1, e3, g13); g22 = fmaf(e2, e2, g22); g23 = fmaf(e2, e3, g23);
                    g33 = fmaf(e3, e3, g33);
                    rp += HW; s00 += HW; s01 += HW; s10 += HW; s11 += HW;
                }
            }
            // ---- every plane whose sample falls into this cell.  A sample within kCellSlack of
            // the cell border stays in the cell with a fraction marginally outside [0,1]: the
            // bilinear interpolant is continuous across cells, so this changes the value by
            // <= slack * |second difference|, the same order as the fp32 rounding of the
            // coordinate itself, and it stops a rectified pair (v = row centre +- 1 ulp, floor
            // flipping between two rows from plane to plane) from opening a new cell per plane.
            const int id = cell.id;
            const float cx0 = (float)tap.x0, cy0 = (float)tap.y0;
            do {
                float nw, ne, sw, se;
                bilinear_weights(tap, nw, ne, sw, se);
                float diag = nw * nw * g00 + ne * ne * g11 + sw * sw * g22 + se * se * g33;
                float off = nw * (ne * g01 + sw * g02 + se * g03) + ne * (sw * g12 + se * g13) +
                            sw * se * g23;
                float val = __fdiv_rn(fmaf(2.0f, off, diag), a.sigma);
                float* slot = cost_s + k * NT + tid;
                *slot = (v == 0) ? val : (*slot + val);
                ++k;
                if (k < nk) {
                    sweep_coord(t1x, t1y, t1z, pt, d_s[k], cx, cy, half_w, half_h, ix, iy);
                    tap = make_tap(ix, iy);
                    cell = cell_taps(tap, a.H, a.W);
                    const float rx0 = ix - cx0, ry0 = iy - cy0;
                    if (id >= 0 && cell.id != id && rx0 >= -kCellSlack && rx0 <= 1.0f + kCellSlack &&
                        ry0 >= -kCellSlack && ry0 <= 1.0f + kCellSlack) {
                        tap.x0 = (int)cx0; tap.y0 = (int)cy0; tap.fx = rx0; tap.fy = ry0;
                        cell.id = id;
                    }
                }
            } while (k < nk && cell.id == id);
        }
    }

    float* out = a.cost + ((long long)b * a.D + k0) * HW + p;
    for (int k = 0; k < nk; ++k) out[(long long)k * HW] = cost_s[k * NT + tid];
    if (a.lsm != nullptr) {   // only launched with PS == 1: this thread holds all D planes
        float m = -INFINITY;
        for (int k = 0; k < nk; ++k) m = fmaxf(m, cost_s[k * NT + tid]);
        float s = 0.f;
        for (int k = 0; k < nk; ++k) s += expf(cost_s[k * NT + tid] - m);
        const float ls = logf(s);
        float* lo = a.lsm + ((long long)b * a.D + k0) * HW + p;
        for (int k = 0; k < nk; ++k) lo[(long long)k * HW] = (cost_s[k * NT + tid] - m) - ls;
    }
    }   // tile loop
}

template <int NT, int DIST>
__global__ void __launch_bounds__(NT) sweep_direct_kernel(const SweepArgs a) {
    extern __shared__ float smem[];
    const int HW = a.H * a.W;
    const int kper = (a.D + a.PS - 1) / a.PS;
    const int k0 = blockIdx.y * kper;
    const int k1 = min(a.D, k0 + kper);
    const int nk = k1 - k0;
    float* cost_s = smem;
    const int b = blockIdx.z;
    const int tid = threadIdx.x;
    const int p = blockIdx.x * NT + tid;
    if (p >= HW || nk <= 0) return;

    const float* rays = a.rays + (long long)b * a.rays_bs;
    const float rx = __ldg(rays + p), ry = __ldg(rays + HW + p), rz = __ldg(rays + 2 * HW + p);
    const float* refp = a.ref + (long long)b * a.ref_bs + p;
    const float half_w = (float)a.W * 0.5f, half_h = (float)a.H * 0.5f;

    for (int v = 0; v < a.V; ++v) {
        const float* src = a.src + (long long)b * a.src_bs + (long long)v * a.src_vs;
        ViewGeom g = load_view_geom(a.K + (long long)b * a.k_bs,
                                    a.pose + (long long)b * a.pose_bs + (long long)v * 16);
        const PixelTerm pt = pixel_term(g, rx, ry, rz);
        for (int k = 0; k < nk; ++k) {
            float ix, iy;
            sweep_coord(g.t1[0], g.t1[1], g.t1[2], pt, __ldg(a.d + k0 + k), g.cx, g.cy, half_w,
                        half_h, ix, iy);
            const Tap tap = make_tap(ix, iy);
            const CellTaps cell = cell_taps(tap, a.H, a.W);
            float nw, ne, sw, se;
            bilinear_weights(tap, nw, ne, sw, se);
            const float* s00 = src + cell.o00;
            const float* s01 = src + cell.o01;
            const float* s10 = src + cell.o10;
            const float* s11 = src + cell.o11;
            const float* rp = refp;
            float acc = 0.f;
#pragma unroll 4
            for (int c = 0; c < a.C; ++c) {
                // same association as ATen: ((nw*a + ne*b) + sw*c) + se*d
                float w = __fmul_rn(cell.v00 ? __ldg(s00) : 0.f, nw);
                w = __fadd_rn(w, __fmul_rn(cell.v01 ? __ldg(s01) : 0.f, ne));
                w = __fadd_rn(w, __fmul_rn(cell.v10 ? __ldg(s10) : 0.f, sw));
                w = __fadd_rn(w, __fmul_rn(cell.v11 ? __ldg(s11) : 0.f, se));
                const float diff = __fsub_rn(w, __ldg(rp));
                acc = (DIST == DPV_DIST_L2) ? __fadd_rn(acc, __fmul_rn(diff, diff))
                                            : __fadd_rn(acc, fabsf(diff));
                rp += HW; s00 += HW; s01 += HW; s10 += HW; s11 += HW;
            }
            const float val = __fdiv_rn(acc, a.sigma);
            float* slot = cost_s + k * NT + tid;
            *slot = (v == 0) ? val : (*slot + val);
        }
    }
    float* out = a.cost + ((long long)b * a.D + k0) * HW + p;
    for (int k = 0; k < nk; ++k) out[(long long)k * HW] = cost_s[k * NT + tid];
    if (a.lsm != nullptr) {
        float m = -INFINITY;
        for (int k = 0; k < nk; ++k) m = fmaxf(m, cost_s[k * NT + tid]);
        float s = 0.f;
        for (int k = 0; k < nk; ++k) s += expf(cost_s[k * NT + tid] - m);
        const float ls = logf(s);
        float* lo = a.lsm + ((long long)b * a.D + k0) * HW + p;
        for (int k = 0; k < nk; ++k) lo[(long long)k * HW] = (cost_s[k * NT + tid] - m) - ls;
    }
}

static int pick_plane_split(int B, int HW, int D, bool need_all_planes) {
    if (need_all_planes) return 1;
    const long long warps = (long long)B * ((HW + 31) / 32);
    int ps = 1;
    while (ps < 8 && warps * ps < 148LL * 8 && D / (ps * 2) >= 8) ps *= 2;
    return ps;
}

}  // namespace dpv

extern "C" int64_t dpv_sweep_workspace_floats(int B, int V, int H, int W) {
    return dpv::sweep_xcorr_workspace_floats(B, V, H, W);
}

extern "C" int dpv_sweep_cost_volume(const float* ref, const float* src, const float* pose,
                                     const float* K, const float* rays, const float* d_candi,
                                     float* cost, float* log_softmax_out, int B, int V, int C,
                                     int D, int H, int W, int64_t ref_bstride, int64_t src_bstride,
                                     int64_t src_vstride, int64_t pose_bstride, int64_t k_bstride,
                                     int64_t rays_bstride, float sigma, int dist, int algo,
                                     void* stream) {
    return dpv_sweep_cost_volume_ws(ref, src, pose, K, rays, d_candi, cost, log_softmax_out, B, V, C, D, H, W,
                                    ref_bstride, src_bstride, src_vstride, pose_bstride, k_bstride, rays_bstride,
                                    sigma, dist, algo, nullptr, stream);
}

extern "C" int dpv_sweep_cost_volume_ws(const float* ref, const float* src, const float* pose,
                                        const float* K, const float* rays, const float* d_candi,
                                        float* cost, float* log_softmax_out, int B, int V, int C,
                                        int D, int H, int W, int64_t ref_bstride, int64_t src_bstride,
                                        int64_t src_vstride, int64_t pose_bstride, int64_t k_bstride,
                                        int64_t rays_bstride, float sigma, int dist, int algo,
                                        float* workspace, void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(ref && src && pose && K && rays && d_candi && cost);
    DPV_CHECK_ARG(B > 0 && V > 0 && C > 0 && D > 0 && H > 0 && W > 0);
    DPV_CHECK_ARG(dist == DPV_DIST_L2 || dist == DPV_DIST_L1);
    DPV_CHECK_ARG(algo >= 0 && algo <= 5);
    DPV_CHECK_ARG(algo != 5 || workspace != nullptr);
    if ((long long)H * W > (1LL << 30) || B > 65535) return DPV_E_UNSUPP;
    if (algo >= 2 && dist != DPV_DIST_L2) return DPV_E_UNSUPP;
    if (algo == 3 && log_softmax_out != nullptr) return DPV_E_UNSUPP;
    constexpr int NT = 32;
    SweepArgs a;
    a.ref = ref; a.src = src; a.pose = pose; a.K = K; a.rays = rays; a.d = d_candi;
    a.cost = cost; a.lsm = log_softmax_out;
    a.B = B; a.V = V; a.C = C; a.D = D; a.H = H; a.W = W;
    a.ref_bs = ref_bstride; a.src_bs = src_bstride; a.src_vs = src_vstride;
    a.pose_bs = pose_bstride; a.k_bs = k_bstride; a.rays_bs = rays_bstride;
    a.sigma = sigma;
    const int HW = H * W;
    a.PS = pick_plane_split(B, HW, D, log_softmax_out != nullptr);
    if (algo == 0) {
        static const int no_xc = [] { const char* v = getenv("DPV_SWEEP_NO_XCORR"); return v ? atoi(v) : 0; }();
        if (dist != DPV_DIST_L2) algo = 1;
        else if (workspace != nullptr && !no_xc && sweep_xcorr_supported(a)) algo = 5;
        else if (sweep_gram_tma_supported(a)) algo = 4;
        else algo = (log_softmax_out != nullptr) ? 2 : 3;
    }
    if (algo == 5) return launch_sweep_xcorr(a, workspace, (cudaStream_t)stream);
    if (algo == 4) return launch_sweep_gram_tma(a, (cudaStream_t)stream);
    if (algo == 3) return launch_sweep_gram_tiled(a, (cudaStream_t)stream);
    const int kper = (D + a.PS - 1) / a.PS;
    const size_t smem = (size_t)kper * (NT + 1) * sizeof(float);
    if (smem > 200 * 1024) return DPV_E_UNSUPP;
    dim3 grid((HW + NT - 1) / NT, a.PS, B), block(NT);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
    if (algo == 2) {
        if (smem > 48 * 1024)
            e = cudaFuncSetAttribute(sweep_gram_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        static const int wps = [] { const char* v = getenv("DPV_SWEEP_WPS"); return v ? atoi(v) : 0; }();
        const int ntiles = grid.x * B;
        int gx = ntiles;
        if (wps > 0) gx = std::min(ntiles, std::max(1, 148 * wps / a.PS));
        dim3 pgrid(gx, a.PS, 1);
        sweep_gram_kernel<NT><<<pgrid, block, smem, st>>>(a);
    } else if (dist == DPV_DIST_L2) {
        if (smem > 48 * 1024)
            e = cudaFuncSetAttribute(sweep_direct_kernel<NT, DPV_DIST_L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        sweep_direct_kernel<NT, DPV_DIST_L2><<<grid, block, smem, st>>>(a);
    } else {
        if (smem > 48 * 1024)
            e = cudaFuncSetAttribute(sweep_direct_kernel<NT, DPV_DIST_L1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        sweep_direct_kernel<NT, DPV_DIST_L1><<<grid, block, smem, st>>>(a);
    }
    DPV_LAUNCH_END();
    return 0;
}
