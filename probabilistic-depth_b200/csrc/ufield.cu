// K5: uncertainty-field (UF) collapse along the road surface.
//
// Replaces gen_ufield (reference utils/img_utils.py:268-358; helpers convert_flowfield
// :170-176, depth_to_pts :111-135, dpv_to_depthmap :52-61).  The reference shifts the whole DPV
// down by `pshift` rows with a nearest grid_sample, takes E[d] of the shifted and unshifted
// volume, thresholds the back-projected height, shifts the mask back up, multiplies the
// exponentiated DPV by the repeated mask and sums over rows: ~15 launches and 3+ passes over
// the 25 MB volume.  A nearest-neighbour shift is a pure index map, so it is expressed here as
// four small look-up tables (built by the host wrapper from the reference's own grid
// construction, so the half-pixel rounding at the borders is reproduced, not re-derived):
//     shifted[y',x'] = src[row_fwd[y'], col_fwd[x']]   (or zero padding when an index is -1)
//     mask'[y,x]     = mask[row_inv[y], col_inv[x]]
// The kernel reads E[d] (393 KB per item), decides per pixel whether it is on the road band,
// and touches the DPV only on rows where the mask is non-zero.  Row chunks are reduced through
// a workspace so the summation order is fixed (no float atomics).
#include "dpv_common.cuh"

namespace dpv {

struct UfArgs {
    const float* dpv; const float* depth; const float* d; const float* intr; const float* mask;
    const int* row_fwd; const int* row_inv; const int* col_fwd; const int* col_inv;
    float* uf; float* depth_zero; float* part; float* cnt;
    int B, D, H, W, mode, rows_per_chunk, nchunk;
    long long intr_bs;
    float zstart, zend, maxd1, mind, pad_depth;
};

// Weight of shifted-frame pixel (ys, xs): 1 when its back-projected point lies in the height
// band and depth range (utils/img_utils.py:316, comparisons kept negated so NaN passes), times
// the shifted ground-truth mask when one is given (:317-322).
__device__ __forceinline__ float band_weight(const UfArgs& a, const float* __restrict__ depth_b,
                                             const float* __restrict__ mask_b, int ys, int xs,
                                             float fy, float cy) {
    const int sy = a.row_fwd[ys], sx = a.col_fwd[xs];
    const bool inside = (sy >= 0) & (sx >= 0);
    const float z = inside ? __ldg(depth_b + sy * a.W + sx) : a.pad_depth;
    const float yf = __fdiv_rn(__fsub_rn((float)ys, cy), fy);
    const float yy = __fmul_rn(yf, z);
    const bool out = (yy > a.zend) || (yy < a.zstart) || (z > a.maxd1) || (z < a.mind);
    float w = out ? 0.f : 1.f;
    if (mask_b != nullptr) w = __fmul_rn(w, inside ? __ldg(mask_b + sy * a.W + sx) : 0.f);
    return w;
}

constexpr int UF_COLS = 32;    // columns per CTA (one warp-width: 128 B rows of the volume)
constexpr int UF_GROUPS = 8;   // warps per CTA, each owning D/8 consecutive bins
constexpr int UF_ROWS = 32;    // image rows per CTA (one partial sum per chunk of rows)

// One CTA = 32 columns x 32 rows of one item.  Step 1: all 256 threads evaluate the per-pixel
// weights once into shared memory (both roles of a row index: as a shifted-frame row for the
// denominator, as an image row for the numerator).  Step 2: warp g accumulates bins
// [g*DB, (g+1)*DB) in registers over the rows whose weight is non-zero; rows off the road band
// cost nothing, and a row on the band is DB coalesced 128 B loads per warp.
template <int DB>
__global__ void __launch_bounds__(UF_COLS * UF_GROUPS) ufield_partial_kernel(const UfArgs a) {
    __shared__ float w_s[UF_ROWS][UF_COLS];     // numerator weight of image pixel (r, x)
    __shared__ float z_s[UF_ROWS][UF_COLS];     // denominator weight of shifted pixel (r, x)
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int x = blockIdx.x * UF_COLS + lane;
    const int chunk = blockIdx.y, b = blockIdx.z;
    const int HW = a.H * a.W;
    const float* depth_b = a.depth + (long long)b * HW;
    const float* mask_b = a.mask ? a.mask + (long long)b * HW : nullptr;
    const float fy = __ldg(a.intr + b * a.intr_bs + 4), cy = __ldg(a.intr + b * a.intr_bs + 5);
    const int r0 = chunk * UF_ROWS;
    const bool col_ok = x < a.W;
    const int xi = col_ok ? a.col_inv[x] : -1;
    for (int rr = grp; rr < UF_ROWS; rr += UF_GROUPS) {
        const int r = r0 + rr;
        float wz = 0.f, w = 0.f;
        if (col_ok && r < a.H) {
            wz = band_weight(a, depth_b, mask_b, r, x, fy, cy);
            const int yi = a.row_inv[r];
            if (yi >= 0 && xi >= 0) w = band_weight(a, depth_b, mask_b, yi, xi, fy, cy);
            if (a.depth_zero != nullptr)
                a.depth_zero[(long long)b * HW + r * a.W + x] = __fmul_rn(__ldg(depth_b + r * a.W + x), w);
        }
        z_s[rr][lane] = wz;
        w_s[rr][lane] = w;
    }
    __syncthreads();
    if (grp == 0 && col_ok) {
        float cnt = 0.f;
        for (int rr = 0; rr < UF_ROWS; ++rr) cnt = __fadd_rn(cnt, z_s[rr][lane]);
        a.cnt[((long long)b * a.nchunk + chunk) * a.W + x] = cnt;
    }
    const int kb = grp * DB;
    if (kb >= a.D) return;
    float acc[DB];
#pragma unroll
    for (int j = 0; j < DB; ++j) acc[j] = 0.f;
    const float* dpv_b = a.dpv + ((long long)b * a.D + kb) * HW + x;
    for (int rr = 0; rr < UF_ROWS; ++rr) {
        const float w = w_s[rr][lane];
        if (!__any_sync(0xffffffffu, w != 0.f)) continue;
        if (w != 0.f) {
            const float* col = dpv_b + (long long)(r0 + rr) * a.W;
            float v[DB];
#pragma unroll
            for (int j = 0; j < DB; ++j) v[j] = (kb + j < a.D) ? ld_stream(col + (long long)j * HW) : 0.f;
#pragma unroll
            for (int j = 0; j < DB; ++j) {
                const float pr = (a.mode == DPV_IN_PROB) ? v[j] : expf(v[j]);
                acc[j] = __fadd_rn(acc[j], __fmul_rn(pr, w));
            }
        }
    }
    if (col_ok) {
        float* part = a.part + (((long long)b * a.nchunk + chunk) * a.D + kb) * a.W + x;
#pragma unroll
        for (int j = 0; j < DB; ++j)
            if (kb + j < a.D) part[(long long)j * a.W] = acc[j];
    }
}

__global__ void __launch_bounds__(128) ufield_finish_kernel(const UfArgs a) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y, b = blockIdx.z;
    if (x >= a.W) return;
    float num = 0.f, den = 0.f;
    for (int c = 0; c < a.nchunk; ++c) {
        num = __fadd_rn(num, a.part[(((long long)b * a.nchunk + c) * a.D + k) * a.W + x]);
        den = __fadd_rn(den, a.cnt[((long long)b * a.nchunk + c) * a.W + x]);
    }
    a.uf[((long long)b * a.D + k) * a.W + x] = __fdiv_rn(num, den);   // 0/0 -> NaN, as the reference
}

}  // namespace dpv

static int ufield_row_chunks(int H) { return (H + dpv::UF_ROWS - 1) / dpv::UF_ROWS; }

extern "C" int64_t dpv_ufield_workspace_floats(int B, int D, int H, int W) {
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
    return (int64_t)B * ufield_row_chunks(H) * (D + 1) * W;
}

extern "C" int dpv_ufield(const float* dpv, const float* depth, const float* d_candi,
                          const float* intr_up, const float* mask, const int* row_fwd,
                          const int* row_inv, const int* col_fwd, const int* col_inv, float* uf,
                          float* depth_zero, float* workspace, int B, int D, int H, int W,
                          int64_t intr_bstride, int in_mode, float zstart, float zend, float maxd,
                          float mind, float pad_depth, void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(dpv && depth && d_candi && intr_up && row_fwd && row_inv && col_fwd && col_inv);
    DPV_CHECK_ARG(uf && workspace);
    DPV_CHECK_ARG(B > 0 && D > 0 && H > 0 && W > 0);
    DPV_CHECK_ARG(in_mode == DPV_IN_LOGPROB || in_mode == DPV_IN_PROB);
    if (B > 65535 || D > 65535) return DPV_E_UNSUPP;
    if (D > 32 * UF_GROUPS) return DPV_E_UNSUPP;
    UfArgs a;
    a.dpv = dpv; a.depth = depth; a.d = d_candi; a.intr = intr_up; a.mask = mask;
    a.row_fwd = row_fwd; a.row_inv = row_inv; a.col_fwd = col_fwd; a.col_inv = col_inv;
    a.uf = uf; a.depth_zero = depth_zero;
    a.B = B; a.D = D; a.H = H; a.W = W; a.mode = in_mode;
    a.nchunk = ufield_row_chunks(H);
    a.rows_per_chunk = UF_ROWS;
    a.part = workspace;
    a.cnt = workspace + (long long)B * a.nchunk * D * W;
    a.intr_bs = intr_bstride;
    a.zstart = zstart; a.zend = zend; a.maxd1 = maxd - 1.0f; a.mind = mind;
    a.pad_depth = pad_depth;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((W + UF_COLS - 1) / UF_COLS, a.nchunk, B), block(UF_COLS * UF_GROUPS);
    const int db = (D + UF_GROUPS - 1) / UF_GROUPS;
    if (db <= 4) ufield_partial_kernel<4><<<grid, block, 0, st>>>(a);
    else if (db <= 8) ufield_partial_kernel<8><<<grid, block, 0, st>>>(a);
    else if (db <= 16) ufield_partial_kernel<16><<<grid, block, 0, st>>>(a);
    else ufield_partial_kernel<32><<<grid, block, 0, st>>>(a);
    DPV_LAUNCH_END();
    dim3 grid2((W + 127) / 128, D, B), block2(128);
    ufield_finish_kernel<<<grid2, block2, 0, st>>>(a);
    DPV_LAUNCH_END();
    return 0;
}
