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
    float* uf; float* depth_zero; float* part; float* cnt; float* wmap; float* colmin;
    int B, D, H, W, mode, rows_per_chunk, nchunk;
    long long intr_bs;
    float zstart, zend, maxd1, mind, pad_depth, quash;
};

// Weight of shifted-frame pixel (ys, xs): 1 when its back-projected point lies in the height
// band and depth range (utils/img_utils.py:316, comparisons kept negated so NaN passes), times
// the shifted ground-truth mask when one is given (:317-322).
// (sy, yf) are properties of the shifted row ys alone -- its source row and (ys - cy) / fy -- and
// are evaluated once per row of the chunk, not once per pixel.
// `zc` receives the "cleaned" shifted depth of the quash_limit branch (:327-328): depth * weight with
// exact zeros replaced by 1000.
__device__ __forceinline__ float band_weight(const UfArgs& a, const float* __restrict__ depth_b,
                                             const float* __restrict__ mask_b, int sy, float yf, int sx,
                                             float* zc = nullptr) {
    const bool inside = (sy >= 0) & (sx >= 0);
    const float z = inside ? __ldg(depth_b + sy * a.W + sx) : a.pad_depth;
    const float yy = __fmul_rn(yf, z);
    const bool out = (yy > a.zend) || (yy < a.zstart) || (z > a.maxd1) || (z < a.mind);
    float w = out ? 0.f : 1.f;
    if (mask_b != nullptr) w = __fmul_rn(w, inside ? __ldg(mask_b + sy * a.W + sx) : 0.f);
    if (zc != nullptr) {
        const float c = __fmul_rn(z, w);
        *zc = (c == 0.f) ? 1000.f : c;
    }
    return w;
}

// quash_limit (:325-332): a shifted-frame pixel keeps its weight only when its cleaned depth lies
// strictly within +/- quash of the minimum of its column.  `cmin` is that column minimum.
__device__ __forceinline__ float quashed_weight(const UfArgs& a, const float* __restrict__ depth_b,
                                                const float* __restrict__ mask_b, int sy, float yf, int sx,
                                                float cmin) {
    float zc;
    const float w = band_weight(a, depth_b, mask_b, sy, yf, sx, &zc);
    const bool keep = (zc > __fsub_rn(cmin, a.quash)) && (zc < __fadd_rn(cmin, a.quash));
    return __fmul_rn(w, keep ? 1.f : 0.f);
}

constexpr int UF_COLS = 32;    // columns per CTA (one warp-width: 128 B rows of the volume)
constexpr int UF_GROUPS = 8;   // warps per CTA, each owning D/8 consecutive bins
constexpr int UF_ROWS = 32;    // image rows per CTA (one partial sum per chunk of rows)

// Kernel 0 (quash_limit only): minimum of the cleaned shifted depth along every shifted-frame column
// (torch.min over axis 0, :329; a NaN in the column makes the minimum NaN, as torch's min does).
// One CTA = 32 columns of one item, the 8 warps stride over the rows.
__global__ void __launch_bounds__(UF_COLS * UF_GROUPS) ufield_colmin_kernel(const UfArgs a) {
    __shared__ float m_s[UF_GROUPS][UF_COLS];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int x = blockIdx.x * UF_COLS + lane, b = blockIdx.y;
    const int HW = a.H * a.W;
    const float* depth_b = a.depth + (long long)b * HW;
    const float* mask_b = a.mask ? a.mask + (long long)b * HW : nullptr;
    const float fy = __ldg(a.intr + b * a.intr_bs + 4), cy = __ldg(a.intr + b * a.intr_bs + 5);
    const bool col_ok = x < a.W;
    const int sx = col_ok ? a.col_fwd[x] : -1;
    float m = INFINITY;
    if (col_ok)
        for (int ys = grp; ys < a.H; ys += UF_GROUPS) {
            float zc;
            band_weight(a, depth_b, mask_b, a.row_fwd[ys], __fdiv_rn(__fsub_rn((float)ys, cy), fy), sx, &zc);
            m = (zc < m || zc != zc) ? zc : m;
        }
    m_s[grp][lane] = m;
    __syncthreads();
    if (grp == 0 && col_ok) {
        for (int g = 1; g < UF_GROUPS; ++g) {
            const float v = m_s[g][lane];
            m = (m != m) ? m : ((v < m || v != v) ? v : m);
        }
        a.colmin[(long long)b * a.W + x] = m;
    }
}

// Kernel 1: per-pixel weights, once.  One CTA = 32 columns x 32 rows of one item.  Every thread
// evaluates both roles of a row index (as a shifted-frame row for the denominator, as an image row
// for the numerator), writes the numerator weight map, depth * weight, and the chunk's per-column
// count.  Touches E[d] only (393 KB per item).
__global__ void __launch_bounds__(UF_COLS * UF_GROUPS) ufield_weights_kernel(const UfArgs a) {
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
    const int sx_d = col_ok ? a.col_fwd[x] : -1;              // source column of shifted column x
    const int sx_n = (xi >= 0) ? a.col_fwd[xi] : -1;          // ... of the numerator's shifted column
    // per row of the chunk: [0] shifted row r itself (denominator), [1] shifted row row_inv[r]
    __shared__ int sy_s[2][UF_ROWS];
    __shared__ float yf_s[2][UF_ROWS];
    __shared__ int yi_s[UF_ROWS];
    if (threadIdx.x < 2 * UF_ROWS) {
        const int which = threadIdx.x / UF_ROWS, rr = threadIdx.x - which * UF_ROWS;
        const int r = r0 + rr;
        int ys = -1;
        if (r < a.H) ys = which ? a.row_inv[r] : r;
        if (which) yi_s[rr] = ys;
        sy_s[which][rr] = (ys >= 0) ? a.row_fwd[ys] : -1;
        yf_s[which][rr] = __fdiv_rn(__fsub_rn((float)ys, cy), fy);
    }
    __syncthreads();
    const bool quash = a.quash > 0.f;
    const float cmin_d = (quash && col_ok) ? a.colmin[(long long)b * a.W + x] : 0.f;
    const float cmin_n = (quash && xi >= 0) ? a.colmin[(long long)b * a.W + xi] : 0.f;
    for (int rr = grp; rr < UF_ROWS; rr += UF_GROUPS) {
        const int r = r0 + rr;
        float wz = 0.f, w = 0.f;
        if (col_ok && r < a.H) {
            if (quash) {
                wz = quashed_weight(a, depth_b, mask_b, sy_s[0][rr], yf_s[0][rr], sx_d, cmin_d);
                if (yi_s[rr] >= 0 && xi >= 0)
                    w = quashed_weight(a, depth_b, mask_b, sy_s[1][rr], yf_s[1][rr], sx_n, cmin_n);
            } else {
                wz = band_weight(a, depth_b, mask_b, sy_s[0][rr], yf_s[0][rr], sx_d);
                if (yi_s[rr] >= 0 && xi >= 0) w = band_weight(a, depth_b, mask_b, sy_s[1][rr], yf_s[1][rr], sx_n);
            }
            const long long pix = (long long)b * HW + r * a.W + x;
            a.wmap[pix] = w;
            if (a.depth_zero != nullptr) a.depth_zero[pix] = __fmul_rn(__ldg(depth_b + r * a.W + x), w);
        }
        z_s[rr][lane] = wz;
    }
    __syncthreads();
    if (grp == 0 && col_ok) {
        float cnt = 0.f;
        for (int rr = 0; rr < UF_ROWS; ++rr) cnt = __fadd_rn(cnt, z_s[rr][lane]);
        a.cnt[((long long)b * a.nchunk + chunk) * a.W + x] = cnt;
    }
}

// Kernel 2: masked column sums.  One CTA = a 32 x 32 pixel tile x 16 bins; warp g owns bins
// [kb, kb + DB) and accumulates them in registers over the rows whose weight is non-zero.  Tiles
// off the road band do no volume traffic at all.
template <int DB, int NB>
__global__ void __launch_bounds__(UF_COLS * UF_GROUPS) ufield_partial_kernel(const UfArgs a) {
    __shared__ float w_s[UF_ROWS][UF_COLS];     // numerator weight of image pixel (r, x)
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int x = blockIdx.x * UF_COLS + lane;
    // blockIdx.z = item * nbg + bin group: the CTA owns UF_GROUPS * DB consecutive bins
    const int nbg = (a.D + UF_GROUPS * DB - 1) / (UF_GROUPS * DB);
    const int chunk = blockIdx.y, b = blockIdx.z / nbg, bg = blockIdx.z - b * nbg;
    const int HW = a.H * a.W;
    const int r0 = chunk * UF_ROWS;
    const bool col_ok = x < a.W;
    // tile-level skip only: a tile with any pixel on the band streams all of its rows (off-band rows
    // are multiplied by a zero weight), which keeps the inner loop free of bookkeeping.
    float wcol[UF_ROWS / UF_GROUPS];
    bool mine = false;
#pragma unroll
    for (int i = 0; i < UF_ROWS / UF_GROUPS; ++i) {
        const int rr = grp + i * UF_GROUPS, r = r0 + rr;
        wcol[i] = (col_ok && r < a.H) ? __ldg(a.wmap + (long long)b * HW + r * a.W + x) : 0.f;
        w_s[rr][lane] = wcol[i];
        mine |= wcol[i] != 0.f;
    }
    const int any = __syncthreads_or(mine ? 1 : 0);
    const int kb = (bg * UF_GROUPS + grp) * DB;
    if (kb >= a.D) return;
    float acc[DB];
#pragma unroll
    for (int j = 0; j < DB; ++j) acc[j] = 0.f;
    if (any) {
        const float* dpv_b = a.dpv + ((long long)b * a.D + kb) * HW + (long long)r0 * a.W + (col_ok ? x : 0);
        const int rows = min(UF_ROWS, a.H - r0);
#pragma unroll
        for (int g0 = 0; g0 < UF_ROWS; g0 += NB) {
            float v[NB][DB];
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                const float* col = dpv_b + (g0 + i) * a.W;
#pragma unroll
                for (int j = 0; j < DB; ++j)
                    v[i][j] = (g0 + i < rows && kb + j < a.D) ? ld_stream(col + (long long)j * HW)
                                                              : ((a.mode == DPV_IN_PROB) ? 0.f : -INFINITY);
            }
            // rows are added in increasing order; a zero weight contributes +0
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                const float w = w_s[g0 + i][lane];
#pragma unroll
                for (int j = 0; j < DB; ++j) {
                    const float pr = (a.mode == DPV_IN_PROB) ? v[i][j] : __expf(v[i][j]);
                    acc[j] = __fadd_rn(acc[j], __fmul_rn(pr, w));
                }
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
    // partial sums [B,chunks,D,W] + counts [B,chunks,W] + weight map [B,H,W] + column minima [B,W]
    return (int64_t)B * ufield_row_chunks(H) * (D + 1) * W + (int64_t)B * H * W + (int64_t)B * W;
}

extern "C" int dpv_ufield(const float* dpv, const float* depth, const float* d_candi,
                          const float* intr_up, const float* mask, const int* row_fwd,
                          const int* row_inv, const int* col_fwd, const int* col_inv, float* uf,
                          float* depth_zero, float* workspace, int B, int D, int H, int W,
                          int64_t intr_bstride, int in_mode, float zstart, float zend, float maxd,
                          float mind, float pad_depth, float quash_range, void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(dpv && depth && d_candi && intr_up && row_fwd && row_inv && col_fwd && col_inv);
    DPV_CHECK_ARG(quash_range >= 0.f);
    DPV_CHECK_ARG(uf && workspace);
    DPV_CHECK_ARG(B > 0 && D > 0 && H > 0 && W > 0);
    DPV_CHECK_ARG(in_mode == DPV_IN_LOGPROB || in_mode == DPV_IN_PROB);
    if (B > 65535 || D > 65535) return DPV_E_UNSUPP;
    UfArgs a;
    a.dpv = dpv; a.depth = depth; a.d = d_candi; a.intr = intr_up; a.mask = mask;
    a.row_fwd = row_fwd; a.row_inv = row_inv; a.col_fwd = col_fwd; a.col_inv = col_inv;
    a.uf = uf; a.depth_zero = depth_zero;
    a.B = B; a.D = D; a.H = H; a.W = W; a.mode = in_mode;
    a.nchunk = ufield_row_chunks(H);
    a.rows_per_chunk = UF_ROWS;
    a.part = workspace;
    a.cnt = workspace + (long long)B * a.nchunk * D * W;
    a.wmap = a.cnt + (long long)B * a.nchunk * W;
    a.colmin = a.wmap + (long long)B * H * W;
    a.quash = quash_range;
    a.intr_bs = intr_bstride;
    a.zstart = zstart; a.zend = zend; a.maxd1 = maxd - 1.0f; a.mind = mind;
    a.pad_depth = pad_depth;
    cudaStream_t st = (cudaStream_t)stream;
    // Every warp owns 4 bins of a 32 x 32 pixel tile and has 16 rows x 4 bins of loads in flight;
    // only the tiles on the road band do any volume traffic (about a third on KITTI-like frames).
    constexpr int DB = 4, NB = 16;
    const int nbg = (D + UF_GROUPS * DB - 1) / (UF_GROUPS * DB);
    if ((long long)B * nbg > 65535) return DPV_E_UNSUPP;
    dim3 grid0((W + UF_COLS - 1) / UF_COLS, a.nchunk, B), block(UF_COLS * UF_GROUPS);
    if (quash_range > 0.f) {
        dim3 gridm((W + UF_COLS - 1) / UF_COLS, B);
        ufield_colmin_kernel<<<gridm, block, 0, st>>>(a);
        DPV_LAUNCH_END();
    }
    ufield_weights_kernel<<<grid0, block, 0, st>>>(a);
    DPV_LAUNCH_END();
    dim3 grid((W + UF_COLS - 1) / UF_COLS, a.nchunk, B * nbg);
    ufield_partial_kernel<DB, NB><<<grid, block, 0, st>>>(a);
    DPV_LAUNCH_END();
    dim3 grid2((W + 127) / 128, D, B), block2(128);
    ufield_finish_kernel<<<grid2, block2, 0, st>>>(a);
    DPV_LAUNCH_END();
    return 0;
}
