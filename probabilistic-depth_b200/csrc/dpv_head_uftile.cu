// K3 + K5 fused, tile form: the scalar head kernel (dpv_head.cu: one thread per pixel, all D bins in
// registers, the volume read once and written once) with gen_ufield's column sums taken while the
// probabilities are still in registers.
//
// Why a second fused kernel: dpv_head followed by dpv_ufield costs 0.075 + 0.040 ms per batch of
// 8 x 256 x 384 (three UF launches that re-read the road-band third of the volume), and the persistent
// TMA-fed kernel (dpv_head_stream.cu) fuses them at 0.110 ms -- its per-row read-modify-write of the
// running sums in global memory and its one-wave schedule cost what the fusion saves.  Here a CTA owns
// a tile of 32 columns x 8 rows; warp w takes rows w and w + 4 of the tile, so the four warps hold
// rows of the SAME 32 columns and their partial column sums meet in shared memory:
//   * per pixel: exactly dpv_head's arithmetic and outputs (log-softmax, E[d], Var, arg-max, 1/4
//     hand-off), then the numerator / denominator weights of gen_ufield in closed form (the two
//     nearest-neighbour shifts cancel; dpv_uf_fused_tables carries the reference's border behaviour);
//   * a warp with a pixel on the road band adds p_k * w into its [D][32] slab of shared memory
//     (conflict-free: lane = column); rows off the band cost one vote;
//   * the CTA adds the four slabs in warp order and writes one partial sum per (tile, bin, column) and a
//     flag; tiles off the band write nothing;
//   * head_uf_tile_finish_kernel adds the flagged tiles of a column in row order and divides: no float
//     atomics, bit-reproducible.
// 3072 short CTAs per batch (4.2 waves of 740 resident) instead of one persistent wave.  (4-row tiles,
// one row per warp, measured the same: 0.094 vs 0.093 ms; packed fp32x2 arithmetic for the soft-max and
// the moments cut the instruction count by a quarter and changed nothing: the kernel waits on memory.)
#include "dpv_common.cuh"

namespace dpv {

constexpr int UT_NT = 128, UT_NW = 4;
constexpr int UT_ROWS = 8;            // rows per tile (two per warp)
constexpr int UT_MAXCH = 1024;        // tiles per column strip (H <= 8192)
constexpr float kUtL2e = 1.4426950408889634f;
constexpr float kUtLn2 = 0.6931471805599453f;

struct UtArgs {
    const float* x; const float* d;
    float* logp; float* depth; float* var; long long* argmax; float* quarter;
    const int4* row_tab; const int* col_tab; const float* intr;
    float* depth_zero; float* part; float* cnt; int* flag; float* uf;
    int B, H, W, nchunk, xtiles;
    long long intr_bs;
    float zstart, zend, maxd1, mind, pad_depth;
};

__device__ __forceinline__ float ut_ex2(float t) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    return r;
}
// utils/img_utils.py:316 -- comparisons kept negated so that NaN passes, as in the reference.
__device__ __forceinline__ float ut_band(const UtArgs& a, float z, float yf) {
    const float yy = __fmul_rn(yf, z);
    const bool out = (yy > a.zend) || (yy < a.zstart) || (z > a.maxd1) || (z < a.mind);
    return out ? 0.f : 1.f;
}
__device__ __forceinline__ float ut_yf(int ys, float fy, float cy) {
    return __fdiv_rn(__fsub_rn((float)ys, cy), fy);
}

// log p / p per bin, E[d], arg-max; log p (and the 1/4 hand-off on kept rows) stored.  Same arithmetic
// as head_kernel's main pass (dpv_head.cu).  WITH_Q is warp-uniform.
template <int D, int MODE, bool LOGP, bool WITH_Q>
__device__ __forceinline__ void ut_main_pass(float (&v)[D], const volatile float* d_v, float* lp_ptr, int HW,
                                             bool live, float ln_s, float log2_s, float top, float* qp, int q4,
                                             bool q_keep, float& mean_out, int& best_out) {
    float mean = 0.f;
    int best_k = 1 << 30;
#pragma unroll
    for (int kk = 0; kk < D; ++kk) {
        const int k = D - 1 - kk;     // last bin first: the smallest index among equal maxima remains
        float lp, pr;
        if (MODE == DPV_IN_LOGPROB) { lp = v[k]; pr = ut_ex2(lp * kUtL2e); }
        else { lp = fmaf(v[k], kUtLn2, -ln_s); pr = ut_ex2(v[k] - log2_s); }
        best_k = (lp == top) ? k : best_k;
        mean = fmaf(d_v[k], pr, mean);
        v[k] = pr;
        if (LOGP) {
            if (live) st_stream(lp_ptr, lp);
            lp_ptr -= HW;
            asm volatile("" : "+l"(lp_ptr));
        }
        if (WITH_Q) {
            if (q_keep) *qp = lp;
            qp -= q4;
            asm volatile("" : "+l"(qp));
        }
    }
    mean_out = mean;
    best_out = best_k;
}

template <int D, int MODE, bool LOGP>
__global__ void __launch_bounds__(UT_NT, 5) head_uf_tile_kernel(const UtArgs a) {
    __shared__ float acc_s[UT_NW][D][32];
    __shared__ float d_s[D];
    __shared__ float cnt_s[UT_NW][32];
    __shared__ int any_s[UT_NW], has_s[UT_NW];
    const volatile float* d_v = d_s;      // keeps the bin depths in shared memory, not in registers
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int HW = a.H * a.W;
    const int chunk = blockIdx.y, b = blockIdx.z;
    const int x_raw = blockIdx.x * 32 + lane;
    const bool live = x_raw < a.W;
    const int x = live ? x_raw : a.W - 1;     // dead lanes shadow the last column (never stored)
    pdl_trigger();    // (PDL) the finish kernel may become resident behind our last CTAs
    for (int k = tid; k < D; k += UT_NT) d_s[k] = __ldg(a.d + k);
    __syncthreads();
    const float fy = __ldg(a.intr + b * a.intr_bs + 4), cy = __ldg(a.intr + b * a.intr_bs + 5);
    const int ct = live ? __ldg(a.col_tab + x) : 0;
    const int h4 = a.H / 4, w4 = a.W / 4, q4 = h4 * w4;
    const long long item = (long long)b * D * HW;
    float cnt = 0.f;
    bool any_band = false, has_acc = false;

    for (int r = warp; r < UT_ROWS; r += UT_NW) {
        const int y = chunk * UT_ROWS + r;
        if (y >= a.H) break;                                     // warp-uniform
        const int pix = y * a.W + x;
        float v[D];
        {
            const float* px = a.x + item + pix;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                v[k] = ld_stream(px);
                px += HW;
                asm volatile("" : "+l"(px));
            }
            // The warp's next row goes to L2 while this one is in flight and being worked on: a warp holds
            // its 64 loads in registers, so bytes in flight are bounded by occupancy (20 warps x 8 KB per SM,
            // and only while a warp is in its load phase); a prefetch holds nothing.  Lane l asks for the 128-byte
            // lines of bins l, l + 32, ... of the tile's 32 columns.  (Measured: 0.0908 -> 0.0898 ms.)
            if (r + UT_NW < UT_ROWS && y + UT_NW < a.H) {
                const float* pn = a.x + item + (long long)lane * HW + (y + UT_NW) * a.W + blockIdx.x * 32;
#pragma unroll
                for (int k = 0; k < D; k += 32) {
                    if (k + lane < D) asm volatile("prefetch.global.L2 [%0];" ::"l"(pn));
                    pn += 32LL * HW;
                }
            }
        }
        float ln_s = 0.f, log2_s = 0.f, top;
        if (MODE == DPV_IN_LOGITS) {
            float m = v[0];
#pragma unroll
            for (int k = 1; k < D; ++k) m = fmaxf(m, v[k]);
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                v[k] = (v[k] - m) * kUtL2e;
                s += ut_ex2(v[k]);
            }
            ln_s = logf(s);
            log2_s = ln_s * kUtL2e;
            top = -ln_s;
        } else {
            top = v[0];
#pragma unroll
            for (int k = 1; k < D; ++k) top = fmaxf(top, v[k]);
        }
        float mean;
        int best_k;
        float* lp_ptr = LOGP ? a.logp + item + (long long)(D - 1) * HW + pix : nullptr;
        // one code path per launch (the fully unrolled pass is ~1000 instructions: a second copy for the
        // rows without hand-off costs more in instruction fetch than its predicated-off stores)
        // (one code path for all warps: the hand-off rows are exactly the rows of warp 0, but giving warps 1-3 a
        // copy of the pass without the hand-off's pointer arithmetic measured 0.0939 vs 0.0896 ms -- the second
        // ~1000-instruction copy costs more in instruction fetch than it saves in issue slots)
        if (a.quarter != nullptr) {
            const bool q_keep = live && ((y & 3) == 0) && ((y >> 2) < h4) && ((x & 3) == 0) && ((x >> 2) < w4);
            float* qp = a.quarter + ((long long)b * D + (D - 1)) * q4 + (q_keep ? (y >> 2) * w4 + (x >> 2) : 0);
            ut_main_pass<D, MODE, LOGP, true>(v, d_v, lp_ptr, HW, live, ln_s, log2_s, top, qp, q4, q_keep, mean, best_k);
        } else {
            ut_main_pass<D, MODE, LOGP, false>(v, d_v, lp_ptr, HW, live, ln_s, log2_s, top, nullptr, 0, false, mean,
                                               best_k);
        }
        const long long opix = (long long)b * HW + pix;
        if (a.var != nullptr) {
            float var = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const float c = d_v[k] - mean;
                var = fmaf(c * c, v[k], var);
            }
            if (live) a.var[opix] = var;
        }
        if (live) {
            if (a.depth != nullptr) a.depth[opix] = mean;
            if (a.argmax != nullptr) a.argmax[opix] = (long long)(best_k == (1 << 30) ? 0 : best_k);
        }

        // ---- uncertainty field: weights of this pixel (same closed form as head_stream_kernel) -------
        const int4 rt = __ldg(a.row_tab + y);
        const float zn = (rt.y | ((ct >> 1) & 1)) ? a.pad_depth : mean;
        const float wn = (live && rt.x >= 0 && (ct & 1)) ? ut_band(a, zn, ut_yf(rt.x, fy, cy)) : 0.f;
        const float wd = (live && rt.z >= 0 && (ct & 4)) ? ut_band(a, mean, ut_yf(rt.z, fy, cy)) : 0.f;
        if (live && a.depth_zero != nullptr) a.depth_zero[opix] = __fmul_rn(mean, wn);
        any_band |= (wn != 0.f) | (wd != 0.f);
        cnt = __fadd_rn(cnt, wd);
        if (__any_sync(0xffffffffu, wn != 0.f)) {
            float* acc = &acc_s[warp][0][lane];
            if (UT_ROWS > UT_NW && has_acc) {      // (more than one row per warp only)
#pragma unroll
                for (int k = 0; k < D; ++k) acc[k * 32] = __fadd_rn(acc[k * 32], __fmul_rn(v[k], wn));
            } else {
#pragma unroll
                for (int k = 0; k < D; ++k) acc[k * 32] = __fmul_rn(v[k], wn);
                has_acc = true;
            }
        }
    }
    cnt_s[warp][lane] = cnt;
    const bool wany = __any_sync(0xffffffffu, any_band) != 0;
    if (lane == 0) { any_s[warp] = wany ? 1 : 0; has_s[warp] = has_acc ? 1 : 0; }
    __syncthreads();
    const bool tile_any = (any_s[0] | any_s[1] | any_s[2] | any_s[3]) != 0;
    const long long tile = (long long)b * a.nchunk + chunk;
    if (tid == 0) a.flag[tile * a.xtiles + blockIdx.x] = tile_any ? 1 : 0;
    if (!tile_any || !live) return;
    for (int k = warp; k < D; k += UT_NW) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < UT_NW; ++w)
            if (has_s[w]) s = __fadd_rn(s, acc_s[w][k][lane]);
        a.part[(tile * D + k) * a.W + x] = s;
    }
    if (warp == 0)
        a.cnt[tile * a.W + x] = __fadd_rn(__fadd_rn(__fadd_rn(cnt_s[0][lane], cnt_s[1][lane]), cnt_s[2][lane]),
                                          cnt_s[3][lane]);
}

// UF[b,k,x] = sum over the flagged tiles of the column (top to bottom) of their partial sums, divided by
// (their counts + the count of the shifted-frame pixels that sample the zero padding).  0/0 = NaN as in
// the reference.  block = 32 columns x 8 bins of one item (768 blocks at the model's shape: one wave).
// The flagged tiles of the strip are compacted into a list (uniform over the block); the partial sums of 16
// tiles are in flight per thread and are added in row order.  Launched with programmatic stream
// serialisation: what does not read partial sums (the padding counts) runs while the tile kernel drains.
template <int D>
__global__ void __launch_bounds__(256) head_uf_tile_finish_kernel(const UtArgs a) {
    __shared__ float den_s[32];
    __shared__ float r0_s[8], r1_s[8];
    __shared__ short list_s[UT_MAXCH];
    __shared__ int nlist_s;
    const int tid = threadIdx.x, c = tid & 31, g = tid >> 5;
    const int x = blockIdx.x * 32 + c, b = blockIdx.z;
    const int k = blockIdx.y * 8 + g;
    const bool ok = x < a.W;
    const int xe = ok ? x : 0;
    // Everything up to pdl_wait() reads only what the host / earlier kernels wrote (intrinsics, tables): it
    // runs while the tile kernel is still draining.
    const float fy = __ldg(a.intr + b * a.intr_bs + 4), cy = __ldg(a.intr + b * a.intr_bs + 5);
    // shifted-frame rows whose pixels sample the zero padding (E[d] = pad_depth there): sums of 0 / 1,
    // exact in any order.  p0: rows that are padding themselves; p1: every row (padded columns).
    float p0 = 0.f, p1 = 0.f;
    for (int ys = tid; ys < a.H; ys += 256) {
        const float w = ut_band(a, a.pad_depth, ut_yf(ys, fy, cy));
        p1 += w;
        if (__ldg(a.row_tab + ys).w) p0 += w;
    }
    p0 = warp_sum(p0); p1 = warp_sum(p1);
    if (c == 0) { r0_s[g] = p0; r1_s[g] = p1; }
    const bool colpad = (__ldg(a.col_tab + xe) >> 3) & 1;
    pdl_wait();       // (PDL) the tile kernel has completed: its partial sums and flags are visible
    // flagged tiles of this column strip, top to bottom, as a compact list (warp 0; ballot + prefix count)
    if (g == 0) {
        int n = 0;
        for (int ch0 = 0; ch0 < a.nchunk; ch0 += 32) {
            const int ch = ch0 + c;
            const bool on = ch < a.nchunk &&
                            __ldcg(a.flag + ((long long)b * a.nchunk + ch) * a.xtiles + blockIdx.x) != 0;
            const unsigned m = __ballot_sync(0xffffffffu, on);
            if (on) list_s[n + __popc(m & ((1u << c) - 1u))] = (short)ch;
            n += __popc(m);
        }
        if (c == 0) nlist_s = n;
    }
    __syncthreads();
    const int nlist = nlist_s;
    if (g == 0) {
        float den = 0.f;
        for (int i = 0; i < 8; ++i) den += colpad ? r1_s[i] : r0_s[i];
        for (int i0 = 0; i0 < nlist; i0 += 8) {
            float cv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                cv[j] = (ok && i0 + j < nlist)
                            ? __ldcg(a.cnt + ((long long)b * a.nchunk + list_s[i0 + j]) * a.W + x) : 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) den = __fadd_rn(den, cv[j]);
        }
        den_s[c] = den;
    }
    float num = 0.f;
    if (ok && k < D) {
        for (int i0 = 0; i0 < nlist; i0 += 16) {
            float pv[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
                pv[j] = (i0 + j < nlist)
                            ? __ldcg(a.part + (((long long)b * a.nchunk + list_s[i0 + j]) * D + k) * a.W + x) : 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) num = __fadd_rn(num, pv[j]);
        }
    }
    __syncthreads();
    if (ok && k < D) a.uf[((long long)b * D + k) * a.W + x] = __fdiv_rn(num, den_s[c]);
}

long long head_uf_tile_workspace_floats(int B, int D, int H, int W) {
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
    const long long nchunk = (H + UT_ROWS - 1) / UT_ROWS, xt = (W + 31) / 32;
    return (long long)B * nchunk * ((long long)D * W + W + xt) + 8;    // part, cnt, flag (int32)
}

template <int D>
static int ut_launch(const UtArgs& a, int mode, cudaStream_t st) {
    dim3 grid(a.xtiles, a.nchunk, a.B), block(UT_NT);
    const bool lp = a.logp != nullptr;
    cudaError_t e;
    if (mode == DPV_IN_LOGITS) {
        if (lp) head_uf_tile_kernel<D, DPV_IN_LOGITS, true><<<grid, block, 0, st>>>(a);
        else head_uf_tile_kernel<D, DPV_IN_LOGITS, false><<<grid, block, 0, st>>>(a);
    } else {
        if (lp) head_uf_tile_kernel<D, DPV_IN_LOGPROB, true><<<grid, block, 0, st>>>(a);
        else head_uf_tile_kernel<D, DPV_IN_LOGPROB, false><<<grid, block, 0, st>>>(a);
    }
    DPV_LAUNCH_END();
    static const int nofinish = [] { const char* v = getenv("DPV_UFTILE_NOFINISH"); return v ? atoi(v) : 0; }();
    if (nofinish) return 0;      // (timing experiments only: the UF is not produced)
    // (PDL) the finish kernel becomes resident behind the tile kernel's last CTAs and waits for them
    e = dpv_launch_pdl(head_uf_tile_finish_kernel<D>, dim3(a.xtiles, (D + 7) / 8, a.B), dim3(256), 0, st, a);
    if (e != cudaSuccess) return (int)e;
    DPV_LAUNCH_END();
    return 0;
}

// dpv_head_ufield through the tile kernel; DPV_E_UNSUPP = shape not handled (take the stream kernel).
int launch_head_uf_tile(const float* x, const float* d, float* logp, float* depth, float* var, long long* argmax,
                        float* quarter, const float* intr, const int* row_tab, const int* col_tab, float* uf,
                        float* depth_zero, float* workspace, int B, int D, int H, int W, long long intr_bs,
                        int mode, float zstart, float zend, float maxd, float mind, float pad_depth,
                        cudaStream_t st) {
    if (D != 32 && D != 64) return DPV_E_UNSUPP;
    if (mode != DPV_IN_LOGITS && mode != DPV_IN_LOGPROB) return DPV_E_UNSUPP;
    if (reinterpret_cast<uintptr_t>(row_tab) & 15) return DPV_E_BADARG;
    UtArgs a = {};
    a.x = x; a.d = d; a.logp = logp; a.depth = depth; a.var = var; a.argmax = argmax; a.quarter = quarter;
    a.row_tab = reinterpret_cast<const int4*>(row_tab); a.col_tab = col_tab; a.intr = intr;
    a.depth_zero = depth_zero; a.uf = uf;
    a.B = B; a.H = H; a.W = W;
    a.nchunk = (H + UT_ROWS - 1) / UT_ROWS;
    a.xtiles = (W + 31) / 32;
    if (B > 65535 || a.nchunk > UT_MAXCH) return DPV_E_UNSUPP;
    a.part = workspace;
    a.cnt = a.part + (long long)B * a.nchunk * D * W;
    a.flag = reinterpret_cast<int*>(a.cnt + (long long)B * a.nchunk * W);
    a.intr_bs = intr_bs;
    a.zstart = zstart; a.zend = zend; a.maxd1 = maxd - 1.0f; a.mind = mind; a.pad_depth = pad_depth;
    return D == 64 ? ut_launch<64>(a, mode, st) : ut_launch<32>(a, mode, st);
}

}  // namespace dpv
