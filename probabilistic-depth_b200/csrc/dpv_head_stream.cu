// K3 (+K5): persistent, bulk-copy staged depth-bin head, optionally fused with the UF collapse.
//
// log-softmax over the bin axis fused with everything the reference derives from it in separate
// passes (E[d], Var[d], arg-max, the 1/4-resolution hand-off; see dpv_head.cu for the sites).
// A thread owns ONE image column of a 128-column strip and keeps the D (<= 64) bins of the current
// pixel in registers: no shuffles, no replicated per-pixel work.  The CTA walks down its run of rows;
// the grid is sized to the machine and the rows of all strips are split evenly over the CTAs.
//
// Data movement is done by the bulk-copy engine (cp.async.bulk, the non-tensor TMA path) instead
// of per-thread loads and stores: one [D][128] tile = D row segments of 512 contiguous bytes, each
// moved by one instruction of one thread, completion signalled on an mbarrier; two tiles are in
// flight per CTA (the next row lands while the current one is computed); log p is written back
// into the tile in place and leaves through cp.async.bulk as well.  The compute threads only see
// shared memory at compile-time offsets.  Why: ncu on the first two kernels (profiles/) showed
// 38-43 % of all issued instructions were global address arithmetic (LEA/IADD3/IMAD) and the
// kernel issue-bound at 73 % issue utilisation, 81 % of the measured HBM roof.
//
// UF = true additionally performs gen_ufield (reference utils/img_utils.py:268-358) on the
// probabilities while they are in registers.  The reference's two nearest-neighbour row shifts
// cancel: for image pixel (y, x) the numerator weight is band(E[d](y, x), shifted row) times border
// predicates (closed form verified to reproduce gen_ufield exactly, NaN pattern
// included); the predicates come from the reference's own grid construction through two small
// tables (dpv_uf_fused_tables).  Each thread accumulates its column over the CTA's run of rows in a
// column-private shared-memory slot (no barrier, no atomics; rows off the road band cost nothing),
// the CTA writes one partial record per (run, strip) if any pixel was on the band, and
// uf_stream_finish_kernel adds the records of a column in run order: bit-reproducible.
#include <algorithm>
#include <cstdlib>

#include "dpv_common.cuh"

namespace dpv {

constexpr int HS_NT = 128;            // threads per CTA = columns per strip
constexpr int HS_NW = HS_NT / 32;     // warps per CTA

struct HeadStreamArgs {
    const float* x; const float* d;
    float* logp; float* depth; float* var; long long* argmax; float* quarter;
    // fused uncertainty field
    const int4* row_tab; const int* col_tab; const float* intr;
    float* depth_zero; float* rec; int* flag;
    int B, H, W, S2, nseg;
    long long units;                  // B * S2 * H rows of strips
    long long intr_bs, rec_floats;
    float zstart, zend, maxd1, mind, pad_depth;
};

__device__ __forceinline__ float hs_ex2(float t) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    return r;
}

// ---- bulk-copy (TMA engine, non-tensor form) and mbarrier primitives ---------------------------
__device__ __forceinline__ unsigned hs_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void hs_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(hs_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void hs_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(hs_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hs_mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n.reg .pred p;\n"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                     "selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(hs_smem(bar)), "r"(parity) : "memory");
    } while (!ok);
}
// global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void hs_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(hs_smem(dst)), "l"(src), "r"(bytes), "r"(hs_smem(bar)) : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void hs_bulk_s2g(void* dst, const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"(hs_smem(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hs_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void hs_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void hs_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void hs_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr float kHsL2e = 1.4426950408889634f;
constexpr float kHsLn2 = 0.6931471805599453f;

// utils/img_utils.py:316 -- comparisons kept negated so that NaN passes, as in the reference.
// yf = (shifted row - cy) / fy is row-uniform.
__device__ __forceinline__ float hs_band(const HeadStreamArgs& a, float z, float yf) {
    const float yy = __fmul_rn(yf, z);
    const bool out = (yy > a.zend) || (yy < a.zstart) || (z > a.maxd1) || (z < a.mind);
    return out ? 0.f : 1.f;
}
__device__ __forceinline__ float hs_yf(int ys, float fy, float cy) {
    return __fdiv_rn(__fsub_rn((float)ys, cy), fy);
}

// first unit of CTA c when `units` rows are split over G CTAs (contiguous, sizes differ by <= 1)
__host__ __device__ __forceinline__ long long hs_first_unit(long long c, long long units, long long G) {
    return (c * units) / G;
}

// Second half of a row: log p / p per bin, E[d], arg-max; log p goes back into the tile in place.
// WITH_Q is CTA-uniform (rows y % 4 == 0 feed the 1/4-resolution hand-off).
template <int D, int MODE, bool LOGP, bool WITH_Q>
__device__ __forceinline__ void hs_main_pass(float (&v)[D], const float* d_s, float* tile, float ln_s,
                                             float log2_s, float top, float* qp, int q4, bool q_keep,
                                             float& mean_out, int& best_out) {
    float mean0 = 0.f, mean1 = 0.f;
    int best_k = 1 << 30;
    // last bin first so that the smallest index among equal maxima is what remains
#pragma unroll
    for (int kk = 0; kk < D; ++kk) {
        const int k = D - 1 - kk;
        float lp, pr;
        if (MODE == DPV_IN_LOGPROB) { lp = v[k]; pr = hs_ex2(lp * kHsL2e); }
        else { lp = fmaf(v[k], kHsLn2, -ln_s); pr = hs_ex2(v[k] - log2_s); }
        best_k = (lp == top) ? k : best_k;
        if (kk & 1) mean1 = fmaf(d_s[k], pr, mean1); else mean0 = fmaf(d_s[k], pr, mean0);
        v[k] = pr;
        if (LOGP && MODE != DPV_IN_LOGPROB) tile[k * HS_NT] = lp;     // immediate-offset STS
        if (WITH_Q) {
            if (q_keep) *qp = lp;
            qp -= q4;
        }
    }
    mean_out = mean0 + mean1;
    best_out = best_k;
}

// One thread = one image column of a 128-column strip; the CTA walks down its run of rows.  Tiles
// ([D][128] floats = one row of the strip, all bins) are brought in by cp.async.bulk into a 2-stage
// ring, transformed in place and written back by cp.async.bulk: the SM's instruction stream has no
// global address arithmetic at all (the first kernels spent 38-43 % of their issue slots on it).
template <int D, int MODE, bool LOGP, bool UF>
__global__ void __launch_bounds__(HS_NT) head_stream_kernel(const HeadStreamArgs a) {
    extern __shared__ __align__(128) unsigned char hs_smem_raw[];
    float* stage = reinterpret_cast<float*>(hs_smem_raw);                 // [2][D][128]
    float* acc_s = stage + 2 * D * HS_NT;                                 // [D][128] (UF only)
    float* d_s = acc_s + (UF ? D * HS_NT : 0);                            // [D]
    unsigned long long* full = reinterpret_cast<unsigned long long*>(d_s + D);   // [2]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int HW = a.H * a.W;
    const long long u0 = hs_first_unit(blockIdx.x, a.units, gridDim.x);
    const long long u1 = hs_first_unit(blockIdx.x + 1, a.units, gridDim.x);
    for (int k = tid; k < D; k += HS_NT) d_s[k] = __ldg(a.d + k);
    if (UF) {
#pragma unroll 8
        for (int k = 0; k < D; ++k) acc_s[k * HS_NT + tid] = 0.f;
    }
    if (tid == 0) {
        hs_mbar_init(&full[0], 1);
        hs_mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (u0 >= u1) return;

    long long bs = u0 / a.H;                     // (item, strip)
    int y = (int)(u0 - bs * a.H);
    int b = (int)(bs / a.S2);
    int ws = (int)(bs - (long long)b * a.S2);
    int x0 = ws * HS_NT;
    int cw = min(HS_NT, a.W - x0);               // live columns of this strip (multiple of 4)
    // Bulk copies are uniform-datapath instructions: issuing them from every lane makes the compiler
    // serialise the warp lane by lane.  Lane 0 of each warp moves D/4 bins' 512-byte row segments.
    constexpr int KPW = D / HS_NW;               // bins per issuing warp
    const bool issuer = (lane == 0);
    const int k_lo = warp * KPW;

    // prologue: first tile into stage 0
    if (issuer) {
        if (warp == 0) hs_mbar_expect_tx(&full[0], (unsigned)(D * cw * 4));
        const float* src = a.x + ((long long)b * D + k_lo) * HW + (long long)y * a.W + x0;
#pragma unroll 4
        for (int k = 0; k < KPW; ++k)
            hs_bulk_g2s(stage + (k_lo + k) * HS_NT, src + (long long)k * HW, (unsigned)(cw * 4), &full[0]);
    }

    float cnt = 0.f, fy = 0.f, cy = 0.f;
    bool seg_any = false;
    int seg = 0, ct = 0;
    bool live = tid < cw;
    if (UF) {
        fy = __ldg(a.intr + b * a.intr_bs + 4); cy = __ldg(a.intr + b * a.intr_bs + 5);
        ct = live ? __ldg(a.col_tab + x0 + tid) : 0;
    }
    const int h4 = a.H / 4, w4 = a.W / 4, q4 = h4 * w4;

    int it = 0;
    for (long long u = u0; u < u1; ++u, ++it) {
        const int s = it & 1;
        float* tile = stage + s * D * HS_NT + tid;
        // ---- where is the next row; bring it in ------------------------------------------------
        const bool has_next = (u + 1 < u1);
        const bool same_strip = (y + 1 < a.H);
        int nb = b, nws = ws, ny = y + 1, nx0 = x0, ncw = cw;
        if (!same_strip) {
            const long long nbs = bs + 1;
            nb = (int)(nbs / a.S2);
            nws = (int)(nbs - (long long)nb * a.S2);
            ny = 0;
            nx0 = nws * HS_NT;
            ncw = min(HS_NT, a.W - nx0);
        }
        if (has_next && issuer) {
            // stage s^1 was written back by this thread's bulk store of the previous tile: wait until
            // the engine has read it out, then overwrite
            hs_bulk_wait_read0();
            if (warp == 0) hs_mbar_expect_tx(&full[s ^ 1], (unsigned)(D * ncw * 4));
            const float* src = a.x + ((long long)nb * D + k_lo) * HW + (long long)ny * a.W + nx0;
            float* dst = stage + (s ^ 1) * D * HS_NT + k_lo * HS_NT;
#pragma unroll 4
            for (int k = 0; k < KPW; ++k)
                hs_bulk_g2s(dst + k * HS_NT, src + (long long)k * HW, (unsigned)(ncw * 4), &full[s ^ 1]);
        }

        // ---- the current tile -------------------------------------------------------------------
        hs_mbar_wait(&full[s], (unsigned)((it >> 1) & 1));
        float v[D];
#pragma unroll
        for (int k = 0; k < D; ++k) v[k] = tile[k * HS_NT];              // immediate-offset LDS
        float ln_s = 0.f, log2_s = 0.f, top;
        if (MODE == DPV_IN_LOGITS) {
            float m = v[0];
#pragma unroll
            for (int k = 1; k < D; ++k) m = fmaxf(m, v[k]);
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int k = 0; k < D; k += 2) {
                v[k] = (v[k] - m) * kHsL2e; s0 += hs_ex2(v[k]);
                v[k + 1] = (v[k + 1] - m) * kHsL2e; s1 += hs_ex2(v[k + 1]);
            }
            ln_s = logf(s0 + s1);
            log2_s = ln_s * kHsL2e;
            top = -ln_s;
        } else {
            top = v[0];
#pragma unroll
            for (int k = 1; k < D; ++k) top = fmaxf(top, v[k]);
        }

        const int x = x0 + tid;
        const int pix = y * a.W + (live ? x : x0);
        float mean;
        int best_k;
        const bool want_q = (a.quarter != nullptr) && ((y & 3) == 0) && ((y >> 2) < h4);   // CTA-uniform
        if (want_q) {
            const bool q_keep = live && ((x & 3) == 0) && ((x >> 2) < w4);
            float* qp = a.quarter + ((long long)b * D + (D - 1)) * q4 + (y >> 2) * w4 + ((live ? x : x0) >> 2);
            hs_main_pass<D, MODE, LOGP, true>(v, d_s, tile, ln_s, log2_s, top, qp, q4, q_keep, mean, best_k);
        } else {
            hs_main_pass<D, MODE, LOGP, false>(v, d_s, tile, ln_s, log2_s, top, nullptr, 0, false, mean, best_k);
        }
        if (LOGP) {
            // hand the transformed tile to the bulk-copy engine
            hs_fence_async();
            __syncthreads();
            if (issuer) {
                float* dst = a.logp + ((long long)b * D + k_lo) * HW + (long long)y * a.W + x0;
                const float* src = stage + s * D * HS_NT + k_lo * HS_NT;
#pragma unroll 4
                for (int k = 0; k < KPW; ++k)
                    hs_bulk_s2g(dst + (long long)k * HW, src + k * HS_NT, (unsigned)(cw * 4));
                hs_bulk_commit();
            }
        } else {
            __syncthreads();          // every thread is done reading stage s before it is refilled
        }

        const long long opix = (long long)b * HW + pix;
        if (a.var != nullptr) {
            float var0 = 0.f, var1 = 0.f;
#pragma unroll
            for (int k = 0; k < D; k += 2) {
                const float c0 = d_s[k] - mean, c1 = d_s[k + 1] - mean;
                var0 = fmaf(c0 * c0, v[k], var0);
                var1 = fmaf(c1 * c1, v[k + 1], var1);
            }
            if (live) a.var[opix] = var0 + var1;
        }
        if (live) {
            if (a.depth != nullptr) a.depth[opix] = mean;
            if (a.argmax != nullptr) a.argmax[opix] = (long long)(best_k == (1 << 30) ? 0 : best_k);
        }

        if (UF) {
            // ---- uncertainty field: weight of this pixel -------------------------------------------
            const int4 rt = __ldg(a.row_tab + y);
            const float zn = (rt.y | ((ct >> 1) & 1)) ? a.pad_depth : mean;
            const float wn = (live && rt.x >= 0 && (ct & 1)) ? hs_band(a, zn, hs_yf(rt.x, fy, cy)) : 0.f;
            const float wd = (live && rt.z >= 0 && (ct & 4)) ? hs_band(a, mean, hs_yf(rt.z, fy, cy)) : 0.f;
            if (live && a.depth_zero != nullptr) a.depth_zero[opix] = __fmul_rn(mean, wn);
            const bool mine = (wn != 0.f) | (wd != 0.f);
            seg_any |= (__any_sync(0xffffffffu, mine) != 0);
            if (wn != 0.f) {
#pragma unroll
                for (int k = 0; k < D; ++k)
                    acc_s[k * HS_NT + tid] = __fadd_rn(acc_s[k * HS_NT + tid], __fmul_rn(v[k], wn));
            }
            cnt = __fadd_rn(cnt, wd);
            // ---- end of this segment (run leaves the strip, or ends): write its record --------------
            if (!has_next || !same_strip) {
                const long long rid = (long long)blockIdx.x * a.nseg + seg;
                if (seg_any) {
                    // record layout: [D][128] partial sums, then [128] counts
                    float* rec = a.rec + rid * a.rec_floats;
#pragma unroll 8
                    for (int k = 0; k < D; ++k) {
                        rec[k * HS_NT + tid] = acc_s[k * HS_NT + tid];
                        acc_s[k * HS_NT + tid] = 0.f;
                    }
                    rec[D * HS_NT + tid] = cnt;
                }
                if (lane == 0) a.flag[rid * HS_NW + warp] = seg_any ? 1 : 0;
                seg_any = false;
                cnt = 0.f;
                ++seg;
            }
        }

        // ---- next row ----------------------------------------------------------------------------
        if (has_next) {
            if (!same_strip) {
                bs += 1;
                live = tid < ncw;
                if (UF) {
                    fy = __ldg(a.intr + nb * a.intr_bs + 4); cy = __ldg(a.intr + nb * a.intr_bs + 5);
                    ct = live ? __ldg(a.col_tab + nx0 + tid) : 0;
                }
            }
            b = nb; ws = nws; y = ny; x0 = nx0; cw = ncw;
        }
    }
    if (LOGP && issuer) hs_bulk_wait_all();      // shared memory must outlive the last bulk store
}

// UF[b,k,x] = sum over the runs covering the column of (partial sums) / (counts + padding rows).
// 0/0 = NaN as in the reference.  block (32 columns, 8 bins); all 32 columns of a block lie in one
// warp's strip of one wide strip (both are multiples of 32 columns wide or the block is clipped).
constexpr int HS_FIN_PIECES = 64;
template <int D>
__global__ void __launch_bounds__(256) uf_stream_finish_kernel(const HeadStreamArgs a, float* uf, int G) {
    constexpr int CWW = 32, CWC = HS_NT;
    __shared__ float den_s[8][32];
    __shared__ long long rid_s[HS_FIN_PIECES];
    __shared__ int nflag_s;
    const int c = threadIdx.x, g = threadIdx.y, tid = g * 32 + c;
    const int x = blockIdx.x * 32 + c, b = blockIdx.z;
    const int k = blockIdx.y * 8 + g;
    const bool ok = x < a.W;
    const int xe = ok ? x : 0;
    const int ws = xe / CWC, xin = xe - ws * CWC, wq = xin / CWW;
    // pieces are per strip; a 32-column block never straddles two 128-column strips
    const int ws0 = (blockIdx.x * 32) / CWC;
    const long long bs = (long long)b * a.S2 + ws0;
    const long long ufirst = bs * a.H, ulast = ufirst + a.H - 1;
    // CTA whose run holds unit u: the largest c with floor(c * units / G) <= u
    const int c_lo = (int)(((ufirst + 1) * G + a.units - 1) / a.units - 1);
    const int c_hi = (int)(((ulast + 1) * G + a.units - 1) / a.units - 1);
    const float fy = __ldg(a.intr + b * a.intr_bs + 4), cy = __ldg(a.intr + b * a.intr_bs + 5);
    // shifted-frame pixels that sample the zero padding (their E[d] is pad_depth): rows split over g
    float den = 0.f;
    {
        const bool colpad = (__ldg(a.col_tab + xe) >> 3) & 1;
        for (int ys = g; ys < a.H; ys += 8)
            if (colpad || __ldg(a.row_tab + ys).w) den += hs_band(a, a.pad_depth, hs_yf(ys, fy, cy));   // small integers: exact
    }
    float num = 0.f;
    for (int p0 = c_lo; p0 <= c_hi; p0 += HS_FIN_PIECES) {
        const int np = min(HS_FIN_PIECES, c_hi - p0 + 1);
        __syncthreads();
        if (tid == 0) nflag_s = 0;
        __syncthreads();
        // the flag of a piece is per (run, strip, warp): uniform over this block's 32 columns.  Keep
        // the flagged pieces only, in run order (ballot-compacted by warp 0).
        if (g == 0) {
            for (int i0 = 0; i0 < np; i0 += 32) {
                const int i = i0 + c;
                long long rid = 0;
                bool on = false;
                if (i < np) {
                    const int cta = p0 + i;
                    const int seg = (int)(bs - hs_first_unit(cta, a.units, G) / a.H);
                    rid = (long long)cta * a.nseg + seg;
                    on = a.flag[rid * HS_NW + wq] != 0;
                }
                const unsigned m = __ballot_sync(0xffffffffu, on);
                const int base = nflag_s;
                if (on) rid_s[base + __popc(m & ((1u << c) - 1u))] = rid;
                __syncwarp();
                if (c == 0) nflag_s = base + __popc(m);
                __syncwarp();
            }
        }
        __syncthreads();
        const int nf = nflag_s;
        if (g == 0) {
            for (int i = 0; i < nf; ++i)
                den = __fadd_rn(den, a.rec[rid_s[i] * a.rec_floats + D * CWC + xin]);
        }
        if (k < D) {
#pragma unroll 8
            for (int i = 0; i < nf; ++i)
                num = __fadd_rn(num, a.rec[rid_s[i] * a.rec_floats + k * CWC + xin]);
        }
    }
    den_s[g][c] = den;
    __syncthreads();
    if (!ok || k >= D) return;
    float dsum = den_s[0][c];
#pragma unroll
    for (int i = 1; i < 8; ++i) dsum = __fadd_rn(dsum, den_s[i][c]);
    uf[((long long)b * D + k) * a.W + x] = __fdiv_rn(num, dsum);
}

// ------------------------------------------------------------------------------------ host side
static int hs_sm_count() {
    static const int n = [] {
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
            cudaGetLastError();
            sms = 148;                       // B200; only the workspace bound depends on it off-device
        }
        return sms;
    }();
    return n;
}

struct HsPlan { int S2, G, nseg; long long units, rec_floats; size_t smem; };

static size_t hs_smem_bytes(int D, bool uf) {
    return (size_t)(2 + (uf ? 1 : 0)) * D * HS_NT * sizeof(float) + (size_t)D * sizeof(float) + 16;
}

// D <= 64: a thread holds all D bins of its pixel in registers.  W % 4 == 0: the bulk copies move
// whole 16-byte units.
static bool hs_plan(int B, int D, int H, int W, bool uf, HsPlan* p) {
    if (D != 16 && D != 32 && D != 64) return false;
    if (B <= 0 || H <= 0 || W <= 0 || (W & 3) != 0) return false;
    p->smem = hs_smem_bytes(D, uf);
    const int per_sm = (int)std::min<size_t>(8, (size_t)(227 * 1024) / (p->smem + 1024));
    p->S2 = (W + HS_NT - 1) / HS_NT;
    p->units = (long long)B * p->S2 * H;
    const long long slots = (long long)hs_sm_count() * per_sm;
    p->G = (int)(p->units < slots ? p->units : slots);
    const long long rpc = (p->units + p->G - 1) / p->G;      // rows per CTA (max)
    p->nseg = (int)((rpc + H - 1) / H + 1);
    p->rec_floats = (long long)D * HS_NT + HS_NT;
    return true;
}

template <int D, int MODE, bool LOGP, bool UF>
static int hs_launch_one(const HeadStreamArgs& a, const HsPlan& p, cudaStream_t st) {
    static bool attr_set = false;            // per instantiation; the attribute is idempotent
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(head_stream_kernel<D, MODE, LOGP, UF>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    head_stream_kernel<D, MODE, LOGP, UF><<<dim3(p.G), dim3(HS_NT), p.smem, st>>>(a);
    DPV_LAUNCH_END();
    return 0;
}

template <int D>
static int hs_launch(const HeadStreamArgs& a, const HsPlan& p, int mode, bool uf, float* uf_out,
                     cudaStream_t st) {
    const bool lp = a.logp != nullptr;
    int rc;
    if (mode == DPV_IN_LOGITS) {
        if (uf) rc = lp ? hs_launch_one<D, DPV_IN_LOGITS, true, true>(a, p, st)
                        : hs_launch_one<D, DPV_IN_LOGITS, false, true>(a, p, st);
        else rc = lp ? hs_launch_one<D, DPV_IN_LOGITS, true, false>(a, p, st)
                     : hs_launch_one<D, DPV_IN_LOGITS, false, false>(a, p, st);
    } else if (mode == DPV_IN_LOGPROB) {
        if (uf) rc = lp ? hs_launch_one<D, DPV_IN_LOGPROB, true, true>(a, p, st)
                        : hs_launch_one<D, DPV_IN_LOGPROB, false, true>(a, p, st);
        else rc = lp ? hs_launch_one<D, DPV_IN_LOGPROB, true, false>(a, p, st)
                     : hs_launch_one<D, DPV_IN_LOGPROB, false, false>(a, p, st);
    } else {
        return DPV_E_UNSUPP;
    }
    if (rc != 0 || !uf) return rc;
    dim3 g2((a.W + 31) / 32, (D + 7) / 8, a.B), b2(32, 8);
    uf_stream_finish_kernel<D><<<g2, b2, 0, st>>>(a, uf_out, p.G);
    DPV_LAUNCH_END();
    return 0;
}

static int hs_dispatch(HeadStreamArgs& a, int D, int mode, bool uf, float* uf_out, cudaStream_t st) {
    HsPlan p;
    if (!hs_plan(a.B, D, a.H, a.W, uf, &p)) return DPV_E_UNSUPP;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (!al16(a.x) || !al16(a.logp)) return DPV_E_UNSUPP;
    if (!al16(a.row_tab)) return DPV_E_BADARG;
    a.S2 = p.S2; a.nseg = p.nseg; a.units = p.units; a.rec_floats = p.rec_floats;
    switch (D) {
        case 16: return hs_launch<16>(a, p, mode, uf, uf_out, st);
        case 32: return hs_launch<32>(a, p, mode, uf, uf_out, st);
        case 64: return hs_launch<64>(a, p, mode, uf, uf_out, st);
        default: return DPV_E_UNSUPP;
    }
}

// dpv_head's scalar-kernel argument block (dpv_head.cu)
struct HeadArgs {
    const float* x; const float* addend; const float* d;
    float* logp; float* prob; float* depth; float* var; long long* argmax; float* quarter;
    int B, D, H, W, mode;
};

// Plain head through the streaming kernel; DPV_E_UNSUPP = shape not handled, take another kernel.
int launch_head_stream_plain(const HeadArgs& h, cudaStream_t st) {
    if (h.addend != nullptr || h.prob != nullptr || h.mode == DPV_IN_PROB) return DPV_E_UNSUPP;
    HeadStreamArgs a = {};
    a.x = h.x; a.d = h.d; a.logp = h.logp; a.depth = h.depth; a.var = h.var;
    a.argmax = h.argmax; a.quarter = h.quarter;
    a.B = h.B; a.H = h.H; a.W = h.W;
    return hs_dispatch(a, h.D, h.mode, false, nullptr, st);
}

}  // namespace dpv

// ------------------------------------------------------------------------------------ C ABI
extern "C" int64_t dpv_head_ufield_workspace_floats(int B, int D, int H, int W) {
    dpv::HsPlan p;
    if (!dpv::hs_plan(B, D, H, W, true, &p)) return 0;
    const int64_t recs = (int64_t)p.G * p.nseg;
    return recs * p.rec_floats + recs * dpv::HS_NW + 8;      // records, then per-warp flags (int32)
}

// Host-side helper (no device work): turn the four nearest-shift index maps of dpv_ufield into the
// two tables the fused kernel reads, checking that the shifts compose to "same pixel or padding"
// (true for the reference's row shifts; DPV_E_UNSUPP otherwise -> use dpv_head + dpv_ufield).
//   row_tab[y] = { yi = row_inv[y] (shifted row the numerator tests, -1 = none),
//                  1 if that shifted row samples the padding,
//                  ys = the shifted row whose source is y (denominator), -1 = none,
//                  1 if shifted row y samples the padding }
//   col_tab[x] = bit0 numerator valid, bit1 numerator samples padding, bit2 denominator valid,
//                bit3 shifted column x samples the padding
extern "C" int dpv_uf_fused_tables(const int* row_fwd, const int* row_inv, const int* col_fwd,
                                   const int* col_inv, int H, int W, int* row_tab, int* col_tab) {
    if (!row_fwd || !row_inv || !col_fwd || !col_inv || !row_tab || !col_tab || H <= 0 || W <= 0)
        return DPV_E_BADARG;
    for (int y = 0; y < H; ++y) {
        row_tab[4 * y] = -1; row_tab[4 * y + 1] = 0; row_tab[4 * y + 2] = -1; row_tab[4 * y + 3] = 0;
    }
    for (int y = 0; y < H; ++y) {
        const int yi = row_inv[y];
        if (yi >= H) return DPV_E_BADARG;
        if (yi >= 0) {
            const int sy = row_fwd[yi];
            if (sy >= 0 && sy != y) return DPV_E_UNSUPP;
            row_tab[4 * y] = yi;
            row_tab[4 * y + 1] = sy < 0;
        }
        const int src = row_fwd[y];          // shifted row y reads source row src
        if (src >= H) return DPV_E_BADARG;
        if (src < 0) row_tab[4 * y + 3] = 1;
        else {
            if (row_tab[4 * src + 2] >= 0) return DPV_E_UNSUPP;   // two shifted rows read one source row
            row_tab[4 * src + 2] = y;
        }
    }
    for (int x = 0; x < W; ++x) {
        int bits = 0;
        const int xi = col_inv[x];
        if (xi >= W) return DPV_E_BADARG;
        if (xi >= 0) {
            const int sx = col_fwd[xi];
            if (sx >= 0 && sx != x) return DPV_E_UNSUPP;
            // the count is kept per shifted column, so the numerator's column must be x itself
            if (xi != x) return DPV_E_UNSUPP;
            bits |= 1;
            if (sx < 0) bits |= 2;
        }
        const int src = col_fwd[x];
        if (src >= 0 && src != x) return DPV_E_UNSUPP;
        if (src == x) bits |= 4;
        if (src < 0) bits |= 8;
        col_tab[x] = bits;
    }
    return 0;
}

extern "C" int dpv_head_ufield(const float* x, const float* d_candi, float* logp, float* depth,
                               float* variance, int64_t* argmax, float* quarter,
                               const float* intr_up, const int* row_tab, const int* col_tab,
                               float* uf, float* depth_zero, float* workspace, int B, int D, int H,
                               int W, int64_t intr_bstride, int in_mode, float zstart, float zend,
                               float maxd, float mind, float pad_depth, void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(x && d_candi && intr_up && row_tab && col_tab && uf && workspace);
    DPV_CHECK_ARG(B > 0 && D > 0 && H > 0 && W > 0);
    DPV_CHECK_ARG(in_mode == DPV_IN_LOGITS || in_mode == DPV_IN_LOGPROB);
    HsPlan p;
    if (!hs_plan(B, D, H, W, true, &p)) return DPV_E_UNSUPP;
    HeadStreamArgs a = {};
    a.x = x; a.d = d_candi; a.logp = logp; a.depth = depth;
    a.var = variance; a.argmax = (long long*)argmax; a.quarter = quarter;
    a.row_tab = reinterpret_cast<const int4*>(row_tab); a.col_tab = col_tab; a.intr = intr_up;
    a.depth_zero = depth_zero;
    a.B = B; a.H = H; a.W = W;
    const int64_t recs = (int64_t)p.G * p.nseg;
    a.rec = workspace;
    a.flag = reinterpret_cast<int*>(workspace + recs * p.rec_floats);
    a.intr_bs = intr_bstride;
    a.zstart = zstart; a.zend = zend; a.maxd1 = maxd - 1.0f; a.mind = mind; a.pad_depth = pad_depth;
    return hs_dispatch(a, D, in_mode, true, uf, (cudaStream_t)stream);
}
