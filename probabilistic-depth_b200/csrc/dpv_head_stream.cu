// K3 (+K5): persistent, bulk-copy fed depth-bin head, optionally fused with the UF collapse.
//
// log-softmax over the bin axis fused with everything the reference derives from it in separate
// passes (E[d], Var[d], arg-max, the 1/4-resolution hand-off; see dpv_head.cu for the sites).
// A thread owns ONE image column of a 128-column strip and keeps the D (<= 64) bins of the current
// pixel in registers: no shuffles, no replicated per-pixel work.  The CTA walks down its run of rows;
// the grid is sized to the machine (5 CTAs of 4 warps per SM) and the rows of all strips are split
// evenly over the CTAs.
//
// The input side is done by the bulk-copy engine (cp.async.bulk, the non-tensor TMA path): one
// [D][128] tile = D row segments of 512 contiguous bytes, each moved by one instruction issued by
// one elected lane, completion counted in bytes on an mbarrier.  The compute threads read the tile
// at compile-time shared-memory offsets, hand the stage back after the arg-max pass over their
// column, and the next row is in flight during the whole soft-max of the current one (5 CTAs x
// 32 KB per SM in flight, no registers spent on it).  Why: ncu on the first kernels (profiles/)
// showed 38-43 % of all issued instructions were global address arithmetic (LEA/IADD3/IMAD), three
// predicated stores per element, and the kernel issue-bound at 81 % of the measured HBM roof.
//
// UF = true additionally performs gen_ufield (reference utils/img_utils.py:268-358) on the
// probabilities while they are in registers.  The reference's two nearest-neighbour row shifts
// cancel: for image pixel (y, x) the numerator weight is band(E[d](y, x), shifted row) times border
// predicates (closed form verified to reproduce gen_ufield exactly, NaN pattern
// included); the predicates come from the reference's own grid construction through two small
// tables (dpv_uf_fused_tables).  Each thread keeps its column's running sums for the CTA's run of
// rows in a private slot of the CTA's record (global memory, L2-resident; rows off the road band
// cost nothing), and uf_stream_finish_kernel adds the records of a column in run order: no
// atomics, bit-reproducible.
#include <algorithm>
#include <cstdlib>

#include <cuda.h>

#include "dpv_common.cuh"

namespace dpv {

constexpr int HS_NT = 128;            // threads per CTA = columns per strip
constexpr int HS_NW = HS_NT / 32;     // warps per CTA
constexpr int HS_CTAS_PER_SM = 5;     // 96 registers x 128 threads, 33 KB of shared memory

struct HeadStreamArgs {
    const float* x; const float* d;
    float* logp; float* depth; float* var; long long* argmax; float* quarter;
    // fused uncertainty field
    const int4* row_tab; const int* col_tab; const float* intr;
    float* depth_zero; float* rec; int* flag;
    int B, H, W, S2, GP;              // S2 strips per row, GP CTAs per (item, strip)
    long long intr_bs, rec_floats;
    float zstart, zend, maxd1, mind, pad_depth;
};

__device__ __forceinline__ void hs_st(float* p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
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
// TMA tile load: box [D][1][128] of the (x, y, item*D + bin) view of the volume -> shared memory,
// completion counted in bytes on the mbarrier; columns past W are zero-filled by the engine.
__device__ __forceinline__ void hs_tma_load(void* dst, const CUtensorMap* map, int x, int y, int kb,
                                            unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
                 "[%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(hs_smem(dst)), "l"(map), "r"(x), "r"(y), "r"(kb), "r"(hs_smem(bar)) : "memory");
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

// Second half of a row: log p / p per bin, E[d], arg-max, log p stored.  WITH_Q is CTA-uniform
// (rows y % 4 == 0 feed the 1/4-resolution hand-off).
template <int D, int MODE, bool LOGP, bool WITH_Q>
__device__ __forceinline__ void hs_main_pass(float (&v)[D], const float* d_s, float* lp_ptr, int HW,
                                             bool live, float ln_s, float log2_s, float top, float* qp,
                                             int q4, bool q_keep, float& mean_out, int& best_out) {
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
        if (LOGP) {
            if (live) hs_st(lp_ptr, lp);
            lp_ptr -= HW;
            asm volatile("" : "+l"(lp_ptr));     // keep one running pointer (a 64-bit add per bin)
        }
        if (WITH_Q) {
            if (q_keep) *qp = lp;
            qp -= q4;
            asm volatile("" : "+l"(qp));
        }
    }
    mean_out = mean0 + mean1;
    best_out = best_k;
}

// One thread = one image column of a 128-column strip; the CTA walks down its run of rows.  The
// input tile ([D][128] floats = one row of the strip, all bins) is brought into shared memory by
// cp.async.bulk (completion on an mbarrier), copied to registers with compile-time-offset LDS, and
// the stage is handed straight back to the copy engine for the next row, which is then in flight
// for the whole soft-max of the current one.  log p leaves through coalesced 128-byte STGs.
template <int D, int MODE, bool LOGP, bool UF>
__global__ void __launch_bounds__(HS_NT, UF ? HS_CTAS_PER_SM - 1 : HS_CTAS_PER_SM)
head_stream_kernel(const HeadStreamArgs a, const __grid_constant__ CUtensorMap tmap_x) {
    __shared__ __align__(128) float stage[D * HS_NT];                     // [D][128]
    __shared__ float d_s[D];
    __shared__ __align__(8) unsigned long long full;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int HW = a.H * a.W;
    // CTA -> (item, strip, j): rows j, j + GP, j + 2 GP, ... of one 128-column strip.  The CTAs that
    // run together therefore stream GP consecutive rows of every strip: contiguous DRAM pages.
    const int bs = blockIdx.x / a.GP, j = blockIdx.x - bs * a.GP;
    const int b = bs / a.S2, ws = bs - b * a.S2;
    const int x0 = ws * HS_NT;
    const bool live = (x0 + tid) < a.W;
    const int x = x0 + (live ? tid : 0);         // dead threads shadow column x0 (never stored)
    for (int k = tid; k < D; k += HS_NT) d_s[k] = __ldg(a.d + k);
    if (tid == 0) {
        hs_mbar_init(&full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (j >= a.H) return;

    // One TMA request per tile, issued by one elected lane.  (Measured: D separate 512-byte
    // cp.async.bulk requests per tile cap the copy engine at ~55 % of the HBM rate.)
    constexpr unsigned kTileBytes = D * HS_NT * 4;
    if (warp == 0 && lane == 0) {
        hs_mbar_expect_tx(&full, kTileBytes);
        hs_tma_load(stage, &tmap_x, x0, j, b * D, &full);
    }

    float cnt = 0.f, fy = 0.f, cy = 0.f;
    bool any_band = false, has_acc = false;
    int ct = 0;
    if (UF) {
        fy = __ldg(a.intr + b * a.intr_bs + 4); cy = __ldg(a.intr + b * a.intr_bs + 5);
        ct = live ? __ldg(a.col_tab + x) : 0;
    }
    const int h4 = a.H / 4, w4 = a.W / 4, q4 = h4 * w4;
    const long long item = (long long)b * D * HW;
    float* const rec = UF ? a.rec + (long long)blockIdx.x * a.rec_floats + tid : nullptr;

    int it = 0;
    for (int y = j; y < a.H; y += a.GP, ++it) {
        // ---- the current tile: shared memory -> registers ----------------------------------------
        hs_mbar_wait(&full, (unsigned)(it & 1));
        float v[D];
#pragma unroll
        for (int k = 0; k < D; ++k) v[k] = stage[k * HS_NT + tid];       // immediate-offset LDS
        float ln_s = 0.f, log2_s = 0.f, top;
        float m = v[0];
#pragma unroll
        for (int k = 1; k < D; ++k) m = fmaxf(m, v[k]);
        // every thread now holds its column (m depends on all D values): give the stage back
        __syncthreads();
        if (y + a.GP < a.H && warp == 0 && lane == 0) {
            hs_mbar_expect_tx(&full, kTileBytes);
            hs_tma_load(stage, &tmap_x, x0, y + a.GP, b * D, &full);
        }
        if (MODE == DPV_IN_LOGITS) {
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
            top = m;
        }

        const int pix = y * a.W + x;
        float mean;
        int best_k;
        float* lp_ptr = LOGP ? a.logp + item + (long long)(D - 1) * HW + pix : nullptr;
        const bool want_q = (a.quarter != nullptr) && ((y & 3) == 0) && ((y >> 2) < h4);   // CTA-uniform
        if (want_q) {
            const bool q_keep = live && ((x & 3) == 0) && ((x >> 2) < w4);
            float* qp = a.quarter + ((long long)b * D + (D - 1)) * q4 + (y >> 2) * w4 + (x >> 2);
            hs_main_pass<D, MODE, LOGP, true>(v, d_s, lp_ptr, HW, live, ln_s, log2_s, top, qp, q4, q_keep,
                                              mean, best_k);
        } else {
            hs_main_pass<D, MODE, LOGP, false>(v, d_s, lp_ptr, HW, live, ln_s, log2_s, top, nullptr, 0,
                                               false, mean, best_k);
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
            any_band |= (wn != 0.f) | (wd != 0.f);
            // The column's running sums live in this CTA's record (global memory, L2-resident, one
            // private slot per thread and bin): rows off the road band cost nothing, no shared memory
            // is tied up, and the order of the additions is fixed.
            if (wn != 0.f) {
                if (has_acc) {
#pragma unroll
                    for (int k0 = 0; k0 < D; k0 += 8) {
                        float r[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) r[k] = rec[(k0 + k) * HS_NT];
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            rec[(k0 + k) * HS_NT] = __fadd_rn(r[k], __fmul_rn(v[k0 + k], wn));
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < D; ++k) rec[k * HS_NT] = __fmul_rn(v[k], wn);
                    has_acc = true;
                }
            }
            cnt = __fadd_rn(cnt, wd);
        }
    }
    if (UF) {
        // record layout: [D][128] sums, then [128] counts; valid when the warp's flag is set
        const bool warp_any = __any_sync(0xffffffffu, any_band) != 0;
        if (warp_any) {
            if (!has_acc) {
#pragma unroll 8
                for (int k = 0; k < D; ++k) rec[k * HS_NT] = 0.f;
            }
            rec[D * HS_NT] = cnt;
        }
        if (lane == 0) a.flag[(long long)blockIdx.x * HS_NW + warp] = warp_any ? 1 : 0;
    }
}

// UF[b,k,x] = sum over the CTAs of the column's strip (in CTA order) of their partial sums, divided
// by (their counts + the padding rows' count).  0/0 = NaN as in the reference.  block (32 columns,
// 8 bins); a 32-column block lies inside one warp's quarter of one strip.
template <int D>
__global__ void __launch_bounds__(256) uf_stream_finish_kernel(const HeadStreamArgs a, float* uf) {
    __shared__ float den_s[8][32];
    __shared__ int on_s[64];
    const int c = threadIdx.x, g = threadIdx.y, tid = g * 32 + c;
    const int x = blockIdx.x * 32 + c, b = blockIdx.z;
    const int k = blockIdx.y * 8 + g;
    const bool ok = x < a.W;
    const int xe = ok ? x : 0;
    const int ws = (blockIdx.x * 32) / HS_NT, xin = xe - ws * HS_NT, wq = (blockIdx.x * 32 - ws * HS_NT) / 32;
    const long long cta0 = ((long long)b * a.S2 + ws) * a.GP;
    const float fy = __ldg(a.intr + b * a.intr_bs + 4), cy = __ldg(a.intr + b * a.intr_bs + 5);
    // shifted-frame pixels that sample the zero padding (their E[d] is pad_depth): rows split over g
    float den = 0.f;
    {
        const bool colpad = (__ldg(a.col_tab + xe) >> 3) & 1;
        for (int ys = g; ys < a.H; ys += 8)
            if (colpad || __ldg(a.row_tab + ys).w) den += hs_band(a, a.pad_depth, hs_yf(ys, fy, cy));   // small integers: exact
    }
    float num = 0.f;
    for (int p0 = 0; p0 < a.GP; p0 += 64) {
        const int np = min(64, a.GP - p0);
        __syncthreads();
        // flags are per (CTA, warp): uniform over this block's 32 columns
        if (tid < np) on_s[tid] = a.flag[(cta0 + p0 + tid) * HS_NW + wq];
        __syncthreads();
        // Unflagged records were never written; they are read anyway (in bounds) and discarded, so
        // that all loads are independent of the flags and of each other.
        if (g == 0) {
#pragma unroll 8
            for (int i = 0; i < np; ++i) {
                const float v = a.rec[(cta0 + p0 + i) * a.rec_floats + D * HS_NT + xin];
                den = __fadd_rn(den, on_s[i] ? v : 0.f);
            }
        }
        if (k < D) {
#pragma unroll 8
            for (int i = 0; i < np; ++i) {
                const float v = a.rec[(cta0 + p0 + i) * a.rec_floats + k * HS_NT + xin];
                num = __fadd_rn(num, on_s[i] ? v : 0.f);
            }
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

struct HsPlan { int S2, GP, G; long long rec_floats; };

// D <= 64: a thread holds all D bins of its pixel in registers.  W % 4 == 0: TMA strides are
// multiples of 16 bytes.
static bool hs_plan(int B, int D, int H, int W, bool uf, HsPlan* p) {
    if (D != 16 && D != 32 && D != 64) return false;
    if (B <= 0 || H <= 0 || W <= 0 || (W & 3) != 0) return false;
    const int per_sm = uf ? HS_CTAS_PER_SM - 1 : HS_CTAS_PER_SM;
    p->S2 = (W + HS_NT - 1) / HS_NT;
    const long long strips = (long long)B * p->S2;
    const long long slots = (long long)hs_sm_count() * per_sm;
    long long gp = slots / strips;              // CTAs per (item, strip): all resident at once
    if (gp < 1) gp = 1;                         // more strips than slots: one CTA each, several waves
    if (gp > H) gp = H;
    if (strips * gp > 0x7fffffffLL) return false;
    p->GP = (int)gp;
    p->G = (int)(strips * gp);
    p->rec_floats = (long long)D * HS_NT + HS_NT;
    return true;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency).
typedef CUresult (*hs_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                 CUtensorMapFloatOOBfill);
static hs_encode_fn hs_encoder() {
    static hs_encode_fn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return (hs_encode_fn)f;
    }();
    return fn;
}

// (x, y, item*D + bin) view of a contiguous [B, D, H, W] fp32 volume, box = one row of a strip.
static int hs_make_map(CUtensorMap* m, const float* base, int B, int D, int H, int W) {
    hs_encode_fn enc = hs_encoder();
    if (!enc) return DPV_E_NODEVICE;
    const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * D};
    const cuuint64_t gstr[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
    const cuuint32_t box[3] = {HS_NT, 1, (cuuint32_t)D};
    const cuuint32_t est[3] = {1, 1, 1};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box,
                           est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : DPV_E_UNSUPP;
}

template <int D, int MODE, bool LOGP, bool UF>
static int hs_launch_one(const HeadStreamArgs& a, const HsPlan& p, cudaStream_t st) {
    CUtensorMap mx;
    const int rc = hs_make_map(&mx, a.x, a.B, D, a.H, a.W);
    if (rc != 0) return rc;
    head_stream_kernel<D, MODE, LOGP, UF><<<dim3(p.G), dim3(HS_NT), 0, st>>>(a, mx);
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
    uf_stream_finish_kernel<D><<<g2, b2, 0, st>>>(a, uf_out);
    DPV_LAUNCH_END();
    return 0;
}

static int hs_dispatch(HeadStreamArgs& a, int D, int mode, bool uf, float* uf_out, cudaStream_t st) {
    HsPlan p;
    if (!hs_plan(a.B, D, a.H, a.W, uf, &p)) return DPV_E_UNSUPP;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (!al16(a.x) || !al16(a.logp)) return DPV_E_UNSUPP;
    if (!al16(a.row_tab)) return DPV_E_BADARG;
    a.S2 = p.S2; a.GP = p.GP; a.rec_floats = p.rec_floats;
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

// dpv_head_uftile.cu: the tile form of the fused kernel (default for D = 32 / 64)
long long head_uf_tile_workspace_floats(int B, int D, int H, int W);
int launch_head_uf_tile(const float* x, const float* d, float* logp, float* depth, float* var, long long* argmax,
                        float* quarter, const float* intr, const int* row_tab, const int* col_tab, float* uf,
                        float* depth_zero, float* workspace, int B, int D, int H, int W, long long intr_bs,
                        int mode, float zstart, float zend, float maxd, float mind, float pad_depth,
                        cudaStream_t st);
static const int g_head_uf_stream = [] { const char* e = getenv("DPV_HEAD_UF_STREAM"); return e ? atoi(e) : 0; }();

}  // namespace dpv

// ------------------------------------------------------------------------------------ C ABI
extern "C" int64_t dpv_head_ufield_workspace_floats(int B, int D, int H, int W) {
    dpv::HsPlan p;
    if (!dpv::hs_plan(B, D, H, W, true, &p)) return 0;
    const int64_t recs = (int64_t)p.G;
    const int64_t stream_ws = recs * p.rec_floats + recs * dpv::HS_NW + 8;      // records, then per-warp flags (int32)
    return std::max<int64_t>(stream_ws, dpv::head_uf_tile_workspace_floats(B, D, H, W));
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
    if (!g_head_uf_stream) {   // tile kernel (dpv_head_uftile.cu) unless the shape is not one of its
        const int rc = launch_head_uf_tile(x, d_candi, logp, depth, variance, (long long*)argmax, quarter, intr_up,
                                           row_tab, col_tab, uf, depth_zero, workspace, B, D, H, W, intr_bstride,
                                           in_mode, zstart, zend, maxd, mind, pad_depth, (cudaStream_t)stream);
        if (rc != DPV_E_UNSUPP) return rc;
    }
    HeadStreamArgs a = {};
    a.x = x; a.d = d_candi; a.logp = logp; a.depth = depth;
    a.var = variance; a.argmax = (long long*)argmax; a.quarter = quarter;
    a.row_tab = reinterpret_cast<const int4*>(row_tab); a.col_tab = col_tab; a.intr = intr_up;
    a.depth_zero = depth_zero;
    a.B = B; a.H = H; a.W = W;
    const int64_t recs = (int64_t)p.G;
    a.rec = workspace;
    a.flag = reinterpret_cast<int*>(workspace + recs * p.rec_floats);
    a.intr_bs = intr_bstride;
    a.zstart = zstart; a.zend = zend; a.maxd1 = maxd - 1.0f; a.mind = mind; a.pad_depth = pad_depth;
    return hs_dispatch(a, D, in_mode, true, uf, (cudaStream_t)stream);
}
