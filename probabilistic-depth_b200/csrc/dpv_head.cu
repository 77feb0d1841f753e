// K3 (+K4b): the depth-bin head -- log-softmax over the D axis fused with everything the
// reference derives from it in separate passes:
//   log_softmax            models/models.py:351,560,637 (and :694 with an addend = K4b)
//   exp(logp)              models/models.py:653,675,697 (decoder input)
//   E[d]                   utils/img_utils.py:52-61
//   Var[d]                 trainer/default_trainer.py:333-336
//   argmax bin             torch.argmax (not in the reference)
//   1/4 nearest hand-off   trainer/default_trainer.py:221-222
// The volume is [B,D,H,W] with the bin axis strided by H*W.  T threads (adjacent lanes) share one
// pixel, each keeping D/T consecutive bins in registers, so the input is read from HBM exactly
// once and every output is written exactly once; a warp-wide load of one register slot is T runs
// of 32/T consecutive pixels (whole 32 B sectors).  Per-pixel reductions (max, sum, mean,
// variance, arg-max) finish with log2(T) warp shuffles.  Splitting the bins over T lanes keeps
// the register footprint small enough for >= 50 % occupancy, which is what hides HBM latency
// here.  HBM-bound: 8*D + 16 bytes per pixel with all outputs on.
#include <cstdlib>

#include "dpv_common.cuh"

namespace dpv {

struct HeadArgs {
    const float* x; const float* addend; const float* d;
    float* logp; float* prob; float* depth; float* var; long long* argmax; float* quarter;
    int B, D, H, W, mode;
};

constexpr int HEAD_NT = 128;

// exp2 on the SFU: one MUFU.EX2, denormal results flushed to zero.
__device__ __forceinline__ float ex2(float t) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    return r;
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

template <int T>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
    for (int o = T / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <int T>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = T / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int T>
__device__ __forceinline__ int group_min(int v) {
#pragma unroll
    for (int o = T / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Per-element budget at the HBM rate is ~25 issue slots (2.9 elements/clk/SM at 6.5 TB/s), so the
// soft-max is arranged around t = (x - max) * log2(e), kept in registers:
//   sum   : S = sum ex2(t)                      (1 MUFU + 1 FADD)
//   log p : t * ln2 - ln S                      (1 FFMA)   = (x - max) - ln S to 1-2 ulp of |x - max|
//   p     : ex2(t - log2 S)                     (1 FADD + 1 MUFU)
// The maximal bin has t = 0 exactly, so its log p is exactly -ln S, and the arg-max (first
// maximum of the log-probabilities, what torch.argmax returns on this kernel's own output) is the
// first bin whose log p equals -ln S.  (A running "best so far" chain makes the compiler keep
// every log p live: +100 registers.)  NaN inputs are not supported.
// LOGP / PROB / QTR say which per-element outputs exist: a store that is merely predicated off still
// costs its issue slot and its address arithmetic (ncu on the first version: three predicated
// stores per element, 43 % of all instructions were LEA/IMAD/IADD3).
template <int D, int T, int MODE, bool ADD, bool LOGP, bool PROB, bool QTR>
__global__ void __launch_bounds__(HEAD_NT) head_kernel(const HeadArgs a) {
    constexpr int DT = D / T;          // bins per thread
    constexpr int PIX = HEAD_NT / T;   // pixels per CTA
    __shared__ float d_s[D];
    // read through a volatile view: keeps the bin depths in shared memory instead of registers
    const volatile float* d_v = d_s;
    const int HW = a.H * a.W;
    const int tid = threadIdx.x;
    for (int k = tid; k < D; k += HEAD_NT) d_s[k] = __ldg(a.d + k);
    __syncthreads();
    const int b = blockIdx.y;
    const int r = tid % T;
    int q = blockIdx.x * PIX + tid / T;
    const bool live = q < HW;          // dead lanes keep running (shuffles) on the last pixel
    q = live ? q : HW - 1;
    const int kb = r * DT;
    const long long base = ((long long)b * D + kb) * HW + q;

    float v[DT];
    {
        const float* px = a.x + base;
        const float* pa = ADD ? a.addend + base : nullptr;
        constexpr int CH = (ADD && DT > 16) ? 16 : DT;   // with an addend: load and add in chunks
#pragma unroll
        for (int k0 = 0; k0 < DT; k0 += CH) {
#pragma unroll
            for (int k = k0; k < k0 + CH; ++k) {
                v[k] = ld_stream(px);
                px += HW;
                asm volatile("" : "+l"(px));   // running pointer: no table of offsets in registers
            }
            if (ADD) {
                float w[CH];
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    w[k] = ld_stream(pa);
                    pa += HW;
                    asm volatile("" : "+l"(pa));
                }
#pragma unroll
                for (int k = 0; k < CH; ++k) v[k0 + k] += w[k];
            }
        }
    }

    float ln_s = 0.f, log2_s = 0.f, top;
    if (MODE == DPV_IN_LOGITS) {
        float m = v[0];
#pragma unroll
        for (int k = 1; k < DT; ++k) m = fmaxf(m, v[k]);
        m = group_max<T>(m);
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < DT; ++k) {
            v[k] = (v[k] - m) * kLog2e;
            s += ex2(v[k]);
        }
        s = group_sum<T>(s);
        ln_s = logf(s);
        log2_s = ln_s * kLog2e;
        top = -ln_s;
    } else {
        top = v[0];
#pragma unroll
        for (int k = 1; k < DT; ++k) top = fmaxf(top, v[k]);
        top = group_max<T>(top);
    }

    // 1/4-resolution hand-off: is this pixel kept, and where does it go
    const int h4 = a.H / 4, w4 = a.W / 4;
    bool q_keep = false;
    float* qp = nullptr;
    if (QTR) {
        const int y = q / a.W, xx = q - y * a.W;
        q_keep = live && ((y & 3) == 0) && ((xx & 3) == 0) && ((y >> 2) < h4) && ((xx >> 2) < w4);
        qp = a.quarter + ((long long)b * D + kb + (DT - 1)) * h4 * w4 + (q_keep ? (y >> 2) * w4 + (xx >> 2) : 0);
    }
    const int q4 = h4 * w4;

    // Main pass, last bin first so that the smallest index among equal maxima is what remains.
    // One running pointer per output (a single 64-bit add per bin).
    float* lp_ptr = LOGP ? a.logp + base + (long long)(DT - 1) * HW : nullptr;
    float* pr_ptr = PROB ? a.prob + base + (long long)(DT - 1) * HW : nullptr;
    float mean = 0.f;
    int best_k = 1 << 30;
#pragma unroll
    for (int kk = 0; kk < DT; ++kk) {
        const int k = DT - 1 - kk;
        const float dk = d_v[kb + k];
        float lp, pr;
        if (MODE == DPV_IN_PROB) {
            pr = v[k];
            lp = pr;                              // arg-max over the values as given
        } else if (MODE == DPV_IN_LOGPROB) {
            lp = v[k];
            pr = ex2(lp * kLog2e);
        } else {
            lp = fmaf(v[k], kLn2, -ln_s);
            pr = ex2(v[k] - log2_s);
        }
        best_k = (lp == top) ? (kb + k) : best_k;
        mean = fmaf(dk, pr, mean);
        v[k] = pr;
        if (LOGP) {
            if (live) st_stream(lp_ptr, (MODE == DPV_IN_PROB) ? logf(pr) : lp);
            lp_ptr -= HW;
            asm volatile("" : "+l"(lp_ptr));
        }
        if (PROB) {
            if (live) st_stream(pr_ptr, pr);
            pr_ptr -= HW;
            asm volatile("" : "+l"(pr_ptr));
        }
        if (QTR) {
            if (q_keep) *qp = (MODE == DPV_IN_PROB) ? pr : lp;
            qp -= q4;
            asm volatile("" : "+l"(qp));
        }
    }
    mean = group_sum<T>(mean);
    best_k = group_min<T>(best_k);

    const long long pix = (long long)b * HW + q;
    const bool writer = live && (r == 0);
    if (a.var != nullptr) {
        float var = 0.f;
#pragma unroll
        for (int k = 0; k < DT; ++k) {
            const float c = d_v[kb + k] - mean;
            var = fmaf(c * c, v[k], var);
        }
        var = group_sum<T>(var);
        if (writer) a.var[pix] = var;
    }
    if (writer) {
        if (a.depth != nullptr) a.depth[pix] = mean;
        if (a.argmax != nullptr) a.argmax[pix] = (long long)(best_k == (1 << 30) ? 0 : best_k);
    }
}

// Any D: one pixel per thread, the bins are re-read from memory (L2 for the later passes).
__global__ void __launch_bounds__(HEAD_NT) head_generic_kernel(const HeadArgs a) {
    const int HW = a.H * a.W, D = a.D;
    const int b = blockIdx.y;
    const int q = blockIdx.x * HEAD_NT + threadIdx.x;
    if (q >= HW) return;
    const long long base = (long long)b * D * HW + q;
    const bool is_prob = (a.mode == DPV_IN_PROB);
    auto in = [&](int k) {
        float t = __ldg(a.x + base + (long long)k * HW);
        if (a.addend != nullptr) t += __ldg(a.addend + base + (long long)k * HW);
        return t;
    };
    float m = 0.f, shift = 0.f;
    if (a.mode == DPV_IN_LOGITS) {
        m = in(0);
        for (int k = 1; k < D; ++k) m = fmaxf(m, in(k));
        float s = 0.f;
        for (int k = 0; k < D; ++k) s += expf(in(k) - m);
        shift = logf(s);
    }
    float mean = 0.f, best = -INFINITY;
    int best_k = 0;
    for (int k = 0; k < D; ++k) {
        float lp, pr;
        if (is_prob) { pr = in(k); lp = pr; } else { lp = (in(k) - m) - shift; pr = expf(lp); }
        const bool take = (lp > best) || (k == 0);
        best_k = take ? k : best_k;
        best = take ? lp : best;
        mean = fmaf(__ldg(a.d + k), pr, mean);
        if (a.logp != nullptr) a.logp[base + (long long)k * HW] = is_prob ? logf(pr) : lp;
        if (a.prob != nullptr) a.prob[base + (long long)k * HW] = pr;
        if (a.quarter != nullptr) {
            const int y = q / a.W, xx = q - y * a.W, h4 = a.H / 4, w4 = a.W / 4;
            if ((y & 3) == 0 && (xx & 3) == 0 && (y >> 2) < h4 && (xx >> 2) < w4)
                a.quarter[((long long)b * D + k) * h4 * w4 + (y >> 2) * w4 + (xx >> 2)] =
                    is_prob ? pr : lp;
        }
    }
    const long long pix = (long long)b * HW + q;
    if (a.depth != nullptr) a.depth[pix] = mean;
    if (a.var != nullptr) {
        float var = 0.f;
        for (int k = 0; k < D; ++k) {
            float pr = is_prob ? in(k) : expf((in(k) - m) - shift);
            const float c = __ldg(a.d + k) - mean;
            var = fmaf(c * c, pr, var);
        }
        a.var[pix] = var;
    }
    if (a.argmax != nullptr) a.argmax[pix] = (long long)best_k;
}

// dpv_head_stream.cu: persistent TMA-fed kernel (D <= 64, W % 4 == 0, no addend / prob output).  It
// is what dpv_head_ufield runs; for the plain head the short-CTA kernel below is faster at the
// model's sizes (measured, profiles/README.md: a persistent CTA gets only ~8 rows each at
// 8 x 256 x 384, so its prologue, tail and 8-vs-9-row imbalance cost ~15 %), so it is opt-in here.
int launch_head_stream_plain(const HeadArgs& a, cudaStream_t st);

static const int g_head_stream = [] { const char* e = getenv("DPV_HEAD_STREAM"); return e ? atoi(e) : 0; }();
static const int g_head_t = [] { const char* e = getenv("DPV_HEAD_T"); return e ? atoi(e) : 0; }();

template <int D, int T>
static int launch_head_t(const HeadArgs& a, cudaStream_t st) {
    const int HW = a.H * a.W;
    constexpr int PIX = HEAD_NT / T;
    dim3 grid((HW + PIX - 1) / PIX, a.B), block(HEAD_NT);
    const bool lp = a.logp != nullptr, pr = a.prob != nullptr, qt = a.quarter != nullptr;
    // the output combinations the callers use; anything else takes head_generic_kernel
    const int combo = (lp && !pr && !qt) ? 0 : (lp && !pr && qt) ? 1 : (lp && pr && !qt) ? 2 :
                      (!lp && !pr && !qt) ? 3 : (lp && pr && qt) ? 4 : -1;
    if (combo < 0) return DPV_E_UNSUPP;
#define DPV_HEAD_GO(MODE_, ADD_)                                                                            \
    do {                                                                                                    \
        if (combo == 0) head_kernel<D, T, MODE_, ADD_, true, false, false><<<grid, block, 0, st>>>(a);      \
        else if (combo == 1) head_kernel<D, T, MODE_, ADD_, true, false, true><<<grid, block, 0, st>>>(a);  \
        else if (combo == 2) head_kernel<D, T, MODE_, ADD_, true, true, false><<<grid, block, 0, st>>>(a);  \
        else if (combo == 3) head_kernel<D, T, MODE_, ADD_, false, false, false><<<grid, block, 0, st>>>(a); \
        else head_kernel<D, T, MODE_, ADD_, true, true, true><<<grid, block, 0, st>>>(a);                   \
    } while (0)
    if (a.mode == DPV_IN_LOGITS && a.addend) DPV_HEAD_GO(DPV_IN_LOGITS, true);
    else if (a.mode == DPV_IN_LOGITS) DPV_HEAD_GO(DPV_IN_LOGITS, false);
    else if (a.mode == DPV_IN_LOGPROB) DPV_HEAD_GO(DPV_IN_LOGPROB, false);
    else DPV_HEAD_GO(DPV_IN_PROB, false);
#undef DPV_HEAD_GO
    DPV_LAUNCH_END();
    return 0;
}

// Threads per pixel.  Measured on B200 at D = 64, 8 x 256 x 384 pixels: T = 1 (96 registers, 5
// CTAs/SM, 128 B runs per request) 5.3 TB/s, T = 2 5.2 TB/s, T = 4 4.1 TB/s, T = 8 2.5 TB/s.  Small
// volumes (the 1/4-resolution DPV) split the bins to put more warps on the chip.
template <int D>
static int launch_head(const HeadArgs& a, cudaStream_t st) {
    const long long pixels = (long long)a.B * a.H * a.W;
    int t = (D > 64) ? D / 64 : 1;
    if (pixels < 148LL * 1024) t = max(t, 2);
    if (pixels < 148LL * 256) t = max(t, 4);
    if (a.addend != nullptr && pixels < 148LL * 1024) t = max(t, 4);   // two volumes to load: 0.0149 -> 0.0106 ms at 8x64x96
    if (g_head_t == 1 || g_head_t == 2 || g_head_t == 4 || g_head_t == 8) t = g_head_t;
    if (D / t > 64) t = D / 64;
    if (D / t < 4) t = D / 4;
    switch (t) {
        case 1: if constexpr (D <= 64) return launch_head_t<D, 1>(a, st); break;
        case 2: if constexpr (D <= 128 && D >= 8) return launch_head_t<D, 2>(a, st); break;
        case 4: if constexpr (D >= 16) return launch_head_t<D, 4>(a, st); break;
        case 8: if constexpr (D >= 32) return launch_head_t<D, 8>(a, st); break;
        default: break;
    }
    return DPV_E_UNSUPP;
}

}  // namespace dpv

extern "C" int dpv_head(const float* x, const float* addend, const float* d_candi, float* logp,
                        float* prob, float* depth, float* variance, int64_t* argmax,
                        float* quarter, int B, int D, int H, int W, int in_mode, void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(x && d_candi);
    DPV_CHECK_ARG(B > 0 && D > 0 && H > 0 && W > 0);
    DPV_CHECK_ARG(in_mode >= DPV_IN_LOGITS && in_mode <= DPV_IN_PROB);
    DPV_CHECK_ARG(addend == nullptr || in_mode == DPV_IN_LOGITS);   // x + addend is a logits-only notion
    if (B > 65535) return DPV_E_UNSUPP;
    HeadArgs a;
    a.x = x; a.addend = addend; a.d = d_candi; a.logp = logp; a.prob = prob; a.depth = depth;
    a.var = variance; a.argmax = (long long*)argmax; a.quarter = quarter;
    a.B = B; a.D = D; a.H = H; a.W = W; a.mode = in_mode;
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = H * W;
    if (g_head_stream) {   // streaming kernel when the shape allows it, else the scalar kernels below
        const int rc = launch_head_stream_plain(a, st);
        if (rc != DPV_E_UNSUPP) return rc;
    }
    int rc = DPV_E_UNSUPP;
    switch (D) {
        case 16: rc = launch_head<16>(a, st); break;
        case 32: rc = launch_head<32>(a, st); break;
        case 64: rc = launch_head<64>(a, st); break;
        case 128: rc = launch_head<128>(a, st); break;
        case 256: rc = launch_head<256>(a, st); break;
        default: break;
    }
    if (rc != DPV_E_UNSUPP) return rc;
    dim3 grid((HW + HEAD_NT - 1) / HEAD_NT, B), block(HEAD_NT);
    head_generic_kernel<<<grid, block, 0, st>>>(a);
    DPV_LAUNCH_END();
    return 0;
}
