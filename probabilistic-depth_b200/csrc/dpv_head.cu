// K3 (+K4b): the depth-bin head -- log-softmax over the D axis fused with everything the
// reference derives from it in separate passes:
//   log_softmax            models/models.py:351,560,637 (and :694 with an addend = K4b)
//   exp(logp)              models/models.py:653,675,697 (decoder input)
//   E[d]                   utils/img_utils.py:52-61
//   Var[d]                 trainer/default_trainer.py:333-336
//   argmax bin             torch.argmax (not in the reference)
//   1/4 nearest hand-off   trainer/default_trainer.py:221-222
// The volume is [B,D,H,W] with the bin axis strided by H*W, so a warp reading one bin of 32 (or
// 64) consecutive pixels is a fully coalesced 128 B (256 B) request.  Each thread keeps all D
// bins of its pixel(s) in registers: the input is read from HBM exactly once and every output
// is written exactly once.  HBM-bound: 8*D + 16 bytes per pixel with all outputs on.
#include "dpv_common.cuh"

namespace dpv {

struct HeadArgs {
    const float* x; const float* addend; const float* d;
    float* logp; float* prob; float* depth; float* var; long long* argmax; float* quarter;
    int B, D, H, W, mode;
};

constexpr int HEAD_NT = 128;

template <int D, int VEC>
__global__ void __launch_bounds__(HEAD_NT, 2) head_kernel(const HeadArgs a) {
    __shared__ float d_s[D];
    const int HW = a.H * a.W;
    const int tid = threadIdx.x;
    for (int k = tid; k < D; k += HEAD_NT) d_s[k] = __ldg(a.d + k);
    __syncthreads();
    const int b = blockIdx.y;
    const int q = (blockIdx.x * HEAD_NT + tid) * VEC;
    if (q >= HW) return;
    const long long base = (long long)b * D * HW + q;

    float v[D][VEC];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        if (VEC == 2) {
            float2 t = ld_stream2(a.x + base + (long long)k * HW);
            v[k][0] = t.x; v[k][VEC - 1] = t.y;
        } else {
            v[k][0] = ld_stream(a.x + base + (long long)k * HW);
        }
    }
    if (a.addend != nullptr) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            if (VEC == 2) {
                float2 t = ld_stream2(a.addend + base + (long long)k * HW);
                v[k][0] += t.x; v[k][VEC - 1] += t.y;
            } else {
                v[k][0] += ld_stream(a.addend + base + (long long)k * HW);
            }
        }
    }

    float shift[VEC];   // what to subtract to get log-probabilities
#pragma unroll
    for (int j = 0; j < VEC; ++j) shift[j] = 0.f;
    if (a.mode == DPV_IN_LOGITS) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float m = v[0][j];
#pragma unroll
            for (int k = 1; k < D; ++k) m = fmaxf(m, v[k][j]);
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                v[k][j] = v[k][j] - m;
                s += expf(v[k][j]);
            }
            shift[j] = logf(s);
        }
    }

    // Pass 3: log-probabilities out, probabilities kept in registers, E[d], arg-max.
    float mean[VEC];
    int best_k[VEC];
    float best[VEC];
    const bool is_prob = (a.mode == DPV_IN_PROB);
#pragma unroll
    for (int j = 0; j < VEC; ++j) { mean[j] = 0.f; best_k[j] = 0; best[j] = -INFINITY; }
#pragma unroll
    for (int k = 0; k < D; ++k) {
        float lp[VEC], pr[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            if (is_prob) {
                pr[j] = v[k][j];
                lp[j] = v[k][j];   // arg-max is taken over the values as given
            } else {
                lp[j] = v[k][j] - shift[j];
                pr[j] = expf(lp[j]);
            }
            // torch.argmax: first maximum wins, NaN counts as the maximum
            const bool take = (lp[j] > best[j]) || (lp[j] != lp[j] && best[j] == best[j]) || (k == 0);
            best_k[j] = take ? k : best_k[j];
            best[j] = take ? lp[j] : best[j];
            mean[j] = fmaf(d_s[k], pr[j], mean[j]);
            v[k][j] = pr[j];
        }
        if (a.logp != nullptr) {
            float* o = a.logp + base + (long long)k * HW;
            if (is_prob) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) lp[j] = logf(pr[j]);
            }
            if (VEC == 2) st_stream2(o, make_float2(lp[0], lp[VEC - 1])); else st_stream(o, lp[0]);
        }
        if (a.prob != nullptr) {
            float* o = a.prob + base + (long long)k * HW;
            if (VEC == 2) st_stream2(o, make_float2(pr[0], pr[VEC - 1])); else st_stream(o, pr[0]);
        }
        if (a.quarter != nullptr) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const int qq = q + j, y = qq / a.W, xx = qq - y * a.W;
                const int h4 = a.H / 4, w4 = a.W / 4;
                if (qq < HW && (y & 3) == 0 && (xx & 3) == 0 && (y >> 2) < h4 && (xx >> 2) < w4) {
                    const float val = is_prob ? pr[j] : lp[j];
                    a.quarter[((long long)b * D + k) * h4 * w4 + (y >> 2) * w4 + (xx >> 2)] = val;
                }
            }
        }
    }

    const long long pix = (long long)b * HW + q;
    if (a.depth != nullptr) {
        if (VEC == 2) st_stream2(a.depth + pix, make_float2(mean[0], mean[VEC - 1]));
        else a.depth[pix] = mean[0];
    }
    if (a.var != nullptr) {
        float var[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            var[j] = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const float c = d_s[k] - mean[j];
                var[j] = fmaf(c * c, v[k][j], var[j]);
            }
        }
        if (VEC == 2) st_stream2(a.var + pix, make_float2(var[0], var[VEC - 1]));
        else a.var[pix] = var[0];
    }
    if (a.argmax != nullptr) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) a.argmax[pix + j] = (long long)best_k[j];
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
        const bool take = (lp > best) || (lp != lp && best == best) || (k == 0);
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

template <int D, int VEC>
static int launch_head(const HeadArgs& a, cudaStream_t st) {
    const int HW = a.H * a.W;
    dim3 grid((HW + HEAD_NT * VEC - 1) / (HEAD_NT * VEC), a.B), block(HEAD_NT);
    head_kernel<D, VEC><<<grid, block, 0, st>>>(a);
    DPV_LAUNCH_END();
    return 0;
}

}  // namespace dpv

extern "C" int dpv_head(const float* x, const float* addend, const float* d_candi, float* logp,
                        float* prob, float* depth, float* variance, int64_t* argmax,
                        float* quarter, int B, int D, int H, int W, int in_mode, void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(x && d_candi);
    DPV_CHECK_ARG(B > 0 && D > 0 && H > 0 && W > 0);
    DPV_CHECK_ARG(in_mode >= DPV_IN_LOGITS && in_mode <= DPV_IN_PROB);
    if (B > 65535) return DPV_E_UNSUPP;
    HeadArgs a;
    a.x = x; a.addend = addend; a.d = d_candi; a.logp = logp; a.prob = prob; a.depth = depth;
    a.var = variance; a.argmax = (long long*)argmax; a.quarter = quarter;
    a.B = B; a.D = D; a.H = H; a.W = W; a.mode = in_mode;
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = H * W;
    const uintptr_t bits = (uintptr_t)x | (uintptr_t)addend | (uintptr_t)logp | (uintptr_t)prob |
                           (uintptr_t)depth | (uintptr_t)variance;
    const bool pair = (HW % 2 == 0) && ((bits & 7) == 0);   // 8-byte vector accesses
    switch (D) {
        case 16: return pair ? launch_head<16, 2>(a, st) : launch_head<16, 1>(a, st);
        case 32: return pair ? launch_head<32, 2>(a, st) : launch_head<32, 1>(a, st);
        case 64: return pair ? launch_head<64, 2>(a, st) : launch_head<64, 1>(a, st);
        default: break;
    }
    dim3 grid((HW + HEAD_NT - 1) / HEAD_NT, B), block(HEAD_NT);
    head_generic_kernel<<<grid, block, 0, st>>>(a);
    DPV_LAUNCH_END();
    return 0;
}
