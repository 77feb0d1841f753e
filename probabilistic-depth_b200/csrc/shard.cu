// Depth-plane sharded soft-max (large-D sweep).
//
// Not in the reference: it only ever soft-maxes D=64 planes on one device
// (models/models.py:351,560; models/packnet.py:394).  When the D planes of a cost volume are
// split over G GPUs, log_softmax over D needs one exchange: the per-pixel maximum (NCCL MAX),
// then the per-pixel sums (NCCL SUM).  Each rank runs
//     shard_max  -> all-reduce(max) -> shard_sums -> all-reduce(sum) -> shard_finish
// and, when the variance is wanted, shard_central -> all-reduce(sum) before finish; the
// collectives are issued by the host wrapper (torch.distributed / NCCL over NVLink) on the
// same stream.  Payload per collective: 4 (or 8) bytes per pixel.  The kernels are streaming
// passes over the local [B, D_local, HW] slice; HBM-bound (the re-reads hit L2 when the slice fits).
#include "dpv_common.cuh"

namespace dpv {

constexpr int SH_NT = 256;

__global__ void __launch_bounds__(SH_NT) shard_max_kernel(const float* __restrict__ x,
                                                          float* __restrict__ m, float* __restrict__ am,
                                                          int D, int HW, int k_offset) {
    const int b = blockIdx.y;
    const int q = blockIdx.x * SH_NT + threadIdx.x;
    if (q >= HW) return;
    const float* p = x + (long long)b * D * HW + q;
    float best = -INFINITY;
    int bk = 0;
    for (int k = 0; k < D; ++k) {
        const float v = __ldg(p + (long long)k * HW);
        const bool take = (v > best) || (k == 0);
        bk = take ? k : bk;
        best = take ? v : best;
    }
    m[(long long)b * HW + q] = best;
    if (am != nullptr) am[(long long)b * HW + q] = (float)(bk + k_offset);
}

// sums[0] = sum exp(x - M), sums[1] = sum d exp(x - M), M the global per-pixel maximum.
__global__ void __launch_bounds__(SH_NT) shard_sums_kernel(const float* __restrict__ x,
                                                           const float* __restrict__ d,
                                                           const float* __restrict__ gmax,
                                                           float* __restrict__ sums, int B, int D,
                                                           int HW) {
    const int b = blockIdx.y;
    const int q = blockIdx.x * SH_NT + threadIdx.x;
    if (q >= HW) return;
    const long long pix = (long long)b * HW + q;
    const float* p = x + (long long)b * D * HW + q;
    const float M = gmax[pix];
    float s = 0.f, t = 0.f;
    for (int k = 0; k < D; ++k) {
        const float e = expf(__ldg(p + (long long)k * HW) - M);
        s += e;
        t = fmaf(__ldg(d + k), e, t);
    }
    sums[pix] = s;
    sums[(long long)B * HW + pix] = t;
}

// central[pix] = sum (d - E)^2 exp(x - M) with E = sums[1]/sums[0] (global sums).
__global__ void __launch_bounds__(SH_NT) shard_central_kernel(const float* __restrict__ x,
                                                              const float* __restrict__ d,
                                                              const float* __restrict__ gmax,
                                                              const float* __restrict__ sums,
                                                              float* __restrict__ central, int B,
                                                              int D, int HW) {
    const int b = blockIdx.y;
    const int q = blockIdx.x * SH_NT + threadIdx.x;
    if (q >= HW) return;
    const long long pix = (long long)b * HW + q;
    const float* p = x + (long long)b * D * HW + q;
    const float M = gmax[pix];
    const float E = __fdiv_rn(sums[(long long)B * HW + pix], sums[pix]);
    float c = 0.f;
    for (int k = 0; k < D; ++k) {
        const float e = expf(__ldg(p + (long long)k * HW) - M);
        const float dd = __ldg(d + k) - E;
        c = fmaf(dd * dd, e, c);
    }
    central[pix] = c;
}

__global__ void __launch_bounds__(SH_NT) shard_finish_kernel(const float* __restrict__ x,
                                                             const float* __restrict__ gmax,
                                                             const float* __restrict__ sums,
                                                             const float* __restrict__ central,
                                                             float* __restrict__ logp,
                                                             float* __restrict__ depth,
                                                             float* __restrict__ var, int B, int D,
                                                             int HW) {
    const int b = blockIdx.y;
    const int q = blockIdx.x * SH_NT + threadIdx.x;
    if (q >= HW) return;
    const long long pix = (long long)b * HW + q;
    const float M = gmax[pix], S = sums[pix];
    const float ls = logf(S);
    if (logp != nullptr) {
        const float* p = x + (long long)b * D * HW + q;
        float* o = logp + (long long)b * D * HW + q;
        for (int k = 0; k < D; ++k)
            st_stream(o + (long long)k * HW, (__ldg(p + (long long)k * HW) - M) - ls);
    }
    if (depth != nullptr) depth[pix] = __fdiv_rn(sums[(long long)B * HW + pix], S);
    if (var != nullptr && central != nullptr) var[pix] = __fdiv_rn(central[pix], S);
}

// First-maximum-wins merge of per-rank (value, index) candidates, ranks in plane order.
__global__ void __launch_bounds__(SH_NT) shard_argmax_merge_kernel(const float* __restrict__ vals,
                                                                   const float* __restrict__ idx,
                                                                   long long* __restrict__ out,
                                                                   int G, long long n) {
    const long long i = (long long)blockIdx.x * SH_NT + threadIdx.x;
    if (i >= n) return;
    float best = vals[i];
    float bi = idx[i];
    for (int g = 1; g < G; ++g) {
        const float v = vals[(long long)g * n + i];
        if (v > best) { best = v; bi = idx[(long long)g * n + i]; }
    }
    out[i] = (long long)bi;
}

}  // namespace dpv

extern "C" int dpv_shard_max(const float* x, float* local_max, float* local_argmax, int B, int D,
                             int HW, int plane_offset, void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(x && local_max && B > 0 && D > 0 && HW > 0);
    dim3 grid((HW + SH_NT - 1) / SH_NT, B);
    shard_max_kernel<<<grid, SH_NT, 0, (cudaStream_t)stream>>>(x, local_max, local_argmax, D, HW,
                                                               plane_offset);
    DPV_LAUNCH_END();
    return 0;
}

extern "C" int dpv_shard_sums(const float* x, const float* d_local, const float* global_max,
                              float* sums, int B, int D, int HW, void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(x && d_local && global_max && sums && B > 0 && D > 0 && HW > 0);
    dim3 grid((HW + SH_NT - 1) / SH_NT, B);
    shard_sums_kernel<<<grid, SH_NT, 0, (cudaStream_t)stream>>>(x, d_local, global_max, sums, B, D, HW);
    DPV_LAUNCH_END();
    return 0;
}

extern "C" int dpv_shard_central(const float* x, const float* d_local, const float* global_max,
                                 const float* global_sums, float* central, int B, int D, int HW,
                                 void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(x && d_local && global_max && global_sums && central && B > 0 && D > 0 && HW > 0);
    dim3 grid((HW + SH_NT - 1) / SH_NT, B);
    shard_central_kernel<<<grid, SH_NT, 0, (cudaStream_t)stream>>>(x, d_local, global_max,
                                                                   global_sums, central, B, D, HW);
    DPV_LAUNCH_END();
    return 0;
}

extern "C" int dpv_shard_finish(const float* x, const float* global_max, const float* global_sums,
                                const float* global_central, float* logp, float* depth,
                                float* variance, int B, int D, int HW, void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(x && global_max && global_sums && B > 0 && D > 0 && HW > 0);
    dim3 grid((HW + SH_NT - 1) / SH_NT, B);
    shard_finish_kernel<<<grid, SH_NT, 0, (cudaStream_t)stream>>>(
        x, global_max, global_sums, global_central, logp, depth, variance, B, D, HW);
    DPV_LAUNCH_END();
    return 0;
}

extern "C" int dpv_shard_argmax_merge(const float* vals, const float* idx, int64_t* out, int G,
                                      int64_t n, void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(vals && idx && out && G > 0 && n > 0);
    shard_argmax_merge_kernel<<<(unsigned)((n + SH_NT - 1) / SH_NT), SH_NT, 0, (cudaStream_t)stream>>>(
        vals, idx, (long long*)out, G, n);
    DPV_LAUNCH_END();
    return 0;
}
