// Depth-plane sharded soft-max (large-D sweep).
//
// Not in the reference: it only ever soft-maxes D=64 planes on one device
// (models/models.py:351,560; models/packnet.py:394).  When the D planes of a cost volume are
// split over G GPUs, log_softmax over D (and E[d], Var, arg-max) needs ONE exchange.  Each rank runs
//     shard_stats  ->  all-gather of the per-pixel statistics  ->  shard_merge_finish
// * shard_stats: one pass over the local [B, D_local, HW] slice with the planes in registers (T adjacent
//   lanes share a pixel): local maximum m, S0 = sum exp(x - m), the local mean mu = sum d exp / S0, the
//   local central second moment M2 = sum (d - mu)^2 exp(x - m) and the index of the first local maximum --
//   five floats per pixel, written as five planes [5][B*HW].
// * the host all-gathers those 20 bytes per pixel (one NCCL collective over NVLink on the same stream).
// * shard_merge_finish: per pixel the statistics of the G ranks are merged with the online-soft-max rule
//   (w_g = S0_g exp(m_g - M)) and the pairwise-variance rule (M2 = sum [M2_g e^(m_g - M) + w_g (mu_g - mu)^2],
//   no cancellation), the arg-max is the first rank attaining M, and the local planes of the log-softmax are
//   written: logp = x - M - log S.
// Two reads and one write of the local slice (the first version read it four times and used three
// collectives: 3 ms per call at world 2; profiles/r01f_nccl_plane_shard.txt).  HBM-bound.
// The all-gather moves 20 B x pixels x G per rank, which at world 8 is more than the local slice itself.  For
// G > 2 the exchange therefore has the reduce-scatter shape: shard_stats writes slice-major records, an
// all-to-all hands rank g the G records of ITS 1/G of the pixels, shard_merge_slice merges them, an all-gather
// replicates the merged records (20 B x pixels per rank in each of the two collectives, independent of G), and
// shard_finish writes the planes.
#include "dpv_common.cuh"

namespace dpv {

constexpr int SH_NT = 128;
constexpr int SH_NSTAT = 5;     // m, S0, mu, M2, arg-max (as float: exact below 2^24 planes)

template <int T>
__device__ __forceinline__ float sh_group_max(float v) {
#pragma unroll
    for (int o = T / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <int T>
__device__ __forceinline__ float sh_group_sum(float v) {
#pragma unroll
    for (int o = T / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int T>
__device__ __forceinline__ int sh_group_min(int v) {
#pragma unroll
    for (int o = T / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// DL local planes, T lanes per pixel, DL / T planes per lane in registers: the slice is read once.
// Where statistic `s` of pixel `pix` (of n) lives: five planes [5][n] (slice == 0), or slice-major
// [n / slice][5][slice] -- the layout an all-to-all wants: block g is what rank g merges.
__device__ __forceinline__ long long sh_stat_index(long long pix, int s, long long n, int slice) {
    if (slice == 0) return (long long)s * n + pix;
    const long long g = pix / slice;
    return (g * SH_NSTAT + s) * slice + (pix - g * slice);
}

template <int DL, int T>
__global__ void __launch_bounds__(SH_NT) shard_stats_kernel(const float* __restrict__ x,
                                                            const float* __restrict__ d,
                                                            float* __restrict__ stats, int B, int HW,
                                                            int k_offset, int slice) {
    constexpr int DT = DL / T, PIX = SH_NT / T;
    __shared__ float d_s[DL];
    for (int k = threadIdx.x; k < DL; k += SH_NT) d_s[k] = __ldg(d + k);
    __syncthreads();
    const int b = blockIdx.y, r = threadIdx.x % T;
    int q = blockIdx.x * PIX + threadIdx.x / T;
    const bool live = q < HW;
    q = live ? q : HW - 1;
    const int kb = r * DT;
    const float* px = x + ((long long)b * DL + kb) * HW + q;
    float v[DT];
#pragma unroll
    for (int k = 0; k < DT; ++k) {
        v[k] = ld_stream(px);
        px += HW;
        asm volatile("" : "+l"(px));
    }
    float m = v[0];
#pragma unroll
    for (int k = 1; k < DT; ++k) m = fmaxf(m, v[k]);
    m = sh_group_max<T>(m);
    int first = 1 << 30;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int kk = 0; kk < DT; ++kk) {
        const int k = DT - 1 - kk;                 // last first: the smallest index among equal maxima remains
        first = (v[k] == m) ? kb + k : first;
        v[k] = __expf(v[k] - m);
        s0 += v[k];
        s1 = fmaf(d_s[kb + k], v[k], s1);
    }
    s0 = sh_group_sum<T>(s0); s1 = sh_group_sum<T>(s1);
    first = sh_group_min<T>(first);
    const float mu = __fdiv_rn(s1, s0);
    float m2 = 0.f;
#pragma unroll
    for (int k = 0; k < DT; ++k) {
        const float c = d_s[kb + k] - mu;
        m2 = fmaf(c * c, v[k], m2);
    }
    m2 = sh_group_sum<T>(m2);
    if (live && r == 0) {
        const long long n = (long long)B * HW, pix = (long long)b * HW + q;
        stats[sh_stat_index(pix, 0, n, slice)] = m; stats[sh_stat_index(pix, 1, n, slice)] = s0;
        stats[sh_stat_index(pix, 2, n, slice)] = mu; stats[sh_stat_index(pix, 3, n, slice)] = m2;
        stats[sh_stat_index(pix, 4, n, slice)] = (float)(first + k_offset);
    }
}

// Any D_local: one thread per pixel, three passes over the pixel's planes (the later ones hit L2).
__global__ void __launch_bounds__(SH_NT) shard_stats_generic_kernel(const float* __restrict__ x,
                                                                    const float* __restrict__ d,
                                                                    float* __restrict__ stats, int B, int D,
                                                                    int HW, int k_offset, int slice) {
    const int b = blockIdx.y;
    const int q = blockIdx.x * SH_NT + threadIdx.x;
    if (q >= HW) return;
    const float* p = x + (long long)b * D * HW + q;
    float m = -INFINITY;
    int first = 0;
    for (int k = 0; k < D; ++k) {
        const float v = __ldg(p + (long long)k * HW);
        const bool take = (v > m) || (k == 0);
        first = take ? k : first;
        m = take ? v : m;
    }
    float s0 = 0.f, s1 = 0.f;
    for (int k = 0; k < D; ++k) {
        const float e = __expf(__ldg(p + (long long)k * HW) - m);
        s0 += e;
        s1 = fmaf(__ldg(d + k), e, s1);
    }
    const float mu = __fdiv_rn(s1, s0);
    float m2 = 0.f;
    for (int k = 0; k < D; ++k) {
        const float e = __expf(__ldg(p + (long long)k * HW) - m);
        const float c = __ldg(d + k) - mu;
        m2 = fmaf(c * c, e, m2);
    }
    const long long n = (long long)B * HW, pix = (long long)b * HW + q;
    stats[sh_stat_index(pix, 0, n, slice)] = m; stats[sh_stat_index(pix, 1, n, slice)] = s0;
    stats[sh_stat_index(pix, 2, n, slice)] = mu; stats[sh_stat_index(pix, 3, n, slice)] = m2;
    stats[sh_stat_index(pix, 4, n, slice)] = (float)(first + k_offset);
}

// The merge of G records of one pixel (ranks in plane order): online-soft-max weights, pairwise variance rule,
// first rank attaining the maximum wins the arg-max.  rec(g, s) reads statistic s of rank g.
template <typename Rec>
__device__ __forceinline__ void sh_merge(const Rec& rec, int G, bool want_var, float& M, float& S, float& mean,
                                         float& var, float& am) {
    M = -INFINITY; am = 0.f;
    for (int g = 0; g < G; ++g) {
        const float mg = rec(g, 0);
        if (mg > M) { M = mg; am = rec(g, 4); }
    }
    S = 0.f;
    float s1 = 0.f;
    for (int g = 0; g < G; ++g) {
        const float w = rec(g, 1) * __expf(rec(g, 0) - M);
        S += w;
        s1 = fmaf(w, rec(g, 2), s1);
    }
    mean = __fdiv_rn(s1, S);
    var = 0.f;
    if (want_var) {
        float m2 = 0.f;
        for (int g = 0; g < G; ++g) {
            const float sc = __expf(rec(g, 0) - M);
            const float c = rec(g, 2) - mean;
            m2 += rec(g, 3) * sc + rec(g, 1) * sc * c * c;
        }
        var = __fdiv_rn(m2, S);
    }
}

// Reduce-scatter form, step 2: this rank merges the G records of ITS slice of the pixels.  recv [G][5][slice]
// (block g = rank g's statistics of my pixels, what the all-to-all delivered) -> merged [5][slice]:
// M, log S, mean, variance, arg-max.
__global__ void __launch_bounds__(SH_NT) shard_merge_slice_kernel(const float* __restrict__ recv,
                                                                  float* __restrict__ merged, int G, int slice) {
    const int i = blockIdx.x * SH_NT + threadIdx.x;
    if (i >= slice) return;
    auto rec = [&](int g, int s) { return recv[((long long)g * SH_NSTAT + s) * slice + i]; };
    float M, S, mean, var, am;
    sh_merge(rec, G, true, M, S, mean, var, am);
    merged[i] = M; merged[slice + i] = logf(S); merged[2 * slice + i] = mean; merged[3 * slice + i] = var;
    merged[4 * slice + i] = am;
}

// Reduce-scatter form, step 3: all [G][5][slice] holds the merged record of every pixel (block g = the slice rank g
// merged, what the all-gather delivered): replicated per-pixel products and the local planes of the log-softmax.
__global__ void __launch_bounds__(SH_NT) shard_finish_kernel(const float* __restrict__ x,
                                                             const float* __restrict__ all,
                                                             float* __restrict__ logp, float* __restrict__ depth,
                                                             float* __restrict__ var, long long* __restrict__ argmax,
                                                             int B, int D, int HW, int slice) {
    const int b = blockIdx.y;
    const int q = blockIdx.x * SH_NT + threadIdx.x;
    if (q >= HW) return;
    const long long n = (long long)B * HW, pix = (long long)b * HW + q;
    const float M = all[sh_stat_index(pix, 0, n, slice)], ls = all[sh_stat_index(pix, 1, n, slice)];
    if (depth != nullptr) depth[pix] = all[sh_stat_index(pix, 2, n, slice)];
    if (var != nullptr) var[pix] = all[sh_stat_index(pix, 3, n, slice)];
    if (argmax != nullptr) argmax[pix] = (long long)all[sh_stat_index(pix, 4, n, slice)];
    if (logp != nullptr) {
        const float* p = x + (long long)b * D * HW + q;
        float* o = logp + (long long)b * D * HW + q;
        int k = 0;
        for (; k + 8 <= D; k += 8) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = ld_stream(p + (long long)(k + j) * HW);
#pragma unroll
            for (int j = 0; j < 8; ++j) st_stream(o + (long long)(k + j) * HW, (v[j] - M) - ls);
        }
        for (; k < D; ++k) st_stream(o + (long long)k * HW, (ld_stream(p + (long long)k * HW) - M) - ls);
    }
}

// gathered [G][5][B*HW] (rank order = plane order).  One thread per pixel merges the G records, writes the
// replicated per-pixel products and streams the local planes of the log-softmax.
__global__ void __launch_bounds__(SH_NT) shard_merge_finish_kernel(const float* __restrict__ x,
                                                                   const float* __restrict__ gathered,
                                                                   float* __restrict__ logp,
                                                                   float* __restrict__ depth,
                                                                   float* __restrict__ var,
                                                                   long long* __restrict__ argmax, int G, int B,
                                                                   int D, int HW) {
    const int b = blockIdx.y;
    const int q = blockIdx.x * SH_NT + threadIdx.x;
    if (q >= HW) return;
    const long long n = (long long)B * HW, pix = (long long)b * HW + q;
    const long long gs = (long long)SH_NSTAT * n;
    auto rec = [&](int g, int s) { return gathered[g * gs + (long long)s * n + pix]; };
    float M, S, mean, vr, am;
    sh_merge(rec, G, var != nullptr, M, S, mean, vr, am);
    if (var != nullptr) var[pix] = vr;
    if (depth != nullptr) depth[pix] = mean;
    if (argmax != nullptr) argmax[pix] = (long long)am;
    if (logp != nullptr) {
        const float ls = logf(S);
        const float* p = x + (long long)b * D * HW + q;
        float* o = logp + (long long)b * D * HW + q;
        int k = 0;
        for (; k + 8 <= D; k += 8) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = ld_stream(p + (long long)(k + j) * HW);
#pragma unroll
            for (int j = 0; j < 8; ++j) st_stream(o + (long long)(k + j) * HW, (v[j] - M) - ls);
        }
        for (; k < D; ++k) st_stream(o + (long long)k * HW, (ld_stream(p + (long long)k * HW) - M) - ls);
    }
}

template <int DL, int T>
static void launch_stats(const float* x, const float* d, float* stats, int B, int HW, int off, int slice,
                         cudaStream_t st) {
    constexpr int PIX = SH_NT / T;
    dim3 grid((HW + PIX - 1) / PIX, B);
    shard_stats_kernel<DL, T><<<grid, SH_NT, 0, st>>>(x, d, stats, B, HW, off, slice);
}

}  // namespace dpv

extern "C" int dpv_shard_stats(const float* x, const float* d_local, float* stats, int B, int D, int HW,
                               int plane_offset, int slice, void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(x && d_local && stats && B > 0 && D > 0 && HW > 0 && plane_offset >= 0 && slice >= 0);
    if (B > 65535) return DPV_E_UNSUPP;
    cudaStream_t st = (cudaStream_t)stream;
    switch (D) {
        case 8: launch_stats<8, 1>(x, d_local, stats, B, HW, plane_offset, slice, st); break;
        case 16: launch_stats<16, 1>(x, d_local, stats, B, HW, plane_offset, slice, st); break;
        case 32: launch_stats<32, 1>(x, d_local, stats, B, HW, plane_offset, slice, st); break;
        case 64: launch_stats<64, 2>(x, d_local, stats, B, HW, plane_offset, slice, st); break;
        case 128: launch_stats<128, 4>(x, d_local, stats, B, HW, plane_offset, slice, st); break;
        default: {
            dim3 grid((HW + SH_NT - 1) / SH_NT, B);
            shard_stats_generic_kernel<<<grid, SH_NT, 0, st>>>(x, d_local, stats, B, D, HW, plane_offset, slice);
        }
    }
    DPV_LAUNCH_END();
    return 0;
}

extern "C" int dpv_shard_merge_finish(const float* x, const float* gathered, float* logp, float* depth,
                                      float* variance, int64_t* argmax, int G, int B, int D, int HW,
                                      void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(x && gathered && G > 0 && B > 0 && D > 0 && HW > 0);
    if (B > 65535) return DPV_E_UNSUPP;
    dim3 grid((HW + SH_NT - 1) / SH_NT, B);
    shard_merge_finish_kernel<<<grid, SH_NT, 0, (cudaStream_t)stream>>>(x, gathered, logp, depth, variance,
                                                                         (long long*)argmax, G, B, D, HW);
    DPV_LAUNCH_END();
    return 0;
}

extern "C" int dpv_shard_merge_slice(const float* recv, float* merged, int G, int slice, void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(recv && merged && G > 0 && slice > 0);
    shard_merge_slice_kernel<<<(slice + SH_NT - 1) / SH_NT, SH_NT, 0, (cudaStream_t)stream>>>(recv, merged, G, slice);
    DPV_LAUNCH_END();
    return 0;
}

extern "C" int dpv_shard_finish(const float* x, const float* all, float* logp, float* depth, float* variance,
                                int64_t* argmax, int B, int D, int HW, int slice, void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(x && all && B > 0 && D > 0 && HW > 0 && slice > 0);
    if (B > 65535) return DPV_E_UNSUPP;
    dim3 grid((HW + SH_NT - 1) / SH_NT, B);
    shard_finish_kernel<<<grid, SH_NT, 0, (cudaStream_t)stream>>>(x, all, logp, depth, variance, (long long*)argmax, B,
                                                                  D, HW, slice);
    DPV_LAUNCH_END();
    return 0;
}
