// K4c: LiDAR prior DPV and the Bayesian multiply-and-renormalise fusion.
//
// dpv_lidar_prior  <- gen_dpv_withmask (reference utils/img_utils.py:360-375) with
//                     gen_soft_label_torch (:31-47) and gen_uniform (:49-50)
// dpv_bayes_fuse   <- models/models.py:669-672
// The reference makes ~12 full passes over [B,D,h,w] tensors (python loop over items, 64-fold
// repeat of d_candi, exp, sum, divide, NaN patch, blend, clamp, log, add, exp, sum, divide,
// clamp, log).  Here each pixel's D bins are produced by one thread; the prior is never
// materialised when fusing.  HBM-bound: 12*D + 8 bytes per pixel.
#include "dpv_common.cuh"

namespace dpv {

constexpr float kEps = 2.220446049250313e-16f;   // float64 eps as fp32 (utils/img_utils.py:12)

// Gaussian bump over the bins, normalised; NaN (0/0 when z is far outside the bin range) -> -1;
// blend with the uniform DPV by the mask; clamp.  Operation order as in the reference.
struct PriorPixel {
    float z, mk, two_sig, sum, uni;
    __device__ __forceinline__ float gauss(float dk) const {
        const float a = fabsf(__fsub_rn(dk, z));
        return expf(__fdiv_rn(-__fmul_rn(a, a), two_sig));
    }
    __device__ __forceinline__ float value(float dk) const {
        float t = __fdiv_rn(gauss(dk), sum);
        if (t != t) t = -1.0f;
        const float mix = __fadd_rn(__fmul_rn(t, mk), __fmul_rn(uni, __fsub_rn(1.0f, mk)));
        return fminf(fmaxf(mix, kEps), 1.0f);
    }
};

__device__ __forceinline__ PriorPixel make_prior(const float* __restrict__ d, int D, float z,
                                                 float mk, float two_sig) {
    PriorPixel p;
    p.z = z; p.mk = mk; p.two_sig = two_sig; p.uni = __fdiv_rn(1.0f, (float)D);
    float s = 0.f;
    for (int k = 0; k < D; ++k) s = __fadd_rn(s, p.gauss(d[k]));
    p.sum = s;
    return p;
}

__global__ void __launch_bounds__(128) lidar_prior_kernel(
    const float* __restrict__ dmaps, const float* __restrict__ masks, const float* __restrict__ dc,
    float* __restrict__ prior, int D, int HW, float two_sig) {
    extern __shared__ float d_s[];
    for (int k = threadIdx.x; k < D; k += blockDim.x) d_s[k] = __ldg(dc + k);
    __syncthreads();
    const int b = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= HW) return;
    const long long pix = (long long)b * HW + q;
    const PriorPixel pp = make_prior(d_s, D, __ldg(dmaps + pix), __ldg(masks + pix), two_sig);
    float* o = prior + (long long)b * D * HW + q;
    for (int k = 0; k < D; ++k) st_stream(o + (long long)k * HW, pp.value(d_s[k]));
}

__global__ void __launch_bounds__(128) bayes_fuse_kernel(
    const float* __restrict__ bv, const float* __restrict__ prior, const float* __restrict__ dmaps,
    const float* __restrict__ masks, const float* __restrict__ dc, float* __restrict__ fused,
    float* __restrict__ log_fused, int D, int HW, float two_sig) {
    extern __shared__ float d_s[];
    for (int k = threadIdx.x; k < D; k += blockDim.x) d_s[k] = __ldg(dc + k);
    __syncthreads();
    const int b = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= HW) return;
    const long long pix = (long long)b * HW + q;
    const long long base = (long long)b * D * HW + q;
    PriorPixel pp;
    if (prior == nullptr) pp = make_prior(d_s, D, __ldg(dmaps + pix), __ldg(masks + pix), two_sig);
    auto joint = [&](int k) {   // exp(BV + log prior), models/models.py:669
        const float pk = (prior != nullptr) ? __ldg(prior + base + (long long)k * HW) : pp.value(d_s[k]);
        return expf(__fadd_rn(__ldg(bv + base + (long long)k * HW), logf(pk)));
    };
    float s = 0.f;
    for (int k = 0; k < D; ++k) s = __fadd_rn(s, joint(k));
    for (int k = 0; k < D; ++k) {
        float f = __fdiv_rn(joint(k), s);
        f = (f != f) ? f : fminf(fmaxf(f, kEps), 1.0f);   // torch.clamp keeps NaN
        if (fused != nullptr) st_stream(fused + base + (long long)k * HW, f);
        if (log_fused != nullptr) st_stream(log_fused + base + (long long)k * HW, logf(f));
    }
}

// ---- T lanes per pixel, D/T bins per lane in registers (the layout of head_kernel) --------------
// One thread per pixel (above) leaves 10 warps per SM at the model's 8 x 64 x 96 pixels and
// evaluates every bin's prior and joint twice with IEEE divisions (74 us measured for 38 MB).  Here
// adjacent lanes share a pixel; sums finish with log2(T) shuffles; the Gaussian keeps expf (its
// underflow decides the reference's NaN -> -1 patch), the rest uses the SFU forms (ex2 / lg2 /
// rcp; relative error ~1e-6, parity budget 1e-4).
template <int T>
__device__ __forceinline__ float fz_group_sum(float v) {
#pragma unroll
    for (int o = T / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int D, int T, bool FUSE>
__global__ void __launch_bounds__(128) prior_fuse_kernel(
    const float* __restrict__ bv, const float* __restrict__ prior_in, const float* __restrict__ dmaps,
    const float* __restrict__ masks, const float* __restrict__ dc, float* __restrict__ out0,
    float* __restrict__ out1, int HW, float two_sig) {
    constexpr int DT = D / T, PIX = 128 / T;
    __shared__ float d_s[D];
    for (int k = threadIdx.x; k < D; k += 128) d_s[k] = __ldg(dc + k);
    __syncthreads();
    const int b = blockIdx.y, r = threadIdx.x % T;
    int q = blockIdx.x * PIX + threadIdx.x / T;
    const bool live = q < HW;
    q = live ? q : HW - 1;
    const int kb = r * DT;
    const long long pix = (long long)b * HW + q;
    const long long base = ((long long)b * D + kb) * HW + q;
    float pk[DT];
    if (prior_in == nullptr) {
        const float z = __ldg(dmaps + pix), mk = __ldg(masks + pix);
        const float uni = __fdiv_rn(1.0f, (float)D);
        const float neg_inv = -__fdiv_rn(1.0f, two_sig);   // exponent differs from -a^2 / two_sig by <= 1 ulp
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < DT; ++k) {
            const float a = __fsub_rn(d_s[kb + k], z);
            pk[k] = expf(__fmul_rn(a, a) * neg_inv);
            s += pk[k];
        }
        s = fz_group_sum<T>(s);
        const float rs = __fdiv_rn(1.0f, s);   // s == 0 -> inf; 0 * inf = NaN -> -1, as 0 / 0 in the reference
        const float um = __fmul_rn(uni, __fsub_rn(1.0f, mk));
#pragma unroll
        for (int k = 0; k < DT; ++k) {
            float t = pk[k] * rs;
            if (t != t) t = -1.0f;
            pk[k] = fminf(fmaxf(__fadd_rn(__fmul_rn(t, mk), um), kEps), 1.0f);
        }
    } else {
#pragma unroll
        for (int k = 0; k < DT; ++k) pk[k] = ld_stream(prior_in + base + (long long)k * HW);
    }
    if (!FUSE) {   // out0 = prior
        if (live) {
#pragma unroll
            for (int k = 0; k < DT; ++k) st_stream(out0 + base + (long long)k * HW, pk[k]);
        }
        return;
    }
    // joint = exp(BV + log prior) (models/models.py:669), normalised, clamped; out0 = fused, out1 = log
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < DT; ++k) {
        const float x = ld_stream(bv + base + (long long)k * HW);
        pk[k] = __expf(x + __logf(pk[k]));
        s += pk[k];
    }
    s = fz_group_sum<T>(s);
    const float rs = __fdiv_rn(1.0f, s);
    if (live) {
#pragma unroll
        for (int k = 0; k < DT; ++k) {
            float f = pk[k] * rs;
            f = (f != f) ? f : fminf(fmaxf(f, kEps), 1.0f);   // torch.clamp keeps NaN
            if (out0 != nullptr) st_stream(out0 + base + (long long)k * HW, f);
            if (out1 != nullptr) st_stream(out1 + base + (long long)k * HW, __logf(f));
        }
    }
}

template <int D, bool FUSE>
static bool launch_prior_fuse(const float* bv, const float* prior, const float* dmaps, const float* masks,
                              const float* dc, float* out0, float* out1, int B, int HW, float two_sig,
                              cudaStream_t st) {
    // T = 4: a warp-wide load of one register slot is 4 runs of 8 consecutive pixels (whole 32 B sectors)
    constexpr int T = (D > 64) ? D / 16 : (D >= 16 ? 4 : 1);
    dim3 grid((HW + 128 / T - 1) / (128 / T), B), block(128);
    prior_fuse_kernel<D, T, FUSE><<<grid, block, 0, st>>>(bv, prior, dmaps, masks, dc, out0, out1, HW, two_sig);
    return true;
}
template <bool FUSE>
static bool dispatch_prior_fuse(int D, const float* bv, const float* prior, const float* dmaps,
                                const float* masks, const float* dc, float* out0, float* out1, int B, int HW,
                                float two_sig, cudaStream_t st) {
    switch (D) {
        case 16: return launch_prior_fuse<16, FUSE>(bv, prior, dmaps, masks, dc, out0, out1, B, HW, two_sig, st);
        case 32: return launch_prior_fuse<32, FUSE>(bv, prior, dmaps, masks, dc, out0, out1, B, HW, two_sig, st);
        case 64: return launch_prior_fuse<64, FUSE>(bv, prior, dmaps, masks, dc, out0, out1, B, HW, two_sig, st);
        case 128: return launch_prior_fuse<128, FUSE>(bv, prior, dmaps, masks, dc, out0, out1, B, HW, two_sig, st);
        case 256: return launch_prior_fuse<256, FUSE>(bv, prior, dmaps, masks, dc, out0, out1, B, HW, two_sig, st);
        default: return false;   // any other D: one thread per pixel (kernels above)
    }
}

}  // namespace dpv

extern "C" int dpv_lidar_prior(const float* dmaps, const float* masks, const float* d_candi,
                               float* prior, int B, int D, int H, int W, float two_sigma_sq,
                               void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(dmaps && masks && d_candi && prior);
    DPV_CHECK_ARG(B > 0 && D > 0 && H > 0 && W > 0);
    if (B > 65535 || D > 8192) return DPV_E_UNSUPP;
    const int HW = H * W;
    if (dispatch_prior_fuse<false>(D, nullptr, nullptr, dmaps, masks, d_candi, prior, nullptr, B, HW,
                                   two_sigma_sq, (cudaStream_t)stream)) {
        DPV_LAUNCH_END();
        return 0;
    }
    dim3 grid((HW + 127) / 128, B), block(128);
    lidar_prior_kernel<<<grid, block, D * sizeof(float), (cudaStream_t)stream>>>(
        dmaps, masks, d_candi, prior, D, HW, two_sigma_sq);
    DPV_LAUNCH_END();
    return 0;
}

extern "C" int dpv_bayes_fuse(const float* bv, const float* prior, const float* dmaps,
                              const float* masks, const float* d_candi, float* fused,
                              float* log_fused, int B, int D, int H, int W, float two_sigma_sq,
                              void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(bv && d_candi && (fused || log_fused));
    DPV_CHECK_ARG(prior || (dmaps && masks));
    DPV_CHECK_ARG(B > 0 && D > 0 && H > 0 && W > 0);
    if (B > 65535 || D > 8192) return DPV_E_UNSUPP;
    const int HW = H * W;
    if (dispatch_prior_fuse<true>(D, bv, prior, dmaps, masks, d_candi, fused, log_fused, B, HW,
                                  two_sigma_sq, (cudaStream_t)stream)) {
        DPV_LAUNCH_END();
        return 0;
    }
    dim3 grid((HW + 127) / 128, B), block(128);
    bayes_fuse_kernel<<<grid, block, D * sizeof(float), (cudaStream_t)stream>>>(
        bv, prior, dmaps, masks, d_candi, fused, log_fused, D, HW, two_sigma_sq);
    DPV_LAUNCH_END();
    return 0;
}
