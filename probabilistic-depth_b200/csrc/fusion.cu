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

}  // namespace dpv

extern "C" int dpv_lidar_prior(const float* dmaps, const float* masks, const float* d_candi,
                               float* prior, int B, int D, int H, int W, float two_sigma_sq,
                               void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(dmaps && masks && d_candi && prior);
    DPV_CHECK_ARG(B > 0 && D > 0 && H > 0 && W > 0);
    if (B > 65535 || D > 8192) return DPV_E_UNSUPP;
    const int HW = H * W;
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
    dim3 grid((HW + 127) / 128, B), block(128);
    bayes_fuse_kernel<<<grid, block, D * sizeof(float), (cudaStream_t)stream>>>(
        bv, prior, dmaps, masks, d_candi, fused, log_fused, D, HW, two_sigma_sq);
    DPV_LAUNCH_END();
    return 0;
}
