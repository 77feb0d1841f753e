// Eval metrics on the device: the reference moves every depth map to the host
// with .cpu().numpy() and loops over its pixels in C++ (trainer/default_trainer.py:246-256); these
// kernels keep the maps where the head kernel left them and return 9 numbers per item.
//
//   dpv_depth_errors  <- utils/img_utils.py:17-22 (depth_error: zeros are invalid) around
//                        depthError, external/deval_lib/src/evaluate_depth.h:19-119, with the
//                        trainer's preparation (clamp of the truth, mask of the prediction,
//                        trainer/default_trainer.py:247-254) as optional fused steps
//   dpv_unc_rmse      <- compute_unc_rmse, utils/img_utils.py:183-194
// Per-pixel arithmetic follows the C++ (float, except the inverse error whose `1.0 / x` is double);
// the sums are accumulated in double in a fixed order (per-thread stride, shuffle tree, chunk order) --
// bit-reproducible, and closer to the exact sums than the reference's sequential float adds, from which
// they differ by ~1e-6 relative.  The final formulas are the reference's, in float.
#include <algorithm>

#include "dpv_common.cuh"

namespace dpv {

constexpr int MT_NT = 256;
constexpr int MT_NS = 11;    // 9 metric sums (slot 6 = signed log sum) + spare + count

__device__ __forceinline__ double mt_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(MT_NT) depth_errors_partial_kernel(
    const float* __restrict__ first, const float* __restrict__ second, const float* __restrict__ mask,
    float clamp_max, int zero_invalid, double* __restrict__ part, int HW, int per_chunk) {
    __shared__ double red_s[MT_NT / 32][MT_NS];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int q0 = chunk * per_chunk, q1 = min(HW, q0 + per_chunk);
    const float* pa = first + (long long)b * HW;
    const float* pb = second + (long long)b * HW;
    const float* pm = mask ? mask + (long long)b * HW : nullptr;
    double s[MT_NS];
#pragma unroll
    for (int i = 0; i < MT_NS; ++i) s[i] = 0.0;
    for (int q = q0 + threadIdx.x; q < q1; q += MT_NT) {
        float gt = __ldg(pa + q);                      // the python wrapper passes the prediction first
        float ip = __ldg(pb + q);
        if (pm != nullptr) gt = __fmul_rn(gt, __ldg(pm + q));          // default_trainer.py:251-254
        if (clamp_max > 0.f && ip >= clamp_max) ip = clamp_max;        // default_trainer.py:247-250
        if (zero_invalid) {                                             // img_utils.py:20-21
            gt = (gt == 0.f) ? -1.f : gt;
            ip = (ip == 0.f) ? -1.f : ip;
        }
        if (gt >= 0.f) {                                                // io_depth.h:99-101
            const float d_err = fabsf(__fsub_rn(gt, ip));
            const float d_sq = __fmul_rn(d_err, d_err);
            const float d_inv = (float)fabs(1.0 / (double)gt - 1.0 / (double)ip);
            const float lg = logf(gt), li = logf(ip);
            const float d_log = fabsf(__fsub_rn(lg, li));
            s[0] += d_err;
            s[1] += d_sq;
            s[2] += d_inv;
            s[3] += __fmul_rn(d_inv, d_inv);
            s[4] += d_log;
            s[5] += __fmul_rn(d_log, d_log);
            s[6] += __fsub_rn(lg, li);
            s[7] += __fdiv_rn(d_err, gt);
            s[8] += __fdiv_rn(d_sq, __fmul_rn(gt, gt));
            s[10] += 1.0;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < MT_NS; ++i) {
        const double v = mt_warp_sum(s[i]);
        if (lane == 0) red_s[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < MT_NS) {
        double v = 0.0;
        for (int w = 0; w < MT_NT / 32; ++w) v += red_s[w][threadIdx.x];
        part[((long long)b * gridDim.x + chunk) * MT_NS + threadIdx.x] = v;
    }
}

// One warp per item: chunk partials in order, then evaluate_depth.h:97-117 in float.
__global__ void depth_errors_finish_kernel(const double* __restrict__ part, float* __restrict__ out,
                                           int* __restrict__ counts, int nchunk) {
    const int b = blockIdx.x, i = threadIdx.x;
    __shared__ float e[MT_NS];
    if (i < MT_NS) {
        double v = 0.0;
        for (int c = 0; c < nchunk; ++c) v += part[((long long)b * nchunk + c) * MT_NS + i];
        e[i] = (float)v;
    }
    __syncthreads();
    if (i == 0) {
        const float n = e[10];
        float* o = out + b * 9;
        if (counts != nullptr) counts[b] = (int)n;
        const float nsl = __fdiv_rn(e[5], n);
        o[0] = __fdiv_rn(e[0], n);
        o[1] = sqrtf(__fdiv_rn(e[1], n));
        o[2] = __fdiv_rn(e[2], n);
        o[3] = sqrtf(__fdiv_rn(e[3], n));
        o[4] = __fdiv_rn(e[4], n);
        o[5] = sqrtf(nsl);
        o[6] = sqrtf(__fsub_rn(nsl, __fdiv_rn(__fmul_rn(e[6], e[6]), __fmul_rn(n, n))));
        o[7] = __fdiv_rn(e[7], n);
        o[8] = __fdiv_rn(e[8], n);
    }
}

// One CTA per item; a thread owns columns x, x + 128, ...
__global__ void __launch_bounds__(128) unc_rmse_kernel(const float* __restrict__ uf_truth,
                                                       const float* __restrict__ uf_pred,
                                                       const float* __restrict__ d, float* __restrict__ out,
                                                       int D, int W) {
    __shared__ float sum_s[4], cnt_s[4];
    const int b = blockIdx.x;
    const float* ut = uf_truth + (long long)b * D * W;
    const float* up = uf_pred + (long long)b * D * W;
    float sum = 0.f, cnt = 0.f;
    for (int x = threadIdx.x; x < W; x += 128) {
        float et = 0.f, ep = 0.f;
        for (int k = 0; k < D; ++k) {                    // dpv_to_depthmap, BV_log=False (:185-186)
            const float dk = __ldg(d + k);
            et = fmaf(dk, __ldg(ut + (long long)k * W + x), et);
            ep = fmaf(dk, __ldg(up + (long long)k * W + x), ep);
        }
        if (x == 0 || x == W - 1) ep = 0.f;              // :187-188
        const bool ok = (et == et) && (ep == ep);        // :189
        if (ok) { sum += fabsf(et - ep); cnt += 1.f; }   // :190-193 (the 'rmse' of :192 is overwritten)
    }
    sum = warp_sum(sum); cnt = warp_sum(cnt);
    if ((threadIdx.x & 31) == 0) { sum_s[threadIdx.x >> 5] = sum; cnt_s[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0)
        out[b] = __fdiv_rn(sum_s[0] + sum_s[1] + sum_s[2] + sum_s[3], cnt_s[0] + cnt_s[1] + cnt_s[2] + cnt_s[3]);
}

static int mt_chunks(int HW) { return std::max(1, std::min(64, (HW + 4095) / 4096)); }

}  // namespace dpv

extern "C" int64_t dpv_depth_errors_workspace_doubles(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return (int64_t)B * dpv::mt_chunks(H * W) * dpv::MT_NS;
}

extern "C" int dpv_depth_errors(const float* first, const float* second, const float* mask, float clamp_max,
                                int zero_invalid, float* out, int* counts, double* workspace, int B, int H,
                                int W, void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(first && second && out && workspace);
    DPV_CHECK_ARG(B > 0 && H > 0 && W > 0);
    if (B > 65535 || (long long)H * W > (1LL << 30)) return DPV_E_UNSUPP;
    const int HW = H * W, nchunk = mt_chunks(HW);
    const int per_chunk = (HW + nchunk - 1) / nchunk;
    cudaStream_t st = (cudaStream_t)stream;
    depth_errors_partial_kernel<<<dim3(nchunk, B), MT_NT, 0, st>>>(first, second, mask, clamp_max, zero_invalid,
                                                                 workspace, HW, per_chunk);
    DPV_LAUNCH_END();
    depth_errors_finish_kernel<<<B, 32, 0, st>>>(workspace, out, counts, nchunk);
    DPV_LAUNCH_END();
    return 0;
}

extern "C" int dpv_unc_rmse(const float* uf_truth, const float* uf_pred, const float* d_candi, float* out,
                            int B, int D, int W, void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(uf_truth && uf_pred && d_candi && out);
    DPV_CHECK_ARG(B > 0 && D > 0 && W > 0);
    unc_rmse_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(uf_truth, uf_pred, d_candi, out, D, W);
    DPV_LAUNCH_END();
    return 0;
}
