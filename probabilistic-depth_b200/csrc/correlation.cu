// K2b: FlowNet/PWC-style local correlation, forward.
//
// Replaces the reference's CUDA extension (models/correlation_package/correlation_cuda.cc:10-87;
// kernels correlation_cuda_kernel.cu:15-39 channels_first and :41-114 correlation_forward) and
// its pure-torch twin (models/correlation_native.py:13-23) for the configuration the reference
// uses (models/pwclite.py:123-125: pad = max_displacement = 4, kernel 1, strides 1):
//     out[b, i*n + j, y, x] = (1/C) sum_c x1[b,c,y,x] * x2[b,c,y+i-r,x+j-r],  n = 2r+1, zero padded.
// The reference kernel first rewrites both inputs to padded NHWC, then runs one 32-thread block
// per output pixel with a serial shared-memory reduction.  Here no re-layout is needed: with
// threads along x, NCHW reads are already coalesced.  Each thread owns 4 x-adjacent pixels and
// the n displacements of one row offset i: per channel it issues 1 + 3 aligned 128-bit loads for
// 4*n FMAs, the window overlap between neighbouring threads and rows is served by L1.
#include "dpv_common.cuh"

namespace dpv {

// r = 4, W % 4 == 0, 16-byte aligned rows.  grid (W/4 tiles of 32 threads, H, B*9).
__global__ void __launch_bounds__(128) corr_r4_kernel(const float* __restrict__ x1,
                                                      const float* __restrict__ x2,
                                                      float* __restrict__ out, int C, int H, int W) {
    constexpr int R = 4, N = 9;
    const int xq = blockIdx.x * blockDim.x + threadIdx.x;   // group of 4 pixels
    const int x = xq * 4;
    const int y = blockIdx.y;
    const int b = blockIdx.z / N, i = blockIdx.z % N;
    if (x >= W) return;
    const long long HW = (long long)H * W;
    const int y2 = y + i - R;
    float acc[4][N];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int j = 0; j < N; ++j) acc[p][j] = 0.f;
    if (y2 >= 0 && y2 < H) {
        const float* a = x1 + (long long)b * C * HW + (long long)y * W + x;
        const float* w = x2 + (long long)b * C * HW + (long long)y2 * W + (x - R);
        const bool in0 = (x - R) >= 0, in2 = (x + R) < W;   // x .. x+3 is always inside
        for (int c = 0; c < C; ++c) {
            const float4 av = __ldg(reinterpret_cast<const float4*>(a));
            float win[12];
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 w0 = in0 ? __ldg(reinterpret_cast<const float4*>(w)) : z;
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + 4));
            const float4 w2 = in2 ? __ldg(reinterpret_cast<const float4*>(w + 8)) : z;
            win[0] = w0.x; win[1] = w0.y; win[2] = w0.z; win[3] = w0.w;
            win[4] = w1.x; win[5] = w1.y; win[6] = w1.z; win[7] = w1.w;
            win[8] = w2.x; win[9] = w2.y; win[10] = w2.z; win[11] = w2.w;
            const float ap[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int j = 0; j < N; ++j) acc[p][j] = fmaf(ap[p], win[p + j], acc[p][j]);
            a += HW; w += HW;
        }
    }
    const float cf = (float)C;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        float4 o = make_float4(__fdiv_rn(acc[0][j], cf), __fdiv_rn(acc[1][j], cf),
                               __fdiv_rn(acc[2][j], cf), __fdiv_rn(acc[3][j], cf));
        *reinterpret_cast<float4*>(out + ((long long)b * N * N + i * N + j) * HW + (long long)y * W + x) = o;
    }
}

// Any radius / width: one thread per output element.
__global__ void __launch_bounds__(128) corr_generic_kernel(const float* __restrict__ x1,
                                                           const float* __restrict__ x2,
                                                           float* __restrict__ out, int C, int H,
                                                           int W, int R) {
    const int N = 2 * R + 1;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int b = blockIdx.z / (N * N), disp = blockIdx.z % (N * N);
    if (x >= W) return;
    const int y2 = y + disp / N - R, x2c = x + disp % N - R;
    const long long HW = (long long)H * W;
    float acc = 0.f;
    if (y2 >= 0 && y2 < H && x2c >= 0 && x2c < W) {
        const float* a = x1 + (long long)b * C * HW + (long long)y * W + x;
        const float* w = x2 + (long long)b * C * HW + (long long)y2 * W + x2c;
        for (int c = 0; c < C; ++c) { acc = fmaf(__ldg(a), __ldg(w), acc); a += HW; w += HW; }
    }
    out[((long long)b * N * N + disp) * HW + (long long)y * W + x] = __fdiv_rn(acc, (float)C);
}

}  // namespace dpv

extern "C" int dpv_correlation(const float* x1, const float* x2, float* out, int B, int C, int H,
                               int W, int max_displacement, void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(x1 && x2 && out);
    DPV_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && max_displacement >= 0);
    const int R = max_displacement, N = 2 * R + 1;
    if (H > 65535 || (long long)B * N * N > 65535) return DPV_E_UNSUPP;
    cudaStream_t st = (cudaStream_t)stream;
    const bool aligned = (((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)out) & 15) == 0;
    if (R == 4 && W % 4 == 0 && aligned) {
        const int groups = W / 4;
        const int nt = groups >= 128 ? 128 : (groups >= 64 ? 64 : 32);
        dim3 grid((groups + nt - 1) / nt, H, B * N), block(nt);
        corr_r4_kernel<<<grid, block, 0, st>>>(x1, x2, out, C, H, W);
    } else {
        dim3 grid((W + 127) / 128, H, B * N * N), block(128);
        corr_generic_kernel<<<grid, block, 0, st>>>(x1, x2, out, C, H, W, R);
    }
    DPV_LAUNCH_END();
    return 0;
}
