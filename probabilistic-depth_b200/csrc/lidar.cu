// LiDAR -> sparse depth maps on the device: what produces `dmaps` / `masks` for the
// upsample mode.  The reference does this on the host, per frame and eye, in the loader process:
//   generate_depth   external/utils_lib/python/utils_lib.cpp:86-160 (upsample = 0, the eval default of
//                    kittiloader/kitti.py:693-697): camera transform, z >= 0.1 cut, projection, z-buffer,
//                    (2f+1)^2 neighbourhood consistency filter
//   minpool          utils/img_utils.py:87-95 via kittiloader/kitti.py:706 (scale 4, zeros ignored)
// Three launches: clear, scatter-min (atomicMin on the bit pattern of the positive depth: the minimum
// does not depend on the order of arrival, so the result is deterministic and equal to the reference's
// sequential "smaller or first" rule), and one gather kernel that filters and pools together -- a
// thread owns one 1/4-resolution pixel, i.e. a 4x4 block of the filtered map.
// The matrix products keep a fixed order (row times column, left to right, separate multiply and add),
// the same as oracle/c/lidar_depthmap.c, so pixel indices and depths are bit-identical to the oracle.
#include "dpv_common.cuh"

namespace dpv {

constexpr unsigned kLidarEmpty = 0xffffffffu;

__device__ __forceinline__ float ld_dot4(const float* __restrict__ m, float x, float y, float z, float w) {
    float s = __fmul_rn(m[0], x);
    s = __fadd_rn(s, __fmul_rn(m[1], y));
    s = __fadd_rn(s, __fmul_rn(m[2], z));
    s = __fadd_rn(s, __fmul_rn(m[3], w));
    return s;
}

__global__ void __launch_bounds__(256) lidar_scatter_kernel(const float4* __restrict__ velo, int n,
                                                            const float* __restrict__ intr,
                                                            const float* __restrict__ m_velo2cam, int width,
                                                            int height, unsigned* __restrict__ zbuf) {
    __shared__ float m_s[16], k_s[12];
    if (threadIdx.x < 16) m_s[threadIdx.x] = __ldg(m_velo2cam + threadIdx.x);
    else if (threadIdx.x < 28) k_s[threadIdx.x - 16] = __ldg(intr + threadIdx.x - 16);
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = __ldg(velo + i);
    float cam[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) cam[r] = ld_dot4(m_s + 4 * r, p.x, p.y, p.z, p.w);       // utils_lib.cpp:95
    if (!(cam[2] >= 0.1f)) return;                                                        // :101
    float proj[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) proj[r] = ld_dot4(k_s + 4 * r, cam[0], cam[1], cam[2], cam[3]);   // :115
    const float px = __fdiv_rn(proj[0], proj[2]), py = __fdiv_rn(proj[1], proj[2]);       // :116-117
    const double ud = (double)px - 0.5, vd = (double)py - 0.5;                            // :123-124
    if (!(ud > -2147483000.0 && ud < 2147483000.0 && vd > -2147483000.0 && vd < 2147483000.0)) return;
    const int u = (int)ud, v = (int)vd;                                                   // truncation
    if (u < 0 || u >= width || v < 0 || v >= height) return;
    atomicMin(zbuf + v * width + u, __float_as_uint(cam[2]));                             // :126-129 (z > 0)
}

__device__ __forceinline__ float ld_raw(const unsigned* __restrict__ zbuf, int idx) {
    const unsigned b = __ldg(zbuf + idx);
    return b == kLidarEmpty ? 0.f : __uint_as_float(b);
}

// thread = one pooled pixel = SCALE x SCALE filtered pixels; also covers the right / bottom remainder
// columns that the pooled map drops (width % scale) through the tail threads.
__global__ void __launch_bounds__(128) lidar_filter_pool_kernel(const unsigned* __restrict__ zbuf, int width,
                                                                int height, int off, float filterdiff,
                                                                int scale, float pool_default,
                                                                float* __restrict__ dmap,
                                                                float* __restrict__ dmap_small,
                                                                float* __restrict__ mask_small) {
    const int w2 = (width + scale - 1) / scale, h2 = (height + scale - 1) / scale;
    const int x2 = blockIdx.x * blockDim.x + threadIdx.x, y2 = blockIdx.y;
    if (x2 >= w2 || y2 >= h2) return;
    float m = 0.f;
    bool first = true;
    for (int dy = 0; dy < scale; ++dy) {
        const int v = y2 * scale + dy;
        if (v >= height) break;
        for (int dx = 0; dx < scale; ++dx) {
            const int u = x2 * scale + dx;
            if (u >= width) break;
            float out = 0.f;
            if (v >= off && v < height - off - 1 && u >= off && u < width - off - 1) {    // :136-137
                const float z = ld_raw(zbuf, v * width + u);
                bool bad = false;
                for (int vv = v - off; vv < v + off + 1; ++vv)
                    for (int uu = u - off; uu < u + off + 1; ++uu) {
                        if (vv == v && uu == u) continue;
                        const float zn = ld_raw(zbuf, vv * width + uu);
                        if (zn != 0.f && __fsub_rn(zn, z) < -filterdiff) bad = true;      // :148-151
                    }
                if (!bad) out = z;
            }
            if (dmap != nullptr) dmap[v * width + u] = out;
            const float pv = (out == 0.f) ? pool_default : out;                           // img_utils.py:89-90
            if (first || pv < m) { m = pv; first = false; }
        }
    }
    // max_pool2d floors the output size: only complete blocks have a pooled pixel
    if (x2 < width / scale && y2 < height / scale) {
        const float small = (m == pool_default) ? 0.f : m;                                // :92
        if (dmap_small != nullptr) dmap_small[y2 * (width / scale) + x2] = small;
        if (mask_small != nullptr) mask_small[y2 * (width / scale) + x2] = (small < 0.01f) ? 0.f : 1.f;
    }
}

// Stand-alone minpool (utils/img_utils.py:87-95) over [N, H, W] planes: default != 0 ignores zeros.
__global__ void __launch_bounds__(128) minpool_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                      int H, int W, int scale, float dflt) {
    const int w2 = W / scale, h2 = H / scale;
    const int x2 = blockIdx.x * blockDim.x + threadIdx.x, y2 = blockIdx.y, nidx = blockIdx.z;
    if (x2 >= w2 || y2 >= h2) return;
    const float* src = in + (long long)nidx * H * W;
    float m = 0.f;
    bool first = true;
    for (int dy = 0; dy < scale; ++dy)
        for (int dx = 0; dx < scale; ++dx) {
            float v = __ldg(src + (y2 * scale + dy) * W + x2 * scale + dx);
            if (dflt != 0.f && v == 0.f) v = dflt;
            if (first || v < m) { m = v; first = false; }
        }
    if (dflt != 0.f && m == dflt) m = 0.f;
    out[((long long)nidx * h2 + y2) * w2 + x2] = m;
}

}  // namespace dpv

extern "C" int dpv_minpool(const float* in, float* out, int N, int H, int W, int scale, float default_value,
                           void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(in && out);
    DPV_CHECK_ARG(N > 0 && H > 0 && W > 0 && scale > 0 && H / scale > 0 && W / scale > 0);
    if (N > 65535 || H / scale > 65535) return DPV_E_UNSUPP;
    minpool_kernel<<<dim3((W / scale + 127) / 128, H / scale, N), 128, 0, (cudaStream_t)stream>>>(
        in, out, H, W, scale, default_value);
    DPV_LAUNCH_END();
    return 0;
}

extern "C" int dpv_lidar_depthmap(const float* velo, int n, const float* intr, const float* m_velo2cam,
                                  int width, int height, int filtering, float filterdiff, int pool_scale,
                                  float pool_default, unsigned* zbuf, float* dmap, float* dmap_small,
                                  float* mask_small, void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(velo && intr && m_velo2cam && zbuf);
    DPV_CHECK_ARG(n >= 0 && width > 0 && height > 0 && filtering >= 0 && pool_scale > 0);
    DPV_CHECK_ARG(dmap || dmap_small || mask_small);
    if ((long long)width * height > (1LL << 30) || ((uintptr_t)velo & 15)) return DPV_E_UNSUPP;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(zbuf, 0xff, sizeof(unsigned) * (size_t)width * height, st);
    if (e != cudaSuccess) return (int)e;
    if (n > 0) {
        lidar_scatter_kernel<<<(n + 255) / 256, 256, 0, st>>>((const float4*)velo, n, intr, m_velo2cam, width,
                                                             height, zbuf);
        DPV_LAUNCH_END();
    }
    const int w2 = (width + pool_scale - 1) / pool_scale, h2 = (height + pool_scale - 1) / pool_scale;
    if (h2 > 65535) return DPV_E_UNSUPP;
    lidar_filter_pool_kernel<<<dim3((w2 + 127) / 128, h2), 128, 0, st>>>(zbuf, width, height, filtering, filterdiff,
                                                                       pool_scale, pool_default, dmap, dmap_small,
                                                                       mask_small);
    DPV_LAUNCH_END();
    return 0;
}
