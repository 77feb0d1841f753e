// SURVEY 8(f) rank 2, second half: Base3D (models/models.py:376-438), the 3-D convolution stack of the feedback
// mode -- comb_volume [B, 4, D, h, w] -> 32 -> 32 -> (two residual blocks of 32 -> 32 -> 32) -> 32 -> 1 channels,
// 3x3x3 filters, stride 1, zero padding 1, no bias, BatchNorm3d after every convolution but the last, ReLU where the
// reference has one.  In eval with running statistics (bn_avg: true in the feedback configs) a BatchNorm is a
// per-channel scale and shift: the scale is folded into the packed weights, the shift is the kernel's bias.
// 2 * 32 * 32 * 27 flops per voxel and layer, 21.7 GFLOP per frame (64 x 64 x 96 voxels) and layer: the dominant cost of the
// feedback mode (826 vs 273 ms on the CPU).
//
// Same machinery as conv_tc.cu (tcgen05.mma kind::tf32, accumulators in TMEM, TF32 x 3 split precision at fp32
// parity), shaped for 32 channels and three spatial dimensions:
//   * Activations travel packed: [B][D+2][H+2][W+2][32], channels innermost (one voxel = one 128-byte swizzle row),
//     a zero border all round, twice (hi / lo).  The A operand of tap (dz, dy, dx) is the SAME matrix
//     [positions][32] shifted by dz*(H+2)*(W+2) + dy*(W+2) + dx rows: one 2-D TMA map serves all 27 taps; rows
//     outside the tensor are zero-filled by the copy engine.  Outputs are computed for all padded positions and
//     border positions are written as zero -- the next layer's padding.
//   * A K-block is one (dz, dy) pair: 9 per tile, each serving its three dx taps from ONE A tile by descriptor
//     start address.  A CTA pass covers 256 positions (two M = 128 tiles) so the 24 KB of weights per K-block are
//     fetched once for both: 92 KB per stage, 2 stages.
//   * Per tap, K step (8 channels) and M tile two MMAs: A_hi x [W_hi ; W_lo] (N = 64) and A_lo x W_hi (N = 32) into
//     the cross-term columns; even and odd K-blocks accumulate into separate TMEM tiles (the tensor core adds into
//     fp32 with truncation: fewer steps per accumulator, conv_tc.cu).  2 M tiles x 2 accumulators x 64 columns =
//     256 columns per pass, double-buffered (512): the epilogue of a pass overlaps the copies and MMAs of the next.
//   * Epilogue (warps 2-5, one voxel per thread, 32 channels in registers): + shift, + residual (the packed
//     activation of an earlier layer, hi + lo), ReLU, then either the re-split into the next layer's packed hi / lo
//     or channel 0 of the last layer to [B][D][H][W].
//   * The first layer (4 real input channels) is z-folded (conv3d_pack_zfold_kernel): the three z-planes become 12 input
//     channels, 3 K-blocks (dy) of two K steps instead of 9 of one.  Unfolded inputs with fewer than 32 channels
//     issue only the K steps that hold data (`ksteps`).
//   * The last layer (32 -> 1) runs on the FP32 pipe (conv3d_c32_to1_kernel below).
#include <cuda.h>

#include <cstdlib>

#include "conv_tc.cuh"

namespace dpv {

constexpr int C3_C = 32;                    // channels in = out of the middle layers; one voxel = 128 bytes
constexpr int C3_M = 128;                   // positions per M tile
constexpr int C3_MT = 2;                    // M tiles per pass (share the weights of a K-block)
constexpr int C3_NKB = 9;                   // (dz, dy) pairs
constexpr int C3_STAGES = 2;
constexpr int C3_BOX = C3_M + 8;            // rows per TMA box
constexpr int C3_A_BYTES = C3_MT * C3_BOX * C3_C * 4;      // 272 rows x 128 B = 34 KB (a multiple of 1024)
constexpr int C3_W_BYTES = C3_C * C3_C * 4;                // 4 KB per tap and half (hi or lo)
constexpr int C3_STAGE_BYTES = 2 * C3_A_BYTES + 6 * C3_W_BYTES;   // 92 KB
constexpr int C3_THREADS = 192;
constexpr int C3_COLS = C3_MT * 2 * 2 * C3_C;              // per pass: 2 M tiles x 2 accumulators x (32 hi*hi + 32 cross) = 256
static_assert(C3_A_BYTES % 1024 == 0 && (C3_BOX * C3_C * 4) % 1024 == 0, "swizzle period");

struct Conv3Maps {
    CUtensorMap a_hi, a_lo;   // [NP][32] packed activations, box {32, 136}, SWIZZLE_128B
    CUtensorMap w_hi, w_lo;   // [27 * 32 (tap, out)][32 in] packed weights, box {32, 32}
};

struct Conv3Args {
    const float* bias;                 // [32] (the BatchNorm shift; zeros for the last layer)
    const float* res_hi; const float* res_lo;   // packed residual, nullable
    float* out_hi; float* out_lo;      // packed output, nullable
    float* out_c0;                     // [B][D][H][W]: channel 0 only, nullable
    float* out_raw;                    // [NP][32] fp32 (the convolution before a batch-statistics BatchNorm), nullable
    double* stats;                     // [64]: per channel sum and sum of squares over the real voxels (+=), nullable
    int B, D, H, W;
    long long NP;                      // B * (D+2) * (H+2) * (W+2)
    int relu, ksteps;
    int nkb, zfold;                    // K-blocks per tile: 9 (dz, dy) pairs, or 3 dy rows when the input is z-folded
};

__global__ void __launch_bounds__(C3_THREADS, 1)
conv3d_c32_tc_kernel(const Conv3Args a, const __grid_constant__ Conv3Maps maps) {
    extern __shared__ __align__(1024) unsigned char c3_smem[];
    __shared__ unsigned long long full_bar[C3_STAGES], empty_bar[C3_STAGES], acc_full[2], acc_empty[2];
    __shared__ unsigned tmem_base_s;
    __shared__ float bias_s[C3_C];
    unsigned char* stage0 = (unsigned char*)(((uintptr_t)c3_smem + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Wp = a.W + 2, Hp = a.H + 2, Dp = a.D + 2;
    const int npass = (int)((a.NP + C3_MT * C3_M - 1) / (C3_MT * C3_M));

    if (threadIdx.x < C3_C) bias_s[threadIdx.x] = a.bias != nullptr ? __ldg(a.bias + threadIdx.x) : 0.f;
    if (threadIdx.x == 0) {
        for (int s = 0; s < C3_STAGES; ++s) { tm_mbar_init(&full_bar[s], 1); tm_mbar_init(&empty_bar[s], 1); }
        for (int i = 0; i < 2; ++i) { tm_mbar_init(&acc_full[i], 1); tm_mbar_init(&acc_empty[i], C3_M); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(tm_smem(&tmem_base_s)), "r"((unsigned)(2 * C3_COLS)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_d = tmem_base_s;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int g = 0;
            for (int pass = blockIdx.x; pass < npass; pass += gridDim.x) {
                const long long p0 = (long long)pass * (C3_MT * C3_M);
                for (int kb = 0; kb < a.nkb; ++kb, ++g) {
                    const int s = g % C3_STAGES, use = g / C3_STAGES;
                    if (use > 0) tm_mbar_wait(&empty_bar[s], (use - 1) & 1);
                    const int dz = a.zfold ? 0 : kb / 3 - 1, dy = a.zfold ? kb - 1 : kb % 3 - 1;
                    unsigned char* st = stage0 + s * C3_STAGE_BYTES;
                    tm_mbar_expect_tx(&full_bar[s], C3_STAGE_BYTES);
                    // the dx = -1 tap's first row; rows < 0 or >= NP are zero-filled.  |row| < 2^31 is checked by the host.
                    const int row = (int)(p0 + (long long)dz * Hp * Wp + dy * Wp - 1);
                    for (int h = 0; h < C3_MT; ++h) {
                        ct_tma_2d(st + h * (C3_BOX * C3_C * 4), &maps.a_hi, 0, row + h * C3_BOX, &full_bar[s]);
                        ct_tma_2d(st + C3_A_BYTES + h * (C3_BOX * C3_C * 4), &maps.a_lo, 0, row + h * C3_BOX, &full_bar[s]);
                    }
                    for (int j = 0; j < 3; ++j) {
                        unsigned char* wb = st + 2 * C3_A_BYTES + j * 2 * C3_W_BYTES;
                        ct_tma_2d(wb, &maps.w_hi, 0, (kb * 3 + j) * C3_C, &full_bar[s]);
                        ct_tma_2d(wb + C3_W_BYTES, &maps.w_lo, 0, (kb * 3 + j) * C3_C, &full_bar[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const unsigned idesc0 = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(C3_M >> 4) << 24);
            const unsigned idesc32 = idesc0 | ((unsigned)(C3_C >> 3) << 17), idesc64 = idesc0 | ((unsigned)(2 * C3_C >> 3) << 17);
            int g = 0, t = 0;
            for (int pass = blockIdx.x; pass < npass; pass += gridDim.x, ++t) {
                const int buf = t & 1;
                if (t >= 2) {
                    tm_mbar_wait(&acc_empty[buf], ((t >> 1) - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const unsigned tmem_t = tmem_d + (unsigned)(buf * C3_COLS);
                for (int kb = 0; kb < a.nkb; ++kb, ++g) {
                    const int s = g % C3_STAGES, use = g / C3_STAGES;
                    tm_mbar_wait(&full_bar[s], use & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    unsigned char* st = stage0 + s * C3_STAGE_BYTES;
#pragma unroll
                    for (int mt = 0; mt < C3_MT; ++mt) {
                        // M tile mt, accumulator kb & 1: 64 columns = hi*hi (32) | cross terms (32)
                        const unsigned d_acc = tmem_t + (unsigned)(mt * 4 * C3_C + (kb & 1) * 2 * C3_C);
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            const unsigned char* ab = st + (mt * C3_M + j) * (C3_C * 4);
                            const unsigned long long a_hi = ct_smem_desc(ab), a_lo = ct_smem_desc(ab + C3_A_BYTES);
                            const unsigned long long b_both = ct_smem_desc(st + 2 * C3_A_BYTES + j * 2 * C3_W_BYTES);
                            for (int k = 0; k < a.ksteps; ++k) {
                                const unsigned long long adv = (unsigned long long)((k * 8 * 4) >> 4);
                                ct_mma_tf32(d_acc, a_hi + adv, b_both + adv, idesc64, ((kb >> 1) | j | k) != 0);
                                ct_mma_tf32(d_acc + C3_C, a_lo + adv, b_both + adv, idesc32, 1);
                            }
                        }
                    }
                    ct_commit(&empty_bar[s]);
                }
                ct_commit(&acc_full[buf]);
            }
        }
    } else {
        // ===== epilogue: warps 2..5, one voxel per thread and M tile =====
        const int q = warp & 3;
        const int m = q * 32 + lane;
        const long long per = (long long)Dp * Hp * Wp;
        int t = 0;
        double st_sum = 0.0, st_sq = 0.0;          // channel `lane`: this warp's share of the batch statistics
        for (int pass = blockIdx.x; pass < npass; pass += gridDim.x, ++t) {
            const int buf = t & 1;
            tm_mbar_wait(&acc_full[buf], (t >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float v[C3_MT][C3_C];
#pragma unroll
            for (int mt = 0; mt < C3_MT; ++mt) {
#pragma unroll
                for (int c = 0; c < C3_C; ++c) v[mt][c] = bias_s[c];
                const unsigned taddr = tmem_d + (unsigned)(buf * C3_COLS + mt * 4 * C3_C) + ((unsigned)(q * 32) << 16);
                // cross terms of both accumulators first, then the two hi*hi sums: small before large
#pragma unroll
                for (int ai = 0; ai < 4; ++ai) {
                    const int col = ai == 0 ? C3_C : ai == 1 ? 3 * C3_C : ai == 2 ? 0 : 2 * C3_C;
#pragma unroll
                    for (int c0 = 0; c0 < C3_C; c0 += 16) {
                        unsigned r[16];
                        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                                       "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),
                                       "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                                     : "r"(taddr + (unsigned)(col + c0)));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[mt][c0 + j] += __uint_as_float(r[j]);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            ct_mbar_arrive(&acc_empty[buf]);
#pragma unroll
            for (int mt = 0; mt < C3_MT; ++mt) {
                const long long p = ((long long)pass * C3_MT + mt) * C3_M + m;
                const bool in_range = p < a.NP;
                const long long b = in_range ? p / per : 0;
                long long rem = p - b * per;
                const int zp = (int)(rem / (Hp * Wp));
                rem -= (long long)zp * (Hp * Wp);
                const int yp = (int)rem / Wp, xp = (int)rem - yp * Wp;
                const bool real = in_range && zp >= 1 && zp <= a.D && yp >= 1 && yp <= a.H && xp >= 1 && xp <= a.W;
                if (a.stats != nullptr) {
                    // sum and sum of squares over the warp's 32 voxels, channel c ending up in lane c: a transposing
                    // butterfly (31 shuffles per quantity), then double accumulation across the CTA's passes
                    float s1[C3_C], s2[C3_C];
#pragma unroll
                    for (int c = 0; c < C3_C; ++c) { s1[c] = real ? v[mt][c] : 0.f; s2[c] = s1[c] * s1[c]; }
#pragma unroll
                    for (int sft = 16; sft >= 1; sft >>= 1) {
                        const bool up = (lane & sft) != 0;
#pragma unroll
                        for (int i = 0; i < sft; ++i) {
                            const float send1 = up ? s1[i] : s1[i + sft], keep1 = up ? s1[i + sft] : s1[i];
                            const float send2 = up ? s2[i] : s2[i + sft], keep2 = up ? s2[i + sft] : s2[i];
                            s1[i] = keep1 + __shfl_xor_sync(0xffffffffu, send1, sft);
                            s2[i] = keep2 + __shfl_xor_sync(0xffffffffu, send2, sft);
                        }
                    }
                    st_sum += (double)s1[0];
                    st_sq += (double)s2[0];
                }
                if (a.out_raw != nullptr && in_range) {
                    float4* o = reinterpret_cast<float4*>(a.out_raw + p * C3_C);
#pragma unroll
                    for (int c = 0; c < C3_C; c += 4)
                        o[c >> 2] = real ? make_float4(v[mt][c], v[mt][c + 1], v[mt][c + 2], v[mt][c + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (a.res_hi != nullptr && real) {
                    const float4* rh = reinterpret_cast<const float4*>(a.res_hi + p * C3_C);
                    const float4* rl = reinterpret_cast<const float4*>(a.res_lo + p * C3_C);
#pragma unroll
                    for (int c = 0; c < C3_C; c += 4) {
                        const float4 h = __ldg(rh + (c >> 2)), l = __ldg(rl + (c >> 2));
                        v[mt][c] += h.x + l.x; v[mt][c + 1] += h.y + l.y; v[mt][c + 2] += h.z + l.z; v[mt][c + 3] += h.w + l.w;
                    }
                }
                if (a.relu) {
#pragma unroll
                    for (int c = 0; c < C3_C; ++c) v[mt][c] = fmaxf(v[mt][c], 0.f);
                }
                if (a.out_hi != nullptr && in_range) {
                    float4* oh = reinterpret_cast<float4*>(a.out_hi + p * C3_C);
                    float4* ol = reinterpret_cast<float4*>(a.out_lo + p * C3_C);
#pragma unroll
                    for (int c = 0; c < C3_C; c += 4) {
                        float4 h, l;
                        h.x = real ? ct_hi(v[mt][c]) : 0.f; h.y = real ? ct_hi(v[mt][c + 1]) : 0.f;
                        h.z = real ? ct_hi(v[mt][c + 2]) : 0.f; h.w = real ? ct_hi(v[mt][c + 3]) : 0.f;
                        l.x = real ? ct_hi(v[mt][c] - h.x) : 0.f; l.y = real ? ct_hi(v[mt][c + 1] - h.y) : 0.f;
                        l.z = real ? ct_hi(v[mt][c + 2] - h.z) : 0.f; l.w = real ? ct_hi(v[mt][c + 3] - h.w) : 0.f;
                        oh[c >> 2] = h; ol[c >> 2] = l;
                    }
                }
                if (a.out_c0 != nullptr && real)
                    a.out_c0[((b * a.D + (zp - 1)) * a.H + (yp - 1)) * a.W + (xp - 1)] = v[mt][0];
            }
        }
        if (a.stats != nullptr) {
            atomicAdd(a.stats + lane, st_sum);
            atomicAdd(a.stats + C3_C + lane, st_sq);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((unsigned)(2 * C3_COLS)) : "memory");
    }
}

// Position decode for the elementwise kernels: the host guarantees fewer than 2^31 positions, so 32-bit divisions do.
struct C3Pos { unsigned b; int zp, yp, xp; };
__device__ __forceinline__ C3Pos c3_decode(unsigned p, int Dp, int Hp, int Wp) {
    C3Pos r;
    const unsigned plane = (unsigned)(Hp * Wp), per = (unsigned)Dp * plane;
    r.b = p / per;
    unsigned rem = p - r.b * per;
    r.zp = (int)(rem / plane);
    rem -= (unsigned)r.zp * plane;
    r.yp = (int)(rem / (unsigned)Wp);
    r.xp = (int)(rem - (unsigned)r.yp * (unsigned)Wp);
    return r;
}

// NCDHW fp32 [B][C][D][H][W] (C <= 32) -> packed hi / lo [B][D+2][H+2][W+2][32], zero border, channels >= C zero.
// One thread per (position, group of 4 channels): the packed stores of a warp are 512 contiguous bytes.
__global__ void __launch_bounds__(256) conv3d_pack_kernel(const float* __restrict__ x, float* __restrict__ hi,
                                                          float* __restrict__ lo, int B, int C, int D, int H, int W) {
    const int Wp = W + 2, Hp = H + 2, Dp = D + 2;
    const long long per = (long long)Dp * Hp * Wp;
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long p = t >> 3;
    const int c = (int)(t & 7) * 4;
    if (p >= (long long)B * per) return;
    const C3Pos q = c3_decode((unsigned)p, Dp, Hp, Wp);
    const long long b = q.b;
    const int zp = q.zp, yp = q.yp, xp = q.xp;
    const bool real = zp >= 1 && zp <= D && yp >= 1 && yp <= H && xp >= 1 && xp <= W;
    const long long DHW = (long long)D * H * W;
    const float* src = x + b * C * DHW + ((long long)(zp - 1) * H + (yp - 1)) * W + (xp - 1);
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (real && c + j < C) ? __ldg(src + (c + j) * DHW) : 0.f;
    float4 h, l;
    h.x = ct_hi(v[0]); h.y = ct_hi(v[1]); h.z = ct_hi(v[2]); h.w = ct_hi(v[3]);
    l.x = ct_hi(v[0] - h.x); l.y = ct_hi(v[1] - h.y); l.z = ct_hi(v[2] - h.z); l.w = ct_hi(v[3] - h.w);
    reinterpret_cast<float4*>(hi)[t] = h;
    reinterpret_cast<float4*>(lo)[t] = l;
}

// The first layer's input, z-folded: channel dzi * C + c of position (z, y, x) holds x[c] at (z + dzi - 1, y, x)
// (3 C <= 32; zero outside the volume).  The convolution over dz then is part of the channel contraction: 3 K-blocks
// (dy) of two K steps instead of 9 of one -- a third of the copies for a layer whose cost is its copies.
__global__ void __launch_bounds__(256) conv3d_pack_zfold_kernel(const float* __restrict__ x, float* __restrict__ hi,
                                                                float* __restrict__ lo, int B, int C, int D, int H, int W) {
    const int Wp = W + 2, Hp = H + 2, Dp = D + 2;
    const long long per = (long long)Dp * Hp * Wp;
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long p = t >> 3;
    const int k0 = (int)(t & 7) * 4;
    if (p >= (long long)B * per) return;
    const C3Pos q = c3_decode((unsigned)p, Dp, Hp, Wp);
    const long long b = q.b;
    const int zp = q.zp, yp = q.yp, xp = q.xp;
    const bool real = zp >= 1 && zp <= D && yp >= 1 && yp <= H && xp >= 1 && xp <= W;
    const long long DHW = (long long)D * H * W, HW = (long long)H * W;
    const float* src = x + b * C * DHW + (long long)(yp - 1) * W + (xp - 1);
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int k = k0 + j, dzi = (k >= C) + (k >= 2 * C), c = k - dzi * C;      // (k < 3 C: no division)
        const int z = zp - 1 + dzi - 1;
        v[j] = (real && k < 3 * C && z >= 0 && z < D) ? __ldg(src + c * DHW + z * HW) : 0.f;
    }
    float4 h, l;
    h.x = ct_hi(v[0]); h.y = ct_hi(v[1]); h.z = ct_hi(v[2]); h.w = ct_hi(v[3]);
    l.x = ct_hi(v[0] - h.x); l.y = ct_hi(v[1] - h.y); l.z = ct_hi(v[2] - h.z); l.w = ct_hi(v[3] - h.w);
    reinterpret_cast<float4*>(hi)[t] = h;
    reinterpret_cast<float4*>(lo)[t] = l;
}

// The matching filter: tap slot dy * 3 + dx (the first 9 of 27), input channel dzi * C_in + c.
__global__ void __launch_bounds__(256) conv3d_pack_weights_zfold_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                                                        float* __restrict__ hi, float* __restrict__ lo,
                                                                        int C_out, int C_in) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= 27 * C3_C * C3_C) return;
    const int k = i % C3_C, o = (i / C3_C) % C3_C, t = i / (C3_C * C3_C);
    const int dzi = k / C_in, c = k - dzi * C_in;
    float v = 0.f;
    if (t < 9 && o < C_out && dzi < 3) {
        v = __ldg(w + ((long long)o * C_in + c) * 27 + dzi * 9 + t);
        if (scale != nullptr) v = __fmul_rn(v, __ldg(scale + o));
    }
    const float h = ct_hi(v);
    hi[i] = h; lo[i] = ct_hi(v - h);
}

// BatchNorm3d with BATCH statistics (training mode, or track_running_stats = False: torch.nn.functional.batch_norm
// with training = True) applied to a raw convolution output: y = (x - mean) / sqrt(var + eps) * gamma + beta with the
// biased variance over the real voxels, then + residual, ReLU, and the split into the next layer's packed hi / lo.
// stats = the sums the convolution kernel accumulated; count = B * D * H * W.
__global__ void __launch_bounds__(256) conv3d_bn_apply_kernel(const float* __restrict__ raw, const double* __restrict__ stats,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              float eps, const float* __restrict__ res_hi,
                                                              const float* __restrict__ res_lo, float* __restrict__ hi,
                                                              float* __restrict__ lo, int B, int D, int H, int W, int relu) {
    __shared__ float sc_s[C3_C], sh_s[C3_C];
    const int Wp = W + 2, Hp = H + 2, Dp = D + 2;
    const long long per = (long long)Dp * Hp * Wp;
    if (threadIdx.x < C3_C) {
        const double n = (double)B * D * H * W;
        const double mean = stats[threadIdx.x] / n;
        double var = stats[C3_C + threadIdx.x] / n - mean * mean;
        var = var > 0.0 ? var : 0.0;
        const double g = gamma != nullptr ? (double)__ldg(gamma + threadIdx.x) : 1.0;
        const double bt = beta != nullptr ? (double)__ldg(beta + threadIdx.x) : 0.0;
        const double sc = g / sqrt(var + (double)eps);
        sc_s[threadIdx.x] = (float)sc;
        sh_s[threadIdx.x] = (float)(bt - mean * sc);
    }
    __syncthreads();
    // one thread per (position, group of 4 channels): every load and store of a warp is 512 contiguous bytes
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long p = t >> 3;
    const int c = (int)(t & 7) * 4;
    if (p >= (long long)B * per) return;
    const C3Pos q = c3_decode((unsigned)p, Dp, Hp, Wp);
    const long long b = q.b;
    const int zp = q.zp, yp = q.yp, xp = q.xp;
    const bool real = zp >= 1 && zp <= D && yp >= 1 && yp <= H && xp >= 1 && xp <= W;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (real) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(raw) + t);
        v[0] = fmaf(x.x, sc_s[c], sh_s[c]); v[1] = fmaf(x.y, sc_s[c + 1], sh_s[c + 1]);
        v[2] = fmaf(x.z, sc_s[c + 2], sh_s[c + 2]); v[3] = fmaf(x.w, sc_s[c + 3], sh_s[c + 3]);
        if (res_hi != nullptr) {
            const float4 h = __ldg(reinterpret_cast<const float4*>(res_hi) + t);
            const float4 l = __ldg(reinterpret_cast<const float4*>(res_lo) + t);
            v[0] += h.x + l.x; v[1] += h.y + l.y; v[2] += h.z + l.z; v[3] += h.w + l.w;
        }
        if (relu) { v[0] = fmaxf(v[0], 0.f); v[1] = fmaxf(v[1], 0.f); v[2] = fmaxf(v[2], 0.f); v[3] = fmaxf(v[3], 0.f); }
    }
    float4 h, l;
    h.x = ct_hi(v[0]); h.y = ct_hi(v[1]); h.z = ct_hi(v[2]); h.w = ct_hi(v[3]);
    l.x = ct_hi(v[0] - h.x); l.y = ct_hi(v[1] - h.y); l.z = ct_hi(v[2] - h.z); l.w = ct_hi(v[3] - h.w);
    reinterpret_cast<float4*>(hi)[t] = h;
    reinterpret_cast<float4*>(lo)[t] = l;
}

// The classifier, Conv3d(32, 1, 3, padding 1, bias=False) (models/models.py:403): one output channel is a dot product
// of 864 terms per voxel -- nothing for a 128-row MMA whose cost does not shrink with N (section 4.6 of DESIGN.md), so it
// runs on the FP32 pipe: 128 consecutive padded positions per CTA; per (dz, dy) pair the 130 rows the three dx taps need
// are staged as value = hi + lo in shared memory, channel-major with a row stride of 131 floats (staging stores and
// compute loads are both bank-conflict-free); the 864 weights travel as LAUNCH PARAMETERS, so every FMA takes its weight
// from the constant bank and the inner loop is one LDS + one FFMA per term.
constexpr int C1_TP = 128, C1_ROWS = C1_TP + 2, C1_STRIDE = 131;
struct Conv1Weights { float w[27 * C3_C]; };      // [tap][input channel]
__global__ void __launch_bounds__(C1_TP) conv3d_c32_to1_kernel(const float* __restrict__ in_hi, const float* __restrict__ in_lo,
                                                               float* __restrict__ out, int B, int D, int H, int W,
                                                               const __grid_constant__ Conv1Weights wt) {
    __shared__ float s[C3_C * C1_STRIDE];
    const int Wp = W + 2, Hp = H + 2, Dp = D + 2;
    const long long per = (long long)Dp * Hp * Wp, NP = (long long)B * per;
    const long long p0 = (long long)blockIdx.x * C1_TP;
    const int tid = threadIdx.x;
    float acc = 0.f;
#pragma unroll
    for (int g = 0; g < 9; ++g) {
        const long long r0 = p0 + (long long)(g / 3 - 1) * Hp * Wp + (g % 3 - 1) * Wp - 1;      // first staged row
        // 130 rows x 8 float4: thread t takes float4 (t % 8) of rows t / 8, t / 8 + 16, ...
        for (int i = tid; i < C1_ROWS * 8; i += C1_TP) {
            const int row = i >> 3, c4 = i & 7;
            const long long r = r0 + row;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r >= 0 && r < NP) {
                const float4 h = __ldg(reinterpret_cast<const float4*>(in_hi + r * C3_C) + c4);
                const float4 l = __ldg(reinterpret_cast<const float4*>(in_lo + r * C3_C) + c4);
                v = make_float4(h.x + l.x, h.y + l.y, h.z + l.z, h.w + l.w);
            }
            float* d = s + (c4 * 4) * C1_STRIDE + row;
            d[0] = v.x; d[C1_STRIDE] = v.y; d[2 * C1_STRIDE] = v.z; d[3 * C1_STRIDE] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < C3_C; ++c) {
            const float* sc = s + c * C1_STRIDE + tid;
            acc = fmaf(sc[0], wt.w[(g * 3 + 0) * C3_C + c], acc);
            acc = fmaf(sc[1], wt.w[(g * 3 + 1) * C3_C + c], acc);
            acc = fmaf(sc[2], wt.w[(g * 3 + 2) * C3_C + c], acc);
        }
        __syncthreads();
    }
    const long long p = p0 + tid;
    if (p >= NP) return;
    const long long b = p / per;
    long long rem = p - b * per;
    const int zp = (int)(rem / (Hp * Wp));
    rem -= (long long)zp * (Hp * Wp);
    const int yp = (int)rem / Wp, xp = (int)rem - yp * Wp;
    if (zp >= 1 && zp <= D && yp >= 1 && yp <= H && xp >= 1 && xp <= W)
        out[((b * D + (zp - 1)) * H + (yp - 1)) * W + (xp - 1)] = acc;
}

// torch weight [C_out][C_in][3][3][3] (C_out, C_in <= 32), optional per-output-channel scale (the folded BatchNorm)
// -> packed hi / lo [27 taps][32 out][32 in], zero where out >= C_out or in >= C_in.
__global__ void __launch_bounds__(256) conv3d_pack_weights_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                                                  float* __restrict__ hi, float* __restrict__ lo,
                                                                  int C_out, int C_in) {
    const int i = blockIdx.x * 256 + threadIdx.x;           // over 27 * 32 * 32
    if (i >= 27 * C3_C * C3_C) return;
    const int c = i % C3_C, o = (i / C3_C) % C3_C, t = i / (C3_C * C3_C);
    float v = 0.f;
    if (o < C_out && c < C_in) {
        v = __ldg(w + ((long long)o * C_in + c) * 27 + t);
        if (scale != nullptr) v = __fmul_rn(v, __ldg(scale + o));
    }
    const float h = ct_hi(v);
    hi[i] = h; lo[i] = ct_hi(v - h);
}

static bool c3_encode(tm_encode_fn enc, CUtensorMap* m, const float* base, cuuint64_t rows, cuuint32_t box_rows) {
    const cuuint64_t gdim[2] = {(cuuint64_t)C3_C, rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)C3_C * 4};
    const cuuint32_t box[2] = {C3_C, box_rows};
    const cuuint32_t est[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, est,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace dpv

extern "C" int64_t dpv_conv3d_packed_floats(int B, int D, int H, int W) {
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
    return (int64_t)B * (D + 2) * (H + 2) * (W + 2) * dpv::C3_C;
}

extern "C" int dpv_conv3d_pack(const float* x, float* packed_hi, float* packed_lo, int B, int C, int D, int H, int W,
                               void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(x && packed_hi && packed_lo && B > 0 && D > 0 && H > 0 && W > 0 && C > 0);
    if (C > C3_C) return DPV_E_UNSUPP;
    const long long n = (long long)B * (D + 2) * (H + 2) * (W + 2);
    if (n > (1LL << 31) - 4096) return DPV_E_UNSUPP;
    conv3d_pack_kernel<<<(unsigned)((n * 8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, packed_hi, packed_lo, B, C, D, H, W);
    DPV_LAUNCH_END();
    return 0;
}

extern "C" int dpv_conv3d_pack_weights(const float* weight, const float* scale, float* w_hi, float* w_lo, int C_out,
                                       int C_in, void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(weight && w_hi && w_lo && C_out > 0 && C_in > 0);
    if (C_out > C3_C || C_in > C3_C) return DPV_E_UNSUPP;
    conv3d_pack_weights_kernel<<<(27 * C3_C * C3_C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(weight, scale, w_hi, w_lo, C_out, C_in);
    DPV_LAUNCH_END();
    return 0;
}

extern "C" int dpv_conv3d_c32(const float* in_hi, const float* in_lo, const float* w_hi, const float* w_lo,
                              const float* shift, const float* res_hi, const float* res_lo, float* out_hi,
                              float* out_lo, float* out_c0, float* out_raw, double* stats, int B, int D, int H, int W,
                              int relu, int c_in, void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(in_hi && in_lo && w_hi && w_lo && B > 0 && D > 0 && H > 0 && W > 0);
    DPV_CHECK_ARG((out_hi == nullptr) == (out_lo == nullptr));
    DPV_CHECK_ARG((res_hi == nullptr) == (res_lo == nullptr));
    DPV_CHECK_ARG(out_hi != nullptr || out_c0 != nullptr || out_raw != nullptr);
    DPV_CHECK_ARG(c_in > 0 && c_in <= C3_C);
    const long long np = (long long)B * (D + 2) * (H + 2) * (W + 2);
    // TMA row coordinates are 32-bit: the lowest / highest row any tap touches must fit
    if (np + (long long)(H + 2) * (W + 2) + (W + 2) + 2 * C3_BOX >= (1LL << 31)) return DPV_E_UNSUPP;
    if (((uintptr_t)in_hi | (uintptr_t)in_lo | (uintptr_t)w_hi | (uintptr_t)w_lo | (uintptr_t)out_hi | (uintptr_t)out_lo |
         (uintptr_t)res_hi | (uintptr_t)res_lo | (uintptr_t)out_raw) & 15)
        return DPV_E_BADARG;
    tm_encode_fn enc = tm_encoder();
    if (enc == nullptr) return DPV_E_UNSUPP;
    Conv3Maps maps;
    if (!c3_encode(enc, &maps.a_hi, in_hi, (cuuint64_t)np, C3_BOX) || !c3_encode(enc, &maps.a_lo, in_lo, (cuuint64_t)np, C3_BOX) ||
        !c3_encode(enc, &maps.w_hi, w_hi, 27 * C3_C, C3_C) || !c3_encode(enc, &maps.w_lo, w_lo, 27 * C3_C, C3_C))
        return DPV_E_UNSUPP;
    Conv3Args a;
    a.bias = shift; a.res_hi = res_hi; a.res_lo = res_lo; a.out_hi = out_hi; a.out_lo = out_lo; a.out_c0 = out_c0;
    a.out_raw = out_raw; a.stats = stats;
    a.B = B; a.D = D; a.H = H; a.W = W; a.NP = np; a.relu = (relu & 1) ? 1 : 0;
    a.zfold = (relu & 4) ? 1 : 0;
    a.nkb = a.zfold ? 3 : C3_NKB;
    a.ksteps = (c_in + 7) / 8;
    const size_t smem = (size_t)C3_STAGES * C3_STAGE_BYTES + 1024;
    cudaError_t e = cudaFuncSetAttribute(conv3d_c32_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            return DPV_E_UNSUPP;
        n_sm = n;
    }
    const long long npass = (np + C3_MT * C3_M - 1) / (C3_MT * C3_M);
    const unsigned grid = npass < n_sm ? (unsigned)npass : (unsigned)n_sm;
    conv3d_c32_tc_kernel<<<grid, C3_THREADS, smem, (cudaStream_t)stream>>>(a, maps);
    DPV_LAUNCH_END();
    return 0;
}

extern "C" int dpv_conv3d_bn_apply(const float* raw, const double* stats, const float* gamma, const float* beta, float eps,
                                   const float* res_hi, const float* res_lo, float* out_hi, float* out_lo, int B, int D,
                                   int H, int W, int relu, void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(raw && stats && out_hi && out_lo && B > 0 && D > 0 && H > 0 && W > 0);
    DPV_CHECK_ARG((res_hi == nullptr) == (res_lo == nullptr));
    if (((uintptr_t)raw | (uintptr_t)out_hi | (uintptr_t)out_lo | (uintptr_t)res_hi | (uintptr_t)res_lo) & 15) return DPV_E_BADARG;
    const long long n = (long long)B * (D + 2) * (H + 2) * (W + 2);
    conv3d_bn_apply_kernel<<<(unsigned)((n * 8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(raw, stats, gamma, beta, eps, res_hi,
                                                                                       res_lo, out_hi, out_lo, B, D, H, W, relu ? 1 : 0);
    DPV_LAUNCH_END();
    return 0;
}

extern "C" int dpv_conv3d_c32_to1(const float* in_hi, const float* in_lo, const float* weight_host, float* out, int B, int D,
                                  int H, int W, int c_in, void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(in_hi && in_lo && weight_host && out && B > 0 && D > 0 && H > 0 && W > 0);
    DPV_CHECK_ARG(c_in > 0 && c_in <= C3_C);
    if (((uintptr_t)in_hi | (uintptr_t)in_lo) & 15) return DPV_E_BADARG;
    const long long np = (long long)B * (D + 2) * (H + 2) * (W + 2);
    if (np > (1LL << 31) - 4096) return DPV_E_UNSUPP;
    Conv1Weights wt;                                 // torch layout [1][c_in][3][3][3] -> [tap][32], zero beyond c_in
    for (int t = 0; t < 27; ++t)
        for (int c = 0; c < C3_C; ++c) wt.w[t * C3_C + c] = c < c_in ? weight_host[c * 27 + t] : 0.f;
    conv3d_c32_to1_kernel<<<(unsigned)((np + C1_TP - 1) / C1_TP), C1_TP, 0, (cudaStream_t)stream>>>(in_hi, in_lo, out, B, D, H, W, wt);
    DPV_LAUNCH_END();
    return 0;
}

extern "C" int dpv_conv3d_pack_zfold(const float* x, float* packed_hi, float* packed_lo, int B, int C, int D, int H, int W,
                                     void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(x && packed_hi && packed_lo && B > 0 && D > 0 && H > 0 && W > 0 && C > 0);
    if (3 * C > C3_C) return DPV_E_UNSUPP;
    const long long n = (long long)B * (D + 2) * (H + 2) * (W + 2);
    if (n > (1LL << 31) - 4096) return DPV_E_UNSUPP;
    conv3d_pack_zfold_kernel<<<(unsigned)((n * 8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, packed_hi, packed_lo, B, C, D, H, W);
    DPV_LAUNCH_END();
    return 0;
}

extern "C" int dpv_conv3d_pack_weights_zfold(const float* weight, const float* scale, float* w_hi, float* w_lo, int C_out,
                                             int C_in, void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(weight && w_hi && w_lo && C_out > 0 && C_in > 0);
    if (C_out > C3_C || 3 * C_in > C3_C) return DPV_E_UNSUPP;
    conv3d_pack_weights_zfold_kernel<<<(27 * C3_C * C3_C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(weight, scale, w_hi, w_lo, C_out, C_in);
    DPV_LAUNCH_END();
    return 0;
}
