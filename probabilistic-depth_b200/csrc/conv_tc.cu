// the three D -> D 3x3 convolutions between the cost volume and the 1/4-res soft-max
// (models/models.py:456-460 conv0 / conv0_1 / conv0_2, applied at :555-560 and :632-637), D = 64 channels in and
// out, stride 1, zero padding 1, bias, LeakyReLU(0.01) after the first two (models/models.py:38-46), log_softmax
// over the channels (= depth bins) after the third.
//
// The only dense contraction on the path: per convolution 2 * 64 * 64 * 9 flops per pixel, 3.6 GFLOP per batch
// of 8 x 64 x 96 -- tensor-core work.  Implicit GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulator
// in TMEM), fp32 parity through split precision:
//   * Activations live in a packed form [B][H+2][W+2][64] (channels innermost, one pixel of zero padding all
//     round), twice: hi = the TF32 value nearest to it, lo = the TF32 value nearest to (value - hi).  In that form the
//     A operand of filter tap (dy, dx) is the SAME matrix [pixels][64] shifted by dy * (W+2) + dx rows, so one 2-D
//     TMA map (box 32 channels x 128 pixels, 128-byte swizzle = the K-major canonical layout of tcgen05) serves all
//     nine taps and any shift; rows before the first / after the last pixel are zero-filled by the copy engine.
//     Outputs are computed for all padded positions (98 % are real pixels); border positions are written as zero,
//     which IS the next convolution's padding.
//   * D[p][o] = sum over 9 taps x 64 channels.  A K-block is (filter row, 32 channels): ONE A tile of 136 pixels
//     serves the row's three taps -- the dx = 0 / +1 taps read the same shared-memory tile one and two 128-byte rows
//     further down (descriptor start address) -- so the activations are streamed 3 times per tile, not 9.  Three
//     product terms  hi*hi + hi*lo + lo*hi  (the dropped lo*lo term is 2^-22 relative) in TWO tcgen05.mma per tap and
//     K step (kind::tf32, M = 128, K = 8): A_hi x [W_hi ; W_lo] with N = 128 (the two weight tiles sit next to each
//     other in the stage, so one descriptor covers both and A_hi is read once) and A_lo x W_hi with N = 64 into the
//     cross-term columns.  144 MMAs per 128-pixel tile, issued by ONE thread.  The tensor core adds into its fp32
//     accumulator with truncation, a bias that grows with the number of accumulation steps at full magnitude
//     (one accumulator for all steps: 1.2e-5 of the logits' scale, measured): even and odd K-blocks therefore
//     accumulate into two separate TMEM tiles (36 steps each), and the epilogue adds cross terms and the two
//     hi*hi sums in fp32 with rounding, small before large.
//   * Persistent CTAs (one per SM, tiles round-robin), accumulators double-buffered in TMEM (2 x 256 columns): the
//     epilogue of a tile overlaps the copies and MMAs of the next.  Measured on the way (B=8, 64x96, three layers
//     + pack, graph-timed): one tile per CTA 0.1245 ms -> persistent + double-buffered 0.0982 -> N = 128 MMAs
//     0.0857.  Timing experiments on the one-tile kernel: without the copies after the first two K-blocks 0.1230
//     (copies were hidden), without the MMAs 0.0794 (= copies + epilogues).  ncu of the persistent kernel: tensor
//     pipe busy 55-62 % of the SM-active time (an M = 128, K = 8 TF32 MMA occupies it ~50 cycles at N <= 64, ~64 at
//     N = 128), 204 MB per layer from L2 at 7 TB/s underneath, a quarter of the launch ramp and tail (405 tiles on
//     148 SMs).
//   * Warp roles: warp 0 = TMA producer (per K-block A_hi, A_lo and W_hi / W_lo of three taps; 82 KB per stage, 2 stages),
//     warp 1 = TMEM allocation + MMA issue (tcgen05.commit releases a stage / publishes the accumulators),
//     warps 2-5 = epilogue: tcgen05.ld gives every thread ONE pixel with all 64 output channels in registers, so
//     bias + LeakyReLU + re-split into the next layer's packed hi / lo, or bias + log-softmax over the 64 depth
//     bins, happens in registers with no cross-thread step; NCHW stores are coalesced across the warp's 32 pixels.
#include <cuda.h>

#include <cstdlib>

#include "conv_tc.cuh"     // TMA / tcgen05 helpers (+ sweep_tma.cuh: mbarrier wrappers, tm_encoder())

namespace dpv {

constexpr int CT_C = 64;                   // channels in = channels out = depth bins
constexpr int CT_M = 128;                  // pixels per tile (UMMA M)
constexpr int CT_KB = 32;                  // channels per K-block: 32 x 4 B = one 128-byte swizzle row
constexpr int CT_NKB = 3 * (CT_C / CT_KB); // 6 K-blocks: (filter row, channel half); each serves the row's 3 taps
constexpr int CT_STAGES = 2;
constexpr int CT_AROWS = CT_M + 8;         // pixels per A tile: 128 + the two neighbours of the dx = -1 / +1 taps (+ pad)
constexpr int CT_A_BYTES = CT_AROWS * CT_KB * 4;    // 17 KB (a multiple of the 1024-byte swizzle period)
constexpr int CT_B_BYTES = CT_C * CT_KB * 4;        // 8 KB per tap
constexpr int CT_STAGE_BYTES = 2 * CT_A_BYTES + 6 * CT_B_BYTES;   // A hi / lo + W hi / lo of three taps = 82 KB
constexpr int CT_THREADS = 192;            // warp 0 producer, warp 1 MMA, warps 2-5 epilogue
constexpr int CT_NACC = 4;                 // 64-column accumulator tiles per buffer: (hi*hi | cross) of the even and of the odd K-blocks
constexpr int CT_TMEM_COLS = CT_NACC * CT_C;   // 256 columns per buffer: four fp32 tiles of 128 lanes x 64 columns; two buffers

struct ConvMaps {
    CUtensorMap a_hi, a_lo;   // [NP pixels][64 ch] packed activations, box {32, 128}, SWIZZLE_128B
    CUtensorMap w_hi, w_lo;   // [9 * 64 (tap, out)][64 in] packed weights, box {32, 64}, SWIZZLE_128B
};

struct ConvArgs {
    const float* bias;        // [64]
    float* out_hi; float* out_lo;   // packed [NP][64] (next layer's input), nullable
    float* out_nchw;          // [B][64][H][W], nullable
    int B, H, W, NP;          // NP = B * (H + 2) * (W + 2)
    int epilogue;             // 0 = bias, 1 = bias + LeakyReLU(slope), 2 = bias + log_softmax over the channels
    float slope;
};

// Persistent: CTA i takes tiles i, i + gridDim.x, ...  The accumulators are double-buffered in TMEM (2 x 256
// columns = all 512), so the epilogue of one tile (TMEM -> registers -> global) runs while the copy engine and the
// tensor core are already on the next; the stage ring keeps running across tiles.
__global__ void __launch_bounds__(CT_THREADS, 1)
conv3x3_d64_tc_kernel(const ConvArgs a, const __grid_constant__ ConvMaps maps) {
    extern __shared__ __align__(1024) unsigned char ct_smem[];
    __shared__ unsigned long long full_bar[CT_STAGES], empty_bar[CT_STAGES], acc_full[2], acc_empty[2];
    __shared__ unsigned tmem_base_s;
    __shared__ float bias_s[CT_C];
    // dynamic shared memory may start at any 16-byte boundary: the swizzled tiles need 1024
    unsigned char* stage0 = (unsigned char*)(((uintptr_t)ct_smem + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Wp = a.W + 2, Hp = a.H + 2;
    const int ntiles = (a.NP + CT_M - 1) / CT_M;

    if (threadIdx.x < CT_C) bias_s[threadIdx.x] = __ldg(a.bias + threadIdx.x);
    if (threadIdx.x == 0) {
        for (int s = 0; s < CT_STAGES; ++s) { tm_mbar_init(&full_bar[s], 1); tm_mbar_init(&empty_bar[s], 1); }
        for (int i = 0; i < 2; ++i) { tm_mbar_init(&acc_full[i], 1); tm_mbar_init(&acc_empty[i], CT_M); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {      // TMEM: one warp allocates (and later frees) the accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(tm_smem(&tmem_base_s)), "r"((unsigned)(2 * CT_TMEM_COLS)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_d = tmem_base_s;

    if (warp == 0) {
        // ===== TMA producer (one elected lane) =====
        if (lane == 0) {
            int g = 0;                                                      // K-blocks issued so far (ring position)
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int p0 = tile * CT_M;                                 // first padded-grid position of the tile
                for (int kb = 0; kb < CT_NKB; ++kb, ++g) {
                    const int s = g % CT_STAGES, use = g / CT_STAGES;
                    if (use > 0) tm_mbar_wait(&empty_bar[s], (use - 1) & 1);    // the MMAs of the previous use are done
                    const int frow = kb >> 1, half = kb & 1;                // filter row dy = frow - 1
                    unsigned char* st = stage0 + s * CT_STAGE_BYTES;
                    tm_mbar_expect_tx(&full_bar[s], CT_STAGE_BYTES);
                    const int row = p0 + (frow - 1) * Wp - 1;               // dx = -1 tap's first pixel; < 0 / past NP: zero-filled
                    ct_tma_2d(st, &maps.a_hi, half * CT_KB, row, &full_bar[s]);
                    ct_tma_2d(st + CT_A_BYTES, &maps.a_lo, half * CT_KB, row, &full_bar[s]);
                    for (int j = 0; j < 3; ++j) {                           // the row's three taps
                        unsigned char* wb = st + 2 * CT_A_BYTES + j * 2 * CT_B_BYTES;
                        ct_tma_2d(wb, &maps.w_hi, half * CT_KB, (frow * 3 + j) * CT_C, &full_bar[s]);
                        ct_tma_2d(wb + CT_B_BYTES, &maps.w_lo, half * CT_KB, (frow * 3 + j) * CT_C, &full_bar[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one elected lane) =====
        if (lane == 0) {
            // instruction descriptors: D fp32, A / B TF32, both K-major, M = 128, N = 64 / 128
            const unsigned idesc0 = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(CT_M >> 4) << 24);
            const unsigned idesc64 = idesc0 | ((unsigned)(CT_C >> 3) << 17), idesc128 = idesc0 | ((unsigned)(2 * CT_C >> 3) << 17);
            int g = 0, t = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
                const int buf = t & 1;
                if (t >= 2) {                      // the epilogue has read this buffer's previous tile out of TMEM
                    tm_mbar_wait(&acc_empty[buf], ((t >> 1) - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const unsigned tmem_t = tmem_d + (unsigned)(buf * CT_TMEM_COLS);
                for (int kb = 0; kb < CT_NKB; ++kb, ++g) {
                    const int s = g % CT_STAGES, use = g / CT_STAGES;
                    tm_mbar_wait(&full_bar[s], use & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    unsigned char* st = stage0 + s * CT_STAGE_BYTES;
                    // even K-blocks accumulate into columns [0, 128), odd ones into [128, 256): in each, the hi*hi
                    // sum in the first 64 columns and the cross terms in the last 64
                    const unsigned d_acc = tmem_t + (unsigned)((kb & 1) * 2 * CT_C);
#pragma unroll
                    for (int j = 0; j < 3; ++j) {                          // dx = j - 1: the A tile, j rows further down
                        const unsigned long long a_hi = ct_smem_desc(st + j * 128), a_lo = ct_smem_desc(st + CT_A_BYTES + j * 128);
                        const unsigned char* wb = st + 2 * CT_A_BYTES + j * 2 * CT_B_BYTES;
                        const unsigned long long b_both = ct_smem_desc(wb);      // W hi (64 rows) followed by W lo (64 rows)
#pragma unroll
                        for (int k = 0; k < CT_KB / 8; ++k) {
                            const unsigned long long adv = (unsigned long long)((k * 8 * 4) >> 4);   // 32 bytes along K
                            ct_mma_tf32(d_acc, a_hi + adv, b_both + adv, idesc128, ((kb >> 1) | j | k) != 0);   // hi*hi | hi*lo
                            ct_mma_tf32(d_acc + CT_C, a_lo + adv, b_both + adv, idesc64, 1);                    // lo*hi
                        }
                    }
                    ct_commit(&empty_bar[s]);      // frees the stage when these MMAs have read it
                }
                ct_commit(&acc_full[buf]);         // this tile's accumulators are complete
            }
        }
    } else {
        // ===== epilogue: warps 2..5, one pixel per thread, 64 channels in registers =====
        const int q = warp & 3;                    // TMEM lane quarter this warp may read
        const int m = q * 32 + lane;               // row of the tile
        const int per = Hp * Wp;
        int t = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
            const int buf = t & 1;
            const int p = tile * CT_M + m;         // padded-grid position
            tm_mbar_wait(&acc_full[buf], (t >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float v[CT_C];
#pragma unroll
            for (int c = 0; c < CT_C; ++c) v[c] = bias_s[c];
            {
                const unsigned taddr = tmem_d + (unsigned)(buf * CT_TMEM_COLS) + ((unsigned)(q * 32) << 16);
                // cross terms first (columns 64.., 192..), then the two hi*hi sums: small before large
#pragma unroll
                for (int ai = 0; ai < CT_NACC; ++ai) {
                    const int acc = ai == 0 ? 1 : ai == 1 ? 3 : ai == 2 ? 0 : 2;
#pragma unroll
                    for (int c0 = 0; c0 < CT_C; c0 += 16) {
                        unsigned r[16];
                        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                                       "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),
                                       "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                                     : "r"(taddr + (unsigned)(acc * CT_C + c0)));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[c0 + j] += __uint_as_float(r[j]);
                    }
                }
            }
            // the accumulators are in registers: hand the TMEM buffer back to the MMA warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            ct_mbar_arrive(&acc_empty[buf]);
            // where is this position on the padded grid
            const bool in_range = p < a.NP;
            const int b = in_range ? p / per : 0;
            const int rem = p - b * per;
            const int yp = rem / Wp, xp = rem - yp * Wp;
            const bool real = in_range && yp >= 1 && yp <= a.H && xp >= 1 && xp <= a.W;
            if (a.epilogue == 1) {
#pragma unroll
                for (int c = 0; c < CT_C; ++c) v[c] = v[c] > 0.f ? v[c] : v[c] * a.slope;
            } else if (a.epilogue == 2) {
                float mx = v[0];
#pragma unroll
                for (int c = 1; c < CT_C; ++c) mx = fmaxf(mx, v[c]);
                float s = 0.f;
#pragma unroll
                for (int c = 0; c < CT_C; ++c) s += __expf(v[c] - mx);
                const float ls = logf(s);
#pragma unroll
                for (int c = 0; c < CT_C; ++c) v[c] = (v[c] - mx) - ls;
            }
            if (a.out_hi != nullptr && in_range) {     // the next layer's packed input; border positions = its zero padding
                float4* oh = reinterpret_cast<float4*>(a.out_hi + (long long)p * CT_C);
                float4* ol = reinterpret_cast<float4*>(a.out_lo + (long long)p * CT_C);
#pragma unroll
                for (int c = 0; c < CT_C; c += 4) {
                    float4 h, l;
                    h.x = real ? ct_hi(v[c]) : 0.f; h.y = real ? ct_hi(v[c + 1]) : 0.f;
                    h.z = real ? ct_hi(v[c + 2]) : 0.f; h.w = real ? ct_hi(v[c + 3]) : 0.f;
                    l.x = real ? ct_hi(v[c] - h.x) : 0.f; l.y = real ? ct_hi(v[c + 1] - h.y) : 0.f;
                    l.z = real ? ct_hi(v[c + 2] - h.z) : 0.f; l.w = real ? ct_hi(v[c + 3] - h.w) : 0.f;
                    oh[c >> 2] = h; ol[c >> 2] = l;
                }
            }
            if (a.out_nchw != nullptr && real) {
                const long long HW = (long long)a.H * a.W;
                float* o = a.out_nchw + (long long)b * CT_C * HW + (long long)(yp - 1) * a.W + (xp - 1);
#pragma unroll
                for (int c = 0; c < CT_C; ++c) o[c * HW] = v[c];
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((unsigned)(2 * CT_TMEM_COLS)) : "memory");
    }
}

// NCHW fp32 [B][64][H][W] -> packed hi / lo [B][H+2][W+2][64] with a zero border.  One thread per padded
// position: 64 strided loads (coalesced across the warp: consecutive positions), 2 x 256 contiguous bytes out.
__global__ void __launch_bounds__(128) conv3x3_pack_kernel(const float* __restrict__ x, float* __restrict__ hi,
                                                           float* __restrict__ lo, int B, int H, int W) {
    const int Wp = W + 2, Hp = H + 2, per = Hp * Wp;
    const long long p = (long long)blockIdx.x * 128 + threadIdx.x;
    if (p >= (long long)B * per) return;
    const int b = (int)(p / per), rem = (int)(p - (long long)b * per);
    const int yp = rem / Wp, xp = rem - yp * Wp;
    const bool real = yp >= 1 && yp <= H && xp >= 1 && xp <= W;
    const long long HW = (long long)H * W;
    const float* src = x + (long long)b * CT_C * HW + (long long)(yp - 1) * W + (xp - 1);
    float4* oh = reinterpret_cast<float4*>(hi + p * CT_C);
    float4* ol = reinterpret_cast<float4*>(lo + p * CT_C);
#pragma unroll 4
    for (int c = 0; c < CT_C; c += 4) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = real ? __ldg(src + (c + j) * HW) : 0.f;
        float4 h, l;
        h.x = ct_hi(v[0]); h.y = ct_hi(v[1]); h.z = ct_hi(v[2]); h.w = ct_hi(v[3]);
        l.x = ct_hi(v[0] - h.x); l.y = ct_hi(v[1] - h.y); l.z = ct_hi(v[2] - h.z); l.w = ct_hi(v[3] - h.w);
        oh[c >> 2] = h; ol[c >> 2] = l;
    }
}

// torch weight [64 out][64 in][3][3] -> packed hi / lo [9 taps][64 out][64 in]
__global__ void __launch_bounds__(256) conv3x3_pack_weights_kernel(const float* __restrict__ w, float* __restrict__ hi,
                                                                   float* __restrict__ lo) {
    const int i = blockIdx.x * 256 + threadIdx.x;           // over 9 * 64 * 64
    if (i >= 9 * CT_C * CT_C) return;
    const int c = i % CT_C, o = (i / CT_C) % CT_C, t = i / (CT_C * CT_C);
    const float v = __ldg(w + ((long long)o * CT_C + c) * 9 + t);
    const float h = ct_hi(v);
    hi[i] = h; lo[i] = ct_hi(v - h);
}

static bool ct_encode_2d(tm_encode_fn enc, CUtensorMap* m, const float* base, cuuint64_t rows, cuuint32_t box_rows) {
    const cuuint64_t gdim[2] = {(cuuint64_t)CT_C, rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)CT_C * 4};
    const cuuint32_t box[2] = {CT_KB, box_rows};
    const cuuint32_t est[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, est,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace dpv

extern "C" int64_t dpv_conv3x3_packed_floats(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return (int64_t)B * (H + 2) * (W + 2) * dpv::CT_C;
}

extern "C" int dpv_conv3x3_pack(const float* x, float* packed_hi, float* packed_lo, int B, int C, int H, int W,
                                void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(x && packed_hi && packed_lo && B > 0 && H > 0 && W > 0);
    if (C != CT_C) return DPV_E_UNSUPP;
    const long long n = (long long)B * (H + 2) * (W + 2);
    conv3x3_pack_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(x, packed_hi, packed_lo, B, H, W);
    DPV_LAUNCH_END();
    return 0;
}

extern "C" int dpv_conv3x3_pack_weights(const float* weight, float* w_hi, float* w_lo, int C_out, int C_in,
                                        void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(weight && w_hi && w_lo);
    if (C_out != CT_C || C_in != CT_C) return DPV_E_UNSUPP;
    conv3x3_pack_weights_kernel<<<(9 * CT_C * CT_C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(weight, w_hi, w_lo);
    DPV_LAUNCH_END();
    return 0;
}

extern "C" int dpv_conv3x3_d64(const float* in_hi, const float* in_lo, const float* w_hi, const float* w_lo,
                               const float* bias, float* out_hi, float* out_lo, float* out_nchw, int B, int H,
                               int W, int epilogue, float slope, void* stream) {
    using namespace dpv;
    DPV_CHECK_ARG(in_hi && in_lo && w_hi && w_lo && bias && B > 0 && H > 0 && W > 0);
    DPV_CHECK_ARG(epilogue >= 0 && epilogue <= 2);
    DPV_CHECK_ARG((out_hi == nullptr) == (out_lo == nullptr));
    DPV_CHECK_ARG(out_hi != nullptr || out_nchw != nullptr);
    const long long np = (long long)B * (H + 2) * (W + 2);
    if (np > (1LL << 30)) return DPV_E_UNSUPP;
    if (((uintptr_t)in_hi | (uintptr_t)in_lo | (uintptr_t)w_hi | (uintptr_t)w_lo | (uintptr_t)out_hi | (uintptr_t)out_lo) & 15)
        return DPV_E_BADARG;
    tm_encode_fn enc = tm_encoder();
    if (enc == nullptr) return DPV_E_UNSUPP;
    ConvMaps maps;
    if (!ct_encode_2d(enc, &maps.a_hi, in_hi, (cuuint64_t)np, CT_AROWS) || !ct_encode_2d(enc, &maps.a_lo, in_lo, (cuuint64_t)np, CT_AROWS) ||
        !ct_encode_2d(enc, &maps.w_hi, w_hi, 9 * CT_C, CT_C) || !ct_encode_2d(enc, &maps.w_lo, w_lo, 9 * CT_C, CT_C))
        return DPV_E_UNSUPP;
    ConvArgs a;
    a.bias = bias; a.out_hi = out_hi; a.out_lo = out_lo; a.out_nchw = out_nchw;
    a.B = B; a.H = H; a.W = W; a.NP = (int)np; a.epilogue = epilogue; a.slope = slope;
    const size_t smem = (size_t)CT_STAGES * CT_STAGE_BYTES + 1024;
    cudaError_t e = cudaFuncSetAttribute(conv3x3_d64_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const unsigned tiles = (unsigned)((np + CT_M - 1) / CT_M);
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            return DPV_E_UNSUPP;
        n_sm = n;
    }
    const unsigned grid = tiles < (unsigned)n_sm ? tiles : (unsigned)n_sm;     // persistent: one CTA per SM
    conv3x3_d64_tc_kernel<<<grid, CT_THREADS, smem, (cudaStream_t)stream>>>(a, maps);
    DPV_LAUNCH_END();
    return 0;
}
