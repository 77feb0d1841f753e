// Device helpers shared by the tensor-core convolution kernels (conv_tc.cu: 2-D, 64 channels; conv3d_tc.cu: 3-D,
// 32 channels): 2-D TMA tile loads, tcgen05 shared-memory descriptors, the TF32 MMA, commit, the TF32 split.
#pragma once
#include <cuda.h>

#include "sweep_tma.cuh"   // mbarrier / TMA wrappers, tm_encoder()

namespace dpv {

__device__ __forceinline__ void ct_tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(tm_smem(dst)), "l"(map), "r"(c0), "r"(c1), "r"(tm_smem(bar)) : "memory");
}
// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), version 1 (sm_100).
// The swizzle is a function of the shared-memory ADDRESS (bits 4-6 xor bits 7-9), for the copy engine that wrote
// the tile and for the tensor core that reads it alike: a descriptor that starts one or two 128-byte rows into a
// tile (the dx = 0 / +1 taps) reads those rows correctly with the base-offset field left at 0 (measured: setting
// it to the row phase gives wrong results).
__device__ __forceinline__ unsigned long long ct_smem_desc(const void* p) {
    const unsigned long long addr = (unsigned long long)(tm_smem(p) >> 4) & 0x3FFFull;
    return addr | (1ull << 16) | ((1024ull >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// D[tmem] (+)= A[smem] * B[smem], M = 128, N = 64, K = 8, TF32 in, fp32 accumulate
__device__ __forceinline__ void ct_mma_tf32(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc,
                                            unsigned accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void ct_commit(unsigned long long* bar) {   // arrives on `bar` when all prior MMAs are done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tm_smem(bar)) : "memory");
}
// The TF32 value nearest to x (round to nearest, ties away: cvt.rna).  Truncation would bias every hi towards zero
// and the sums of 576 products with it (measured: 3e-5 of the logits' scale instead of 3e-6).
__device__ __forceinline__ float ct_hi(float x) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__device__ __forceinline__ void ct_mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tm_smem(bar)) : "memory");
}

}  // namespace dpv
