// Shared declarations of the plane-sweep kernels (sweep_cost.cu, sweep_tiled.cu).
#pragma once
#include "dpv_common.cuh"

namespace dpv {

struct SweepArgs {
    const float* ref; const float* src; const float* pose; const float* K; const float* rays;
    const float* d; float* cost; float* lsm;
    int B, V, C, D, H, W, PS;
    long long ref_bs, src_bs, src_vs, pose_bs, k_bs, rays_bs;
    float sigma;
};

constexpr float kCellSlack = 1.5e-5f;   // ~4 ulp of a coordinate of 50 px

struct CellTaps {
    int o00, o01, o10, o11;      // element offsets inside one channel plane (0 when invalid)
    bool v00, v01, v10, v11;
    int id;                      // cell identity for run detection
};

__device__ __forceinline__ CellTaps cell_taps(const Tap& t, int H, int W) {
    CellTaps c;
    bool xl = (t.x0 >= 0) & (t.x0 < W), xr = (t.x0 + 1 >= 0) & (t.x0 + 1 < W);
    bool yt = (t.y0 >= 0) & (t.y0 < H), yb = (t.y0 + 1 >= 0) & (t.y0 + 1 < H);
    c.v00 = xl & yt; c.v01 = xr & yt; c.v10 = xl & yb; c.v11 = xr & yb;
    int base = t.y0 * W + t.x0;
    c.o00 = c.v00 ? base : 0;
    c.o01 = c.v01 ? base + 1 : 0;
    c.o10 = c.v10 ? base + W : 0;
    c.o11 = c.v11 ? base + W + 1 : 0;
    bool any = c.v00 | c.v01 | c.v10 | c.v11;
    c.id = any ? (t.y0 + 2) * (W + 4) + (t.x0 + 2) : -1;   // all-outside cells are one cell
    return c;
}


// Defined in sweep_tiled.cu: shared-memory staged variant of the Gram formulation.
int launch_sweep_gram_tiled(const SweepArgs& a, cudaStream_t st);
// Defined in sweep_tma.cu: TMA-fed variant, four lanes per reference pixel (production path).
bool sweep_gram_tma_supported(const SweepArgs& a);
int launch_sweep_gram_tma(const SweepArgs& a, cudaStream_t st);
// Defined in sweep_xcorr.cu: cross-correlation form (source-only products by a pre-pass into a workspace).
long long sweep_xcorr_workspace_floats(int B, int V, int H, int W);
bool sweep_xcorr_supported(const SweepArgs& a);
int launch_sweep_xcorr(const SweepArgs& a, float* workspace, cudaStream_t st);

}  // namespace dpv
