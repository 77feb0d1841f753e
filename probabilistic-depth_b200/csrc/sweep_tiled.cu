// K1 + K2a, shared-memory staged Gram kernel (the production path for the L2 cost volume).
//
// Same mathematics as sweep_gram_kernel (sweep_cost.cu): per reference pixel the D sampling
// positions fall into a handful of source 2x2 cells ("runs" of consecutive planes); for each run
// the channel contraction is a 4x4 Gram matrix of (tap - ref) differences and each plane of the
// run is a 10-term quadratic form in its bilinear weights.  The per-thread global gathers of that
// kernel are latency-bound (5 dependent-stride loads per channel through L1/L2); here a CTA owns a
// 4x32 tile of reference pixels and
//   1. every thread walks its D planes once and records its runs (cell, first plane) in shared
//      memory; the CTA reduces the bounding box of all cells -> the source window of the tile
//      (the image of tile x [d_min, d_max] under a homography is convex, so the window is small:
//      <= 10 x 48 source pixels at the model's 1/4 resolution);
//   2. the window and the reference tile are streamed through shared memory in chunks of 8
//      channels with a 2-stage cp.async pipeline (out-of-image taps are zero-filled by the copy:
//      this is grid_sample's zeros padding);
//   3. each thread keeps the Gram matrices of up to NSLOT runs in registers and accumulates them
//      from shared memory with immediate-offset loads (4 LDS + 4 FFMA + 10 FFMA per channel and
//      run, no address arithmetic); pixels with more runs take another pass;
//   4. the planes of every run are evaluated from its Gram matrix and written to the cost volume.
// Tiles whose window does not fit (very large motion) fall back to the gather kernel.
#include <cstdlib>

#include "sweep_common.cuh"

namespace dpv {

constexpr int TL_TH = 4, TL_TW = 32, TL_NT = TL_TH * TL_TW;
constexpr int TL_WR = 10, TL_WC = 48;          // source window capacity (rows, cols)
constexpr int TL_CK = 8;                       // channels per stage
constexpr int TL_CS = TL_WR * TL_WC;           // channel stride inside a stage
constexpr int TL_MAXRUN = 24;                  // runs recorded per pixel and view
constexpr int TL_STAGE_FLOATS = TL_CK * TL_CS + TL_CK * TL_NT;

__device__ __forceinline__ void cp_async_f32(float* smem_dst, const float* gsrc, bool pred) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int bytes = pred ? 4 : 0;   // src-size 0: nothing is read, the 4 bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ int pack_cell(int x0, int y0) { return ((y0 + 32768) << 16) | ((x0 + 32768) & 0xffff); }
__device__ __forceinline__ int cell_x(int p) { return (p & 0xffff) - 32768; }
__device__ __forceinline__ int cell_y(int p) { return ((p >> 16) & 0xffff) - 32768; }
constexpr int kOutsideCell = -1;   // pack_cell never yields -1 for coordinates the taps accept

struct TileShared {
    int bbox[4];       // x0 min, x0 max, y0 min, y0 max over all recorded cells
    int max_runs;
    int overflow;
};

template <int NSLOT>
__global__ void __launch_bounds__(TL_NT) sweep_gram_tiled_kernel(const SweepArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* stage0 = smem;                                   // 2 x [CK][WR][WC] + [CK][NT]
    int* cell_s = (int*)(smem + 2 * TL_STAGE_FLOATS);       // [MAXRUN][NT]
    short* kst_s = (short*)(cell_s + TL_MAXRUN * TL_NT);    // [MAXRUN + 1][NT]
    float* d_s = (float*)(kst_s + (TL_MAXRUN + 1) * TL_NT + (TL_NT & 1));   // [kper]
    __shared__ TileShared ts;

    const int HW = a.H * a.W;
    const int kper = (a.D + a.PS - 1) / a.PS;
    const int k0 = blockIdx.y * kper;
    const int nk = min(a.D, k0 + kper) - k0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles_x = (a.W + TL_TW - 1) / TL_TW;
    const int tiles_y = (a.H + TL_TH - 1) / TL_TH;
    const int b = blockIdx.z;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    (void)tiles_y;
    const int y = ty * TL_TH + warp, x = tx * TL_TW + lane;
    const bool active = (y < a.H) && (x < a.W);
    const int p = active ? y * a.W + x : 0;
    for (int k = tid; k < nk; k += TL_NT) d_s[k] = __ldg(a.d + k0 + k);
    if (nk <= 0) return;

    const float* rays = a.rays + (long long)b * a.rays_bs;
    const float rx = __ldg(rays + p), ry = __ldg(rays + HW + p), rz = __ldg(rays + 2 * HW + p);
    const float* ref = a.ref + (long long)b * a.ref_bs;
    const float half_w = (float)a.W * 0.5f, half_h = (float)a.H * 0.5f;
    float* out = a.cost + ((long long)b * a.D + k0) * HW + p;
    const int nchunk = (a.C + TL_CK - 1) / TL_CK;

    for (int v = 0; v < a.V; ++v) {
        const float* src = a.src + (long long)b * a.src_bs + (long long)v * a.src_vs;
        const ViewGeom g = load_view_geom(a.K + (long long)b * a.k_bs,
                                          a.pose + (long long)b * a.pose_bs + (long long)v * 16);
        const PixelTerm pt = pixel_term(g, rx, ry, rz);
        const float t1x = g.t1[0], t1y = g.t1[1], t1z = g.t1[2], cx = g.cx, cy = g.cy;
        const float inv_cx = __frcp_rn(cx), inv_cy = __frcp_rn(cy), inv_sigma = __frcp_rn(a.sigma);

        // ---------------- 1. runs of this pixel, bounding box of the tile ------------------
        if (tid == 0) {
            ts.bbox[0] = 1 << 30; ts.bbox[1] = -(1 << 30); ts.bbox[2] = 1 << 30; ts.bbox[3] = -(1 << 30);
            ts.max_runs = 0; ts.overflow = 0;
        }
        __syncthreads();   // also orders d_s / previous view's use of cell_s
        int nrun = 0;
        {
            int bx0 = 1 << 30, bx1 = -(1 << 30), by0 = 1 << 30, by1 = -(1 << 30);
            int cur_id = kOutsideCell, cur_x = 0, cur_y = 0;
            bool over = false;
            if (active) {
                for (int k = 0; k < nk; ++k) {
                    float ix, iy;
                    sweep_coord_fast(t1x, t1y, t1z, pt, d_s[k], cx, cy, inv_cx, inv_cy, half_w, half_h, ix, iy);
                    const Tap tap = make_tap(ix, iy);
                    // a cell is "outside" when none of its four taps lies in the image
                    const bool inside = (tap.x0 >= -1) & (tap.x0 < a.W) & (tap.y0 >= -1) & (tap.y0 < a.H);
                    const int id = inside ? pack_cell(tap.x0, tap.y0) : kOutsideCell;
                    bool cont = (k > 0) && (id == cur_id);
                    if (!cont && k > 0 && cur_id != kOutsideCell) {
                        const float fx = ix - (float)cur_x, fy = iy - (float)cur_y;
                        cont = fx >= -kCellSlack && fx <= 1.0f + kCellSlack && fy >= -kCellSlack &&
                               fy <= 1.0f + kCellSlack;
                    }
                    if (!cont) {
                        cur_id = id; cur_x = tap.x0; cur_y = tap.y0;
                        if (nrun < TL_MAXRUN) {
                            cell_s[nrun * TL_NT + tid] = id;
                            kst_s[nrun * TL_NT + tid] = (short)k;
                        } else {
                            over = true;
                        }
                        ++nrun;
                        if (inside) {
                            bx0 = min(bx0, tap.x0); bx1 = max(bx1, tap.x0);
                            by0 = min(by0, tap.y0); by1 = max(by1, tap.y0);
                        }
                    }
                }
                if (!over) kst_s[nrun * TL_NT + tid] = (short)nk;
            }
            bx0 = __reduce_min_sync(0xffffffffu, bx0); bx1 = __reduce_max_sync(0xffffffffu, bx1);
            by0 = __reduce_min_sync(0xffffffffu, by0); by1 = __reduce_max_sync(0xffffffffu, by1);
            const int mr = __reduce_max_sync(0xffffffffu, nrun);
            const int ov = __any_sync(0xffffffffu, over);
            if (lane == 0) {
                atomicMin(&ts.bbox[0], bx0); atomicMax(&ts.bbox[1], bx1);
                atomicMin(&ts.bbox[2], by0); atomicMax(&ts.bbox[3], by1);
                atomicMax(&ts.max_runs, mr);
                if (ov) atomicOr(&ts.overflow, 1);
            }
        }
        __syncthreads();
        int wx0 = ts.bbox[0], wy0 = ts.bbox[2];
        int ww = ts.bbox[1] + 2 - wx0, wh = ts.bbox[3] + 2 - wy0;   // taps reach x0+1, y0+1
        if (ts.bbox[1] < ts.bbox[0]) { wx0 = 0; wy0 = 0; ww = 0; wh = 0; }   // every cell outside
        const bool fits = (ww <= TL_WC) && (wh <= TL_WR) && !ts.overflow;
        const int max_runs = ts.max_runs;

        if (!fits) {
            // ---------------- fallback: per-thread gathers from global memory ---------------
            if (active) {
                const float* refp = ref + p;
                int k = 0;
                float ix, iy;
                sweep_coord(t1x, t1y, t1z, pt, d_s[0], cx, cy, half_w, half_h, ix, iy);
                Tap tap = make_tap(ix, iy);
                CellTaps cell = cell_taps(tap, a.H, a.W);
                while (k < nk) {
                    float q[10];
#pragma unroll
                    for (int i = 0; i < 10; ++i) q[i] = 0.f;
                    for (int c = 0; c < a.C; ++c) {
                        const float r = __ldg(refp + (long long)c * HW);
                        const float* sc = src + (long long)c * HW;
                        const float e0 = (cell.v00 ? __ldg(sc + cell.o00) : 0.f) - r;
                        const float e1 = (cell.v01 ? __ldg(sc + cell.o01) : 0.f) - r;
                        const float e2 = (cell.v10 ? __ldg(sc + cell.o10) : 0.f) - r;
                        const float e3 = (cell.v11 ? __ldg(sc + cell.o11) : 0.f) - r;
                        q[0] = fmaf(e0, e0, q[0]); q[1] = fmaf(e0, e1, q[1]); q[2] = fmaf(e0, e2, q[2]);
                        q[3] = fmaf(e0, e3, q[3]); q[4] = fmaf(e1, e1, q[4]); q[5] = fmaf(e1, e2, q[5]);
                        q[6] = fmaf(e1, e3, q[6]); q[7] = fmaf(e2, e2, q[7]); q[8] = fmaf(e2, e3, q[8]);
                        q[9] = fmaf(e3, e3, q[9]);
                    }
                    const int id = cell.id;
                    const float cx0 = (float)tap.x0, cy0 = (float)tap.y0;
                    do {
                        float nw, ne, sw, se;
                        bilinear_weights(tap, nw, ne, sw, se);
                        const float diag = nw * nw * q[0] + ne * ne * q[4] + sw * sw * q[7] + se * se * q[9];
                        const float off = nw * (ne * q[1] + sw * q[2] + se * q[3]) +
                                          ne * (sw * q[5] + se * q[6]) + sw * se * q[8];
                        const float val = __fdiv_rn(fmaf(2.0f, off, diag), a.sigma);
                        float* o = out + (long long)k * HW;
                        *o = (v == 0) ? val : (*o + val);
                        ++k;
                        if (k < nk) {
                            sweep_coord(t1x, t1y, t1z, pt, d_s[k], cx, cy, half_w, half_h, ix, iy);
                            tap = make_tap(ix, iy);
                            cell = cell_taps(tap, a.H, a.W);
                            const float rx0 = ix - cx0, ry0 = iy - cy0;
                            if (id >= 0 && cell.id != id && rx0 >= -kCellSlack && rx0 <= 1.0f + kCellSlack &&
                                ry0 >= -kCellSlack && ry0 <= 1.0f + kCellSlack) {
                                tap.x0 = (int)cx0; tap.y0 = (int)cy0; tap.fx = rx0; tap.fy = ry0;
                                cell.id = id;
                            }
                        }
                    } while (k < nk && cell.id == id);
                }
            }
            continue;   // next view (uniform across the CTA)
        }

        // ---------------- 2-4. passes of NSLOT runs each -----------------------------------
        for (int first = 0; first < max_runs; first += NSLOT) {
            float G[NSLOT][10];
            int coff[NSLOT];
            float msk[NSLOT];
#pragma unroll
            for (int j = 0; j < NSLOT; ++j) {
#pragma unroll
                for (int i = 0; i < 10; ++i) G[j][i] = 0.f;
                coff[j] = 0; msk[j] = 0.f;
                if (first + j < nrun) {
                    const int pc = cell_s[(first + j) * TL_NT + tid];
                    if (pc != kOutsideCell) {
                        coff[j] = (cell_y(pc) - wy0) * TL_WC + (cell_x(pc) - wx0);
                        msk[j] = 1.0f;
                    }
                }
            }
            const int nloc = min(NSLOT, nrun - first);   // <= 0: nothing for this thread

            auto issue = [&](int chunk) {
                float* st = stage0 + (chunk & 1) * TL_STAGE_FLOATS;
                const int c0 = chunk * TL_CK;
                // window: warp w copies channels w and w + 4 of the chunk, lanes along columns
#pragma unroll
                for (int cc = 0; cc < TL_CK / TL_TH; ++cc) {
                    const int c = warp + cc * TL_TH;
                    const bool cok = (c0 + c) < a.C;
                    const float* sc = src + (long long)(c0 + c) * HW;
                    float* dst = st + c * TL_CS;
                    for (int row = 0; row < wh; ++row) {
                        const int sy = wy0 + row;
                        const bool rok = cok && sy >= 0 && sy < a.H;
                        for (int col = lane; col < ww; col += 32) {
                            const int sx = wx0 + col;
                            const bool ok = rok && sx >= 0 && sx < a.W;
                            cp_async_f32(dst + row * TL_WC + col, ok ? sc + sy * a.W + sx : src, ok);
                        }
                    }
                }
                // reference tile: every thread fetches its own pixel for the chunk's channels
                float* rdst = st + TL_CK * TL_CS;
#pragma unroll
                for (int c = 0; c < TL_CK; ++c) {
                    const bool ok = active && (c0 + c) < a.C;
                    cp_async_f32(rdst + c * TL_NT + tid, ok ? ref + (long long)(c0 + c) * HW + p : ref, ok);
                }
            };

            issue(0);
            cp_async_commit();
            for (int chunk = 0; chunk < nchunk; ++chunk) {
                if (chunk + 1 < nchunk) issue(chunk + 1);
                cp_async_commit();
                cp_async_wait<1>();
                __syncthreads();
                const float* st = stage0 + (chunk & 1) * TL_STAGE_FLOATS;
                float r[TL_CK];
#pragma unroll
                for (int c = 0; c < TL_CK; ++c) r[c] = st[TL_CK * TL_CS + c * TL_NT + tid];
#pragma unroll
                for (int j = 0; j < NSLOT; ++j) {
                    if (j < nloc) {
                        const float* w = st + coff[j];
                        const float m = msk[j];
#pragma unroll
                        for (int c = 0; c < TL_CK; ++c) {
                            const float e0 = fmaf(w[c * TL_CS], m, -r[c]);
                            const float e1 = fmaf(w[c * TL_CS + 1], m, -r[c]);
                            const float e2 = fmaf(w[c * TL_CS + TL_WC], m, -r[c]);
                            const float e3 = fmaf(w[c * TL_CS + TL_WC + 1], m, -r[c]);
                            G[j][0] = fmaf(e0, e0, G[j][0]); G[j][1] = fmaf(e0, e1, G[j][1]);
                            G[j][2] = fmaf(e0, e2, G[j][2]); G[j][3] = fmaf(e0, e3, G[j][3]);
                            G[j][4] = fmaf(e1, e1, G[j][4]); G[j][5] = fmaf(e1, e2, G[j][5]);
                            G[j][6] = fmaf(e1, e3, G[j][6]); G[j][7] = fmaf(e2, e2, G[j][7]);
                            G[j][8] = fmaf(e2, e3, G[j][8]); G[j][9] = fmaf(e3, e3, G[j][9]);
                        }
                    }
                }
                __syncthreads();
            }

            // ---------------- 4. planes of each run --------------------------------------
#pragma unroll
            for (int j = 0; j < NSLOT; ++j) {
                if (j < nloc) {
                    const int run = first + j;
                    const int pc = cell_s[run * TL_NT + tid];
                    const int ka = kst_s[run * TL_NT + tid], kb = kst_s[(run + 1) * TL_NT + tid];
                    const float fx0 = (float)cell_x(pc), fy0 = (float)cell_y(pc);
                    for (int k = ka; k < kb; ++k) {
                        float val;
                        if (pc == kOutsideCell) {
                            val = G[j][0];
                        } else {
                            float ix, iy;
                            sweep_coord_fast(t1x, t1y, t1z, pt, d_s[k], cx, cy, inv_cx, inv_cy, half_w, half_h, ix, iy);
                            Tap tap;
                            tap.x0 = 0; tap.y0 = 0;
                            tap.fx = ix - fx0; tap.fy = iy - fy0;
                            float nw, ne, sw, se;
                            bilinear_weights(tap, nw, ne, sw, se);
                            const float diag = nw * nw * G[j][0] + ne * ne * G[j][4] + sw * sw * G[j][7] +
                                               se * se * G[j][9];
                            const float off = nw * (ne * G[j][1] + sw * G[j][2] + se * G[j][3]) +
                                              ne * (sw * G[j][5] + se * G[j][6]) + sw * se * G[j][8];
                            val = fmaf(2.0f, off, diag);
                        }
                        val *= inv_sigma;
                        float* o = out + (long long)k * HW;
                        *o = (v == 0) ? val : (*o + val);
                    }
                }
            }
        }
    }
}

static size_t tiled_smem_bytes(int kper) {
    size_t n = (size_t)2 * TL_STAGE_FLOATS * sizeof(float) + (size_t)TL_MAXRUN * TL_NT * sizeof(int) +
               ((size_t)(TL_MAXRUN + 1) * TL_NT + (TL_NT & 1)) * sizeof(short) + (size_t)kper * sizeof(float);
    return (n + 15) & ~(size_t)15;
}

int launch_sweep_gram_tiled(const SweepArgs& a, cudaStream_t st) {
    static const int nslot = [] { const char* e = getenv("DPV_SWEEP_NSLOT"); return e ? atoi(e) : 0; }();
    const int kper = (a.D + a.PS - 1) / a.PS;
    if (kper > 32767) return DPV_E_UNSUPP;
    const size_t smem = tiled_smem_bytes(kper);
    const int tiles = ((a.W + TL_TW - 1) / TL_TW) * ((a.H + TL_TH - 1) / TL_TH);
    dim3 grid(tiles, a.PS, a.B), block(TL_NT);
    cudaError_t e;
    if (nslot == 12) {
        e = cudaFuncSetAttribute(sweep_gram_tiled_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        sweep_gram_tiled_kernel<12><<<grid, block, smem, st>>>(a);
    } else if (nslot == 8) {
        e = cudaFuncSetAttribute(sweep_gram_tiled_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        sweep_gram_tiled_kernel<8><<<grid, block, smem, st>>>(a);
    } else {
        e = cudaFuncSetAttribute(sweep_gram_tiled_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        sweep_gram_tiled_kernel<6><<<grid, block, smem, st>>>(a);
    }
    DPV_LAUNCH_END();
    return 0;
}

}  // namespace dpv
