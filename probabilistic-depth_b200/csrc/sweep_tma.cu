// K1 + K2a, TMA-fed Gram kernel with four threads per reference pixel (production path for the L2
// cost volume; algo = 4).
//
// Same mathematics as sweep_gram_tiled_kernel (sweep_tiled.cu): per reference pixel the D sampling
// positions fall into a handful of source 2x2 cells ("runs" of consecutive planes); for each run
// the channel contraction is the 4x4 Gram matrix of (tap - ref) differences, and every plane of the
// run is a 10-term quadratic form in its bilinear weights.  What changed is the mapping onto the
// machine (ncu on the tiled kernel, profiles/r01_*: one thread per pixel gives 10 warps per SM at
// the model's 8 x 64 x 96 pixels, and a quarter of its instructions were 4-byte cp.async copies):
//   * a CTA owns one row segment of 32 reference pixels with FOUR threads per pixel: lane = pixel,
//     warp = quarter.  The warps split the planes for run detection (each walks D/4 planes of its
//     32 pixels; the four run lists of a pixel are stitched through shared memory), split the runs
//     round-robin for the Gram accumulation (NSLOT Gram matrices per thread instead of 6, updated
//     two at a time with fp32x2 FFMA2), exchange the matrices through shared memory and evaluate
//     the planes of their own quarter.  Lanes of a warp do the same job on neighbouring pixels, so
//     run boundaries and trip counts are (nearly) warp-uniform.  4x the threads at fewer registers:
//     20 warps per SM instead of 10.
//   * the source window of the tile and the reference pixels are fetched by the TMA engine
//     (cp.async.bulk.tensor, 5-D / 4-D maps over the caller's strided [B,V,C,H,W] / [B,C,H,W]
//     tensors): one elected thread issues one box per window row and chunk of 8 channels,
//     completion is counted in bytes on an mbarrier, and out-of-image taps / channels past C are
//     zero-filled by the engine -- grid_sample's zeros padding costs no instruction.
//   * results are collected in shared memory ([plane][pixel]) and written with 128-byte rows.
// Shared memory: stage layout [row][channel][col] so that the four taps of a cell and the channels
// of a chunk are compile-time offsets from one per-run base (LDS with immediate offsets).
// Tiles whose window does not fit (very large motion) gather from global memory instead.
#include <cstdlib>

#include "sweep_tma.cuh"

namespace dpv {

constexpr int TM_T = 4, TM_NT = TM_PX * TM_T;
constexpr int TM_WC = 52, TM_WR = 6;             // source window capacity (cols, rows)
constexpr int TM_CK = 8;                         // channels per stage
constexpr int TM_NSTAGE = 2;
constexpr int TM_MAXRUN = 24;                    // runs recorded per pixel and view
constexpr int TM_ROW = TM_CK * TM_WC;            // floats of one window row (all channels of a chunk)
constexpr int TM_WIN = TM_WR * TM_ROW;
constexpr int TM_STAGE = TM_WIN + TM_CK * TM_PX; // + the reference pixels of the chunk

template <int NSLOT, bool EXACT>
__global__ void __launch_bounds__(TM_NT, 6)
sweep_gram_tma_kernel(const SweepArgs a, const __grid_constant__ CUtensorMap map_src,
                      const __grid_constant__ CUtensorMap map_ref) {
    __shared__ __align__(128) float stage0[TM_NSTAGE * TM_STAGE];      // TMA destinations
    __shared__ int cell_s[TM_MAXRUN * TM_PX];                          // [MAXRUN][PX] cell of each run
    __shared__ short kst_s[(TM_MAXRUN + 1) * TM_PX];                   // [MAXRUN + 1][PX] first plane
    extern __shared__ __align__(16) float out_s[];                     // [kper][PX] result tile
    __shared__ float geo_s[16];                                        // K R (9), K t (3), cx, cy of the view
    __shared__ unsigned long long full_bar[TM_NSTAGE];
    __shared__ TmShared ts;

    const int HW = a.H * a.W;
    const int kper = (a.D + a.PS - 1) / a.PS;
    float* d_s = out_s + kper * TM_OS;                                 // [kper]
    const int k0 = blockIdx.y * kper;
    const int nk = min(a.D, k0 + kper) - k0;
    const int tid = threadIdx.x, lane = tid & 31;
    const int px = tid & 31, t = tid >> 5;   // lane = pixel, warp = plane quarter / run residue
    const int tiles_x = (a.W + TM_PX - 1) / TM_PX;
    const int b = blockIdx.z;
    const int y = blockIdx.x / tiles_x, tx = blockIdx.x - y * tiles_x;
    const int x = tx * TM_PX + px;
    const bool active = x < a.W;
    const int p = active ? y * a.W + x : y * a.W;
    if (nk <= 0) return;
    for (int k = tid; k < nk; k += TM_NT) d_s[k] = __ldg(a.d + k0 + k);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < TM_NSTAGE; ++s) tm_mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    const float* rays = a.rays + (long long)b * a.rays_bs;
    const float rx = __ldg(rays + p), ry = __ldg(rays + HW + p), rz = __ldg(rays + 2 * HW + p);
    const float* ref = a.ref + (long long)b * a.ref_bs;
    const int nchunk = (a.C + TM_CK - 1) / TM_CK;
    const float inv_sigma = __frcp_rn(a.sigma);
    // planes walked by this lane in the run detection
    const int kpt = (nk + TM_T - 1) / TM_T;
    const int wa = min(nk, t * kpt), wb = min(nk, wa + kpt);
    unsigned q_issue = 0, q_use = 0;   // chunk sequence numbers (stage = q % NSTAGE, parity = q / NSTAGE)

    for (int v = 0; v < a.V; ++v) {
        const float* src = a.src + (long long)b * a.src_bs + (long long)v * a.src_vs;
        // ---------------- 0. view geometry: K R and K t once per CTA ---------------------------
        if (tid < 12) {   // warping/homography.py:119-121; same dot3 order as load_view_geom
            const float* Kp = a.K + (long long)b * a.k_bs + (tid < 9 ? tid / 3 : tid - 9) * 3;
            const float* pp = a.pose + (long long)b * a.pose_bs + (long long)v * 16 + (tid < 9 ? tid % 3 : 3);
            geo_s[tid] = dot3(__ldg(Kp), __ldg(Kp + 1), __ldg(Kp + 2), __ldg(pp), __ldg(pp + 4), __ldg(pp + 8));
        } else if (tid < 14) {
            geo_s[tid] = __ldg(a.K + (long long)b * a.k_bs + (tid == 12 ? 2 : 5));   // cx, cy
        }
        if (tid == 0) {
            ts.bbox[0] = 1 << 30; ts.bbox[1] = -(1 << 30); ts.bbox[2] = 1 << 30; ts.bbox[3] = -(1 << 30);
            ts.max_runs = 0; ts.overflow = 0;
        }
        __syncthreads();   // geo_s, d_s, barrier init, ts; previous view done with cell_s / kst_s
        PixelTerm pt;
        TmGeom g;
        pt.x = dot3(geo_s[0], geo_s[1], geo_s[2], rx, ry, rz);
        pt.y = dot3(geo_s[3], geo_s[4], geo_s[5], rx, ry, rz);
        pt.z = dot3(geo_s[6], geo_s[7], geo_s[8], rx, ry, rz);
        g.t1x = geo_s[9]; g.t1y = geo_s[10]; g.t1z = geo_s[11]; g.cx = geo_s[12]; g.cy = geo_s[13];
        g.inv_cx = __frcp_rn(g.cx); g.inv_cy = __frcp_rn(g.cy);
        g.half_w = (float)a.W * 0.5f; g.half_h = (float)a.H * 0.5f;

        // ---------------- 1. runs of this pixel: each warp walks its quarter of the planes ----
        int nrun;          // runs of this pixel (all four lanes agree)
        int my_run0;       // run that holds my first plane
        {
            unsigned long long starts = 0ull;
            int n = 0, first_id = kTmNone, cur_id = kTmNone;
            int bx0 = 1 << 30, bx1 = -(1 << 30), by0 = 1 << 30, by1 = -(1 << 30);
            if (active) {
                float cur_xf = 0.f, cur_yf = 0.f;   // origin of the current run's cell
                bool in_cell = false;               // ... when it has one (not "outside")
                for (int k = wa; k < wb; ++k) {
                    float ix, iy;
                    tm_coord<EXACT>(g, pt, d_s[k], ix, iy);
                    // fast path: still inside the current cell (with the border slack)?  Covers "same
                    // floor cell" too, so the floor / bounds / packing below run only where a run may start.
                    const float fx = ix - cur_xf, fy = iy - cur_yf;
                    if (in_cell && fx >= -kCellSlack && fx <= 1.0f + kCellSlack && fy >= -kCellSlack &&
                        fy <= 1.0f + kCellSlack)
                        continue;
                    const Tap tap = make_tap(ix, iy);
                    const bool inside = (tap.x0 >= -1) & (tap.x0 < a.W) & (tap.y0 >= -1) & (tap.y0 < a.H);
                    const int id = inside ? tm_pack(tap.x0, tap.y0) : kTmOutside;
                    if (id == cur_id) continue;     // outside -> outside
                    cur_id = id; cur_xf = (float)tap.x0; cur_yf = (float)tap.y0; in_cell = inside;
                    starts |= 1ull << (k - wa);
                    if (n == 0) first_id = id;
                    ++n;
                    if (inside) {
                        bx0 = min(bx0, tap.x0); bx1 = max(bx1, tap.x0);
                        by0 = min(by0, tap.y0); by1 = max(by1, tap.y0);
                    }
                }
            }
            // stitch the four lists (one per warp) through shared memory: a lane's first run
            // continues the previous quarter's last run when both start from the same cell
            int* st_n = (int*)stage0;                 // the stages are idle during run detection
            int* st_first = st_n + TM_T * TM_PX;
            int* st_last = st_first + TM_T * TM_PX;
            st_n[t * TM_PX + px] = n; st_first[t * TM_PX + px] = first_id; st_last[t * TM_PX + px] = cur_id;
            __syncthreads();
            int merge = 0, mine = 0, incl = 0;
            nrun = 0;
#pragma unroll
            for (int u = 0; u < TM_T; ++u) {
                const int nu = st_n[u * TM_PX + px];
                const int mu = (u > 0 && nu > 0 && st_first[u * TM_PX + px] == st_last[(u > 0 ? u - 1 : 0) * TM_PX + px]) ? 1 : 0;
                nrun += nu - mu;
                if (u == t) { merge = mu; mine = nu - mu; incl = nrun; }
            }
            my_run0 = incl - mine - merge;
            bool over = nrun > TM_MAXRUN;
            if (active && !over) {
                int gi = incl - mine - merge;      // global index of my first local run
                unsigned long long m = starts;
                while (m) {
                    const int i = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    if (gi >= incl - mine) {       // not the merged one
                        const int k = wa + i;
                        const Tap tap = tm_tap<EXACT>(g, pt, d_s[k]);
                        const bool inside = (tap.x0 >= -1) & (tap.x0 < a.W) & (tap.y0 >= -1) & (tap.y0 < a.H);
                        cell_s[gi * TM_PX + px] = inside ? tm_pack(tap.x0, tap.y0) : kTmOutside;
                        kst_s[gi * TM_PX + px] = (short)k;
                    }
                    ++gi;
                }
                if (wb == nk && wa < wb) kst_s[nrun * TM_PX + px] = (short)nk;
            }
            bx0 = __reduce_min_sync(0xffffffffu, bx0); bx1 = __reduce_max_sync(0xffffffffu, bx1);
            by0 = __reduce_min_sync(0xffffffffu, by0); by1 = __reduce_max_sync(0xffffffffu, by1);
            const int mr = __reduce_max_sync(0xffffffffu, active ? nrun : 0);
            const int ov = __any_sync(0xffffffffu, active && over);
            if (lane == 0) {
                atomicMin(&ts.bbox[0], bx0); atomicMax(&ts.bbox[1], bx1);
                atomicMin(&ts.bbox[2], by0); atomicMax(&ts.bbox[3], by1);
                atomicMax(&ts.max_runs, mr);
                if (ov) atomicOr(&ts.overflow, 1);
            }
        }
        __syncthreads();
        // TMA needs every box row to start on a 16-byte boundary of global memory (an unaligned
        // innermost coordinate raises "illegal instruction", tools/probe/tma_probe.cu): the window
        // starts at a multiple of 4 columns (W % 4 == 0, so rows keep that alignment)
        int wx0 = ts.bbox[0] & ~3, wy0 = ts.bbox[2];
        int ww = ts.bbox[1] + 2 - wx0, wh = ts.bbox[3] + 2 - wy0;   // taps reach x0 + 1, y0 + 1
        if (ts.bbox[1] < ts.bbox[0]) { wx0 = 0; wy0 = 0; ww = 0; wh = 0; }   // every cell outside
        const bool fits = (ww <= TM_WC) && (wh <= TM_WR) && !ts.overflow;
        const int max_runs = ts.max_runs;
        __syncthreads();   // ts is re-initialised by thread 0 at the top of the next view

        if (!fits) {
            if (active && wa < wb)
                tm_gather_planes<EXACT>(a.C, a.H, a.W, src, ref + p, g, pt, d_s, wa, wb, inv_sigma, out_s + px, v == 0);
            continue;   // next view (uniform across the CTA)
        }

        // ---------------- 2-4. passes of TM_T * NSLOT runs ----------------------------------
        for (int first = 0; first < max_runs; first += TM_T * NSLOT) {
            static_assert(NSLOT % 2 == 0, "runs are accumulated in fp32x2 pairs");
            tm_f2 G2[NSLOT / 2][10];   // Gram matrices of slots (2q, 2q+1) in the (lo, hi) halves
            int coff[NSLOT];   // window offset of the run's cell; runs whose cell is outside the image
                               // accumulate from offset 0 and are evaluated from rr instead
            float rr = 0.f;    // sum_c ref^2: the cost of a plane with no tap inside the image
            int nloc = 0;
#pragma unroll
            for (int j = 0; j < NSLOT; ++j) {
                if ((j & 1) == 0) {
#pragma unroll
                    for (int i = 0; i < 10; ++i) G2[j / 2][i] = 0ull;
                }
                coff[j] = 0;
                const int run = first + t + TM_T * j;
                if (active && run < nrun) {
                    nloc = j + 1;
                    const int pc = cell_s[run * TM_PX + px];
                    if (pc != kTmOutside) coff[j] = (tm_cell_y(pc) - wy0) * TM_ROW + (tm_cell_x(pc) - wx0);
                }
            }

            auto issue = [&](int chunk) {
                const int s = q_issue % TM_NSTAGE;
                ++q_issue;
                float* st = stage0 + s * TM_STAGE;
                // The stage area is also written through the generic proxy (run-list stitching, the Gram
                // exchange, the soft-max partials) and read by all lanes; those accesses are ordered before
                // this point by __syncthreads, and this fence orders them before the async-proxy writes of
                // the bulk copies below.
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tm_mbar_expect_tx(&full_bar[s], (unsigned)((wh * TM_ROW + TM_CK * TM_PX) * sizeof(float)));
                for (int r = 0; r < wh; ++r)
                    tm_load_5d(st + r * TM_ROW, &map_src, wx0, wy0 + r, chunk * TM_CK, v, b, &full_bar[s]);
                tm_load_4d(st + TM_WIN, &map_ref, tx * TM_PX, y, chunk * TM_CK, b, &full_bar[s]);
            };
            if (tid == 0) {
                for (int c = 0; c < min(TM_NSTAGE, nchunk); ++c) issue(c);
            }
            for (int chunk = 0; chunk < nchunk; ++chunk) {
                const int s = q_use % TM_NSTAGE;
                tm_mbar_wait(&full_bar[s], (q_use / TM_NSTAGE) & 1);
                ++q_use;
                const float* st = stage0 + s * TM_STAGE;
                const int nc = min(TM_CK, a.C - chunk * TM_CK);
                if (nc == TM_CK) {
                    float r[TM_CK];
#pragma unroll
                    for (int c = 0; c < TM_CK; ++c) {
                        r[c] = st[TM_WIN + c * TM_PX + px];
                        rr = fmaf(r[c], r[c], rr);
                    }
#pragma unroll
                    for (int q = 0; q < NSLOT / 2; ++q) {
                        if (2 * q < nloc) {
                            const float* wa_ = st + coff[2 * q];
                            const float* wb_ = st + coff[2 * q + 1];
#pragma unroll
                            for (int c = 0; c < TM_CK; ++c) {
                                const tm_f2 R = tm_pk(r[c], r[c]);
                                const tm_f2 e0 = tm_sub2(tm_pk(wa_[c * TM_WC], wb_[c * TM_WC]), R);
                                const tm_f2 e1 = tm_sub2(tm_pk(wa_[c * TM_WC + 1], wb_[c * TM_WC + 1]), R);
                                const tm_f2 e2 = tm_sub2(tm_pk(wa_[c * TM_WC + TM_ROW], wb_[c * TM_WC + TM_ROW]), R);
                                const tm_f2 e3 = tm_sub2(tm_pk(wa_[c * TM_WC + TM_ROW + 1], wb_[c * TM_WC + TM_ROW + 1]), R);
                                tm_fma2(G2[q][0], e0, e0); tm_fma2(G2[q][1], e0, e1); tm_fma2(G2[q][2], e0, e2);
                                tm_fma2(G2[q][3], e0, e3); tm_fma2(G2[q][4], e1, e1); tm_fma2(G2[q][5], e1, e2);
                                tm_fma2(G2[q][6], e1, e3); tm_fma2(G2[q][7], e2, e2); tm_fma2(G2[q][8], e2, e3);
                                tm_fma2(G2[q][9], e3, e3);
                            }
                        }
                    }
                } else {
                    // last, partial chunk: only the channels that exist
                    for (int c = 0; c < nc; ++c) {
                        const float rc = st[TM_WIN + c * TM_PX + px];
                        rr = fmaf(rc, rc, rr);
                    }
#pragma unroll
                    for (int q = 0; q < NSLOT / 2; ++q) {
                        if (2 * q < nloc) {
                            const float* wa_ = st + coff[2 * q];
                            const float* wb_ = st + coff[2 * q + 1];
#pragma unroll 1
                            for (int c = 0; c < nc; ++c) {
                                const float rc = st[TM_WIN + c * TM_PX + px];
                                const tm_f2 R = tm_pk(rc, rc);
                                const tm_f2 e0 = tm_sub2(tm_pk(wa_[c * TM_WC], wb_[c * TM_WC]), R);
                                const tm_f2 e1 = tm_sub2(tm_pk(wa_[c * TM_WC + 1], wb_[c * TM_WC + 1]), R);
                                const tm_f2 e2 = tm_sub2(tm_pk(wa_[c * TM_WC + TM_ROW], wb_[c * TM_WC + TM_ROW]), R);
                                const tm_f2 e3 = tm_sub2(tm_pk(wa_[c * TM_WC + TM_ROW + 1], wb_[c * TM_WC + TM_ROW + 1]), R);
                                tm_fma2(G2[q][0], e0, e0); tm_fma2(G2[q][1], e0, e1); tm_fma2(G2[q][2], e0, e2);
                                tm_fma2(G2[q][3], e0, e3); tm_fma2(G2[q][4], e1, e1); tm_fma2(G2[q][5], e1, e2);
                                tm_fma2(G2[q][6], e1, e3); tm_fma2(G2[q][7], e2, e2); tm_fma2(G2[q][8], e2, e3);
                                tm_fma2(G2[q][9], e3, e3);
                            }
                        }
                    }
                }
                __syncthreads();   // every lane is done with this stage
                if (tid == 0 && chunk + TM_NSTAGE < nchunk) issue(chunk + TM_NSTAGE);
            }

            // ---------------- 4. planes -> result tile ------------------------------------------
            // The Gram matrices go through shared memory (the stage buffers are idle now) so that
            // every lane evaluates an equal share of the planes -- its own quarter, the same planes
            // it walked in step 1 -- whichever lane accumulated their run.  (Evaluating "the planes
            // of my runs" instead leaves most lanes idle while one works through a 25-plane run.)
            float* Gs = stage0;   // [run - first][10][PX]
            constexpr int GS_RUN = 10 * TM_PX;
            static_assert(TM_T * NSLOT * GS_RUN <= TM_NSTAGE * TM_STAGE, "Gram exchange fits in the stages");
#pragma unroll
            for (int j = 0; j < NSLOT; ++j) {
                if (j < nloc) {
                    float* gdst = Gs + (t + TM_T * j) * GS_RUN + px;
#pragma unroll
                    for (int i = 0; i < 10; ++i) {
                        float lo, hi;
                        tm_upk(G2[j / 2][i], lo, hi);
                        gdst[i * TM_PX] = (j & 1) ? hi : lo;
                    }
                }
            }
            __syncthreads();
            if (active && first < nrun) {   // (a pixel with fewer runs has nothing in this pass)
                const int last = min(first + TM_T * NSLOT, nrun);   // runs [first, last) are in Gs
                int k = max(wa, (int)kst_s[first * TM_PX + px]);
                const int kend = min(wb, (int)kst_s[last * TM_PX + px]);
                int run = max(my_run0, first);
                int knext = k;    // first plane of the run after `run`; forces a load on entry
                float gq[10];
                float fx0 = 0.f, fy0 = 0.f;
                bool outside = true;
                --run;
                for (; k < kend; ++k) {
                    if (k >= knext) {
                        do {
                            ++run;
                            knext = kst_s[(run + 1) * TM_PX + px];
                        } while (k >= knext);
                        const int pc = cell_s[run * TM_PX + px];
                        outside = (pc == kTmOutside);
                        fx0 = (float)tm_cell_x(pc); fy0 = (float)tm_cell_y(pc);
                        const float* gsrc = Gs + (run - first) * GS_RUN + px;
#pragma unroll
                        for (int i = 0; i < 10; ++i) gq[i] = gsrc[i * TM_PX];
                    }
                    float val = rr;
                    if (!outside) {
                        float ix, iy;
                        tm_coord<EXACT>(g, pt, d_s[k], ix, iy);
                        val = tm_quad(gq, ix - fx0, iy - fy0);
                    }
                    val *= inv_sigma;
                    float* o = out_s + k * TM_OS + px;
                    *o = (v == 0) ? val : (*o + val);
                }
            }
            __syncthreads();   // Gs (= the stages) is free again for the next pass / view
        }
    }

    // ---------------- 5. result tile -> global memory, one 128-byte row per warp-instruction ----
    __syncthreads();
    float lsm_m = 0.f, lsm_l = 0.f;   // per-column max and log-sum (lane = column in the copy-out too)
    if (a.lsm != nullptr) {   // log_softmax over the planes (host guarantees PS == 1)
        float* red = stage0;   // [2][T][PX] partials; the stages are idle
        float m = -INFINITY;
        for (int k = t; k < nk; k += TM_T) m = fmaxf(m, out_s[k * TM_OS + px]);
        red[t * TM_PX + px] = m;
        __syncthreads();
#pragma unroll
        for (int u = 0; u < TM_T; ++u) m = fmaxf(m, red[u * TM_PX + px]);
        // exp on the SFU (ex2.approx of (v - m) log2 e: relative error ~2e-7 for the |v - m| <~ 30 that
        // contribute to the sum at all); the logarithm of the sum stays logf
        float sum = 0.f;
        for (int k = t; k < nk; k += TM_T) sum += __expf(out_s[k * TM_OS + px] - m);
        red[(TM_T + t) * TM_PX + px] = sum;
        __syncthreads();
        sum = 0.f;
#pragma unroll
        for (int u = 0; u < TM_T; ++u) sum += red[(TM_T + u) * TM_PX + px];   // same order in every quarter
        lsm_m = m; lsm_l = logf(sum);
    }
    {
        const int col = px, x_out = tx * TM_PX + col;
        if (x_out < a.W) {
            const long long base = ((long long)b * a.D + k0) * HW + (long long)y * a.W + x_out;
            for (int k = t; k < nk; k += TM_T) {
                const float val = out_s[k * TM_OS + col];
                a.cost[base + (long long)k * HW] = val;
                if (a.lsm != nullptr) a.lsm[base + (long long)k * HW] = (val - lsm_m) - lsm_l;
            }
        }
    }
}

// ---- host side --------------------------------------------------------------------------------
tm_encode_fn tm_encoder() {
    static tm_encode_fn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return (tm_encode_fn)f;
    }();
    return fn;
}

static size_t tm_smem_bytes(int kper) {
    // dynamic part only: result tile + plane depths (stages and run tables are static, ~33 KB)
    const size_t n = (size_t)kper * TM_OS * sizeof(float) + (size_t)kper * sizeof(float);
    return (n + 15) & ~(size_t)15;
}

// Can the TMA kernel take this call?  (16-byte aligned bases and strides, plane block small enough
// for the 64-bit run-start mask.)
bool sweep_gram_tma_supported(const SweepArgs& a) {
    const int kper = (a.D + a.PS - 1) / a.PS;
    if (a.W % 4 != 0 || kper > 4 * 64 || a.W >= 32767 || a.H >= 32767) return false;
    if (((uintptr_t)a.ref | (uintptr_t)a.src) & 15) return false;
    if ((a.ref_bs | a.src_bs | a.src_vs) & 3) return false;
    if (a.ref_bs < 0 || a.src_bs < 0 || a.src_vs < 0) return false;
    if (a.B > 1 && (a.ref_bs == 0 || a.src_bs == 0)) return false;
    if (a.V > 1 && a.src_vs == 0) return false;
    if (tm_smem_bytes(kper) > 160 * 1024) return false;
    return tm_encoder() != nullptr;
}

int launch_sweep_gram_tma(const SweepArgs& a, cudaStream_t st) {
    static const int exact_env = [] { const char* e = getenv("DPV_SWEEP_TMA_EXACT"); return e ? atoi(e) : -1; }();
    if (!sweep_gram_tma_supported(a)) return DPV_E_UNSUPP;
    tm_encode_fn enc = tm_encoder();
    const int kper = (a.D + a.PS - 1) / a.PS;
    const cuuint64_t chw = (cuuint64_t)a.C * a.H * a.W;
    CUtensorMap msrc, mref;
    {
        const cuuint64_t vs = a.V > 1 ? (cuuint64_t)a.src_vs : chw;
        const cuuint64_t bs = a.B > 1 ? (cuuint64_t)a.src_bs : vs * a.V;
        const cuuint64_t gdim[5] = {(cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.C, (cuuint64_t)a.V, (cuuint64_t)a.B};
        const cuuint64_t gstr[4] = {(cuuint64_t)a.W * 4, (cuuint64_t)a.H * a.W * 4, vs * 4, bs * 4};
        const cuuint32_t box[5] = {TM_WC, 1, TM_CK, 1, 1};
        const cuuint32_t est[5] = {1, 1, 1, 1, 1};
        if (enc(&msrc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(a.src), gdim, gstr, box, est,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return DPV_E_UNSUPP;
    }
    {
        const cuuint64_t bs = a.B > 1 ? (cuuint64_t)a.ref_bs : chw;
        const cuuint64_t gdim[4] = {(cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.C, (cuuint64_t)a.B};
        const cuuint64_t gstr[3] = {(cuuint64_t)a.W * 4, (cuuint64_t)a.H * a.W * 4, bs * 4};
        const cuuint32_t box[4] = {TM_PX, 1, TM_CK, 1};
        const cuuint32_t est[4] = {1, 1, 1, 1};
        if (enc(&mref, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a.ref), gdim, gstr, box, est,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return DPV_E_UNSUPP;
    }
    const size_t smem = tm_smem_bytes(kper);
    const int tiles = ((a.W + TM_PX - 1) / TM_PX) * a.H;
    dim3 grid(tiles, a.PS, a.B), block(TM_NT);
    // coordinates: reference operation order for wide images (see tm_coord), SFU form otherwise
    const bool exact = exact_env >= 0 ? (exact_env != 0) : (a.W > 192 || a.H > 192);
    cudaError_t e;
    if (exact) {
        e = cudaFuncSetAttribute(sweep_gram_tma_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        sweep_gram_tma_kernel<4, true><<<grid, block, smem, st>>>(a, msrc, mref);
    } else {
        e = cudaFuncSetAttribute(sweep_gram_tma_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        sweep_gram_tma_kernel<4, false><<<grid, block, smem, st>>>(a, msrc, mref);
    }
    DPV_LAUNCH_END();
    return 0;
}

}  // namespace dpv
