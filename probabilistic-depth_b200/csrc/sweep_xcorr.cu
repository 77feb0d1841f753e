// K1 + K2a, cross-correlation form of the L2 cost volume (algo = 5; production path when the caller
// provides a workspace).  warping/homography.py:80-86,98-135 (est_swp_volume_v4) fused with :170-198
// (_back_warp_homo_parallel).
//
// Same geometry machinery as the per-cell Gram kernel (sweep_tma.cu): a CTA owns a row segment of 32
// reference pixels, four threads per pixel (lane = pixel, warp = plane quarter); the D sampling positions of a
// pixel fall into a handful of source 2x2 cells ("runs" of consecutive planes), and every plane of a run is
// the 10-term quadratic form of its bilinear weights in the cell's Gram matrix G_kl = <s_k - r, s_l - r>
// (s_k: the four taps, r: the reference pixel, both C-vectors).  What changed is how G is obtained.  The
// Gram kernel accumulates all ten products per cell and channel (14 FMA + 8 LDS per cell and channel; 41 %
// of its instructions).  Expanding the differences,
//     G_kl = <s_k, s_l>  -  <s_k, r>  -  <s_l, r>  +  <r, r>
// separates what depends on the reference pixel from what does not:
//   * <s_k, s_l> involves source pixels only: five maps per source image (norm, and the products with the
//     right / lower / lower-right / lower-left-of-right neighbour), and <r, r> one map per reference image,
//     computed ONCE per call by a small pre-pass kernel (sweep_smaps_kernel: 5 C FMAs per source pixel
//     instead of 10 C per cell and reference pixel) into a caller-provided workspace, zero-padded by one
//     pixel so that grid_sample's zeros padding needs no test;
//   * <s, r> is ONE product per (reference pixel, source pixel) pair and channel.  All pixels of a tile need
//     the same range of column offsets dx = x_src - x_ref (the disparity range of the tile) on the window's
//     rows, so the tile computes the banded correlation P[row][dx][pixel] = sum_c r_c[pixel] * s_c[row][pixel
//     + dx] with register tiles of 4 pixels x 8 offsets: per channel 4 x LDS.128 (4 reference values, 12
//     source values) feed 32 FFMA -- 1.1 instructions per product instead of the Gram form's 2.7 per
//     product-equivalent, and 3-10x fewer products (one per tap instead of ten per cell).
// The source window and the reference pixels arrive by TMA (cp.async.bulk.tensor) in chunks of 8 channels
// through a ring of up to 8 stages (as many as the window height allows: the per-chunk arithmetic is now
// too short to hide a copy behind one other stage); out-of-image taps and channels past C are zero-filled
// by the copy engine.  Warps split the (row, dx-group, pixel-group) items and, when a tile has few items,
// the channels of a chunk; the partial correlations meet in shared memory in a fixed order.
// Precision: G is formed from fp32 sums of ~C terms by three subtractions, i.e. with an absolute error of a
// few ulp of <r, r> (~1e-5 for unit-variance features, C = 67) where the Gram form has a relative one; the
// cost volume agrees with the reference to <= 2e-5 relative on the goldens, and a source equal to the
// reference under the identity pose gives |cost| <= 1e-5 instead of exactly 0
// (tests/test_gpu_parity.py::test_sweep_identity_pose_is_near_zero: bar 6.7e-4).
// Tiles whose window does not fit (more than 6 rows, a disparity range above 23 columns, more than 24 runs)
// gather from global memory like the Gram kernel's.
#include <cstdlib>

#include "sweep_tma.cuh"

namespace dpv {

constexpr int XC_T = 4, XC_NT = TM_PX * XC_T;
constexpr int XC_WC = 56, XC_WR = 6;             // source window capacity (cols, rows)
constexpr int XC_WCW = 96, XC_WRW = 3;           // wide window for few rows: 32 pixels + up to 64 column offsets
constexpr int XC_CK = 8;                         // channels per chunk
constexpr int XC_ROW = XC_CK * XC_WC;            // floats of one window row of a chunk
constexpr int XC_AREA = 2 * (XC_WR * XC_ROW + XC_CK * TM_PX);   // stage ring / correlation tile (floats)
constexpr int XC_MAXSTAGE = 8;
constexpr int XC_MAXITEM = 128;                  // (row, dx-group, pixel-group) items per tile: one per thread
constexpr int XC_NMAP = 5;                       // source-only product maps per source image
constexpr int XC_PRE_NT = 256, XC_PRE_SL = XC_PRE_NT / 32;   // pre-pass: 32 positions x 8 channel slices

struct XcMaps {
    CUtensorMap src[XC_WR];   // box {XC_WC cols, h rows, XC_CK channels} for h = 1 .. XC_WR: one copy per chunk
    CUtensorMap wide[XC_WRW]; // box {XC_WCW cols, h rows, XC_CK channels} for h = 1 .. XC_WRW (large images:
                              // a long, flat disparity range, e.g. rectified stereo at full resolution)
    CUtensorMap ref;          // box {32 pixels, 1 row, XC_CK channels}
};

struct XcShared {
    int dx[2];         // min / max over the tile of (cell x0 - pixel x)
    int yy[2];         // min / max of cell y0
    int wide;          // a pixel whose sampling positions are not an ordered segment (pole of the projection)
};

// ---------------------------------------------------------------------------------------------------
// Pre-pass: the products that do not involve the reference pixel.  grid (ceil((H+2)(W+2)/32), V + 1, B),
// 256 threads = 32 grid positions (lanes) x 8 channel slices (warps): every thread takes the channels
// c = slice, slice + 8, ... (independent loads, several in flight), the slices meet in shared memory and
// are added in slice order.  blockIdx.y < V: the five maps of source image (b, y) on the grid padded by one
// pixel on every side (entry (yp, xp) <-> pixel (yp - 1, xp - 1); products with a pixel outside the image
// are zero, which is what sampling zeros padding means); blockIdx.y == V: <r, r> of the reference image.
//   map 0  N (x, y) = <s(x, y), s(x, y)>          map 1  Hp(x, y) = <s(x, y), s(x + 1, y)>
//   map 2  Vp(x, y) = <s(x, y), s(x, y + 1)>      map 3  Dp(x, y) = <s(x, y), s(x + 1, y + 1)>
//   map 4  Ap(x, y) = <s(x + 1, y), s(x, y + 1)>
__global__ void __launch_bounds__(XC_PRE_NT) sweep_smaps_kernel(const SweepArgs a, float* __restrict__ smaps,
                                                                float* __restrict__ refn) {
    __shared__ float red[XC_PRE_SL][XC_NMAP][32];
    pdl_trigger();     // the sweep kernel proper may become resident; it waits before it reads these maps
    const int HW = a.H * a.W, Wp = a.W + 2, MP = (a.H + 2) * Wp;
    const int b = blockIdx.z, v = blockIdx.y;
    const int lane = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int q = blockIdx.x * 32 + lane;
    if (v == a.V) {
        if (blockIdx.x * 32 >= HW) return;
        float s = 0.f;
        if (q < HW) {
            const float* r = a.ref + (long long)b * a.ref_bs + q;
#pragma unroll 4
            for (int c = sl; c < a.C; c += XC_PRE_SL) {
                const float t = __ldg(r + (long long)c * HW);
                s = fmaf(t, t, s);
            }
        }
        red[sl][0][lane] = s;
        __syncthreads();
        if (sl == 0 && q < HW) {
            float t = red[0][0][lane];
#pragma unroll
            for (int i = 1; i < XC_PRE_SL; ++i) t += red[i][0][lane];
            refn[(long long)b * HW + q] = t;
        }
        return;
    }
    float n = 0.f, hp = 0.f, vp = 0.f, dp = 0.f, ap = 0.f;
    if (q < MP) {
        const int yp = q / Wp, xp = q - yp * Wp;
        const int x = xp - 1, y = yp - 1;
        const bool x0 = (x >= 0) & (x < a.W), x1 = (x + 1 >= 0) & (x + 1 < a.W);
        const bool y0 = (y >= 0) & (y < a.H), y1 = (y + 1 >= 0) & (y + 1 < a.H);
        const bool v00 = x0 & y0, v01 = x1 & y0, v10 = x0 & y1, v11 = x1 & y1;
        const float* s = a.src + (long long)b * a.src_bs + (long long)v * a.src_vs;
        const int o = y * a.W + x;
#pragma unroll 3
        for (int c = sl; c < a.C; c += XC_PRE_SL) {
            const float* sc = s + (long long)c * HW;
            const float s00 = v00 ? __ldg(sc + o) : 0.f;
            const float s01 = v01 ? __ldg(sc + o + 1) : 0.f;
            const float s10 = v10 ? __ldg(sc + o + a.W) : 0.f;
            const float s11 = v11 ? __ldg(sc + o + a.W + 1) : 0.f;
            n = fmaf(s00, s00, n); hp = fmaf(s00, s01, hp); vp = fmaf(s00, s10, vp);
            dp = fmaf(s00, s11, dp); ap = fmaf(s01, s10, ap);
        }
    }
    red[sl][0][lane] = n; red[sl][1][lane] = hp; red[sl][2][lane] = vp; red[sl][3][lane] = dp; red[sl][4][lane] = ap;
    __syncthreads();
    if (sl < XC_NMAP && q < MP) {      // warp m adds map m over the slices, in slice order
        float t = red[0][sl][lane];
#pragma unroll
        for (int i = 1; i < XC_PRE_SL; ++i) t += red[i][sl][lane];
        smaps[(((long long)b * a.V + v) * XC_NMAP + sl) * MP + q] = t;
    }
}

long long sweep_xcorr_workspace_floats(int B, int V, int H, int W) {
    if (B <= 0 || V <= 0 || H <= 0 || W <= 0) return 0;
    return (long long)B * V * XC_NMAP * (H + 2) * (W + 2) + (long long)B * H * W + 8;
}

// ---------------------------------------------------------------------------------------------------
// grid (tiles per row, H, B * PS): blockIdx.z = b * PS + plane block.
template <bool EXACT, int MINB>
__global__ void __launch_bounds__(XC_NT, MINB)
sweep_xcorr_kernel(const SweepArgs a, const float* smaps, const float* refn,
                   const __grid_constant__ XcMaps maps) {
    __shared__ __align__(128) float area[XC_AREA];                     // TMA stage ring, then P[row][dx][pixel]
    extern __shared__ __align__(16) float out_s[];                     // [kper][PX] result tile
    __shared__ float geo_s[20];          // K R (9), K t (3), cx, cy, 1/cx, 1/cy of the view; d range
    __shared__ unsigned long long full_bar[XC_MAXSTAGE];
    __shared__ XcShared ts;

    const int HW = a.H * a.W;
    const int kper = (a.PS == 1) ? a.D : (a.D + a.PS - 1) / a.PS;
    float* d_s = out_s + kper * TM_OS;                                 // [kper]
    const int b = (a.PS == 1) ? blockIdx.z : blockIdx.z / a.PS;
    const int k0 = ((a.PS == 1) ? 0 : blockIdx.z - b * a.PS) * kper;
    const int nk = min(a.D, k0 + kper) - k0;
    const int tid = threadIdx.x, lane = tid & 31;
    const int px = tid & 31, t = tid >> 5;   // lane = pixel, warp = plane quarter
    const int y = blockIdx.y, tx = blockIdx.x;
    const int x = tx * TM_PX + px;
    const bool active = x < a.W;
    const int p = active ? y * a.W + x : y * a.W;
    if (nk <= 0) return;
    for (int k = tid; k < nk; k += XC_NT) d_s[k] = __ldg(a.d + k0 + k);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < XC_MAXSTAGE; ++s) tm_mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid == 64) geo_s[18] = __frcp_rn(a.sigma);

    const float* rays = a.rays + (long long)b * a.rays_bs;
    const float rx = __ldg(rays + p), ry = __ldg(rays + HW + p), rz = __ldg(rays + 2 * HW + p);
    const float* ref = a.ref + (long long)b * a.ref_bs;
    const int nchunk = (a.C + XC_CK - 1) / XC_CK;
    const int Wp = a.W + 2, MP = (a.H + 2) * Wp;
    unsigned stage_phase = 0;          // bit s: parity the next wait on stage s expects
    float rr = 0.f;                    // <r, r> of this pixel (read after the pre-pass has completed)
    bool have_rr = false;

    for (int v = 0; v < a.V; ++v) {
        const float* src = a.src + (long long)b * a.src_bs + (long long)v * a.src_vs;
        // ---------------- 0. view geometry: K R and K t once per CTA ---------------------------
        if (tid < 12) {   // warping/homography.py:119-121; same dot3 order as load_view_geom
            const float* Kp = a.K + (long long)b * a.k_bs + (tid < 9 ? tid / 3 : tid - 9) * 3;
            const float* pp = a.pose + (long long)b * a.pose_bs + (long long)v * 16 + (tid < 9 ? tid % 3 : 3);
            geo_s[tid] = dot3(__ldg(Kp), __ldg(Kp + 1), __ldg(Kp + 2), __ldg(pp), __ldg(pp + 4), __ldg(pp + 8));
        } else if (tid < 14) {
            const float c = __ldg(a.K + (long long)b * a.k_bs + (tid == 12 ? 2 : 5));   // cx, cy
            geo_s[tid] = c;
            geo_s[tid + 2] = __frcp_rn(c);
        }
        __syncthreads();   // geo_s, d_s, barrier init; previous view done with area
        // (the per-pixel term and the view constants are rebuilt from shared memory where they are used, so that
        // they do not occupy registers across the correlation loop)
        auto load_geom = [&](PixelTerm& pt, TmGeom& g) {
            pt.x = dot3(geo_s[0], geo_s[1], geo_s[2], rx, ry, rz);
            pt.y = dot3(geo_s[3], geo_s[4], geo_s[5], rx, ry, rz);
            pt.z = dot3(geo_s[6], geo_s[7], geo_s[8], rx, ry, rz);
            g.t1x = geo_s[9]; g.t1y = geo_s[10]; g.t1z = geo_s[11]; g.cx = geo_s[12]; g.cy = geo_s[13];
            g.inv_cx = geo_s[14]; g.inv_cy = geo_s[15];
            g.half_w = (float)a.W * 0.5f; g.half_h = (float)a.H * 0.5f;
        };

        // The planes are taken in blocks [cur, cur + len): a block is shrunk (halved) until its band fits the
        // shared-memory window -- at the model's 64 x 96 the first block is all D planes; on a 256 x 384 or
        // 384 x 1280 image the disparity range of all planes (40+ columns) does not fit and the near planes
        // go in blocks of 4-8, the far ones in blocks of 16-32.  After a block that fitted, the next one is
        // tried at twice its length.
        int cur = 0, len = nk;
        while (cur < nk) {
        len = min(len, nk - cur);
        if (tid == 32) {
            ts.dx[0] = 1 << 30; ts.dx[1] = -(1 << 30); ts.yy[0] = 1 << 30; ts.yy[1] = -(1 << 30);
            ts.wide = 0;
        }
        __syncthreads();   // ts; the previous block is done with the correlation tile in `area`
        const int kpt = (len + XC_T - 1) >> 2;
        const int wa = min(cur + len, cur + t * kpt), wb = min(cur + len, wa + kpt);   // planes of this thread

        // ---------------- 1. the band of the tile --------------------------------------------------
        // Along a ray the source coordinate is a linear-fractional function of the depth, hence monotone
        // between the smallest and the largest plane depth as long as the projective denominator keeps its
        // sign there: the cells of ALL planes of a pixel lie between the cells of those two depths (floor is
        // monotone; a plane kept in its predecessor's cell by the border slack is in a cell already counted).
        // Warp 0 takes the smallest depth, warp 1 the largest; a pixel whose denominator changes sign or
        // whose end points are not finite sends the tile to the gather path.
        if (t < 2) {
            int cx0 = 1 << 30, cx1 = -(1 << 30), cy0 = 1 << 30, cy1 = -(1 << 30);
            bool wide = false;
            // smallest and largest plane depth of the block (the planes need not be sorted)
            float dmin = INFINITY, dmax = -INFINITY;
            for (int k = cur + lane; k < cur + len; k += 32) {
                const float dk = d_s[k];
                dmin = fminf(dmin, dk); dmax = fmaxf(dmax, dk);
            }
            dmin = -warp_max(-dmin); dmax = warp_max(dmax);
            if (active) {
                PixelTerm pt;
                TmGeom g;
                load_geom(pt, g);
                const float dk = t == 0 ? dmin : dmax;
                const float den = fmaf(pt.z, dk, g.t1z);
                const float den_o = fmaf(pt.z, t == 0 ? dmax : dmin, g.t1z);
                wide = !(den > 1e-6f && den_o > 1e-6f);
                float ix, iy;
                tm_coord<EXACT>(g, pt, dk, ix, iy);
                const Tap tap = make_tap(ix, iy);
                const bool inside = (tap.x0 >= -1) & (tap.x0 < a.W) & (tap.y0 >= -1) & (tap.y0 < a.H);
                // an end point outside the image: clamp its cell to the image border cells -- every in-image
                // cell of the segment still lies between the two clamped end cells
                if (tap.x0 > -1000000) {
                    // a position within the border slack below an integer counts for the upper cell (phase 3
                    // puts every plane into a band cell it is within the slack of): a coordinate that is
                    // constant along the ray up to rounding -- the row, in rectified stereo -- then costs one
                    // cell row instead of two
                    const int sx = tap.x0 + (tap.fx > 1.0f - kCellSlack ? 1 : 0);
                    const int sy = tap.y0 + (tap.fy > 1.0f - kCellSlack ? 1 : 0);
                    const int ex = min(max(sx, -1), a.W - 1), ey = min(max(sy, -1), a.H - 1);
                    cx0 = cx1 = ex - x; cy0 = cy1 = ey;
                    (void)inside;
                } else {
                    wide = true;
                }
            }
            cx0 = __reduce_min_sync(0xffffffffu, cx0); cx1 = __reduce_max_sync(0xffffffffu, cx1);
            cy0 = __reduce_min_sync(0xffffffffu, cy0); cy1 = __reduce_max_sync(0xffffffffu, cy1);
            const int wd = __any_sync(0xffffffffu, wide);
            if (lane == 0) {
                atomicMin(&ts.dx[0], cx0); atomicMax(&ts.dx[1], cx1);
                atomicMin(&ts.yy[0], cy0); atomicMax(&ts.yy[1], cy1);
                if (wd) atomicOr(&ts.wide, 1);
            }
        }
        __syncthreads();
        // dx in [dlo, dlo + 8 ng8) with dlo a multiple of 4 (then every 128-bit shared-memory access below is
        // aligned, and so is the TMA box: global column tx * 32 + dlo), rows [wy0, wy0 + wh).
        const bool any_cell = ts.dx[0] <= ts.dx[1];
        const int dlo = any_cell ? (ts.dx[0] & ~3) : 0;
        const int ng8 = any_cell ? (ts.dx[1] + 2 - dlo + 7) >> 3 : 0;      // taps reach x0 + 1
        const int wy0 = any_cell ? ts.yy[0] : 0;
        const int wh = any_cell ? ts.yy[1] + 2 - wy0 : 0;                   // taps reach y0 + 1
        const int nitem = wh * ng8 * 8;
        const int ndxp = 8 * ng8;
        // window width: 56 columns (up to 6 rows), or 96 columns when the band is long and flat (up to 3 rows)
        const bool use_wide = (ng8 > (XC_WC - TM_PX) / 8) && (wh <= XC_WRW);
        const int wc = use_wide ? XC_WCW : XC_WC;
        const bool fits = (ng8 * 8 + TM_PX <= wc) && (wh <= XC_WR) && (nitem <= XC_MAXITEM) && !ts.wide;
        const bool pole = ts.wide != 0;
        __syncthreads();   // ts is re-initialised at the top of the next block

        if (!fits) {       // (uniform across the CTA)
            if (!pole && len > 1) {
                len = (len + 1) >> 1;      // same start, half the planes
                continue;
            }
            // a single plane whose band does not fit, or a pole of the projection inside the block: gather
            if (active && wa < wb) {
                PixelTerm pt;
                TmGeom g;
                load_geom(pt, g);
                tm_gather_planes<EXACT>(a.C, a.H, a.W, src, ref + p, g, pt, d_s, wa, wb, geo_s[18], out_s + px, v == 0);
            }
            cur += len;
            len = 2 * len;
            continue;
        }

        // ---------------- 2. banded correlation of the tile ---------------------------------------
        if (nitem > 0) {
            const int win_floats = wh * XC_CK * wc;
            const int stage_floats = win_floats + XC_CK * TM_PX;
            // stages that fit the ring: XC_AREA / stage_floats for wh = 1..6 (56 columns), 1..3 (96 columns)
            const int ns = min(use_wide ? (wh == 1 ? 5 : wh == 2 ? 3 : 2) : (wh == 1 ? 8 : wh == 2 ? 5 : wh == 3 ? 3 : 2),
                               nchunk);
            const int wx0 = tx * TM_PX + dlo;
            // warps split the items and, when there are few, the channels of a chunk
            const int iw_n = nitem <= 32 ? 1 : (nitem <= 64 ? 2 : 4);   // warps side by side on items
            const int cpp = 2 * iw_n;                                     // channels per warp and chunk: 2, 4, 8
            const int iw = t & (iw_n - 1), cw = (iw_n == 1) ? t : (iw_n == 2 ? t >> 1 : 0);
            const int item = iw * 32 + lane;
            const bool has_item = item < nitem;
            const int it = has_item ? item : 0;
            const int ig = it & 7, rest = it >> 3;                        // rest = row * ng8 + dx group, < 16
            const int ir = ng8 == 1 ? rest : (ng8 == 2 ? rest >> 1 : (ng8 == 3 ? (rest * 11) >> 5 : rest / ng8));
            const int im = rest - ir * ng8;
            // stage layout (the order a bulk tensor copy writes its box): [channel][row][col], then the
            // reference pixels [channel][pixel]
            const int chs = wh * wc;                                      // channel stride of the window
            const int s_off = cw * cpp * chs + ir * wc + 4 * ig + 8 * im;
            const int r_off = win_floats + 4 * ig + cw * cpp * TM_PX;
            float acc[4][8];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int u = 0; u < 8; ++u) acc[j][u] = 0.f;

            // chunk c is issued by lane 0 of warp c % 4 into stage c % ns
            auto issue = [&](int chunk) {
                const int s = chunk % ns;
                float* st = area + s * stage_floats;
                // the ring is also used through the generic proxy (the correlation tile, the soft-max partials):
                // those accesses are ordered before this point by __syncthreads; this fence orders them before
                // the async-proxy writes of the bulk copies
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tm_mbar_expect_tx(&full_bar[s], (unsigned)(stage_floats * sizeof(float)));
                tm_load_5d(st, use_wide ? &maps.wide[wh - 1] : &maps.src[wh - 1], wx0, wy0, chunk * XC_CK, v, b,
                           &full_bar[s]);
                tm_load_4d(st + win_floats, &maps.ref, tx * TM_PX, y, chunk * XC_CK, b, &full_bar[s]);
            };
            if (lane == 0) {
                for (int c = t; c < ns; c += XC_T) issue(c);
            }
            int s_use = 0;
            for (int chunk = 0; chunk < nchunk; ++chunk) {
                const int s = s_use;
                s_use = (s_use + 1 == ns) ? 0 : s_use + 1;
                tm_mbar_wait(&full_bar[s], (stage_phase >> s) & 1u);
                stage_phase ^= 1u << s;
                if (has_item) {
                    const float* sp = area + s * stage_floats + s_off;
                    const float* rp = area + s * stage_floats + r_off;
#pragma unroll 2
                    for (int c = 0; c < cpp; ++c) {
                        const float4 r4 = *reinterpret_cast<const float4*>(rp + c * TM_PX);
                        const float4 s0 = *reinterpret_cast<const float4*>(sp + c * chs);
                        const float4 s1 = *reinterpret_cast<const float4*>(sp + c * chs + 4);
                        const float4 s2 = *reinterpret_cast<const float4*>(sp + c * chs + 8);
                        const float rv[4] = {r4.x, r4.y, r4.z, r4.w};
                        const float sv[12] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w, s2.x, s2.y, s2.z, s2.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j)
#pragma unroll
                            for (int u = 0; u < 8; ++u) acc[j][u] = fmaf(rv[j], sv[j + u], acc[j][u]);
                    }
                }
                __syncthreads();   // every warp is done with this stage
                if (lane == 0 && chunk + ns < nchunk && ((chunk + ns) & (XC_T - 1)) == t) issue(chunk + ns);
            }
            // partial correlations -> shared memory, one copy per channel part: [cw][row][dx][pixel]
            const int psize = nitem * TM_PX;      // = wh * ndxp * 32
            if (has_item) {
                float* pd = area + cw * psize + (ir * ndxp + 8 * im) * TM_PX + 4 * ig;
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    *reinterpret_cast<float4*>(pd + u * TM_PX) = make_float4(acc[0][u], acc[1][u], acc[2][u], acc[3][u]);
            }
            __syncthreads();
            if (iw_n < 4) {     // add the channel parts in a fixed order into copy 0
                const int ncopy = iw_n == 1 ? 4 : 2;
                for (int i4 = tid; i4 < (psize >> 2); i4 += XC_NT) {
                    float4 s4 = *reinterpret_cast<const float4*>(area + 4 * i4);
                    for (int c = 1; c < ncopy; ++c) {
                        const float4 o4 = *reinterpret_cast<const float4*>(area + c * psize + 4 * i4);
                        s4.x += o4.x; s4.y += o4.y; s4.z += o4.z; s4.w += o4.w;
                    }
                    *reinterpret_cast<float4*>(area + 4 * i4) = s4;
                }
                __syncthreads();
            }
        }

        // ---------------- 3. planes -> result tile ----------------------------------------------
        // Each thread walks its quarter of the planes.  A plane stays in the cell of its predecessor while its
        // position is within the cell (plus the border slack: the two bilinear forms agree on a shared edge);
        // on entering a new cell the Gram matrix of the cell is assembled from the correlation tile, the
        // source-only maps and <r, r>.
        if (!have_rr) {
            pdl_wait();        // the pre-pass (source-only maps, <r, r>) has completed
            rr = active ? *(refn + (long long)b * HW + p) : 0.f;
            have_rr = true;
        }
        if (active && wa < wb) {
            PixelTerm pt;
            TmGeom g;
            load_geom(pt, g);
            const float inv_sigma = geo_s[18];
            const float* sm = smaps + ((long long)b * a.V + v) * XC_NMAP * MP;
            float gq[10];
            float fx0 = 0.f, fy0 = 0.f;
            bool in_cell = false, outside = false;     // in_cell: gq / fx0 / fy0 describe an in-image cell
            for (int k = wa; k < wb; ++k) {
                float ix, iy;
                tm_coord<EXACT>(g, pt, d_s[k], ix, iy);
                float fx = ix - fx0, fy = iy - fy0;
                if (!(in_cell && fx >= -kCellSlack && fx <= 1.0f + kCellSlack && fy >= -kCellSlack &&
                      fy <= 1.0f + kCellSlack)) {
                    const Tap tap = make_tap(ix, iy);
                    const bool inside = (tap.x0 >= -1) & (tap.x0 < a.W) & (tap.y0 >= -1) & (tap.y0 < a.H);
                    in_cell = inside; outside = !inside;
                    if (inside) {
                        // Monotonicity holds in exact arithmetic; a coordinate that does not change along the ray
                        // (rectified stereo: the row) wobbles by an ulp around an integer, so a middle plane may
                        // floor to the cell next to the end points' cells.  It is within the border slack of
                        // the band, where both cells give the same value: take the band's cell (and never
                        // index outside the correlation tile).
                        const int cx0 = min(max(tap.x0, x + dlo), x + dlo + ndxp - 2);
                        const int cy0 = min(max(tap.y0, wy0), wy0 + wh - 2);
                        fx0 = (float)cx0; fy0 = (float)cy0;
                        fx = ix - fx0; fy = iy - fy0;
                        const float* P = area + ((cy0 - wy0) * ndxp + (cx0 - x - dlo)) * TM_PX + px;
                        const float p00 = P[0], p01 = P[TM_PX], p10 = P[ndxp * TM_PX], p11 = P[ndxp * TM_PX + TM_PX];
                        // (plain loads: written by the pre-pass while this kernel was already resident, so not
                        // the read-only path; L1 cannot hold them from before the wait)
                        const float* q = sm + (cy0 + 1) * Wp + (cx0 + 1);
                        const float n00 = q[0], n01 = q[1], n10 = q[Wp], n11 = q[Wp + 1];
                        const float h0 = q[MP], h1 = q[MP + Wp];
                        const float v0 = q[2 * MP], v1 = q[2 * MP + 1];
                        const float dg = q[3 * MP], ad = q[4 * MP];
                        // G_kl = <s_k, s_l> - <s_l, r> + (<r, r> - <s_k, r>)
                        const float a0 = rr - p00, a1 = rr - p01, a2 = rr - p10, a3 = rr - p11;
                        gq[0] = (n00 - p00) + a0; gq[1] = (h0 - p01) + a0; gq[2] = (v0 - p10) + a0; gq[3] = (dg - p11) + a0;
                        gq[4] = (n01 - p01) + a1; gq[5] = (ad - p10) + a1; gq[6] = (v1 - p11) + a1;
                        gq[7] = (n10 - p10) + a2; gq[8] = (h1 - p11) + a2;
                        gq[9] = (n11 - p11) + a3;
                    }
                }
                float val = outside ? rr : tm_quad(gq, fx, fy);
                val *= inv_sigma;
                float* o = out_s + k * TM_OS + px;
                *o = (v == 0) ? val : (*o + val);
            }
        }
        cur += len;
        len = 2 * len;
        }   // plane blocks (the next block's / view's first __syncthreads orders these reads of `area`)
    }
    // ---------------- 4. result tile -> global memory, one 128-byte row per warp-instruction ----
    if (!have_rr) pdl_wait();   // (tiles that only gathered) never complete ahead of the pre-pass
    __syncthreads();
    float lsm_m = 0.f, lsm_l = 0.f;   // per-column max and log-sum (lane = column in the copy-out too)
    if (a.lsm != nullptr) {   // log_softmax over the planes (host guarantees PS == 1)
        float* red = area;    // [2][T][PX] partials; the ring is idle
        float m = -INFINITY;
        for (int k = t; k < nk; k += XC_T) m = fmaxf(m, out_s[k * TM_OS + px]);
        red[t * TM_PX + px] = m;
        __syncthreads();
#pragma unroll
        for (int u = 0; u < XC_T; ++u) m = fmaxf(m, red[u * TM_PX + px]);
        float sum = 0.f;
        for (int k = t; k < nk; k += XC_T) sum += __expf(out_s[k * TM_OS + px] - m);
        red[(XC_T + t) * TM_PX + px] = sum;
        __syncthreads();
        sum = 0.f;
#pragma unroll
        for (int u = 0; u < XC_T; ++u) sum += red[(XC_T + u) * TM_PX + px];   // same order in every quarter
        lsm_m = m; lsm_l = logf(sum);
    }
    {
        const int col = px, x_out = tx * TM_PX + col;
        if (x_out < a.W) {
            const long long base = ((long long)b * a.D + k0) * HW + (long long)y * a.W + x_out;
            for (int k = t; k < nk; k += XC_T) {
                const float val = out_s[k * TM_OS + col];
                a.cost[base + (long long)k * HW] = val;
                if (a.lsm != nullptr) a.lsm[base + (long long)k * HW] = (val - lsm_m) - lsm_l;
            }
        }
    }
}

// ---- host side --------------------------------------------------------------------------------
static size_t xc_smem_bytes(int kper) {
    // dynamic part only: result tile + plane depths
    const size_t n = (size_t)kper * TM_OS * sizeof(float) + (size_t)kper * sizeof(float);
    return (n + 15) & ~(size_t)15;
}

bool sweep_xcorr_supported(const SweepArgs& a) {
    const int kper = (a.D + a.PS - 1) / a.PS;
    if (a.W % 4 != 0 || a.W >= 32767 || a.H >= 32767 || (long long)a.B * a.PS > 65535) return false;
    if (((uintptr_t)a.ref | (uintptr_t)a.src) & 15) return false;
    if ((a.ref_bs | a.src_bs | a.src_vs) & 3) return false;
    if (a.ref_bs < 0 || a.src_bs < 0 || a.src_vs < 0) return false;
    if (a.B > 1 && (a.ref_bs == 0 || a.src_bs == 0)) return false;
    if (a.V > 1 && a.src_vs == 0) return false;
    if (xc_smem_bytes(kper) > 160 * 1024) return false;
    return tm_encoder() != nullptr;
}

int launch_sweep_xcorr(const SweepArgs& a, float* workspace, cudaStream_t st) {
    static const int exact_env = [] { const char* e = getenv("DPV_SWEEP_TMA_EXACT"); return e ? atoi(e) : -1; }();
    if (!workspace || !sweep_xcorr_supported(a)) return DPV_E_UNSUPP;
    tm_encode_fn enc = tm_encoder();
    const int kper = (a.D + a.PS - 1) / a.PS;
    const cuuint64_t chw = (cuuint64_t)a.C * a.H * a.W;
    XcMaps maps;
    {
        const cuuint64_t vs = a.V > 1 ? (cuuint64_t)a.src_vs : chw;
        const cuuint64_t bs = a.B > 1 ? (cuuint64_t)a.src_bs : vs * a.V;
        const cuuint64_t gdim[5] = {(cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.C, (cuuint64_t)a.V, (cuuint64_t)a.B};
        const cuuint64_t gstr[4] = {(cuuint64_t)a.W * 4, (cuuint64_t)a.H * a.W * 4, vs * 4, bs * 4};
        const cuuint32_t est[5] = {1, 1, 1, 1, 1};
        for (int hgt = 1; hgt <= XC_WR + XC_WRW; ++hgt) {
            const bool wide = hgt > XC_WR;
            const cuuint32_t box[5] = {(cuuint32_t)(wide ? XC_WCW : XC_WC), (cuuint32_t)(wide ? hgt - XC_WR : hgt), XC_CK, 1, 1};
            if (enc(wide ? &maps.wide[hgt - XC_WR - 1] : &maps.src[hgt - 1], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5,
                    const_cast<float*>(a.src), gdim, gstr, box, est,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return DPV_E_UNSUPP;
        }
    }
    {
        const cuuint64_t bs = a.B > 1 ? (cuuint64_t)a.ref_bs : chw;
        const cuuint64_t gdim[4] = {(cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.C, (cuuint64_t)a.B};
        const cuuint64_t gstr[3] = {(cuuint64_t)a.W * 4, (cuuint64_t)a.H * a.W * 4, bs * 4};
        const cuuint32_t box[4] = {TM_PX, 1, XC_CK, 1};
        const cuuint32_t est[4] = {1, 1, 1, 1};
        if (enc(&maps.ref, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a.ref), gdim, gstr, box, est,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return DPV_E_UNSUPP;
    }
    float* smaps = workspace;
    float* refn = workspace + (long long)a.B * a.V * XC_NMAP * (a.H + 2) * (a.W + 2);
    {   // pre-pass: source-only product maps and <r, r>
        const int mp = (a.H + 2) * (a.W + 2);
        dim3 grid((mp + 31) / 32, a.V + 1, a.B), block(XC_PRE_NT);
        sweep_smaps_kernel<<<grid, block, 0, st>>>(a, smaps, refn);
        DPV_LAUNCH_END();
    }
    const size_t smem = xc_smem_bytes(kper);
    const int tiles = ((a.W + TM_PX - 1) / TM_PX) * a.H;
    (void)tiles;
    dim3 grid((a.W + TM_PX - 1) / TM_PX, a.H, a.B * a.PS), block(XC_NT);
    // coordinates: reference operation order for wide images (see tm_coord), SFU form otherwise
    const bool exact = exact_env >= 0 ? (exact_env != 0) : (a.W > 192 || a.H > 192);
    cudaError_t e;
    static const int minb = [] { const char* v = getenv("DPV_XC_MINB"); return v ? atoi(v) : 5; }();
    // (PDL) the sweep kernel becomes resident behind the pre-pass and waits for it only where it first reads
    // the maps: the band and the correlation do not need them
#define DPV_XC_GO(EX_, MB_)                                                                                      \
    do {                                                                                                         \
        e = cudaFuncSetAttribute(sweep_xcorr_kernel<EX_, MB_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return (int)e;                                                                     \
        e = dpv_launch_pdl(sweep_xcorr_kernel<EX_, MB_>, grid, block, smem, st, a, (const float*)smaps,          \
                           (const float*)refn, maps);                                                      \
    } while (0)
    if (exact && minb == 5) DPV_XC_GO(true, 5);
    else if (exact) DPV_XC_GO(true, 6);
    else if (minb == 5) DPV_XC_GO(false, 5);
    else DPV_XC_GO(false, 6);
#undef DPV_XC_GO
    if (e != cudaSuccess) return (int)e;
    DPV_LAUNCH_END();
    return 0;
}

}  // namespace dpv
