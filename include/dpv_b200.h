/*
 * dpv_b200.h -- C ABI of libdpv_sm100a.so, the B200 (sm_100a) implementation of the
 * depth-probability-volume (DPV) hot path of soulslicer/probabilistic-depth.
 *
 * The reference has no FFI for this path (it is Python calling PyTorch ATen ops); the one
 * native binding it does have is correlation_cuda.forward(...)
 * (models/correlation_package/correlation_cuda.cc:10-16,169-172).  The entry points below
 * are what a cffi / ctypes / pybind binding for the path would bind, one per reference
 * function; each comment names the reference site it replaces (paths relative to the
 * reference tree).  INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions
 *  - All buffers are DEVICE pointers to contiguous fp32 (row-major, NCHW-style) unless a
 *    parameter says otherwise; the caller owns every allocation, the library never
 *    allocates on the data path (dpv_pipeline_* owns its staging buffers) and never syncs.
 *  - `stream` is a cudaStream_t passed as void*; work is enqueued on it and the call returns.
 *  - Return value: 0 on success, a positive cudaError_t on a CUDA failure, a negative
 *    DPV_E_* code on a bad argument.  dpv_error_string() explains either.
 *  - Inputs are never modified.  No global mutable state besides one atomic launch counter
 *    (dpv_launch_count); safe to call from one host thread per device/stream.
 *  - There is no CPU implementation behind any of these.
 */
#ifndef DPV_B200_H_
#define DPV_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPV_E_BADARG   (-1)   /* null pointer, non-positive dimension, ... */
#define DPV_E_UNSUPP   (-2)   /* dimension outside what the kernels were built for */
#define DPV_E_NODEVICE (-3)   /* no sm_100 device / image not loadable */

#define DPV_DIST_L2 0         /* warping/homography.py:80-82 */
#define DPV_DIST_L1 1         /* warping/homography.py:84-86 */

/* Input interpretation of the depth-bin axis for dpv_head / dpv_ufield. */
#define DPV_IN_LOGITS 0       /* un-normalised scores: log-softmax is applied        */
#define DPV_IN_LOGPROB 1      /* already log-probabilities (BV_log=True)             */
#define DPV_IN_PROB 2         /* linear probabilities (BV_log=False)                 */

int dpv_abi_version(void);
const char* dpv_error_string(int code);
/* Number of kernel launches this library has enqueued since load (for gpu_launches). */
long long dpv_launch_count(void);

/* ---- K1 + K2a : plane-sweep cost volume ------------------------------------------------
 * Replaces est_swp_volume_v4 (warping/homography.py:98-135) including
 * _back_warp_homo_parallel (:170-198), the D-fold .repeat (:123), F.grid_sample (:197) and
 * img_dis_L2_pard / img_dis_L1_pard (:80-86), batched over B items and V source views in
 * one launch (the reference loops over items in Python, models/models.py:528-550).
 *
 *   cost[b,k,y,x] = sum_v ( sum_c dist( bilinear(src[b,v,c], P(b,v,k,y,x)) - ref[b,c,y,x] ) ) / sigma
 *   P = K t + (K R ray) d_k ;  u,v = P_xy / (P_z + 1e-10) ;  grid = ((u-cx)/cx, (v-cy)/cy) ;
 *   bilinear, zeros padding, align_corners=False.  Non-finite coordinates sample 0.
 *
 * ref   : item b at ref + b*ref_bstride,                      [C,H,W]
 * src   : view v of item b at src + b*src_bstride + v*src_vstride, [C,H,W]
 * pose  : 4x4 row-major [R|t] of view v of item b at pose + b*pose_bstride + v*16
 * K     : 3x3 row-major intrinsics of item b at K + b*k_bstride   (k_bstride may be 0)
 * rays  : [3, H*W] unit rays of item b at rays + b*rays_bstride   (rays_bstride may be 0)
 * d_candi : [D] fp32 depth of each plane
 * cost  : [B, D, H, W] out
 * log_softmax_out : optional [B, D, H, W]; when non-null additionally receives
 *         log_softmax(cost, dim=1) (models/packnet.py:394 places them back to back).
 * algo  : 0 = choose, 1 = direct per-plane gather (L1 or L2), 2 = per-cell Gram form with global
 *         gathers (L2), 3 = per-cell Gram form staged through shared memory with cp.async (L2),
 *         4 = per-cell Gram form, TMA-fed, four lanes per pixel (L2, production path; needs W % 4 == 0
 *         and 16-byte aligned ref/src bases and strides -- "choose" falls back to 3/2 otherwise),
 *         5 = cross-correlation form (L2; dpv_sweep_cost_volume_ws only): the products between source
 *         pixels are taken once per call by a pre-pass into `workspace`, the tile kernel computes one
 *         banded correlation <ref pixel, source pixel> per tap; same shape conditions as 4.  "choose"
 *         takes 5 whenever a workspace is given.
 * dpv_sweep_cost_volume_ws: the same call with a caller-owned scratch buffer of
 *         dpv_sweep_workspace_floats(B, V, H, W) floats (16-byte aligned, contents irrelevant before and
 *         after the call, not shared by calls that run concurrently); dpv_sweep_cost_volume == workspace NULL.
 */
int dpv_sweep_cost_volume(const float* ref, const float* src, const float* pose, const float* K,
                          const float* rays, const float* d_candi, float* cost,
                          float* log_softmax_out,
                          int B, int V, int C, int D, int H, int W,
                          int64_t ref_bstride, int64_t src_bstride, int64_t src_vstride,
                          int64_t pose_bstride, int64_t k_bstride, int64_t rays_bstride,
                          float sigma, int dist, int algo, void* stream);
int64_t dpv_sweep_workspace_floats(int B, int V, int H, int W);
int dpv_sweep_cost_volume_ws(const float* ref, const float* src, const float* pose, const float* K,
                             const float* rays, const float* d_candi, float* cost,
                             float* log_softmax_out,
                             int B, int V, int C, int D, int H, int W,
                             int64_t ref_bstride, int64_t src_bstride, int64_t src_vstride,
                             int64_t pose_bstride, int64_t k_bstride, int64_t rays_bstride,
                             float sigma, int dist, int algo, float* workspace, void* stream);

/* ---- K1 stand-alone : per-plane warp ----------------------------------------------------
 * Replaces _back_warp_homo_parallel (warping/homography.py:170-198) for callers that want
 * the warped stack itself.  img [N or 1, C, H, W] (img_nstride = 0 broadcasts one image over
 * the planes instead of the reference's .repeat), term1 [3], term2 [3, H*W], d [N];
 * out [N, C, H, W].  cx, cy as the reference reads them from intrinsic_M.
 */
int dpv_warp_planes(const float* img, const float* d, const float* term1, const float* term2,
                    float* out, int N, int C, int H, int W, int64_t img_nstride,
                    float cx, float cy, void* stream);

/* ---- K4a : diagonal feature warp --------------------------------------------------------
 * Replaces warp_feature (warping/homography.py:137-168): out[b,v,k] = channel k of view v
 * warped onto plane k.  feat [B,V,D,H,W] -> out [B,V,D,H,W]; pose/K/rays as above.
 */
int dpv_warp_feature(const float* feat, const float* pose, const float* K, const float* rays,
                     const float* d_candi, float* out, int B, int V, int D, int H, int W,
                     int64_t pose_bstride, int64_t k_bstride, int64_t rays_bstride, void* stream);

/* ---- K3 (+K4b) : depth-bin head ---------------------------------------------------------
 * One pass over x [B,D,H,W] producing any subset of (null pointer = skip):
 *   logp     [B,D,H,W]  log_softmax(x (+ addend), dim=1)     models/models.py:351,560,637; :694 with addend
 *   prob     [B,D,H,W]  exp(logp)                             models/models.py:653,675,697
 *   depth    [B,H,W]    sum_k d_k exp(logp_k), fp32           utils/img_utils.py:52-61
 *   variance [B,H,W]    sum_k (d_k-mean)^2 p_k, fp32          trainer/default_trainer.py:333-336 (f64 there)
 *   argmax   [B,H,W]    int64 first maximal bin of logp       (torch.argmax; not in the reference)
 *   quarter  [B,D,H/4,W/4] logp at rows/cols 0,4,8,...        trainer/default_trainer.py:221-222
 * in_mode: DPV_IN_LOGITS / DPV_IN_LOGPROB / DPV_IN_PROB (the last two skip the soft-max;
 * logp then echoes log-probabilities).  addend may be null and must be null unless in_mode is
 * DPV_IN_LOGITS (DPV_E_BADARG otherwise).  d_candi [D] fp32.
 */
int dpv_head(const float* x, const float* addend, const float* d_candi,
             float* logp, float* prob, float* depth, float* variance, int64_t* argmax,
             float* quarter, int B, int D, int H, int W, int in_mode, void* stream);

/* ---- K4c : LiDAR prior and Bayesian fusion ----------------------------------------------
 * dpv_lidar_prior replaces gen_dpv_withmask (utils/img_utils.py:360-375, :31-50):
 *   dmaps [B,H,W], masks [B,H,W] -> prior [B,D,H,W] clamped to [2.2e-16, 1].
 * two_sigma_sq is 2*pow(sqrt(var),2) evaluated in fp32 by the caller as the reference does.
 * dpv_bayes_fuse replaces models/models.py:669-672: fused = clamp(normalise(exp(bv + log prior)));
 *   the prior is either given (prior != null) or generated on the fly from dmaps/masks.
 *   Outputs fused [B,D,H,W] and/or log_fused [B,D,H,W].
 */
int dpv_lidar_prior(const float* dmaps, const float* masks, const float* d_candi, float* prior,
                    int B, int D, int H, int W, float two_sigma_sq, void* stream);
int dpv_bayes_fuse(const float* bv, const float* prior, const float* dmaps, const float* masks,
                   const float* d_candi, float* fused, float* log_fused,
                   int B, int D, int H, int W, float two_sigma_sq, void* stream);

/* ---- K5 : uncertainty-field collapse ----------------------------------------------------
 * Replaces gen_ufield (utils/img_utils.py:268-358), batched.  dpv [B,D,H,W] in `in_mode`
 * (DPV_IN_LOGPROB or DPV_IN_PROB); depth [B,H,W] = E[d] of that DPV (dpv_head's `depth`
 * output); intr_up 3x3 per item (bstride may be 0); mask [B,H,W] optional.
 * pad_depth: E[d] of an all-zero (padding) DPV column as the reference would compute it:
 * sum_k d_k in log mode (exp(0) = 1 per bin), 0 in linear mode.
 * row_fwd/row_inv [H], col_fwd/col_inv [W]: int32 source index of the +pshift / -pshift
 * nearest-neighbour shifts (-1 = samples zero padding), built once per shape by the host
 * wrapper from the reference's own grid construction.
 * quash_range > 0 selects the quash_limit branch (utils/img_utils.py:325-332, taken for ILIM data and
 * for every `cfgx` caller, ros/ros_net.py:279): a shifted-frame pixel keeps its weight only when its
 * masked depth (zeros -> 1000) lies strictly within +/- quash_range (the reference uses 1.0) of the
 * minimum of its column; 0 = the KITTI branch.
 * Outputs uf [B,D,W] (0/0 = NaN as in the reference) and depth_zero [B,H,W].
 * workspace: dpv_ufield_workspace_floats(B, D, H, W) floats.
 */
int64_t dpv_ufield_workspace_floats(int B, int D, int H, int W);
int dpv_ufield(const float* dpv, const float* depth, const float* d_candi, const float* intr_up,
               const float* mask, const int* row_fwd, const int* row_inv, const int* col_fwd,
               const int* col_inv, float* uf, float* depth_zero, float* workspace,
               int B, int D, int H, int W, int64_t intr_bstride, int in_mode,
               float zstart, float zend, float maxd, float mind, float pad_depth,
               float quash_range, void* stream);

/* ---- K3 + K5 fused : depth-bin head with the uncertainty field accumulated in the same pass ----
 * dpv_head (log-softmax, E[d], Var, arg-max, 1/4 hand-off; utils/img_utils.py:52-61,
 * models/models.py:351, trainer/default_trainer.py:221-222,333-336) and gen_ufield
 * (utils/img_utils.py:268-358) in one read of x [B,D,H,W] (in_mode DPV_IN_LOGITS or DPV_IN_LOGPROB):
 * the trainer calls them back to back on the refined DPV (trainer/default_trainer.py:232-244).
 * No ground-truth mask (that variant is dpv_ufield).  D in {16,32,64}, W % 4 == 0, else
 * DPV_E_UNSUPP (D = 32 / 64 run the tile kernel of dpv_head_uftile.cu, D = 16 the persistent
 * TMA-fed kernel of dpv_head_stream.cu).  row_tab [H][4] / col_tab [W] int32 DEVICE tables: build them on the host with
 * dpv_uf_fused_tables from the same four index maps dpv_ufield takes (HOST pointers there), which
 * returns DPV_E_UNSUPP when the shifts do not compose to "same pixel or padding".
 * Any of logp / depth / variance / argmax / quarter / depth_zero may be null.
 * workspace: dpv_head_ufield_workspace_floats(B, D, H, W) floats, 16-byte aligned.
 */
int64_t dpv_head_ufield_workspace_floats(int B, int D, int H, int W);
int dpv_uf_fused_tables(const int* row_fwd, const int* row_inv, const int* col_fwd,
                        const int* col_inv, int H, int W, int* row_tab, int* col_tab);
int dpv_head_ufield(const float* x, const float* d_candi, float* logp, float* depth,
                    float* variance, int64_t* argmax, float* quarter, const float* intr_up,
                    const int* row_tab, const int* col_tab, float* uf, float* depth_zero,
                    float* workspace, int B, int D, int H, int W, int64_t intr_bstride, int in_mode,
                    float zstart, float zend, float maxd, float mind, float pad_depth, void* stream);

/* ---- K2b : local correlation ------------------------------------------------------------
 * Replaces correlation_cuda.forward (models/correlation_package/correlation_cuda.cc:10-87,
 * kernel correlation_cuda_kernel.cu:41-114) and models/correlation_native.py:13-23 for
 * kernel_size=1, stride1=stride2=1, pad_size=max_displacement:
 *   out[b,(i*(2r+1)+j),y,x] = (1/C) sum_c x1[b,c,y,x] * x2[b,c,y+i-r,x+j-r]   (zero padded)
 * x1, x2 [B,C,H,W] -> out [B,(2r+1)^2,H,W].
 */
int dpv_correlation(const float* x1, const float* x2, float* out, int B, int C, int H, int W,
                    int max_displacement, void* stream);

/* ---- depth-plane sharding (large D) -----------------------------------------------------
 * Soft-max over planes that live on several GPUs (not in the reference).  Rank g
 * owns planes [plane_offset, plane_offset + D) of x [B, D, HW].  Sequence per rank, with the host issuing
 * ONE collective (NCCL all-gather over NVLink) on the same stream between the two calls:
 *   dpv_shard_stats        -> stats [5][B*HW]: local max m, S0 = sum exp(x - m), local mean mu = sum d exp / S0,
 *                             local central moment M2 = sum (d - mu)^2 exp(x - m), index of the first local max
 *   all-gather             -> gathered [G][5][B*HW], ranks in plane order
 *   dpv_shard_merge_finish -> merges the G records per pixel (online-soft-max weights, pairwise variance rule,
 *                             first-maximum-wins arg-max) and writes logp = x - M - log S for the local planes;
 *                             depth / variance / argmax [B*HW] come out replicated on every rank.  Any of
 *                             logp / depth / variance / argmax may be null.
 * Reduce-scatter form (pays from G > 2: both collectives move 20 B x pixels per rank, independent of G):
 *   dpv_shard_stats with slice = ceil(B*HW / G): records slice-major, [G][5][slice]
 *   all-to-all             -> recv [G][5][slice]: rank g's statistics of MY slice of the pixels
 *   dpv_shard_merge_slice  -> merged [5][slice]: M, log S, mean, variance, arg-max of my pixels
 *   all-gather             -> all [G][5][slice]: the merged record of every pixel
 *   dpv_shard_finish       -> logp of the local planes, depth / variance / argmax replicated
 */
int dpv_shard_stats(const float* x, const float* d_local, float* stats, int B, int D, int HW,
                    int plane_offset, int slice, void* stream);
int dpv_shard_merge_slice(const float* recv, float* merged, int G, int slice, void* stream);
int dpv_shard_finish(const float* x, const float* all, float* logp, float* depth, float* variance,
                     int64_t* argmax, int B, int D, int HW, int slice, void* stream);
int dpv_shard_merge_finish(const float* x, const float* gathered, float* logp, float* depth,
                           float* variance, int64_t* argmax, int G, int B, int D, int HW, void* stream);

/* ---- the D -> D 3x3 convolutions before the 1/4-res soft-max -----------------------
 * conv0 / conv0_1 (Conv2d(D, D, 3, 1, 1, bias) + LeakyReLU) and conv0_2 (Conv2d) followed by
 * F.log_softmax(dim=1) (models/models.py:456-460, applied at :555-560 and :632-637; LeakyReLU models/models.py:38-46),
 * D = 64.  Implicit GEMM on the tcgen05 tensor cores with the accumulator in TMEM, fp32 parity through split
 * precision (TF32 x 3: hi*hi + lo*hi + hi*lo).  Activations travel between the layers in a PACKED form
 * [B][H+2][W+2][64] (channels innermost, zero border = the convolution's padding), twice: hi (the TF32 value
 * nearest to it) and lo (the TF32 value nearest to value - hi); dpv_conv3x3_packed_floats gives the size of one of the two.
 *   dpv_conv3x3_pack          NCHW [B,64,H,W] -> packed hi / lo
 *   dpv_conv3x3_pack_weights  torch weight [64,64,3,3] -> packed hi / lo [9 taps][64 out][64 in] (once per model)
 *   dpv_conv3x3_d64           one convolution: packed in -> bias + epilogue (0 = none, 1 = LeakyReLU(slope),
 *                             2 = log_softmax over the 64 channels) -> packed hi / lo for the next layer (out_hi /
 *                             out_lo, both or neither) and / or NCHW [B,64,H,W] (out_nchw).
 * C != 64 returns DPV_E_UNSUPP.  All buffers 16-byte aligned.
 */
int64_t dpv_conv3x3_packed_floats(int B, int H, int W);
int dpv_conv3x3_pack(const float* x, float* packed_hi, float* packed_lo, int B, int C, int H, int W, void* stream);
int dpv_conv3x3_pack_weights(const float* weight, float* w_hi, float* w_lo, int C_out, int C_in, void* stream);
int dpv_conv3x3_d64(const float* in_hi, const float* in_lo, const float* w_hi, const float* w_lo,
                    const float* bias, float* out_hi, float* out_lo, float* out_nchw, int B, int H, int W,
                    int epilogue, float slope, void* stream);

/* ---- SURVEY.md 8f rank 2, second half: Base3D, the 3-D convolutions of the feedback mode ---------------
 * models/models.py:376-438: Conv3d(., 32, 3, 1, 1, bias=False) + BatchNorm3d (models/models.py:31-36) [+ ReLU]
 * stacks over the volume [B, C, D, h, w] (applied at models/models.py:693).  Same tcgen05 / TMEM / TF32 x 3
 * machinery as dpv_conv3x3_d64, for 32 channels: activations packed as [B][D+2][H+2][W+2][32] hi / lo
 * (dpv_conv3d_packed_floats floats each), channels innermost, zero border.
 *   dpv_conv3d_pack          NCDHW [B,C<=32,D,H,W] -> packed hi / lo (channels >= C zero)
 *   dpv_conv3d_pack_weights  torch weight [C_out<=32][C_in<=32][3][3][3], optionally times scale[C_out] (a BatchNorm in
 *                            eval with running statistics folded into the filter) -> packed hi / lo [27][32][32]
 *   dpv_conv3d_c32           one convolution.  shift [32] (nullable = 0; the folded BatchNorm's shift) is added, then the
 *                            packed residual (res_hi / res_lo, nullable), then ReLU (relu != 0).  Outputs, any
 *                            combination: packed hi / lo for the next layer; out_c0 [B,D,H,W] = channel 0 (the
 *                            32 -> 1 classifier, models/models.py:403); out_raw [positions][32] fp32 together with
 *                            stats (double[64], caller-zeroed: per channel sum and sum of squares over the real
 *                            voxels, +=) for a BatchNorm that uses BATCH statistics.  c_in = real input channels
 *                            (K steps beyond them are skipped).  relu: bit 0 = ReLU; bit 2 = the input and the filter
 *                            are z-folded (below; c_in = 3 x the layer's input channels).
 *   dpv_conv3d_pack_zfold, dpv_conv3d_pack_weights_zfold
 *                            a first layer with 3 C_in <= 32 input channels: the packed input holds, per position, the
 *                            C_in channels of the three z-planes z-1, z, z+1 (channel dzi * C_in + c), the filter is
 *                            packed to match ([dy*3+dx][out][dzi * C_in + c]); the convolution over dz becomes part of
 *                            the channel contraction (3 K-blocks instead of 9: a third of the copies).
 *   dpv_conv3d_bn_apply      BatchNorm with batch statistics (F.batch_norm(training=True): the unregistered
 *                            dres_modules of Base3D never leave training mode, models/models.py:394-399): raw, stats
 *                            -> (x - mean) / sqrt(biased var + eps) * gamma + beta, + residual, ReLU -> packed hi / lo.
 *                            Running statistics are not updated.
 *   dpv_conv3d_c32_to1       the classifier Conv3d(c_in <= 32, 1, 3, 1, 1, bias=False) (models/models.py:403) on the FP32
 *                            pipe: packed in -> [B,D,H,W].  weight_host: the torch weight [1][c_in][3][3][3] in HOST
 *                            memory (864 floats at most; they travel as launch parameters).
 */
int64_t dpv_conv3d_packed_floats(int B, int D, int H, int W);
int dpv_conv3d_pack(const float* x, float* packed_hi, float* packed_lo, int B, int C, int D, int H, int W, void* stream);
int dpv_conv3d_pack_weights(const float* weight, const float* scale, float* w_hi, float* w_lo, int C_out, int C_in,
                            void* stream);
int dpv_conv3d_c32(const float* in_hi, const float* in_lo, const float* w_hi, const float* w_lo, const float* shift,
                   const float* res_hi, const float* res_lo, float* out_hi, float* out_lo, float* out_c0,
                   float* out_raw, double* stats, int B, int D, int H, int W, int relu, int c_in, void* stream);
int dpv_conv3d_pack_zfold(const float* x, float* packed_hi, float* packed_lo, int B, int C, int D, int H, int W,
                          void* stream);
int dpv_conv3d_pack_weights_zfold(const float* weight, const float* scale, float* w_hi, float* w_lo, int C_out, int C_in,
                                  void* stream);
int dpv_conv3d_c32_to1(const float* in_hi, const float* in_lo, const float* weight_host, float* out, int B, int D, int H,
                       int W, int c_in, void* stream);
int dpv_conv3d_bn_apply(const float* raw, const double* stats, const float* gamma, const float* beta, float eps,
                        const float* res_hi, const float* res_lo, float* out_hi, float* out_lo, int B, int D, int H,
                        int W, int relu, void* stream);

/* ---- eval metrics on the device ------------------------------------
 * dpv_depth_errors replaces img_utils.depth_error (utils/img_utils.py:17-22) around depthError
 * (external/deval_lib/src/evaluate_depth.h:19-119), batched: first/second [B,H,W] in the python
 * wrapper's argument order (prediction first; pixels with first >= 0 count).  zero_invalid != 0 maps
 * zeros to -1 as the wrapper does.  Optional fused preparation of trainer/default_trainer.py:247-254:
 * mask [B,H,W] multiplies `first`; clamp_max > 0 clamps `second` from above.  out [B,9] in the order
 * mae, rmse, inverse mae, inverse rmse, log mae, log rmse, scale invariant log, abs relative,
 * squared relative; counts [B] (optional) = valid pixels (0 => the reference throws; out is NaN).
 * workspace: dpv_depth_errors_workspace_doubles(B, H, W) doubles.
 * dpv_unc_rmse replaces compute_unc_rmse (utils/img_utils.py:183-194): uf_* [B,D,W] linear
 * uncertainty fields -> out [B] mean |E_truth - E_pred| over the columns where both are finite.
 */
int64_t dpv_depth_errors_workspace_doubles(int B, int H, int W);
int dpv_depth_errors(const float* first, const float* second, const float* mask, float clamp_max,
                     int zero_invalid, float* out, int* counts, double* workspace,
                     int B, int H, int W, void* stream);
int dpv_unc_rmse(const float* uf_truth, const float* uf_pred, const float* d_candi, float* out,
                 int B, int D, int W, void* stream);

/* ---- LiDAR -> sparse depth maps -----------------------------------
 * dpv_lidar_depthmap replaces generate_depth (external/utils_lib/python/utils_lib.cpp:86-160, the
 * upsample = 0 branch that kittiloader/kitti.py:693-697 uses for evaluation) followed by
 * minpool(., pool_scale, pool_default) (utils/img_utils.py:87-95 via kittiloader/kitti.py:706):
 * velo [n,4] (x, y, z, 1; 16-byte aligned), intr [3,4], m_velo2cam [4,4] row-major ->
 *   dmap       [height, width]                         filtered z-buffer (0 = no return)       optional
 *   dmap_small [height/pool_scale, width/pool_scale]   minimum of the non-zero depths per block optional
 *   mask_small same shape: 1 where dmap_small >= 0.01 (kittiloader/kitti.py:713-714 builds the
 *              complement), the `masks` input of dpv_bayes_fuse                               optional
 * zbuf: width*height uint32 scratch.  filtering = neighbourhood half-width, filterdiff = tolerated
 * depth step (utils_lib.cpp:89-92).
 * dpv_minpool is the stand-alone minpool over N planes (default_value 0 = plain minimum).
 */
int dpv_lidar_depthmap(const float* velo, int n, const float* intr, const float* m_velo2cam,
                       int width, int height, int filtering, float filterdiff, int pool_scale,
                       float pool_default, unsigned* zbuf, float* dmap, float* dmap_small,
                       float* mask_small, void* stream);
int dpv_minpool(const float* in, float* out, int N, int H, int W, int scale, float default_value,
                void* stream);

/* ---- whole-frame host pipeline (end-to-end with HOST buffers) ---------------------------
 * What a non-PyTorch host (the reference's ROS node, ros/ros_net.py:241-303) would call: one
 * object owning device staging buffers, streams and events; run() takes pinned or pageable
 * HOST pointers, overlaps H2D copies, kernels and D2H copies item by item and returns when
 * the results are in host memory.  Stages: cost volume -> log-softmax (1/4 res) -> full-res
 * head -> uncertainty field (the `default` / stereo frame of SURVEY.md section 8).
 */
typedef struct dpv_pipeline dpv_pipeline;
int dpv_pipeline_create(dpv_pipeline** out, int device, int B, int V, int C, int D,
                        int h, int w, int H, int W);
int dpv_pipeline_destroy(dpv_pipeline* p);
/* Host inputs: feats [B,V+1,C,h,w] (reference view last, models/models.py:531-535),
 * poses [B,V+1,4,4], K [B,3,3], rays [B,3,h*w], d_candi [D], logits_full [B,D,H,W],
 * intr_up [B,3,3], shift LUTs as for dpv_ufield.
 * Host outputs (null = skip): bv [B,D,h,w], depth [B,H,W], variance [B,H,W],
 * argmax [B,H,W] int64, uf [B,D,W], depth_zero [B,H,W], quarter [B,D,H/4,W/4].
 * Device-resident result kept for the next frame: the refined log-DPV. */
int dpv_pipeline_run(dpv_pipeline* p, const float* feats, const float* poses, const float* K,
                     const float* rays, const float* d_candi, const float* logits_full,
                     const float* intr_up, const int* row_fwd, const int* row_inv,
                     const int* col_fwd, const int* col_inv, float sigma,
                     float* bv, float* depth, float* variance, int64_t* argmax, float* uf,
                     float* depth_zero, float* quarter);
/* The same call split in two, so that consecutive batches overlap: submit() enqueues the copies and the
 * kernels of one batch and returns; wait() blocks until the OLDEST outstanding submission has its results in
 * host memory.  The handle stages two batches on the device: batch n + 1 is copied in while batch n computes
 * and batch n - 1 is read back; a third submit() first waits for the oldest.  The host buffers of a submission
 * (inputs and outputs) must stay untouched until its wait() has returned; pinned memory makes the copies
 * asynchronous.  run() = submit() + wait() for everything outstanding.  wait() with nothing outstanding
 * returns DPV_E_BADARG. */
int dpv_pipeline_submit(dpv_pipeline* p, const float* feats, const float* poses, const float* K,
                        const float* rays, const float* d_candi, const float* logits_full,
                        const float* intr_up, const int* row_fwd, const int* row_inv,
                        const int* col_fwd, const int* col_inv, float sigma,
                        float* bv, float* depth, float* variance, int64_t* argmax, float* uf,
                        float* depth_zero, float* quarter);
int dpv_pipeline_wait(dpv_pipeline* p);
/* Bytes copied host->device and device->host by the last submission. */
int dpv_pipeline_last_bytes(const dpv_pipeline* p, int64_t* h2d, int64_t* d2h);

#ifdef __cplusplus
}
#endif
#endif /* DPV_B200_H_ */
