// Whole-frame host pipeline: HOST buffers in, HOST buffers out.
//
// What a non-PyTorch caller of the reference's hot path (its ROS node, ros/ros_net.py:241-303,
// or an eval loop like trainer/default_trainer.py:189-260) would bind: one object that owns the
// device staging buffers, three streams and per-item events, and pushes a batch of frames
// through
//     cost volume (K1+K2a) -> log-softmax at 1/4 res (K3) -> full-res head (K3: log-DPV, E[d],
//     Var, arg-max, 1/4 hand-off) -> uncertainty field (K5)
// item by item, so that the H2D copy of item i+1, the kernels of item i and the D2H copy of
// item i-1 overlap -- within a call and, with dpv_pipeline_submit / dpv_pipeline_wait, across calls: the
// device staging exists twice (two slots), so batch n + 1 is copied in while batch n computes and batch
// n - 1 is read back.  dpv_pipeline_run = submit + wait for everything outstanding.  The refined log-DPV stays resident on the device (it is the next frame's
// feedback input, trainer/default_trainer.py:221-222); only the per-frame products travel back.
#include <algorithm>
#include <cstdlib>
#include <new>
#include <vector>

#include "dpv_common.cuh"

constexpr int kPipeSlots = 2;      // submissions in flight: batch n + 1 is copied in while batch n computes

struct PipeSlot {
    std::vector<cudaEvent_t> e_in, e_done;     // per item: inputs on the device / kernels done
    cudaEvent_t e_run_done, e_out_done;        // per submission: last kernel done / last result on the host
    // inputs
    float *feats, *poses, *K, *rays, *d, *logits, *intr;
    int *row_fwd, *row_inv, *col_fwd, *col_inv;
    // results
    float *cost, *bv, *refined, *depth, *var, *uf, *dz, *quarter, *ws, *sweep_ws;
    long long* argmax;
    // fused head + UF (dpv_head_ufield): device tables
    int *row_tab, *col_tab;
    bool used;
};

struct dpv_pipeline {
    int device, B, V, C, D, h, w, H, W;
    cudaStream_t s_in, s_run, s_out;
    PipeSlot slot[kPipeSlots];
    int64_t ws_floats_per_item, sweep_ws_floats_per_item;
    int64_t h2d, d2h;
    int64_t submitted, completed;              // submission counters: slot = n % kPipeSlots
    std::vector<int> h_row_tab, h_col_tab;     // host staging of the fused tables
};

namespace {

template <typename T>
cudaError_t dalloc(T** p, int64_t n) { return cudaMalloc((void**)p, (size_t)n * sizeof(T)); }

#define PIPE_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return (int)e__; } while (0)
#define PIPE_RC(expr) do { int r__ = (expr); if (r__ != 0) return r__; } while (0)

struct DeviceGuard {
    int prev;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); cudaSetDevice(dev); }
    ~DeviceGuard() { cudaSetDevice(prev); }
};

int slot_create(dpv_pipeline* p, PipeSlot* s) {
    const int B = p->B, V = p->V, C = p->C, D = p->D, H = p->H, W = p->W;
    const int64_t hw = (int64_t)p->h * p->w, HW = (int64_t)H * W;
    s->used = false;
    s->e_in.resize(B); s->e_done.resize(B);
    for (int i = 0; i < B; ++i) {
        PIPE_TRY(cudaEventCreateWithFlags(&s->e_in[i], cudaEventDisableTiming));
        PIPE_TRY(cudaEventCreateWithFlags(&s->e_done[i], cudaEventDisableTiming));
    }
    PIPE_TRY(cudaEventCreateWithFlags(&s->e_run_done, cudaEventDisableTiming));
    PIPE_TRY(cudaEventCreateWithFlags(&s->e_out_done, cudaEventDisableTiming));
    PIPE_TRY(dalloc(&s->feats, (int64_t)B * (V + 1) * C * hw));
    PIPE_TRY(dalloc(&s->poses, (int64_t)B * (V + 1) * 16));
    PIPE_TRY(dalloc(&s->K, (int64_t)B * 9));
    PIPE_TRY(dalloc(&s->rays, (int64_t)B * 3 * hw));
    PIPE_TRY(dalloc(&s->d, D));
    PIPE_TRY(dalloc(&s->logits, (int64_t)B * D * HW));
    PIPE_TRY(dalloc(&s->intr, (int64_t)B * 9));
    PIPE_TRY(dalloc(&s->row_fwd, H)); PIPE_TRY(dalloc(&s->row_inv, H));
    PIPE_TRY(dalloc(&s->col_fwd, W)); PIPE_TRY(dalloc(&s->col_inv, W));
    PIPE_TRY(dalloc(&s->cost, (int64_t)B * D * hw));
    PIPE_TRY(dalloc(&s->bv, (int64_t)B * D * hw));
    PIPE_TRY(dalloc(&s->refined, (int64_t)B * D * HW));
    PIPE_TRY(dalloc(&s->depth, (int64_t)B * HW));
    PIPE_TRY(dalloc(&s->var, (int64_t)B * HW));
    PIPE_TRY(dalloc(&s->argmax, (int64_t)B * HW));
    PIPE_TRY(dalloc(&s->uf, (int64_t)B * D * W));
    PIPE_TRY(dalloc(&s->dz, (int64_t)B * HW));
    PIPE_TRY(dalloc(&s->quarter, (int64_t)B * D * (H / 4) * (W / 4) + 1));
    PIPE_TRY(dalloc(&s->row_tab, (int64_t)H * 4));
    PIPE_TRY(dalloc(&s->col_tab, W));
    PIPE_TRY(dalloc(&s->ws, p->ws_floats_per_item * B));
    PIPE_TRY(dalloc(&s->sweep_ws, p->sweep_ws_floats_per_item * B));
    return 0;
}

void slot_destroy(PipeSlot* s) {
    void* bufs[] = {s->feats, s->poses, s->K, s->rays, s->d, s->logits, s->intr, s->row_fwd, s->row_inv,
                    s->col_fwd, s->col_inv, s->cost, s->bv, s->refined, s->depth, s->var, s->argmax, s->uf,
                    s->dz, s->quarter, s->ws, s->sweep_ws, s->row_tab, s->col_tab};
    for (void* b : bufs) if (b) cudaFree(b);
    for (auto e : s->e_in) cudaEventDestroy(e);
    for (auto e : s->e_done) cudaEventDestroy(e);
    cudaEventDestroy(s->e_run_done);
    cudaEventDestroy(s->e_out_done);
}

}  // namespace

extern "C" int dpv_pipeline_create(dpv_pipeline** out, int device, int B, int V, int C, int D,
                                   int h, int w, int H, int W) {
    if (!out || B <= 0 || V <= 0 || C <= 0 || D <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0)
        return DPV_E_BADARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return DPV_E_NODEVICE;
    DeviceGuard guard(device);
    dpv_pipeline* p = new (std::nothrow) dpv_pipeline();
    if (!p) return DPV_E_BADARG;
    p->device = device; p->B = B; p->V = V; p->C = C; p->D = D; p->h = h; p->w = w; p->H = H; p->W = W;
    p->h2d = p->d2h = 0;
    p->submitted = p->completed = 0;
    PIPE_TRY(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
    PIPE_TRY(cudaStreamCreateWithFlags(&p->s_run, cudaStreamNonBlocking));
    PIPE_TRY(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
    p->ws_floats_per_item = std::max(dpv_ufield_workspace_floats(1, D, H, W),
                                     (dpv_head_ufield_workspace_floats(1, D, H, W) + 3) & ~(int64_t)3);
    p->sweep_ws_floats_per_item = (dpv_sweep_workspace_floats(1, V, h, w) + 3) & ~(int64_t)3;
    p->h_row_tab.resize((size_t)H * 4);
    p->h_col_tab.resize(W);
    for (int i = 0; i < kPipeSlots; ++i) PIPE_RC(slot_create(p, &p->slot[i]));
    *out = p;
    return 0;
}

extern "C" int dpv_pipeline_destroy(dpv_pipeline* p) {
    if (!p) return DPV_E_BADARG;
    DeviceGuard guard(p->device);
    cudaStreamSynchronize(p->s_in); cudaStreamSynchronize(p->s_run); cudaStreamSynchronize(p->s_out);
    for (int i = 0; i < kPipeSlots; ++i) slot_destroy(&p->slot[i]);
    cudaStreamDestroy(p->s_in); cudaStreamDestroy(p->s_run); cudaStreamDestroy(p->s_out);
    delete p;
    return 0;
}

extern "C" int dpv_pipeline_wait(dpv_pipeline* p) {
    if (!p) return DPV_E_BADARG;
    if (p->completed >= p->submitted) return DPV_E_BADARG;      // nothing outstanding
    DeviceGuard guard(p->device);
    PipeSlot& s = p->slot[p->completed % kPipeSlots];
    PIPE_TRY(cudaEventSynchronize(s.e_out_done));
    PIPE_TRY(cudaEventSynchronize(s.e_run_done));
    ++p->completed;
    return 0;
}

extern "C" int dpv_pipeline_submit(dpv_pipeline* p, const float* feats, const float* poses,
                                   const float* K, const float* rays, const float* d_candi,
                                   const float* logits_full, const float* intr_up, const int* row_fwd,
                                   const int* row_inv, const int* col_fwd, const int* col_inv,
                                   float sigma, float* bv, float* depth, float* variance,
                                   int64_t* argmax, float* uf, float* depth_zero, float* quarter) {
    if (!p || !feats || !poses || !K || !rays || !d_candi || !logits_full || !intr_up || !row_fwd ||
        !row_inv || !col_fwd || !col_inv)
        return DPV_E_BADARG;
    while (p->submitted - p->completed >= kPipeSlots) PIPE_RC(dpv_pipeline_wait(p));   // every slot is in flight
    DeviceGuard guard(p->device);
    PipeSlot& s = p->slot[p->submitted % kPipeSlots];
    const int B = p->B, V = p->V, C = p->C, D = p->D, h = p->h, w = p->w, H = p->H, W = p->W;
    const int64_t hw = (int64_t)h * w, HW = (int64_t)H * W;
    const int64_t q4 = (int64_t)(H / 4) * (W / 4);
    int64_t in_bytes = 0, out_bytes = 0;
    auto up = [&](void* dst, const void* src, int64_t bytes) {
        in_bytes += bytes;
        return cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, p->s_in);
    };
    auto down = [&](void* dst, const void* src, int64_t bytes) {
        out_bytes += bytes;
        return cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, p->s_out);
    };
    // the slot's previous submission (two back) has been waited for by the host (see the loop above), so its
    // buffers are free; the stream-side waits below order the streams among themselves for the submission
    // that is still in flight in the OTHER slot (streams are FIFO: nothing to add)
    // per-call constants
    PIPE_TRY(up(s.poses, poses, (int64_t)B * (V + 1) * 16 * 4));
    PIPE_TRY(up(s.K, K, (int64_t)B * 9 * 4));
    PIPE_TRY(up(s.rays, rays, (int64_t)B * 3 * hw * 4));
    PIPE_TRY(up(s.d, d_candi, (int64_t)D * 4));
    PIPE_TRY(up(s.intr, intr_up, (int64_t)B * 9 * 4));
    PIPE_TRY(up(s.row_fwd, row_fwd, (int64_t)H * 4)); PIPE_TRY(up(s.row_inv, row_inv, (int64_t)H * 4));
    PIPE_TRY(up(s.col_fwd, col_fwd, (int64_t)W * 4)); PIPE_TRY(up(s.col_inv, col_inv, (int64_t)W * 4));
    // K3 + K5 in one pass when the shifts allow it (they do for the reference's row shifts)
    // (the tile form of the fused kernel has no per-item penalty; DPV_PIPELINE_FUSED_UF=0 switches it off)
    static const bool want_fused = [] { const char* e = getenv("DPV_PIPELINE_FUSED_UF"); return !e || atoi(e) != 0; }();
    const bool fused = want_fused && (W % 4 == 0) && dpv_head_ufield_workspace_floats(1, D, H, W) > 0 &&
                       dpv_uf_fused_tables(row_fwd, row_inv, col_fwd, col_inv, H, W,
                                           p->h_row_tab.data(), p->h_col_tab.data()) == 0;
    if (fused) {
        // pageable staging owned by the handle: the copy call returns once the bytes are staged
        PIPE_TRY(up(s.row_tab, p->h_row_tab.data(), (int64_t)H * 16));
        PIPE_TRY(up(s.col_tab, p->h_col_tab.data(), (int64_t)W * 4));
    }
    // sum of the bins = E[d] of a zero-padded log-DPV column (see dpv_ufield)
    float pad_depth = 0.f;
    for (int k = 0; k < D; ++k) pad_depth += d_candi[k];

    const int64_t feat_item = (int64_t)(V + 1) * C * hw;
    for (int i = 0; i < B; ++i) {
        PIPE_TRY(up(s.feats + i * feat_item, feats + i * feat_item, feat_item * 4));
        PIPE_TRY(up(s.logits + i * D * HW, logits_full + i * D * HW, (int64_t)D * HW * 4));
        PIPE_TRY(cudaEventRecord(s.e_in[i], p->s_in));
        PIPE_TRY(cudaStreamWaitEvent(p->s_run, s.e_in[i], 0));
        const float* fi = s.feats + i * feat_item;
        // reference view is the last one (models/models.py:531-535)
        PIPE_RC(dpv_sweep_cost_volume_ws(fi + (int64_t)V * C * hw, fi, s.poses + (int64_t)i * (V + 1) * 16,
                                         s.K + i * 9, s.rays + (int64_t)i * 3 * hw, s.d,
                                         s.cost + (int64_t)i * D * hw, nullptr, 1, V, C, D, h, w,
                                         0, 0, (int64_t)C * hw, 0, 0, 0, sigma, DPV_DIST_L2, 0,
                                         s.sweep_ws + i * p->sweep_ws_floats_per_item, p->s_run));
        PIPE_RC(dpv_head(s.cost + (int64_t)i * D * hw, nullptr, s.d, s.bv + (int64_t)i * D * hw,
                         nullptr, nullptr, nullptr, nullptr, nullptr, 1, D, h, w, DPV_IN_LOGITS,
                         p->s_run));
        if (fused) {
            PIPE_RC(dpv_head_ufield(s.logits + i * D * HW, s.d, s.refined + i * D * HW,
                                    s.depth + i * HW, s.var + i * HW, (int64_t*)(s.argmax + i * HW),
                                    s.quarter + i * D * q4, s.intr + i * 9, s.row_tab, s.col_tab,
                                    s.uf + (int64_t)i * D * W, s.dz + i * HW,
                                    s.ws + i * p->ws_floats_per_item, 1, D, H, W, 0, DPV_IN_LOGITS,
                                    0.6f, 0.6f + 0.3f, 100.f, 0.f, pad_depth, p->s_run));
        } else {
            PIPE_RC(dpv_head(s.logits + i * D * HW, nullptr, s.d, s.refined + i * D * HW, nullptr,
                             s.depth + i * HW, s.var + i * HW, (int64_t*)(s.argmax + i * HW),
                             s.quarter + i * D * q4, 1, D, H, W, DPV_IN_LOGITS, p->s_run));
            PIPE_RC(dpv_ufield(s.refined + i * D * HW, s.depth + i * HW, s.d, s.intr + i * 9, nullptr,
                               s.row_fwd, s.row_inv, s.col_fwd, s.col_inv, s.uf + (int64_t)i * D * W,
                               s.dz + i * HW, s.ws + i * p->ws_floats_per_item, 1, D, H, W, 0,
                               DPV_IN_LOGPROB, 0.6f, 0.6f + 0.3f, 100.f, 0.f, pad_depth, 0.f, p->s_run));
        }
        PIPE_TRY(cudaEventRecord(s.e_done[i], p->s_run));
        PIPE_TRY(cudaStreamWaitEvent(p->s_out, s.e_done[i], 0));
        if (bv) PIPE_TRY(down(bv + (int64_t)i * D * hw, s.bv + (int64_t)i * D * hw, (int64_t)D * hw * 4));
        if (depth) PIPE_TRY(down(depth + i * HW, s.depth + i * HW, HW * 4));
        if (variance) PIPE_TRY(down(variance + i * HW, s.var + i * HW, HW * 4));
        if (argmax) PIPE_TRY(down(argmax + i * HW, s.argmax + i * HW, HW * 8));
        if (uf) PIPE_TRY(down(uf + (int64_t)i * D * W, s.uf + (int64_t)i * D * W, (int64_t)D * W * 4));
        if (depth_zero) PIPE_TRY(down(depth_zero + i * HW, s.dz + i * HW, HW * 4));
        if (quarter) PIPE_TRY(down(quarter + i * D * q4, s.quarter + i * D * q4, (int64_t)D * q4 * 4));
    }
    PIPE_TRY(cudaEventRecord(s.e_run_done, p->s_run));
    PIPE_TRY(cudaEventRecord(s.e_out_done, p->s_out));
    s.used = true;
    p->h2d = in_bytes; p->d2h = out_bytes;
    ++p->submitted;
    return 0;
}

extern "C" int dpv_pipeline_run(dpv_pipeline* p, const float* feats, const float* poses,
                                const float* K, const float* rays, const float* d_candi,
                                const float* logits_full, const float* intr_up, const int* row_fwd,
                                const int* row_inv, const int* col_fwd, const int* col_inv,
                                float sigma, float* bv, float* depth, float* variance,
                                int64_t* argmax, float* uf, float* depth_zero, float* quarter) {
    const int rc = dpv_pipeline_submit(p, feats, poses, K, rays, d_candi, logits_full, intr_up, row_fwd, row_inv,
                                       col_fwd, col_inv, sigma, bv, depth, variance, argmax, uf, depth_zero, quarter);
    if (rc != 0) return rc;
    while (p->completed < p->submitted) {      // this submission and anything submitted before it
        const int rw = dpv_pipeline_wait(p);
        if (rw != 0) return rw;
    }
    return 0;
}

extern "C" int dpv_pipeline_last_bytes(const dpv_pipeline* p, int64_t* h2d, int64_t* d2h) {
    if (!p || !h2d || !d2h) return DPV_E_BADARG;
    *h2d = p->h2d; *d2h = p->d2h;
    return 0;
}
