// extern "C" surface of libunetb200.so (declared in include/unetb200.h).
#include <string.h>

#include "../../include/unetb200.h"
#include "train.cuh"

using namespace ub;

struct unetb200_ctx {
    Ctx* c;
};

static thread_local std::string g_create_error;
static const NetSpec& spec() {
    static NetSpec s;
    return s;
}

extern "C" {

int unetb200_create(unetb200_ctx** out, int device, int max_batch, int H, int W) {
    if (!out) return 1;
    *out = nullptr;
    Ctx* c = nullptr;
    std::string err;
    if (ctx_create(&c, device, max_batch, H, W, &err)) {
        g_create_error = err;
        return 1;
    }
    *out = new unetb200_ctx{c};
    return 0;
}

void unetb200_destroy(unetb200_ctx* ctx) {
    if (!ctx) return;
    delete ctx->c;
    delete ctx;
}

const char* unetb200_last_error(unetb200_ctx* ctx) {
    return ctx ? ctx->c->last_error.c_str() : g_create_error.c_str();
}

int unetb200_check_device_error(unetb200_ctx* h, int* flag_out) {
    Ctx* ctx = h->c;
    UB_CUDA(cudaDeviceSynchronize());
    int f = 0;
    UB_CUDA(cudaMemcpy(&f, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag_out) *flag_out = f;
    return 0;
}

int unetb200_num_tensors(void) { return (int)spec().tensors.size(); }
int unetb200_tensor_info(int index, char* name_out, int name_cap, int* ndim_out, int shape_out[4],
                         long long* offset_out, int* kind_out) {
    const NetSpec& S = spec();
    if (index < 0 || index >= (int)S.tensors.size()) return 1;
    const TensorInfo& t = S.tensors[index];
    if (name_out && name_cap > 0) {
        strncpy(name_out, t.name.c_str(), name_cap - 1);
        name_out[name_cap - 1] = 0;
    }
    if (ndim_out) *ndim_out = t.ndim;
    if (shape_out)
        for (int i = 0; i < 4; ++i) shape_out[i] = t.shape[i];
    if (offset_out) *offset_out = t.offset;
    if (kind_out) *kind_out = t.kind;
    return 0;
}
long long unetb200_num_params(void) { return spec().n_params; }
long long unetb200_num_buffers(void) { return spec().n_buffers; }
int unetb200_num_counters(void) { return spec().n_counters; }

int unetb200_load_weights(unetb200_ctx* h, const float* params_dev, const float* buffers_dev, void* stream) {
    return unetb200_load_weights_ex(h, params_dev, buffers_dev, 0, stream);
}

int unetb200_load_weights_ex(unetb200_ctx* h, const float* params_dev, const float* buffers_dev, int skip_bn_fold,
                             void* stream) {
    Ctx* ctx = h->c;
    if (!params_dev || !buffers_dev) return ctx_fail(ctx, "load_weights: null pointer");
    UB_CUDA(cudaSetDevice(ctx->device));
    return ctx_load_weights(ctx, params_dev, buffers_dev, (cudaStream_t)stream, !skip_bn_fold);
}

int unetb200_forward_infer(unetb200_ctx* h, const float* x_dev, float* logits_dev, float* prob_dev, uint8_t* mask_dev,
                           float thresh, int N, void* stream) {
    Ctx* ctx = h->c;
    if (!x_dev) return ctx_fail(ctx, "forward_infer: x is null");
    if (!logits_dev && !prob_dev && !mask_dev) return ctx_fail(ctx, "forward_infer: no output requested");
    UB_CUDA(cudaSetDevice(ctx->device));
    return ctx_forward_infer(ctx, x_dev, logits_dev, prob_dev, mask_dev, thresh, N, (cudaStream_t)stream);
}

// ---- host-buffer inference: two staging slots, three streams (H2D / compute / D2H) ---------------------------------
static int io_prepare(Ctx* ctx) {
    if (ctx->io_stream) return 0;
    const size_t mpx = (size_t)ctx->max_batch * ctx->H * ctx->W;
    for (Ctx::IoSlot& sl : ctx->io) {
        UB_CUDA(cudaMalloc(&sl.x, mpx * 3 * sizeof(float)));
        UB_CUDA(cudaMalloc(&sl.x8, mpx * 3));
        UB_CUDA(cudaMalloc(&sl.f, mpx * 2 * sizeof(float)));
        UB_CUDA(cudaMalloc(&sl.m, mpx));
        UB_CUDA(cudaEventCreateWithFlags(&sl.h2d_done, cudaEventDisableTiming));
        UB_CUDA(cudaEventCreateWithFlags(&sl.compute_done, cudaEventDisableTiming));
        UB_CUDA(cudaEventCreateWithFlags(&sl.d2h_done, cudaEventDisableTiming));
    }
    UB_CUDA(cudaStreamCreateWithFlags(&ctx->io_in, cudaStreamNonBlocking));
    UB_CUDA(cudaStreamCreateWithFlags(&ctx->io_out, cudaStreamNonBlocking));
    UB_CUDA(cudaStreamCreateWithFlags(&ctx->io_stream, cudaStreamNonBlocking));
    return 0;
}

static int io_submit(Ctx* ctx, int slot, const float* x_host, const uint8_t* img_host, int bgr, const float* mean,
                     const float* stdv, float* logits_host, float* prob_host, uint8_t* mask_host, float thresh, int N) {
    if (slot < 0 || slot > 1) return ctx_fail(ctx, "infer_host: slot must be 0 or 1");
    if (!x_host && !img_host) return ctx_fail(ctx, "infer_host: input is null");
    if (N < 1 || N > ctx->max_batch) return ctx_fail(ctx, "infer_host: batch outside [1, max_batch]");
    if (!logits_host && !prob_host && !mask_host) return ctx_fail(ctx, "infer_host: no output requested");
    if (!ctx->weights_ready) return ctx_fail(ctx, "infer_host: weights not loaded");
    UB_CUDA(cudaSetDevice(ctx->device));
    if (io_prepare(ctx)) return 1;
    Ctx::IoSlot& sl = ctx->io[slot];
    if (sl.busy) return ctx_fail(ctx, "infer_host: slot still in flight (call unetb200_infer_host_wait first)");
    const size_t px = (size_t)N * ctx->H * ctx->W;
    NormParams np;
    if (img_host) {
        for (int c = 0; c < 3; ++c) {
            np.mean[c] = mean ? mean[c] : 0.f;
            np.inv_std[c] = 1.f / (stdv ? stdv[c] : 1.f);
        }
        UB_CUDA(cudaMemcpyAsync(sl.x8, img_host, px * 3, cudaMemcpyHostToDevice, ctx->io_in));
    } else {
        UB_CUDA(cudaMemcpyAsync(sl.x, x_host, px * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->io_in));
    }
    UB_CUDA(cudaEventRecord(sl.h2d_done, ctx->io_in));
    cudaStream_t st = ctx->io_stream;
    UB_CUDA(cudaStreamWaitEvent(st, sl.h2d_done, 0));
    if (ctx->weights_event) UB_CUDA(cudaStreamWaitEvent(st, ctx->weights_event, 0));
    float* dl = logits_host ? sl.f : nullptr;
    float* dp = prob_host ? sl.f + (size_t)ctx->max_batch * ctx->H * ctx->W : nullptr;
    uint8_t* dm = mask_host ? sl.m : nullptr;
    if (ctx_forward_infer(ctx, img_host ? nullptr : sl.x, dl, dp, dm, thresh, N, st, nullptr, img_host ? sl.x8 : nullptr,
                          bgr, &np))
        return 1;
    UB_CUDA(cudaEventRecord(sl.compute_done, st));
    UB_CUDA(cudaStreamWaitEvent(ctx->io_out, sl.compute_done, 0));
    if (dl) UB_CUDA(cudaMemcpyAsync(logits_host, dl, px * sizeof(float), cudaMemcpyDeviceToHost, ctx->io_out));
    if (dp) UB_CUDA(cudaMemcpyAsync(prob_host, dp, px * sizeof(float), cudaMemcpyDeviceToHost, ctx->io_out));
    if (dm) UB_CUDA(cudaMemcpyAsync(mask_host, dm, px, cudaMemcpyDeviceToHost, ctx->io_out));
    UB_CUDA(cudaEventRecord(sl.d2h_done, ctx->io_out));
    sl.busy = true;
    return 0;
}

int unetb200_infer_host_submit(unetb200_ctx* h, int slot, const float* x_host, float* logits_host, float* prob_host,
                               uint8_t* mask_host, float thresh, int N) {
    return io_submit(h->c, slot, x_host, nullptr, 0, nullptr, nullptr, logits_host, prob_host, mask_host, thresh, N);
}

int unetb200_infer_host_u8_submit(unetb200_ctx* h, int slot, const uint8_t* img_host, int bgr, const float* mean3,
                                  const float* std3, float* logits_host, float* prob_host, uint8_t* mask_host,
                                  float thresh, int N) {
    return io_submit(h->c, slot, nullptr, img_host, bgr, mean3, std3, logits_host, prob_host, mask_host, thresh, N);
}

int unetb200_infer_host_wait(unetb200_ctx* h, int slot) {
    Ctx* ctx = h->c;
    if (slot < 0 || slot > 1) return ctx_fail(ctx, "infer_host_wait: slot must be 0 or 1");
    Ctx::IoSlot& sl = ctx->io[slot];
    if (!sl.busy) return 0;
    sl.busy = false;
    UB_CUDA(cudaEventSynchronize(sl.d2h_done));
    return 0;
}

int unetb200_infer_host(unetb200_ctx* h, const float* x_host, float* logits_host, float* prob_host, uint8_t* mask_host,
                        float thresh, int N) {
    if (unetb200_infer_host_wait(h, 0)) return 1;
    if (unetb200_infer_host_submit(h, 0, x_host, logits_host, prob_host, mask_host, thresh, N)) return 1;
    return unetb200_infer_host_wait(h, 0);
}

int unetb200_infer_host_u8(unetb200_ctx* h, const uint8_t* img_host, int bgr, const float* mean3, const float* std3,
                           float* logits_host, float* prob_host, uint8_t* mask_host, float thresh, int N) {
    if (unetb200_infer_host_wait(h, 0)) return 1;
    if (unetb200_infer_host_u8_submit(h, 0, img_host, bgr, mean3, std3, logits_host, prob_host, mask_host, thresh, N))
        return 1;
    return unetb200_infer_host_wait(h, 0);
}

/* uint8 HWC device input (frames already on the GPU): same fused pre-processing, caller's stream */
int unetb200_forward_infer_u8(unetb200_ctx* h, const uint8_t* img_dev, int bgr, const float* mean3, const float* std3,
                              float* logits_dev, float* prob_dev, uint8_t* mask_dev, float thresh, int N, void* stream) {
    Ctx* ctx = h->c;
    if (!img_dev) return ctx_fail(ctx, "forward_infer_u8: img is null");
    if (!logits_dev && !prob_dev && !mask_dev) return ctx_fail(ctx, "forward_infer_u8: no output requested");
    UB_CUDA(cudaSetDevice(ctx->device));
    NormParams np;
    for (int c = 0; c < 3; ++c) {
        np.mean[c] = mean3 ? mean3[c] : 0.f;
        np.inv_std[c] = 1.f / (std3 ? std3[c] : 1.f);
    }
    return ctx_forward_infer(ctx, nullptr, logits_dev, prob_dev, mask_dev, thresh, N, (cudaStream_t)stream, nullptr,
                             img_dev, bgr, &np);
}

int unetb200_infer_launch_count(unetb200_ctx* h, int N) {
    auto it = h->c->infer_plans.find(N);
    return it == h->c->infer_plans.end() ? -1 : it->second.launches;
}

int unetb200_infer_debug_count(unetb200_ctx* h, int N) {
    auto it = h->c->infer_plans.find(N);
    return it == h->c->infer_plans.end() ? -1 : (int)it->second.dbg.size();
}
int unetb200_infer_debug_info(unetb200_ctx* h, int N, int index, char* name_out, int name_cap, int shape_out[4]) {
    auto it = h->c->infer_plans.find(N);
    if (it == h->c->infer_plans.end() || index < 0 || index >= (int)it->second.dbg.size()) return 1;
    const Ctx::InferPlan::Dbg& d = it->second.dbg[index];
    if (name_out && name_cap > 0) {
        strncpy(name_out, d.name.c_str(), name_cap - 1);
        name_out[name_cap - 1] = 0;
    }
    if (shape_out) { shape_out[0] = d.n; shape_out[1] = d.h; shape_out[2] = d.w; shape_out[3] = d.c; }
    return 0;
}
int unetb200_infer_debug_copy(unetb200_ctx* h, int N, int index, void* dst_dev, long long cap_bytes, void* stream) {
    Ctx* ctx = h->c;
    auto it = ctx->infer_plans.find(N);
    if (it == ctx->infer_plans.end() || index < 0 || index >= (int)it->second.dbg.size())
        return ctx_fail(ctx, "infer_debug_copy: bad index (run a forward at this batch size first)");
    const Ctx::InferPlan::Dbg& d = it->second.dbg[index];
    const long long bytes = (long long)d.n * d.h * d.w * d.c * 2;
    if (bytes > cap_bytes) return ctx_fail(ctx, "infer_debug_copy: destination too small");
    UB_CUDA(cudaMemcpyAsync(dst_dev, d.ptr, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

int unetb200_profile_infer(unetb200_ctx* h, const float* x_dev, float* logits_dev, int N, void* stream, float* ms_out,
                           int* is_igemm_out, int cap, int* n_out) {
    Ctx* ctx = h->c;
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<cudaEvent_t> ev;
    if (ctx_forward_infer(ctx, x_dev, logits_dev, nullptr, nullptr, 0.5f, N, st, &ev)) return 1;
    UB_CUDA(cudaStreamSynchronize(st));
    const int n = (int)ev.size() - 1;
    auto& plan = ctx->infer_plans[N];
    for (int i = 0; i < n && i < cap; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
        ms_out[i] = ms;
        // launch 0 = input pack, last = seg head (a tensor-core conv launch), others = plan steps
        is_igemm_out[i] = (i >= 1 && i <= (int)plan.steps.size()) ? plan.steps[i - 1].is_igemm : (i == n - 1 ? 1 : 0);
    }
    for (auto e : ev) cudaEventDestroy(e);
    if (n_out) *n_out = n;
    return 0;
}

int unetb200_profile_name(unetb200_ctx* h, int N, int i, char* out, int cap) {
    auto it = h->c->infer_plans.find(N);
    if (it == h->c->infer_plans.end() || cap < 1) return 1;
    const int ns = (int)it->second.steps.size();
    std::string nm = i == 0 ? "input_pack" : (i <= ns ? it->second.steps[i - 1].name : "segmentation_head");
    strncpy(out, nm.c_str(), cap - 1);
    out[cap - 1] = 0;
    return 0;
}

int unetb200_conv_nhwc(unetb200_ctx* h, const void* in, const float* w, const float* scale, const float* shift,
                       const void* residual, int relu, void* out, float* stats, int N, int H, int W, int cin, int cout,
                       int k, int stride, void* stream) {
    Ctx* ctx = h->c;
    cudaStream_t st = (cudaStream_t)stream;
    if (!(k == 1 || k == 3) || !(stride == 1 || stride == 2)) return ctx_fail(ctx, "conv_nhwc: k in {1,3}, stride in {1,2}");
    if (cin % 16 || cout % 16) return ctx_fail(ctx, "conv_nhwc: channel counts must be multiples of 16");
    UB_CUDA(cudaSetDevice(ctx->device));
    ConvRef c;
    c.cin = cin; c.cout = cout; c.k = k; c.stride = stride;
    const size_t wn = (size_t)cout * cin * k * k;
    __nv_bfloat16* wpk = nullptr;
    UB_CUDA(cudaMallocAsync(&wpk, wn * 2, st));
    launch_k(pack_conv_w_kernel, ew_grid(wn, 256, ctx->num_sms), 256, 0, st, w, wpk, cout, cin, k, k, 0);
    EpilogueDesc ep;
    ep.scale = scale; ep.shift = shift; ep.relu = relu;
    const int Ho = H / stride, Wo = W / stride;
    if (residual) ep.residual = nhwc_view(residual, N, Ho, Wo, cout);
    float* part = nullptr;
    if (stats) {
        UB_CUDA(cudaMallocAsync(&part, (size_t)ctx->num_sms * cout * 2 * sizeof(float), st));
        ep.stats = part;
    }
    IgemmLaunch L;
    std::string e = build_conv(ctx, L, c, wpk, in, N, H, W, out, ep);
    if (!e.empty()) return ctx_fail(ctx, "conv_nhwc: " + e);
    UB_CUDA(igemm_launch(L, st));
    if (stats) {
        launch_k(reduce_partials_kernel, (cout * 2 + 127) / 128, 128, 0, st, part, stats, L.grid, cout * 2);
        UB_CUDA(cudaGetLastError());
        UB_CUDA(cudaFreeAsync(part, st));
    }
    UB_CUDA(cudaFreeAsync(wpk, st));
    return 0;
}

// ------------------------------------------------------------------------------------------------ training step
int unetb200_train_forward(unetb200_ctx* h, const float* x_dev, float* logits_dev, const float* params_dev,
                           float* buffers_dev, long long* counters_dev, float* grads_dev, int N, void* stream) {
    Ctx* ctx = h->c;
    if (!x_dev || !logits_dev || !params_dev || !buffers_dev || !counters_dev || !grads_dev)
        return ctx_fail(ctx, "train_forward: null pointer");
    if (N < 1 || N > ctx->max_batch) return ctx_fail(ctx, "train_forward: batch outside [1, max_batch]");
    UB_CUDA(cudaSetDevice(ctx->device));
    return ctx_train_forward(ctx, x_dev, logits_dev, params_dev, buffers_dev, counters_dev, grads_dev, N,
                             (cudaStream_t)stream);
}

int unetb200_train_forward_u8(unetb200_ctx* h, const uint8_t* img_dev, int bgr, const float* mean3, const float* std3,
                              float* logits_dev, const float* params_dev, float* buffers_dev, long long* counters_dev,
                              float* grads_dev, int N, void* stream) {
    Ctx* ctx = h->c;
    if (!img_dev || !logits_dev || !params_dev || !buffers_dev || !counters_dev || !grads_dev)
        return ctx_fail(ctx, "train_forward_u8: null pointer");
    if (N < 1 || N > ctx->max_batch) return ctx_fail(ctx, "train_forward_u8: batch outside [1, max_batch]");
    UB_CUDA(cudaSetDevice(ctx->device));
    NormParams np;
    for (int c = 0; c < 3; ++c) {
        np.mean[c] = mean3 ? mean3[c] : 0.f;
        np.inv_std[c] = 1.f / (std3 ? std3[c] : 1.f);
    }
    return ctx_train_forward(ctx, nullptr, logits_dev, params_dev, buffers_dev, counters_dev, grads_dev, N,
                             (cudaStream_t)stream, img_dev, bgr, &np);
}

int unetb200_train_backward(unetb200_ctx* h, const float* dlogits_dev, int N, int stage_first, int stage_last,
                            void* stream) {
    Ctx* ctx = h->c;
    UB_CUDA(cudaSetDevice(ctx->device));
    return ctx_train_backward(ctx, dlogits_dev, N, stage_first, stage_last, (cudaStream_t)stream);
}

int unetb200_grad_bucket_range(int stage, long long* begin_out, long long* end_out) {
    if (stage < 0 || stage > 3 || !begin_out || !end_out) return 1;
    grad_bucket_range(spec(), stage, begin_out, end_out);
    return 0;
}

int unetb200_train_launch_count(unetb200_ctx* h, int N, int* fwd_out, int* bwd_out) {
    TrainState* T = static_cast<TrainState*>(h->c->train);
    if (!T) return 1;
    auto it = T->plans.find(N);
    if (it == T->plans.end()) return 1;
    if (fwd_out) *fwd_out = it->second->n_fwd;
    if (bwd_out) *bwd_out = it->second->n_bwd;
    return 0;
}

int unetb200_train_debug_count(unetb200_ctx* h, int N) {
    TrainState* T = static_cast<TrainState*>(h->c->train);
    if (!T) return -1;
    auto it = T->plans.find(N);
    return it == T->plans.end() ? -1 : (int)it->second->dbg.size();
}
int unetb200_train_debug_info(unetb200_ctx* h, int N, int index, char* name_out, int name_cap, int shape_out[4],
                              int* is_bf16_out) {
    TrainState* T = static_cast<TrainState*>(h->c->train);
    if (!T) return 1;
    auto it = T->plans.find(N);
    if (it == T->plans.end() || index < 0 || index >= (int)it->second->dbg.size()) return 1;
    const TrainPlan::Dbg& d = it->second->dbg[index];
    if (name_out && name_cap > 0) {
        strncpy(name_out, d.name.c_str(), name_cap - 1);
        name_out[name_cap - 1] = 0;
    }
    if (shape_out) { shape_out[0] = d.n; shape_out[1] = d.h; shape_out[2] = d.w; shape_out[3] = d.c; }
    if (is_bf16_out) *is_bf16_out = d.bf16;
    return 0;
}
int unetb200_train_debug_copy(unetb200_ctx* h, int N, int index, void* dst_dev, long long cap_bytes, void* stream) {
    Ctx* ctx = h->c;
    TrainState* T = static_cast<TrainState*>(ctx->train);
    if (!T) return ctx_fail(ctx, "train_debug_copy: no training state");
    auto it = T->plans.find(N);
    if (it == T->plans.end() || index < 0 || index >= (int)it->second->dbg.size())
        return ctx_fail(ctx, "train_debug_copy: bad index");
    const TrainPlan::Dbg& d = it->second->dbg[index];
    const long long bytes = (long long)d.n * d.h * d.w * d.c * (d.bf16 ? 2 : 4);
    if (bytes > cap_bytes) return ctx_fail(ctx, "train_debug_copy: destination too small");
    UB_CUDA(cudaMemcpyAsync(dst_dev, d.ptr, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

int unetb200_profile_enable(unetb200_ctx* h, int on) {
    Ctx* ctx = h->c;
    for (auto& pe : ctx->prof_ev) cudaEventDestroy(pe.second);
    ctx->prof_ev.clear();
    ctx->prof_on = on != 0;
    return 0;
}
int unetb200_profile_dump(unetb200_ctx* h, const char* path) {
    Ctx* ctx = h->c;
    UB_CUDA(cudaDeviceSynchronize());
    FILE* f = fopen(path, "w");
    if (!f) return ctx_fail(ctx, "profile_dump: cannot open file");
    fprintf(f, "launch,kind,layer,ms\n");
    for (size_t i = 0; i + 1 < ctx->prof_ev.size(); ++i) {
        const std::string& nm = ctx->prof_ev[i].first;
        if (nm.rfind("end:", 0) == 0) continue;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->prof_ev[i].second, ctx->prof_ev[i + 1].second);
        const size_t c = nm.find(':');
        fprintf(f, "%zu,%s,%s,%.5f\n", i, nm.substr(0, c).c_str(), nm.substr(c + 1).c_str(), ms);
    }
    fclose(f);
    return 0;
}

static int loss_fail(const char* msg) {
    g_create_error = msg;
    return 1;
}
int unetb200_loss_scratch_floats(void) { return 4 * 1024; }

int unetb200_loss_bce_dice_forward(const float* logits_dev, const float* target_dev, long long n, float eps,
                                   float* scratch_dev, float* result_dev, void* stream) {
    if (!logits_dev || !target_dev || !result_dev || !scratch_dev || n < 1) return loss_fail("loss_forward: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    long long nb = (n + 1023) / 1024;
    if (nb > 1024) nb = 1024;
    launch_k(loss_partial_kernel, (int)nb, 256, 0, st, logits_dev, target_dev, scratch_dev, n);
    launch_k(loss_finalize_kernel, 1, 32, 0, st, scratch_dev, (int)nb, (double)n, eps, result_dev);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : loss_fail(cudaGetErrorString(e));
}

int unetb200_loss_bce_dice_backward(const float* logits_dev, const float* target_dev, const float* result_dev,
                                    const float* g_bce_dev, const float* g_dice_dev, float gscale, float eps,
                                    float* dlogits_dev, long long n, void* stream) {
    if (!logits_dev || !target_dev || !result_dev || !dlogits_dev || n < 1) return loss_fail("loss_backward: bad argument");
    long long nb = (n + 1023) / 1024;
    if (nb > 148 * 16) nb = 148 * 16;
    launch_k(loss_bwd_kernel, (int)nb, 256, 0, (cudaStream_t)stream, logits_dev, target_dev, result_dev, g_bce_dev, g_dice_dev,
                                                               gscale, eps, dlogits_dev, n);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : loss_fail(cudaGetErrorString(e));
}

int unetb200_seg_metrics_scratch_floats(int N) { return N > 0 ? N * kMetricBlocks * 3 : 0; }

int unetb200_seg_metrics(const float* pred_dev, const float* target_dev, int N, long long hw, float thresh, float eps,
                         float* scratch_dev, float* out2_dev, void* stream) {
    if (!pred_dev || !target_dev || !scratch_dev || !out2_dev || N < 1 || hw < 1) return loss_fail("seg_metrics: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    launch_k(seg_metrics_partial_kernel, dim3(kMetricBlocks, N), 256, 0, st, pred_dev, target_dev, hw, thresh, scratch_dev);
    launch_k(seg_metrics_finalize_kernel, 1, 256, 0, st, scratch_dev, N, eps, out2_dev);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : loss_fail(cudaGetErrorString(e));
}

int unetb200_adamw_step(unetb200_ctx* h, float* params_dev, float* grads_dev, float* exp_avg_dev, float* exp_avg_sq_dev,
                        long long n, float lr, float beta1, float beta2, float eps, float weight_decay, long long step,
                        float grad_scale, int zero_grad, void* stream) {
    Ctx* ctx = h->c;
    if (!params_dev || !grads_dev || !exp_avg_dev || !exp_avg_sq_dev || n < 1 || step < 1)
        return ctx_fail(ctx, "adamw_step: bad argument");
    UB_CUDA(cudaSetDevice(ctx->device));
    const float bc1 = 1.f - (float)pow((double)beta1, (double)step);
    const float bc2s = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    launch_k(adamw_kernel, ew_grid(n / 4 + 4, 256, ctx->num_sms), 256, 0, (cudaStream_t)stream, 
        params_dev, grads_dev, exp_avg_dev, exp_avg_sq_dev, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2s, grad_scale,
        zero_grad);
    UB_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
