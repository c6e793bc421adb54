// extern "C" surface of libunetb200.so (declared in include/unetb200.h).
#include <string.h>

#include "../../include/unetb200.h"
#include "unet.cuh"

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
    Ctx* ctx = h->c;
    if (!params_dev || !buffers_dev) return ctx_fail(ctx, "load_weights: null pointer");
    UB_CUDA(cudaSetDevice(ctx->device));
    return ctx_load_weights(ctx, params_dev, buffers_dev, (cudaStream_t)stream);
}

int unetb200_forward_infer(unetb200_ctx* h, const float* x_dev, float* logits_dev, float* prob_dev, uint8_t* mask_dev,
                           float thresh, int N, void* stream) {
    Ctx* ctx = h->c;
    if (!x_dev) return ctx_fail(ctx, "forward_infer: x is null");
    if (!logits_dev && !prob_dev && !mask_dev) return ctx_fail(ctx, "forward_infer: no output requested");
    UB_CUDA(cudaSetDevice(ctx->device));
    return ctx_forward_infer(ctx, x_dev, logits_dev, prob_dev, mask_dev, thresh, N, (cudaStream_t)stream);
}

int unetb200_infer_host(unetb200_ctx* h, const float* x_host, float* logits_host, float* prob_host, uint8_t* mask_host,
                        float thresh, int N) {
    Ctx* ctx = h->c;
    if (!x_host) return ctx_fail(ctx, "infer_host: x is null");
    if (N < 1 || N > ctx->max_batch) return ctx_fail(ctx, "infer_host: batch outside [1, max_batch]");
    UB_CUDA(cudaSetDevice(ctx->device));
    const size_t px = (size_t)N * ctx->H * ctx->W;
    if (!ctx->io_x) {
        const size_t mpx = (size_t)ctx->max_batch * ctx->H * ctx->W;
        UB_CUDA(cudaMalloc(&ctx->io_x, mpx * 3 * sizeof(float)));
        UB_CUDA(cudaMalloc(&ctx->io_f, mpx * 2 * sizeof(float)));
        UB_CUDA(cudaMalloc(&ctx->io_m, mpx));
        UB_CUDA(cudaStreamCreateWithFlags(&ctx->io_stream, cudaStreamNonBlocking));
    }
    cudaStream_t st = ctx->io_stream;
    if (ctx->weights_event) UB_CUDA(cudaStreamWaitEvent(st, ctx->weights_event, 0));
    UB_CUDA(cudaMemcpyAsync(ctx->io_x, x_host, px * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    float* dl = logits_host ? ctx->io_f : nullptr;
    float* dp = prob_host ? ctx->io_f + (size_t)ctx->max_batch * ctx->H * ctx->W : nullptr;
    uint8_t* dm = mask_host ? ctx->io_m : nullptr;
    if (!dl && !dp && !dm) return ctx_fail(ctx, "infer_host: no output requested");
    if (ctx_forward_infer(ctx, ctx->io_x, dl, dp, dm, thresh, N, st)) return 1;
    if (dl) UB_CUDA(cudaMemcpyAsync(logits_host, dl, px * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (dp) UB_CUDA(cudaMemcpyAsync(prob_host, dp, px * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (dm) UB_CUDA(cudaMemcpyAsync(mask_host, dm, px, cudaMemcpyDeviceToHost, st));
    UB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int unetb200_infer_launch_count(unetb200_ctx* h, int N) {
    auto it = h->c->infer_plans.find(N);
    return it == h->c->infer_plans.end() ? -1 : it->second.launches;
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
        // launch 0 = input pack, last = head, others = plan steps
        is_igemm_out[i] = (i >= 1 && i <= (int)plan.steps.size()) ? plan.steps[i - 1].is_igemm : 0;
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
    pack_conv_w_kernel<<<ew_grid(wn, 256, ctx->num_sms), 256, 0, st>>>(w, wpk, cout, cin, k, k, 0);
    EpilogueDesc ep;
    ep.scale = scale; ep.shift = shift; ep.relu = relu;
    const int Ho = H / stride, Wo = W / stride;
    if (residual) ep.residual = nhwc_view(residual, N, Ho, Wo, cout);
    float* part = nullptr;
    const int mt = igemm_m_tiles(nhwc_view(out, N, Ho, Wo, cout));
    if (stats) {
        UB_CUDA(cudaMallocAsync(&part, (size_t)mt * cout * 2 * sizeof(float), st));
        ep.stats = part;
    }
    IgemmLaunch L;
    std::string e = build_conv(ctx, L, c, wpk, in, N, H, W, out, ep);
    if (!e.empty()) return ctx_fail(ctx, "conv_nhwc: " + e);
    UB_CUDA(igemm_launch(L, st));
    if (stats) {
        reduce_partials_kernel<<<(cout * 2 + 127) / 128, 128, 0, st>>>(part, stats, mt, cout * 2);
        UB_CUDA(cudaGetLastError());
        UB_CUDA(cudaFreeAsync(part, st));
    }
    UB_CUDA(cudaFreeAsync(wpk, st));
    return 0;
}

}  // extern "C"
