// U-Net(ResNet-34) network description + inference orchestration on top of the igemm / elementwise kernels.
// Tensor table order == the state_dict order of smp.Unet("resnet34"), parameters and
// buffers each in their own flat fp32 array owned by the caller (PyTorch).
#pragma once
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "conv_plan.cuh"
#include "ops.cuh"
#include "pack.cuh"

namespace ub {

struct TensorInfo {
    std::string name;
    int ndim;
    int shape[4];
    long long offset;  // element offset inside its flat array
    int kind;          // 0 = parameter (fp32 flat params), 1 = fp32 buffer (flat buffers), 2 = int64 counter (index)
    long long numel() const {
        long long n = 1;
        for (int i = 0; i < ndim; ++i) n *= shape[i];
        return n;
    }
};

struct BnRef {
    int c = 0;
    long long gamma = -1, beta = -1;  // offsets into flat params
    long long mean = -1, var = -1;    // offsets into flat buffers
    long long fold = -1;              // offset into the folded scale/shift arrays
    int counter = -1;
};
struct ConvRef {
    std::string name;
    int cin = 0, cout = 0, k = 0, stride = 1;
    long long w = -1;     // offset into flat params
    long long bias = -1;  // head only
    int bn = -1;          // index into bns
    long long wpk = -1;   // offset (elements) into the packed bf16 weight arena
    long long wpk_elems = 0;
    int hc = 0;           // > 0: runs on the halo-resident kernel (hconv.cuh) with this many pipeline stages
    int hc_cup = 0;       // hconv: channels taken from the nearest-2x up-sampled low-res source (decoder conv1)
    int tc = 0;           // 1: runs on the TMA halo kernel (tconv.cuh), 2: its parity mode (up-sampled input, no skip),
                          // 3: decoder conv1 split into two tconv launches: parity mode over the up-sampled channels
                          //    (scaled, bf16) then plain mode over the skip channels with that as the residual
                          // 4: the same split for the wide blocks: wpconv (parity, PK_DEC1 matrix) + wconv or tconv
    long long wpk2 = -1;  // tc == 3: offset of the second (skip-channel) operand
};

struct NetSpec {
    std::vector<TensorInfo> tensors;
    std::vector<ConvRef> convs;
    std::vector<BnRef> bns;
    long long n_params = 0, n_buffers = 0;
    int n_counters = 0;
    long long fold_total = 0, wpk_total = 0;
    // indices
    int stem = -1, head = -1;
    struct Block { int c1, c2, ds; };
    std::vector<Block> enc_blocks[4];
    struct Dec { int c1, c2; int cup, cskip, cout; };
    std::vector<Dec> dec;

    int add_bn(const std::string& prefix, int c) {
        BnRef b;
        b.c = c;
        b.gamma = n_params;
        tensors.push_back({prefix + ".weight", 1, {c, 0, 0, 0}, n_params, 0});
        n_params += c;
        b.beta = n_params;
        tensors.push_back({prefix + ".bias", 1, {c, 0, 0, 0}, n_params, 0});
        n_params += c;
        b.mean = n_buffers;
        tensors.push_back({prefix + ".running_mean", 1, {c, 0, 0, 0}, n_buffers, 1});
        n_buffers += c;
        b.var = n_buffers;
        tensors.push_back({prefix + ".running_var", 1, {c, 0, 0, 0}, n_buffers, 1});
        n_buffers += c;
        b.counter = n_counters;
        tensors.push_back({prefix + ".num_batches_tracked", 0, {0, 0, 0, 0}, n_counters, 2});
        n_counters += 1;
        b.fold = fold_total;
        fold_total += c;
        bns.push_back(b);
        return int(bns.size()) - 1;
    }
    int add_conv(const std::string& wname, int cin, int cout, int k, int stride) {
        ConvRef c;
        c.name = wname;
        c.cin = cin; c.cout = cout; c.k = k; c.stride = stride;
        c.w = n_params;
        tensors.push_back({wname, 4, {cout, cin, k, k}, n_params, 0});
        n_params += (long long)cout * cin * k * k;
        convs.push_back(c);
        return int(convs.size()) - 1;
    }

    NetSpec() {
        stem = add_conv("encoder.conv1.weight", 3, 64, 7, 2);
        convs[stem].bn = add_bn("encoder.bn1", 64);
        const int nblocks[4] = {3, 4, 6, 3};
        const int planes[4] = {64, 128, 256, 512};
        int inpl = 64;
        for (int l = 0; l < 4; ++l) {
            for (int b = 0; b < nblocks[l]; ++b) {
                const std::string p = "encoder.layer" + std::to_string(l + 1) + "." + std::to_string(b);
                const int stride = (b == 0 && l > 0) ? 2 : 1;
                Block blk;
                blk.c1 = add_conv(p + ".conv1.weight", inpl, planes[l], 3, stride);
                convs[blk.c1].bn = add_bn(p + ".bn1", planes[l]);
                blk.c2 = add_conv(p + ".conv2.weight", planes[l], planes[l], 3, 1);
                convs[blk.c2].bn = add_bn(p + ".bn2", planes[l]);
                blk.ds = -1;
                if (stride != 1 || inpl != planes[l]) {
                    blk.ds = add_conv(p + ".downsample.0.weight", inpl, planes[l], 1, stride);
                    convs[blk.ds].bn = add_bn(p + ".downsample.1", planes[l]);
                }
                inpl = planes[l];
                enc_blocks[l].push_back(blk);
            }
        }
        const int cin[5] = {512, 256, 128, 64, 32}, cskip[5] = {256, 128, 64, 64, 0}, co[5] = {256, 128, 64, 32, 16};
        for (int i = 0; i < 5; ++i) {
            const std::string p = "decoder.blocks." + std::to_string(i);
            Dec d;
            d.cup = cin[i]; d.cskip = cskip[i]; d.cout = co[i];
            d.c1 = add_conv(p + ".conv1.0.weight", cin[i] + cskip[i], co[i], 3, 1);
            convs[d.c1].bn = add_bn(p + ".conv1.1", co[i]);
            d.c2 = add_conv(p + ".conv2.0.weight", co[i], co[i], 3, 1);
            convs[d.c2].bn = add_bn(p + ".conv2.1", co[i]);
            dec.push_back(d);
        }
        head = add_conv("segmentation_head.0.weight", 16, 1, 3, 1);
        convs[head].bias = n_params;
        tensors.push_back({"segmentation_head.0.bias", 1, {1, 0, 0, 0}, n_params, 0});
        n_params += 1;

        // packed bf16 weight arena (forward operands)
        for (size_t i = 0; i < convs.size(); ++i) {
            ConvRef& c = convs[i];
            if ((int)i == head) {  // 16 -> 16 tconv operand, rows 0 / 1 = bf16 high / low parts of the fp32 weights (PK_HEAD)
                c.wpk = wpk_total;
                c.wpk_elems = 9 * 16 * 16;
                wpk_total += c.wpk_elems;
                continue;
            }
            long long n;
            if ((int)i == stem) n = 64 * 224;
            else n = (long long)c.cout * c.cin * c.k * c.k;
            bool is_dec1 = false;
            for (auto& d : dec) if (d.c1 == (int)i) { n = 4ll * d.cout * (9 * d.cskip + 4 * d.cup); is_dec1 = true; }
            if (c.k == 3 && c.stride == 1) {
                int cup = 0;
                for (auto& d : dec) if (d.c1 == (int)i) cup = d.cup;
                c.hc = hconv_stages(cup, c.cin - cup, c.cout);
                c.hc_cup = c.hc ? cup : 0;
                if (c.hc && 9ll * c.cin * c.cout > n) n = 9ll * c.cin * c.cout;
                if (c.hc && cup == 0 && tconv_ok(c.cin, c.cout, false)) c.tc = 1;
                if (c.hc && cup == c.cin && tconv_ok(c.cin, c.cout, true)) {
                    c.tc = 2;
                    if (tconv_w_elems(c.cin, c.cout, true) > n) n = tconv_w_elems(c.cin, c.cout, true);
                }
                if (c.hc && cup > 0 && cup < c.cin && tconv_ok(cup, c.cout, true) && tconv_ok(c.cin - cup, c.cout, false)) {
                    c.tc = 3;
                    n = tconv_w_elems(cup, c.cout, true) + tconv_w_elems(c.cin - cup, c.cout, false);
                }
            }
            if (is_dec1 && !c.hc) {
                int cup = 0;
                for (auto& d : dec) if (d.c1 == (int)i) cup = d.cup;
                const int cskip = c.cin - cup;
                if (cskip > 0 && wpconv_ok(cup, c.cout) && (wconv_ok(cskip, c.cout) || tconv_ok(cskip, c.cout, false))) {
                    c.tc = 4;
                    c.hc_cup = cup;
                    if (!wconv_ok(cskip, c.cout)) {
                        c.wpk2 = n;  // relative for now: the skip-slice tconv operand follows the PK_DEC1 matrix
                        n += tconv_w_elems(cskip, c.cout, false);
                    }
                }
            }
            c.wpk = wpk_total;
            c.wpk_elems = n;
            if (c.tc == 3) c.wpk2 = c.wpk + tconv_w_elems(c.hc_cup, c.cout, true);
            if (c.tc == 4 && c.wpk2 >= 0) c.wpk2 += c.wpk;
            wpk_total += (n + 63) & ~63ll;  // keep every matrix 128 B aligned
        }
    }
};

// ------------------------------------------------------------------------------------------------ runtime context
struct Ctx {
    NetSpec spec;
    int device = 0, num_sms = 148;
    int max_batch = 0, H = 0, W = 0;
    std::string last_error;
    int* d_err = nullptr;
    // caches owned by the library
    __nv_bfloat16* wpk = nullptr;    // packed forward weights
    float* fold_scale = nullptr;     // [fold_total]
    float* fold_shift = nullptr;
    float* head_w = nullptr;         // fp32 copy [144] + bias [1]
    bool weights_ready = false;
    PackTable fwd_pack;              // forward operand layouts of all convs (one launch)
    int* fold_tab = nullptr;         // [n_bn][5]: gamma, beta, mean, var offsets, fold offset (one launch for all BN folds)
    // activation arena
    uint8_t* arena = nullptr;
    size_t arena_bytes = 0;

    struct InferPlan {
        int N = 0;
        __nv_bfloat16* xp = nullptr;
        __nv_bfloat16* head_in = nullptr;
        TconvLaunch head;   // seg head on the tensor core (outputs patched per call)
        struct Step {
            std::function<cudaError_t(cudaStream_t)> fn;
            std::string name;
            int is_igemm;
        };
        std::vector<Step> steps;  // everything between input pack and head
        int launches = 0;
        // named activations of this plan (NHWC bf16) for the layer-local parity tests (unetb200_infer_debug_*)
        struct Dbg { std::string name; const void* ptr; int n, h, w, c; };
        std::vector<Dbg> dbg;
    };
    std::map<int, InferPlan> infer_plans;  // keyed by batch size
    // staging for the host-buffer entry points (unetb200_infer_host / _submit / _wait): two slots, so that the H2D
    // copy of request k+1 and the D2H copy of request k-1 run on the copy engines while request k computes
    struct IoSlot {
        float* x = nullptr;       // fp32 NCHW input staging [max_batch,3,H,W]
        uint8_t* x8 = nullptr;    // uint8 HWC input staging [max_batch,H,W,3]
        float* f = nullptr;       // logits | prob staging [2][max_batch,H,W]
        uint8_t* m = nullptr;     // mask staging
        cudaEvent_t h2d_done = nullptr, compute_done = nullptr, d2h_done = nullptr;
        bool busy = false;
    };
    IoSlot io[2];
    cudaStream_t io_in = nullptr, io_stream = nullptr, io_out = nullptr;  // H2D, compute, D2H
    cudaEvent_t weights_event = nullptr;  // recorded after every weight re-pack (cross-stream ordering for io_stream)
    // the activation arena is shared by every forward of this context: a forward enqueued on a different stream than
    // the previous one first waits for that one to finish (caller's stream vs the library's io_stream)
    cudaEvent_t arena_event = nullptr;
    cudaStream_t arena_stream = nullptr;
    bool arena_used = false;
    // optional per-launch profiling of the train step: (name, event recorded BEFORE the launch)
    bool prof_on = false;
    std::vector<std::pair<std::string, cudaEvent_t>> prof_ev;
    void prof_mark(const std::string& name, cudaStream_t st) {
        if (!prof_on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        prof_ev.emplace_back(name, e);
    }
    // training state (train.cuh), type-erased so that inference-only translation units need not see it
    void* train = nullptr;
    void (*train_free)(void*) = nullptr;

    ~Ctx() {
        if (train && train_free) train_free(train);
        if (weights_event) cudaEventDestroy(weights_event);
        if (arena_event) cudaEventDestroy(arena_event);
        for (IoSlot& sl : io) {
            cudaFree(sl.x);
            cudaFree(sl.x8);
            cudaFree(sl.f);
            cudaFree(sl.m);
            if (sl.h2d_done) cudaEventDestroy(sl.h2d_done);
            if (sl.compute_done) cudaEventDestroy(sl.compute_done);
            if (sl.d2h_done) cudaEventDestroy(sl.d2h_done);
        }
        if (io_in) cudaStreamDestroy(io_in);
        if (io_stream) cudaStreamDestroy(io_stream);
        if (io_out) cudaStreamDestroy(io_out);
        cudaFree(d_err);
        cudaFree(wpk);
        cudaFree(fold_scale);
        cudaFree(fold_shift);
        cudaFree(head_w);
        cudaFree(fold_tab);
        cudaFree(arena);
    }
};

#define UB_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            ctx->last_error = std::string(#expr) + ": " + cudaGetErrorString(_e);                  \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

inline int ctx_fail(Ctx* ctx, const std::string& msg) {
    ctx->last_error = msg;
    return 1;
}

// all eval-mode BatchNorm folds in one launch: tab[b] = {gamma, beta, mean, var, fold} offsets, channel count
__global__ void bn_fold_all_kernel(const int* __restrict__ tab, int nbn, const float* __restrict__ params,
                                   const float* __restrict__ buffers, float eps, float* __restrict__ scale,
                                   float* __restrict__ shift) {
    griddep_launch();
    griddep_wait();
    const int b = blockIdx.x;
    if (b >= nbn) return;
    const int* t = tab + b * 6;
    for (int c = threadIdx.x; c < t[5]; c += blockDim.x) {
        const float s = params[t[0] + c] * rsqrtf(buffers[t[3] + c] + eps);
        scale[t[4] + c] = s;
        shift[t[4] + c] = params[t[1] + c] - buffers[t[2] + c] * s;
    }
}

inline std::string ctx_build_pack_tables(Ctx* ctx) {
    const NetSpec& S = ctx->spec;
    PackTable& T = ctx->fwd_pack;
    for (size_t i = 0; i < S.convs.size(); ++i) {
        const ConvRef& c = S.convs[i];
        PackEntry e;
        if ((int)i == S.head) {
            T.add(pk_entry(PK_HEAD, c.w, c.wpk, 9 * 16 * 16));
            continue;
        }
        if (c.tc == 3) {
            const int cup = c.hc_cup, cskip = c.cin - cup;
            e = pk_entry(PK_HPAR, c.w, c.wpk, tconv_w_elems(cup, c.cout, true));
            e.cout = c.cout; e.cin = cup; e.a = c.cin;
            T.add(e);
            e = pk_entry(PK_HCONV, c.w, c.wpk2, tconv_w_elems(cskip, c.cout, false));
            e.cout = c.cout; e.cin = cskip; e.a = c.cin; e.b = cup; e.c = 0;
        } else if (c.tc == 2) {
            e = pk_entry(PK_HPAR, c.w, c.wpk, tconv_w_elems(c.cin, c.cout, true));
            e.cout = c.cout; e.cin = c.cin; e.a = c.cin; e.b = 0; e.c = 0;
        } else if (c.hc) {
            e = pk_entry(PK_HCONV, c.w, c.wpk, 9ll * c.cin * c.cout);
            e.cout = c.cout; e.cin = c.cin; e.a = c.cin; e.b = 0; e.c = 0;
        } else if ((int)i == S.stem) {
            e = pk_entry(PK_STEM2, c.w, c.wpk, 64 * 224);
        } else {
            const NetSpec::Dec* dd = nullptr;
            for (auto& d : S.dec) if (d.c1 == (int)i) dd = &d;
            if (dd) {
                e = pk_entry(PK_DEC1, c.w, c.wpk, 4ll * dd->cout * (9 * dd->cskip + 4 * dd->cup));
                e.cout = dd->cout; e.cin = dd->cup; e.a = dd->cskip;
                if (c.tc == 4 && c.wpk2 >= 0) {  // + the skip-channel slice as a tconv operand
                    T.add(e);
                    e = pk_entry(PK_HCONV, c.w, c.wpk2, tconv_w_elems(dd->cskip, dd->cout, false));
                    e.cout = dd->cout; e.cin = dd->cskip; e.a = c.cin; e.b = dd->cup; e.c = 0;
                }
            } else {
                e = pk_entry(PK_CONV, c.w, c.wpk, (long long)c.cout * c.cin * c.k * c.k);
                e.cout = c.cout; e.cin = c.cin; e.a = c.k; e.b = c.k; e.c = 0;
            }
        }
        T.add(e);
    }
    if (T.upload() != cudaSuccess) return "pack table upload failed";
    std::vector<int> ft;
    for (const BnRef& b : S.bns) {
        ft.push_back((int)b.gamma); ft.push_back((int)b.beta); ft.push_back((int)b.mean); ft.push_back((int)b.var);
        ft.push_back((int)b.fold); ft.push_back(b.c);
    }
    if (cudaMalloc(&ctx->fold_tab, ft.size() * sizeof(int)) != cudaSuccess ||
        cudaMemcpy(ctx->fold_tab, ft.data(), ft.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess)
        return "fold table upload failed";
    return "";
}

// Re-pack fp32 master weights (flat params / buffers in state-dict order) into the bf16 operand caches and (unless
// fold_bn == false: training uses batch statistics) fold eval-mode BatchNorm.  Three launches in total.
inline int ctx_load_weights(Ctx* ctx, const float* params, const float* buffers, cudaStream_t st, bool fold_bn = true) {
    const NetSpec& S = ctx->spec;
    const ConvRef& hc = S.convs[S.head];
    // a forward enqueued on another stream (the host-slot pipeline's io_stream) may still be reading the operand caches
    if (ctx->arena_used && ctx->arena_stream != st) UB_CUDA(cudaStreamWaitEvent(st, ctx->arena_event, 0));
    UB_CUDA(cudaMemcpyAsync(ctx->head_w, params + hc.w, 145 * sizeof(float), cudaMemcpyDeviceToDevice, st));  // weight + bias
    UB_CUDA(ctx->fwd_pack.launch(params, ctx->wpk, st));
    if (fold_bn) {
        launch_k(bn_fold_all_kernel, (int)S.bns.size(), 128, 0, st, ctx->fold_tab, (int)S.bns.size(), params, buffers, 1e-5f,
                                                              ctx->fold_scale, ctx->fold_shift);
        UB_CUDA(cudaGetLastError());
    }
    if (!ctx->weights_event) UB_CUDA(cudaEventCreateWithFlags(&ctx->weights_event, cudaEventDisableTiming));
    UB_CUDA(cudaEventRecord(ctx->weights_event, st));
    // inference kernels read the packed weights / folded constants BEFORE griddepcontrol.wait (see tconv.cuh): the packs
    // must have completed before any of them can be launched.  Weight loads are rare (once per checkpoint / validation).
    if (fold_bn) UB_CUDA(cudaStreamSynchronize(st));
    ctx->weights_ready = true;
    return 0;
}

// -------------------------------------------------------------------------------------------- conv launch builders
inline void taps_3x3(std::vector<IgemmTap>& taps, int src, int cin, int c0, int chunk, int off_h, int off_w) {
    for (int r = 0; r < 3; ++r)
        for (int s = 0; s < 3; ++s) {
            IgemmTap t;
            t.dw = int16_t(s - 1 + off_w);
            t.dh = int16_t(r - 1 + off_h);
            t.c0 = int16_t(c0);
            t.nchunks = int16_t(cin / chunk);
            t.src = src;
            taps.push_back(t);
        }
}
inline int chunk_for(int cin) { return cin >= 64 ? 64 : cin; }

// plain k x k conv (k = 1 or 3), stride 1 or 2, NHWC in -> NHWC out
inline std::string build_conv(Ctx* ctx, IgemmLaunch& L, const ConvRef& c, const __nv_bfloat16* wpk, const void* in,
                              int N, int Hin, int Win, void* out, const EpilogueDesc& ep) {
    const int chunk = chunk_for(c.cin);
    if (c.cin % chunk) return "cin not divisible by chunk";
    SrcDesc s;
    s.v = nhwc_view(in, N, Hin, Win, c.cin);
    s.es_w = s.es_h = c.stride;
    std::vector<IgemmTap> taps;
    if (c.k == 3) {
        taps_3x3(taps, 0, c.cin, 0, chunk, 0, 0);
    } else {
        IgemmTap t;
        t.dw = 0; t.dh = 0; t.c0 = 0; t.nchunks = int16_t(c.cin / chunk); t.src = 0;
        taps.push_back(t);
    }
    const int Ho = Hin / c.stride, Wo = Win / c.stride;
    View4 o = nhwc_view(out, N, Ho, Wo, c.cout);
    return igemm_build(L, &s, 1, taps.data(), (int)taps.size(), chunk, wpk, c.cin * c.k * c.k, c.cout, o, ep,
                       ctx->d_err, ctx->num_sms);
}

// 7x7 s2 stem over the packed input xp [N][H][W+8][4]: 7 row taps, each a 32-element (8 px x 4 ch) window.
inline std::string build_stem(Ctx* ctx, IgemmLaunch& L, const __nv_bfloat16* wpk, const void* xp, int N, int H, int W,
                              void* out, const EpilogueDesc& ep) {
    SrcDesc s;
    s.v.ptr = xp;
    s.v.C = 32;
    s.v.W = W / 2;           // one window per output column, 16 B apart (overlapping)
    s.v.H = H;
    s.v.N = N;
    s.v.sW = 8;
    s.v.sH = (long long)(W + 8) * 4;
    s.v.sN = (long long)H * (W + 8) * 4;
    s.es_w = 1;
    s.es_h = 2;
    std::vector<IgemmTap> taps;
    for (int r = 0; r < 7; ++r) {
        IgemmTap t;
        t.dw = 0; t.dh = int16_t(r - 3); t.c0 = 0; t.nchunks = 1; t.src = 0;
        taps.push_back(t);
    }
    View4 o = nhwc_view(out, N, H / 2, W / 2, 64);
    return igemm_build(L, &s, 1, taps.data(), 7, 32, wpk, 224, 64, o, ep, ctx->d_err, ctx->num_sms);
}

// decoder conv1 for one output parity: sources = skip (full res, traversal stride 2) and low-res x (stride 1).
inline std::string build_dec1(Ctx* ctx, IgemmLaunch& L, const NetSpec::Dec& d, const __nv_bfloat16* wpk_all, int par,
                              const void* low, const void* skip, int N, int Hl, int Wl, void* out,
                              const EpilogueDesc& ep) {
    const int ph = par >> 1, pw = par & 1;
    const int minc = (d.cskip && d.cskip < d.cup) ? d.cskip : d.cup;
    const int chunk = chunk_for(minc);
    const int kt = 9 * d.cskip + 4 * d.cup;
    SrcDesc s[2];
    std::vector<IgemmTap> taps;
    int nsrc = 1;
    // source 0: low-res x ; source 1: skip
    s[0].v = nhwc_view(low, N, Hl, Wl, d.cup);
    if (d.cskip) {
        s[1].v = nhwc_view(skip, N, Hl * 2, Wl * 2, d.cskip);
        s[1].es_w = s[1].es_h = 2;
        nsrc = 2;
        taps_3x3(taps, 1, d.cskip, 0, chunk, ph, pw);
    }
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
            IgemmTap t;
            t.dh = int16_t(a - 1 + ph);
            t.dw = int16_t(b - 1 + pw);
            t.c0 = 0;
            t.nchunks = int16_t(d.cup / chunk);
            t.src = 0;
            taps.push_back(t);
        }
    // strided output view: pixels (2i+ph, 2j+pw) of the [N, 2Hl, 2Wl, cout] tensor
    View4 o;
    o.ptr = reinterpret_cast<const __nv_bfloat16*>(out) + ((long long)ph * (Wl * 2) + pw) * d.cout;
    o.C = d.cout; o.W = Wl; o.H = Hl; o.N = N;
    o.sW = 2ll * d.cout;
    o.sH = 2ll * (Wl * 2) * d.cout;
    o.sN = (long long)(Hl * 2) * (Wl * 2) * d.cout;
    return igemm_build(L, s, nsrc, taps.data(), (int)taps.size(), chunk, wpk_all + (long long)par * d.cout * kt, kt,
                       d.cout, o, ep, ctx->d_err, ctx->num_sms);
}

// ------------------------------------------------------------------------------------------------ inference plan
struct ArenaCarver {
    uint8_t* base;
    size_t off = 0, cap;
    ArenaCarver(uint8_t* b, size_t c) : base(b), cap(c) {}
    __nv_bfloat16* take(long long elems) {
        size_t bytes = (size_t(elems) * 2 + 1023) & ~size_t(1023);
        uint8_t* p = base ? base + off : nullptr;
        off += bytes;
        return reinterpret_cast<__nv_bfloat16*>(p);
    }
};

// Builds (or sizes, when ctx->arena == nullptr) the launch list for batch N.  Returns "" on success.
inline std::string build_infer_plan(Ctx* ctx, int N, Ctx::InferPlan& plan, size_t* arena_needed) {
    TconvConstWeightsScope const_weights_scope;
    const NetSpec& S = ctx->spec;
    const int H = ctx->H, W = ctx->W;
    ArenaCarver A(ctx->arena, ctx->arena_bytes);
    const bool dry = (ctx->arena == nullptr);
    plan.N = N;
    plan.steps.clear();
    plan.dbg.clear();
    auto reg = [&](const std::string& nm, const void* p, int hh, int ww, int c) {
        if (!dry) plan.dbg.push_back({nm, p, N, hh, ww, c});
    };
    auto fold = [&](int bn, int relu) {
        EpilogueDesc ep;
        ep.scale = ctx->fold_scale + S.bns[bn].fold;
        ep.shift = ctx->fold_shift + S.bns[bn].fold;
        ep.relu = relu;
        return ep;
    };
    std::string err;
    auto add_igemm = [&](const IgemmLaunch& L, const std::string& name) {
        plan.steps.push_back({[L](cudaStream_t st) { return igemm_launch(L, st); }, name, 1});
    };

    // 3x3 / 1x1 conv of the encoder or a decoder conv2: halo-resident kernel where it applies, tap-table kernel otherwise
    auto add_conv = [&](const ConvRef& c, const void* in, int hin, int win, void* out, const EpilogueDesc& ep) -> std::string {
        if (c.tc == 1) {
            TconvLaunch TL;
            std::string e = tconv_build(TL, in, c.cin, false, ctx->wpk + c.wpk, c.cout, N, hin, win, out, ep, ctx->d_err,
                                        ctx->num_sms);
            if (!e.empty()) return c.name + ": " + e;
            plan.steps.push_back({[TL](cudaStream_t st) { return tconv_launch(TL, st); }, c.name, 1});
            return "";
        }
        if (c.hc) {
            HconvLaunch HL;
            std::string e = hconv_build(HL, nullptr, 0, in, c.cin, ctx->wpk + c.wpk, c.cout, N, hin, win, out, ep, ctx->d_err,
                                        ctx->num_sms);
            if (!e.empty()) return c.name + ": " + e;
            plan.steps.push_back({[HL](cudaStream_t st) { return hconv_launch(HL, st); }, c.name, 1});
            return "";
        }
        if (c.k == 3 && c.stride == 1 && wconv_ok(c.cin, c.cout)) {
            // wide layers: halo-resident operand + streamed weights (wconv.cuh); same packed weight matrix as igemm
            WconvLaunch WL;
            std::string e = wconv_build(WL, in, c.cin, ctx->wpk + c.wpk, c.cout, N, hin, win, out, ep, ctx->d_err,
                                        ctx->num_sms);
            if (!e.empty()) return c.name + ": " + e;
            plan.steps.push_back({[WL](cudaStream_t st) { return wconv_launch(WL, st); }, c.name, 1});
            return "";
        }
        IgemmLaunch L;
        std::string e = build_conv(ctx, L, c, ctx->wpk + c.wpk, in, N, hin, win, out, ep);
        if (!e.empty()) return c.name + ": " + e;
        add_igemm(L, c.name);
        return "";
    };

    plan.xp = A.take((long long)N * H * (W + 8) * 4 + 64);  // + 128 B of slack: see tconv_build_stem
    __nv_bfloat16* f1 = A.take((long long)N * (H / 2) * (W / 2) * 64);
    reg("xp", plan.xp, H, W + 8, 4);
    reg("encoder.conv1.weight/out", f1, H / 2, W / 2, 64);
    if (!dry) {
        TconvLaunch TL;
        err = tconv_build_stem(TL, plan.xp, ctx->wpk + S.convs[S.stem].wpk, N, H, W, f1, fold(S.convs[S.stem].bn, 1),
                               ctx->d_err, ctx->num_sms);
        if (!err.empty()) return "stem: " + err;
        plan.steps.push_back({[TL](cudaStream_t st) { return tconv_launch(TL, st); }, "encoder.conv1", 1});
    }
    int h = H / 4, w = W / 4;
    __nv_bfloat16* cur = A.take((long long)N * h * w * 64);
    reg("encoder.maxpool/out", cur, h, w, 64);
    if (!dry) {
        const int num_sms = ctx->num_sms;
        const int Hh = H / 2, Wh = W / 2;
        plan.steps.push_back({[=](cudaStream_t st) {
            launch_k(maxpool3x3s2_kernel, ew_grid((long long)N * (Hh / 2) * (Wh / 2) * 8, 256, num_sms), 256, 0, st, 
                f1, cur, N, Hh, Wh, 64);
            return cudaGetLastError();
        }, "encoder.maxpool", 0});
    }
    __nv_bfloat16* feats[5] = {f1, nullptr, nullptr, nullptr, nullptr};  // f1, layer1..4 outputs
    for (int l = 0; l < 4; ++l) {
        for (size_t b = 0; b < S.enc_blocks[l].size(); ++b) {
            const NetSpec::Block& blk = S.enc_blocks[l][b];
            const ConvRef& c1 = S.convs[blk.c1];
            const ConvRef& c2 = S.convs[blk.c2];
            const int ho = h / c1.stride, wo = w / c1.stride;
            __nv_bfloat16* t = A.take((long long)N * ho * wo * c1.cout);
            __nv_bfloat16* o = A.take((long long)N * ho * wo * c1.cout);
            __nv_bfloat16* ident = cur;
            if (blk.ds >= 0) ident = A.take((long long)N * ho * wo * c1.cout);
            reg(c1.name + "/out", t, ho, wo, c1.cout);
            reg(c2.name + "/out", o, ho, wo, c1.cout);
            if (blk.ds >= 0) reg(S.convs[blk.ds].name + "/out", ident, ho, wo, c1.cout);
            if (!dry) {
                if (!(err = add_conv(c1, cur, h, w, t, fold(c1.bn, 1))).empty()) return err;
                if (blk.ds >= 0) {
                    const ConvRef& cd = S.convs[blk.ds];
                    if (!(err = add_conv(cd, cur, h, w, ident, fold(cd.bn, 0))).empty()) return err;
                }
                EpilogueDesc ep = fold(c2.bn, 1);
                ep.residual = nhwc_view(ident, N, ho, wo, c1.cout);
                if (!(err = add_conv(c2, t, ho, wo, o, ep)).empty()) return err;
            }
            cur = o;
            h = ho;
            w = wo;
        }
        feats[l + 1] = cur;
    }
    // decoder: skips = layer3, layer2, layer1, f1, none
    __nv_bfloat16* skips[5] = {feats[3], feats[2], feats[1], feats[0], nullptr};
    for (int i = 0; i < 5; ++i) {
        const NetSpec::Dec& d = S.dec[i];
        const ConvRef& c1 = S.convs[d.c1];
        const ConvRef& c2 = S.convs[d.c2];
        __nv_bfloat16* t = A.take((long long)N * (2 * h) * (2 * w) * d.cout);
        __nv_bfloat16* o = A.take((long long)N * (2 * h) * (2 * w) * d.cout);
        __nv_bfloat16* r = (c1.tc == 3 || c1.tc == 4) ? A.take((long long)N * (2 * h) * (2 * w) * d.cout) : nullptr;
        reg(c1.name + "/out", t, 2 * h, 2 * w, d.cout);
        reg(c2.name + "/out", o, 2 * h, 2 * w, d.cout);
        if (r) reg(c1.name + "/up", r, 2 * h, 2 * w, d.cout);   // scale * conv over the up-sampled channels (bf16)
        if (!dry) {
            if (c1.tc == 4) {
                // wide blocks: launch 1 = wpconv (scale * parity-folded conv of the up-sampled channels on the low-res
                // tensor -> r), launch 2 = wconv / tconv over the skip channels with r as the residual
                const int kt = 9 * d.cskip + 4 * d.cup;
                EpilogueDesc e2 = fold(c1.bn, 1);
                e2.residual = nhwc_view(r, N, 2 * h, 2 * w, d.cout);
                WpconvLaunch WP;
                err = wpconv_build(WP, cur, d.cup, ctx->wpk + c1.wpk, kt, 9 * d.cskip, d.cout, N, h, w, r, e2.scale,
                                   ctx->d_err, ctx->num_sms);
                if (!err.empty()) return c1.name + ": " + err;
                plan.steps.push_back({[WP](cudaStream_t st) { return wpconv_launch(WP, st); }, c1.name + "[up]", 1});
                if (wconv_ok(d.cskip, d.cout)) {
                    WconvLaunch WL;
                    err = wconv_build(WL, skips[i], d.cskip, ctx->wpk + c1.wpk, d.cout, N, 2 * h, 2 * w, t, e2, ctx->d_err,
                                      ctx->num_sms, kt);
                    if (!err.empty()) return c1.name + ": " + err;
                    plan.steps.push_back({[WL](cudaStream_t st) { return wconv_launch(WL, st); }, c1.name + "[skip]", 1});
                } else {
                    TconvLaunch TL;
                    err = tconv_build(TL, skips[i], d.cskip, false, ctx->wpk + c1.wpk2, d.cout, N, 2 * h, 2 * w, t, e2,
                                      ctx->d_err, ctx->num_sms);
                    if (!err.empty()) return c1.name + ": " + err;
                    plan.steps.push_back({[TL](cudaStream_t st) { return tconv_launch(TL, st); }, c1.name + "[skip]", 1});
                }
            } else if (c1.tc == 3) {
                // launch 1: scale * conv(up-sampled channels) by parity folding on the low-res tensor -> r (bf16);
                // launch 2: relu(scale * conv(skip channels) + shift + r).  Nothing up-sampled or concatenated exists.
                EpilogueDesc e1, e2 = fold(c1.bn, 1);
                e1.scale = e2.scale;
                e2.residual = nhwc_view(r, N, 2 * h, 2 * w, d.cout);
                TconvLaunch T1, T2;
                err = tconv_build(T1, cur, d.cup, true, ctx->wpk + c1.wpk, d.cout, N, 2 * h, 2 * w, r, e1, ctx->d_err,
                                  ctx->num_sms);
                if (err.empty())
                    err = tconv_build(T2, skips[i], d.cskip, false, ctx->wpk + c1.wpk2, d.cout, N, 2 * h, 2 * w, t, e2,
                                      ctx->d_err, ctx->num_sms);
                if (!err.empty()) return c1.name + ": " + err;
                plan.steps.push_back({[T1](cudaStream_t st) { return tconv_launch(T1, st); }, c1.name + "[up]", 1});
                plan.steps.push_back({[T2](cudaStream_t st) { return tconv_launch(T2, st); }, c1.name + "[skip]", 1});
            } else if (c1.tc == 2) {
                // nearest-2x upsample folded into four 2x2-tap parity convolutions on the low-res tensor: ONE launch
                TconvLaunch TL;
                err = tconv_build(TL, cur, d.cup, true, ctx->wpk + c1.wpk, d.cout, N, 2 * h, 2 * w, t, fold(c1.bn, 1),
                                  ctx->d_err, ctx->num_sms);
                if (!err.empty()) return c1.name + ": " + err;
                plan.steps.push_back({[TL](cudaStream_t st) { return tconv_launch(TL, st); }, c1.name, 1});
            } else if (c1.hc) {
                // fused nearest-2x upsample + concat inside the halo loader: ONE launch, original 3x3 weights
                HconvLaunch HL;
                err = hconv_build(HL, cur, d.cup, skips[i], d.cskip, ctx->wpk + c1.wpk, d.cout, N, 2 * h, 2 * w, t,
                                  fold(c1.bn, 1), ctx->d_err, ctx->num_sms);
                if (!err.empty()) return c1.name + ": " + err;
                plan.steps.push_back({[HL](cudaStream_t st) { return hconv_launch(HL, st); }, c1.name, 1});
            } else {
                for (int par = 0; par < 4; ++par) {
                    IgemmLaunch L;
                    err = build_dec1(ctx, L, d, ctx->wpk + c1.wpk, par, cur, skips[i], N, h, w, t, fold(c1.bn, 1));
                    if (!err.empty()) return c1.name + ": " + err;
                    add_igemm(L, c1.name + "[parity " + std::to_string(par) + "]");
                }
            }
            if (!(err = add_conv(c2, t, 2 * h, 2 * w, o, fold(c2.bn, 1))).empty()) return err;
        }
        cur = o;
        h *= 2;
        w *= 2;
    }
    plan.head_in = cur;
    if (!dry) {
        err = tconv_build_head(plan.head, cur, ctx->wpk + S.convs[S.head].wpk, N, h, w, ctx->d_err, ctx->num_sms);
        if (!err.empty()) return "segmentation_head: " + err;
        plan.head.p.head_bias = ctx->head_w + 144;
    }
    plan.launches = (int)plan.steps.size() + 2;  // + input pack + head
    if (arena_needed) *arena_needed = A.off;
    return "";
}

// logits / prob / mask: any may be null (at least one non-null). x: fp32 NCHW [N,3,H,W] device pointer.
// ev != nullptr: record an event before every launch and one at the end (per-launch timing for bench.py).
// x8 != nullptr: the input is uint8 HWC [N,H,W,3] instead (pack_input_u8_kernel; norm / bgr describe the pre-processing).
inline int ctx_forward_infer(Ctx* ctx, const float* x, float* logits, float* prob, uint8_t* mask, float thresh, int N,
                             cudaStream_t st, std::vector<cudaEvent_t>* ev = nullptr, const uint8_t* x8 = nullptr,
                             int bgr = 0, const NormParams* norm = nullptr) {
    auto mark = [&]() {
        if (!ev) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        ev->push_back(e);
    };
    if (!ctx->weights_ready) return ctx_fail(ctx, "forward_infer: weights not loaded");
    if (N < 1 || N > ctx->max_batch) return ctx_fail(ctx, "forward_infer: batch outside [1, max_batch]");
    auto it = ctx->infer_plans.find(N);
    if (it == ctx->infer_plans.end()) {
        Ctx::InferPlan plan;
        std::string e = build_infer_plan(ctx, N, plan, nullptr);
        if (!e.empty()) return ctx_fail(ctx, "plan: " + e);
        it = ctx->infer_plans.emplace(N, std::move(plan)).first;
    }
    Ctx::InferPlan& P = it->second;
    const int H = ctx->H, W = ctx->W;
    if (!ctx->arena_event) UB_CUDA(cudaEventCreateWithFlags(&ctx->arena_event, cudaEventDisableTiming));
    if (ctx->arena_used && ctx->arena_stream != st) UB_CUDA(cudaStreamWaitEvent(st, ctx->arena_event, 0));
    mark();
    if (x8)
        launch_k(pack_input_u8_kernel, ew_grid((long long)N * H * ((W + 8) / 2), 256, ctx->num_sms), 256, 0, st, x8, P.xp, N, H,
                                                                                                          W, bgr, *norm);
    else
        launch_k(pack_input_kernel, ew_grid((long long)N * H * ((W + 8) / 2), 256, ctx->num_sms), 256, 0, st, x, P.xp, N, H, W);
    UB_CUDA(cudaGetLastError());
    for (auto& s : P.steps) {
        mark();
        UB_CUDA(s.fn(st));
    }
    mark();
    float tl = 0.f;
    if (thresh <= 0.f) tl = -INFINITY;
    else if (thresh >= 1.f) tl = INFINITY;
    else tl = logf(thresh / (1.f - thresh));
    UB_CUDA(tconv_launch_head(P.head, logits, prob, mask, tl, st));
    mark();
    UB_CUDA(cudaEventRecord(ctx->arena_event, st));
    ctx->arena_stream = st;
    ctx->arena_used = true;
    return 0;
}

inline int ctx_create(Ctx** out, int device, int max_batch, int H, int W, std::string* err) {
    if (H % 32 || W % 32 || H < 32 || W < 32) {
        *err = "H and W must be positive multiples of 32";
        return 1;
    }
    if (max_batch < 1) {
        *err = "max_batch must be >= 1";
        return 1;
    }
    if ((long long)max_batch * H * W >= (1ll << 31)) {   // the pool / pack / head kernels index elements with 32 bits
        *err = "max_batch * H * W must be below 2^31";
        return 1;
    }
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        *err = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return 1;
    }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        *err = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e);
        return 1;
    }
    if (prop.major != 10) {
        *err = "this library contains sm_100a code only; device is sm_" + std::to_string(prop.major) +
               std::to_string(prop.minor);
        return 1;
    }
    Ctx* ctx = new Ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    ctx->max_batch = max_batch;
    ctx->H = H;
    ctx->W = W;
    const NetSpec& S = ctx->spec;
    bool ok = cudaMalloc(&ctx->d_err, sizeof(int)) == cudaSuccess &&
              cudaMemset(ctx->d_err, 0, sizeof(int)) == cudaSuccess &&
              cudaMalloc(&ctx->wpk, S.wpk_total * 2) == cudaSuccess &&
              cudaMalloc(&ctx->fold_scale, S.fold_total * 4) == cudaSuccess &&
              cudaMalloc(&ctx->fold_shift, S.fold_total * 4) == cudaSuccess &&
              cudaMalloc(&ctx->head_w, 145 * 4) == cudaSuccess;
    if (ok) {
        Ctx::InferPlan tmp;
        size_t need = 0;
        std::string pe = build_infer_plan(ctx, max_batch, tmp, &need);  // dry run: arena == nullptr
        ok = pe.empty() && cudaMalloc(&ctx->arena, need) == cudaSuccess &&
             cudaMemset(ctx->arena, 0, need) == cudaSuccess;
        ctx->arena_bytes = need;
    }
    if (!ok) {
        *err = std::string("allocation failed: ") + cudaGetErrorString(cudaGetLastError());
        delete ctx;
        return 1;
    }
    {
        std::string pe = ctx_build_pack_tables(ctx);
        if (!pe.empty()) {
            *err = pe;
            delete ctx;
            return 1;
        }
    }
    *out = ctx;
    return 0;
}

}  // namespace ub
