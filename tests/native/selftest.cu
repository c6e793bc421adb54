// Native GPU self-test for the igemm conv kernel and helper kernels (no torch needed; runs in seconds on a B200).
// Each case compares against a straightforward CPU convolution (fp32 accumulate over bf16-rounded operands).
// Usage: selftest [case-filter-substring]
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../vickers_hardness_unet_b200/csrc/unet.cuh"
#include "wconv2_glue.cuh"

using namespace ub;

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
            exit(2);                                                                  \
        }                                                                             \
    } while (0)

static uint32_t rng_state = 12345;
static float frand() {  // uniform in [-1, 1)
    rng_state = rng_state * 1664525u + 1013904223u;
    return ((rng_state >> 8) & 0xFFFF) / 32768.0f - 1.0f;
}
static float bf16r(float x) { return __bfloat162float(__float2bfloat16(x)); }

struct HostT {  // NHWC fp32 mirror of a bf16 device tensor
    int N, H, W, C;
    std::vector<float> v;
    HostT(int n, int h, int w, int c) : N(n), H(h), W(w), C(c), v((size_t)n * h * w * c) {}
    float& at(int n, int h, int w, int c) { return v[(((size_t)n * H + h) * W + w) * C + c]; }
    float get(int n, int h, int w, int c) const {
        if (h < 0 || h >= H || w < 0 || w >= W) return 0.f;
        return v[(((size_t)n * H + h) * W + w) * C + c];
    }
};
static __nv_bfloat16* to_dev_bf16(const std::vector<float>& v) {
    std::vector<__nv_bfloat16> h(v.size());
    for (size_t i = 0; i < v.size(); ++i) h[i] = __float2bfloat16(v[i]);
    __nv_bfloat16* d;
    CK(cudaMalloc(&d, v.size() * 2 + 256));
    CK(cudaMemcpy(d, h.data(), v.size() * 2, cudaMemcpyHostToDevice));
    return d;
}
static float* to_dev_f32(const std::vector<float>& v) {
    float* d;
    CK(cudaMalloc(&d, v.size() * 4 + 256));
    CK(cudaMemcpy(d, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
    return d;
}
static std::vector<float> from_dev_bf16(const __nv_bfloat16* d, size_t n) {
    std::vector<__nv_bfloat16> h(n);
    CK(cudaMemcpy(h.data(), d, n * 2, cudaMemcpyDeviceToHost));
    std::vector<float> v(n);
    for (size_t i = 0; i < n; ++i) v[i] = __bfloat162float(h[i]);
    return v;
}
static void fill_rand_bf16(std::vector<float>& v, float scale) {
    for (auto& x : v) x = bf16r(frand() * scale);
}

// CPU conv: in NHWC, w OIHW (already rounded as the GPU sees it), out NHWC fp32
static HostT cpu_conv(const HostT& in, const std::vector<float>& w, int cout, int k, int stride, int pad) {
    const int Ho = (in.H + 2 * pad - k) / stride + 1, Wo = (in.W + 2 * pad - k) / stride + 1;
    HostT out(in.N, Ho, Wo, cout);
    // re-layout weights to [co][r][s][ci] for speed
    std::vector<float> wr((size_t)cout * k * k * in.C);
    for (int co = 0; co < cout; ++co)
        for (int ci = 0; ci < in.C; ++ci)
            for (int r = 0; r < k; ++r)
                for (int s = 0; s < k; ++s)
                    wr[(((size_t)co * k + r) * k + s) * in.C + ci] = w[(((size_t)co * in.C + ci) * k + r) * k + s];
#pragma omp parallel for collapse(2)
    for (int n = 0; n < in.N; ++n)
        for (int ho = 0; ho < Ho; ++ho)
            for (int wo = 0; wo < Wo; ++wo)
                for (int co = 0; co < cout; ++co) {
                    double acc = 0;
                    for (int r = 0; r < k; ++r) {
                        const int h = ho * stride - pad + r;
                        if (h < 0 || h >= in.H) continue;
                        for (int s = 0; s < k; ++s) {
                            const int ww = wo * stride - pad + s;
                            if (ww < 0 || ww >= in.W) continue;
                            const float* ip = &in.v[(((size_t)n * in.H + h) * in.W + ww) * in.C];
                            const float* wp = &wr[(((size_t)co * k + r) * k + s) * in.C];
                            float a = 0;
                            for (int ci = 0; ci < in.C; ++ci) a += ip[ci] * wp[ci];
                            acc += a;
                        }
                    }
                    out.at(n, ho, wo, co) = (float)acc;
                }
    return out;
}

struct Cmp {
    double max_abs = 0, max_ref = 0, sum_abs = 0;
    size_t n = 0, worst = 0;
};
static Cmp compare(const std::vector<float>& got, const std::vector<float>& ref) {
    Cmp c;
    c.n = ref.size();
    for (size_t i = 0; i < ref.size(); ++i) {
        const double d = fabs((double)got[i] - ref[i]);
        if (!(d <= c.max_abs)) {  // also catches NaN
            c.max_abs = d;
            c.worst = i;
        }
        c.sum_abs += d;
        c.max_ref = fmax(c.max_ref, fabs(ref[i]));
    }
    return c;
}

static int g_fail = 0, g_run = 0;
static Ctx* g_ctx = nullptr;

static bool report(const char* name, const Cmp& c, double tol_rel, const std::vector<float>& got,
                   const std::vector<float>& ref, int C) {
    const double rel = c.max_abs / (c.max_ref + 1e-30);
    const bool ok = rel <= tol_rel;
    printf("[%s] %-44s max_abs %.4g (ref max %.4g, rel %.3g) mean_abs %.3g\n", ok ? "PASS" : "FAIL", name, c.max_abs,
           c.max_ref, rel, c.sum_abs / (double)c.n);
    if (!ok) {
        g_fail++;
        printf("       worst idx %zu (pixel %zu, ch %zu): got %g ref %g\n", c.worst, c.worst / C, c.worst % C,
               got[c.worst], ref[c.worst]);
        // a few samples to help diagnose layout problems
        for (size_t i = 0; i < 6 && i < ref.size(); ++i) printf("       [%zu] got %g ref %g\n", i, got[i], ref[i]);
        size_t nbad = 0;
        for (size_t i = 0; i < ref.size(); ++i)
            if (!(fabs((double)got[i] - ref[i]) <= tol_rel * c.max_ref)) nbad++;
        printf("       %zu / %zu elements out of tolerance\n", nbad, ref.size());
    }
    g_run++;
    return ok;
}
static int check_err_flag(const char* name) {
    int e = 0;
    CK(cudaMemcpy(&e, g_ctx->d_err, 4, cudaMemcpyDeviceToHost));
    if (e) {
        printf("[FAIL] %s: pipeline timeout flag = %d (1 producer, 2 mma-tmem, 3 mma-full, 4 epilogue)\n", name, e);
        CK(cudaMemset(g_ctx->d_err, 0, 4));
        g_fail++;
    }
    return e;
}

// ------------------------------------------------------------------------------------------------ cases
static void case_conv(const char* name, int N, int H, int W, int cin, int cout, int k, int stride, bool residual,
                      bool relu, bool scale_shift, bool stats) {
    HostT in(N, H, W, cin);
    fill_rand_bf16(in.v, 1.0f);
    std::vector<float> w((size_t)cout * cin * k * k);
    const float ws = 1.0f / sqrtf((float)cin * k * k);
    for (auto& x : w) x = frand() * ws;
    std::vector<float> wq(w.size());
    for (size_t i = 0; i < w.size(); ++i) wq[i] = bf16r(w[i]);
    HostT ref = cpu_conv(in, wq, cout, k, stride, k / 2);
    std::vector<float> sc(cout, 1.f), sh(cout, 0.f);
    if (scale_shift)
        for (int c = 0; c < cout; ++c) {
            sc[c] = 0.5f + 0.5f * fabsf(frand());
            sh[c] = 0.3f * frand();
        }
    HostT res(N, ref.H, ref.W, cout);
    if (residual) fill_rand_bf16(res.v, 1.0f);
    for (size_t i = 0; i < ref.v.size(); ++i) {
        const int c = i % cout;
        float v = ref.v[i] * sc[c] + sh[c];
        if (residual) v += res.v[i];
        if (relu) v = fmaxf(v, 0.f);
        ref.v[i] = v;
    }
    __nv_bfloat16* d_in = to_dev_bf16(in.v);
    float* d_w = to_dev_f32(w);
    __nv_bfloat16* d_wpk;
    CK(cudaMalloc(&d_wpk, w.size() * 2));
    pack_conv_w_kernel<<<64, 256>>>(d_w, d_wpk, cout, cin, k, k, 0);
    __nv_bfloat16* d_out;
    CK(cudaMalloc(&d_out, ref.v.size() * 2));
    CK(cudaMemset(d_out, 0xFF, ref.v.size() * 2));  // NaN pattern: unwritten outputs are caught
    float *d_sc = to_dev_f32(sc), *d_sh = to_dev_f32(sh);
    __nv_bfloat16* d_res = residual ? to_dev_bf16(res.v) : nullptr;
    ConvRef c;
    c.cin = cin; c.cout = cout; c.k = k; c.stride = stride;
    EpilogueDesc ep;
    if (scale_shift) { ep.scale = d_sc; ep.shift = d_sh; }
    ep.relu = relu;
    if (residual) ep.residual = nhwc_view(d_res, N, ref.H, ref.W, cout);
    View4 ov = nhwc_view(d_out, N, ref.H, ref.W, cout);
    const int mt = igemm_m_tiles(ov);
    float* d_stats = nullptr;
    if (stats) {
        CK(cudaMalloc(&d_stats, (size_t)mt * cout * 2 * 4));
        CK(cudaMemset(d_stats, 0, (size_t)mt * cout * 2 * 4));
        ep.stats = d_stats;
    }
    IgemmLaunch L;
    std::string e = build_conv(g_ctx, L, c, d_wpk, d_in, N, H, W, d_out, ep);
    if (!e.empty()) {
        printf("[FAIL] %s: build: %s\n", name, e.c_str());
        g_fail++;
        return;
    }
    printf("       %s: grid %d smem %u stages %d tile %dx%dx%d ntile %d chunk %d taps %d chunks %d\n", name, L.grid,
           L.smem, L.p.stages, L.p.bw, L.p.bh, L.p.bn, L.p.ntile, L.p.chunk_elems, L.p.num_taps, L.p.total_chunks);
    CK(igemm_launch(L, 0));
    CK(cudaDeviceSynchronize());
    if (check_err_flag(name)) return;
    std::vector<float> got = from_dev_bf16(d_out, ref.v.size());
    Cmp cm = compare(got, ref.v);
    report(name, cm, 0.012, got, ref.v, cout);
    if (stats) {
        std::vector<float> hs((size_t)mt * cout * 2);
        CK(cudaMemcpy(hs.data(), d_stats, hs.size() * 4, cudaMemcpyDeviceToHost));
        std::vector<float> s_got(cout * 2, 0.f), s_ref(cout * 2, 0.f);
        for (int t = 0; t < mt; ++t)
            for (int ch = 0; ch < cout * 2; ++ch) s_got[ch] += hs[(size_t)t * cout * 2 + ch];
        for (size_t i = 0; i < got.size(); ++i) {  // statistics are defined over the bf16 output the kernel wrote
            s_ref[(i % cout) * 2] += got[i];
            s_ref[(i % cout) * 2 + 1] += got[i] * got[i];
        }
        Cmp cs = compare(s_got, s_ref);
        std::string nm = std::string(name) + " [stats]";
        report(nm.c_str(), cs, 2e-3, s_got, s_ref, 2);
    }
    cudaFree(d_in); cudaFree(d_w); cudaFree(d_wpk); cudaFree(d_out); cudaFree(d_sc); cudaFree(d_sh);
    if (d_res) cudaFree(d_res);
    if (d_stats) cudaFree(d_stats);
}

static void case_stem(const char* name, int N, int H, int W) {
    // input fp32 NCHW
    std::vector<float> x((size_t)N * 3 * H * W);
    for (auto& v : x) v = frand() * 2.f;
    HostT in(N, H, W, 3);
    for (int n = 0; n < N; ++n)
        for (int c = 0; c < 3; ++c)
            for (int h = 0; h < H; ++h)
                for (int w = 0; w < W; ++w) in.at(n, h, w, c) = bf16r(x[(((size_t)n * 3 + c) * H + h) * W + w]);
    std::vector<float> w((size_t)64 * 3 * 49), wq(w.size());
    for (auto& v : w) v = frand() * 0.08f;
    for (size_t i = 0; i < w.size(); ++i) wq[i] = bf16r(w[i]);
    HostT ref = cpu_conv(in, wq, 64, 7, 2, 3);
    for (auto& v : ref.v) v = fmaxf(v, 0.f);
    float* d_x = to_dev_f32(x);
    float* d_w = to_dev_f32(w);
    __nv_bfloat16 *d_xp, *d_wpk, *d_out;
    CK(cudaMalloc(&d_xp, (size_t)N * H * (W + 8) * 4 * 2));
    CK(cudaMemset(d_xp, 0xFF, (size_t)N * H * (W + 8) * 4 * 2));
    CK(cudaMalloc(&d_wpk, 64 * 224 * 2));
    CK(cudaMalloc(&d_out, ref.v.size() * 2));
    CK(cudaMemset(d_out, 0xFF, ref.v.size() * 2));
    pack_input_kernel<<<ew_grid((long long)N * H * ((W + 8) / 2), 256, 148), 256>>>(d_x, d_xp, N, H, W);
    pack_stem_w_kernel<<<(64 * 224 + 255) / 256, 256>>>(d_w, d_wpk);
    EpilogueDesc ep;
    ep.relu = 1;
    IgemmLaunch L;
    std::string e = build_stem(g_ctx, L, d_wpk, d_xp, N, H, W, d_out, ep);
    if (!e.empty()) {
        printf("[FAIL] %s: build: %s\n", name, e.c_str());
        g_fail++;
        return;
    }
    CK(igemm_launch(L, 0));
    CK(cudaDeviceSynchronize());
    if (check_err_flag(name)) return;
    std::vector<float> got = from_dev_bf16(d_out, ref.v.size());
    report(name, compare(got, ref.v), 0.012, got, ref.v, 64);
    cudaFree(d_x); cudaFree(d_w); cudaFree(d_xp); cudaFree(d_wpk); cudaFree(d_out);
}

// wide decoder conv1 as two halo-resident launches: wpconv (parity-folded up-sampled channels -> bf16 partial) and wconv
// over the skip channels with that partial as the residual; same reference as case_dec1.  iters > 0: time both.
static void case_dec1_split(const char* name, int N, int Hl, int Wl, int cup, int cskip, int cout, int iters = 0) {
    HostT low(N, Hl, Wl, cup);
    fill_rand_bf16(low.v, 1.0f);
    HostT skip(N, 2 * Hl, 2 * Wl, cskip);
    fill_rand_bf16(skip.v, 1.0f);
    const int cin = cup + cskip;
    std::vector<float> w((size_t)cout * cin * 9);
    const float ws = 1.0f / sqrtf((float)cin * 9);
    for (auto& x : w) x = frand() * ws;
    __nv_bfloat16* d_low = to_dev_bf16(low.v);
    __nv_bfloat16* d_skip = to_dev_bf16(skip.v);
    float* d_w = to_dev_f32(w);
    const long long kt = 9 * cskip + 4 * cup;
    const size_t out_e = (size_t)N * 4 * Hl * Wl * cout;
    __nv_bfloat16 *d_wpk, *d_part, *d_out;
    CK(cudaMalloc(&d_wpk, 4 * cout * kt * 2));
    CK(cudaMalloc(&d_part, out_e * 2));
    CK(cudaMalloc(&d_out, out_e * 2));
    CK(cudaMemset(d_part, 0xFF, out_e * 2));
    CK(cudaMemset(d_out, 0xFF, out_e * 2));
    pack_dec1_w_kernel<<<64, 256>>>(d_w, d_wpk, cout, cup, cskip);
    WpconvLaunch L1;
    std::string e = wpconv_build(L1, d_low, cup, d_wpk, (int)kt, 9 * cskip, cout, N, Hl, Wl, d_part, nullptr, g_ctx->d_err,
                                 g_ctx->num_sms);
    EpilogueDesc ep;
    ep.relu = 1;
    ep.residual = nhwc_view(d_part, N, 2 * Hl, 2 * Wl, cout);
    WconvLaunch L2;
    if (e.empty())
        e = wconv_build(L2, d_skip, cskip, d_wpk, cout, N, 2 * Hl, 2 * Wl, d_out, ep, g_ctx->d_err, g_ctx->num_sms, kt);
    if (!e.empty()) {
        printf("[FAIL] %s: build: %s\n", name, e.c_str());
        g_fail++;
        return;
    }
    printf("       %s: wpconv grid %d bstages %d tiles %dx%dx%d n_tiles %d | wconv grid %d\n", name, L1.grid, L1.p.bstages,
           L1.p.tiles_w, L1.p.tiles_h, N, L1.p.n_tiles, L2.grid);
    CK(wpconv_launch(L1, 0));
    CK(wconv_launch(L2, 0));
    CK(cudaDeviceSynchronize());
    if (check_err_flag(name)) return;
    if (iters > 0) {
        cudaEvent_t e0, e1, e2;
        cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
        float m1 = 0, m2 = 0;
        for (int i = 0; i < iters; ++i) {
            cudaEventRecord(e0);
            CK(wpconv_launch(L1, 0));
            cudaEventRecord(e1);
            CK(wconv_launch(L2, 0));
            cudaEventRecord(e2);
            CK(cudaDeviceSynchronize());
            float a, b;
            cudaEventElapsedTime(&a, e0, e1);
            cudaEventElapsedTime(&b, e1, e2);
            m1 += a; m2 += b;
        }
        printf("[BENCH-D] %-34s wpconv %7.1f us + wconv(skip,+res) %7.1f us\n", name, m1 / iters * 1e3, m2 / iters * 1e3);
    } else {
        HostT cat(N, 2 * Hl, 2 * Wl, cin);
        for (int n = 0; n < N; ++n)
            for (int h = 0; h < 2 * Hl; ++h)
                for (int ww = 0; ww < 2 * Wl; ++ww) {
                    for (int c = 0; c < cup; ++c) cat.at(n, h, ww, c) = low.get(n, h / 2, ww / 2, c);
                    for (int c = 0; c < cskip; ++c) cat.at(n, h, ww, cup + c) = skip.get(n, h, ww, c);
                }
        HostT ref = cpu_conv(cat, w, cout, 3, 1, 1);
        for (auto& v : ref.v) v = fmaxf(v, 0.f);
        std::vector<float> got = from_dev_bf16(d_out, ref.v.size());
        report(name, compare(got, ref.v), 0.02, got, ref.v, cout);
    }
    cudaFree(d_low); cudaFree(d_skip); cudaFree(d_w); cudaFree(d_wpk); cudaFree(d_part); cudaFree(d_out);
}

// the same stem through tconv's overlapped-row (non-swizzled descriptor) mode; iters > 0: also time it
static void case_tstem(const char* name, int N, int H, int W, bool stats, int iters = 0) {
    std::vector<float> x((size_t)N * 3 * H * W);
    for (auto& v : x) v = frand() * 2.f;
    std::vector<float> w((size_t)64 * 3 * 49), wq(w.size());
    for (auto& v : w) v = frand() * 0.08f;
    for (size_t i = 0; i < w.size(); ++i) wq[i] = bf16r(w[i]);
    float* d_x = to_dev_f32(x);
    float* d_w = to_dev_f32(w);
    const size_t out_e = (size_t)N * (H / 2) * (W / 2) * 64;
    __nv_bfloat16 *d_xp, *d_wpk, *d_out;
    CK(cudaMalloc(&d_xp, (size_t)N * H * (W + 8) * 4 * 2 + 128));
    CK(cudaMemset(d_xp, 0xFF, (size_t)N * H * (W + 8) * 4 * 2 + 128));
    CK(cudaMalloc(&d_wpk, 64 * 224 * 2));
    CK(cudaMalloc(&d_out, out_e * 2));
    CK(cudaMemset(d_out, 0xFF, out_e * 2));
    pack_input_kernel<<<ew_grid((long long)N * H * ((W + 8) / 2), 256, 148), 256>>>(d_x, d_xp, N, H, W);
    {
        PackTable T;
        T.add(pk_entry(PK_STEM2, 0, 0, 64 * 224));
        CK(T.upload());
        CK(T.launch(d_w, d_wpk, 0));
        CK(cudaDeviceSynchronize());
    }
    EpilogueDesc ep;
    ep.relu = 1;
    float* d_stats = nullptr;
    if (stats) {
        CK(cudaMalloc(&d_stats, (size_t)g_ctx->num_sms * 64 * 2 * 4));
        CK(cudaMemset(d_stats, 0, (size_t)g_ctx->num_sms * 64 * 2 * 4));
        ep.stats = d_stats;
    }
    TconvLaunch L;
    std::string e = tconv_build_stem(L, d_xp, d_wpk, N, H, W, d_out, ep, g_ctx->d_err, g_ctx->num_sms);
    if (!e.empty()) {
        printf("[FAIL] %s: build: %s\n", name, e.c_str());
        g_fail++;
        return;
    }
    printf("       %s: grid %d smem %u stages %d tiles %dx%dx%d\n", name, L.grid, L.smem, L.p.stages, L.p.tiles_w, L.p.tiles_h, N);
    CK(tconv_launch(L, 0));
    CK(cudaDeviceSynchronize());
    if (check_err_flag(name)) return;
    if (iters > 0) {
        long long* d_prof;
        CK(cudaMalloc(&d_prof, (size_t)L.grid * 16 * 8));
        L.p.prof = d_prof;
        const int modes[] = {0, 8, 9, 12, 13, 15};
        for (int mode : modes) {
            L.p.dbg = mode;
            CK(cudaMemset(d_prof, 0, (size_t)L.grid * 16 * 8));
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            CK(tconv_launch(L, 0));
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            for (int i = 0; i < iters; ++i) CK(tconv_launch(L, 0));
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            ms /= iters;
            const double bytes = (double)N * H * (W + 8) * 8 + out_e * 2.0;
            const int tiles = L.p.tiles_w * L.p.tiles_h * N;
            printf("[BENCH-S] %-30s skip[%s%s%s] %8.1f us %7.1f GB/s (min traffic) %6.0f cyc/tile/CTA\n", name,
                   mode & 1 ? "L" : "-", mode & 2 ? "M" : "-", mode & 4 ? "E" : "-", ms * 1e3, bytes / ms * 1e-6,
                   ms * 1e-3 * 1.965e9 / ((double)tiles / L.grid));
            if (mode & 8) {
                std::vector<long long> hp((size_t)L.grid * 16);
                CK(cudaMemcpy(hp.data(), d_prof, hp.size() * 8, cudaMemcpyDeviceToHost));
                const double per = (double)tiles / L.grid;
                const char* nm[12] = {"P:wait_empty", "P:issue", "", "", "M:wait_tempty", "M:wait_full", "M:issue", "M:commit",
                                      "E:prefetch", "E:wait_tfull", "E:ld+arrive", "E:work"};
                printf("          cycles/tile (CTA 0):");
                for (int k = 0; k < 12; ++k)
                    if (nm[k][0]) printf(" %s=%.0f", nm[k], hp[k] / per);
                printf("\n");
            }
        }
        cudaFree(d_prof);
    } else {
        HostT in(N, H, W, 3);
        for (int n = 0; n < N; ++n)
            for (int c = 0; c < 3; ++c)
                for (int h = 0; h < H; ++h)
                    for (int ww = 0; ww < W; ++ww) in.at(n, h, ww, c) = bf16r(x[(((size_t)n * 3 + c) * H + h) * W + ww]);
        HostT ref = cpu_conv(in, wq, 64, 7, 2, 3);
        for (auto& v : ref.v) v = fmaxf(v, 0.f);
        std::vector<float> got = from_dev_bf16(d_out, ref.v.size());
        report(name, compare(got, ref.v), 0.012, got, ref.v, 64);
        if (stats) {
            std::vector<float> hs((size_t)L.grid * 128);
            CK(cudaMemcpy(hs.data(), d_stats, hs.size() * 4, cudaMemcpyDeviceToHost));
            std::vector<float> gs(128, 0.f), rs(128, 0.f);
            for (int b = 0; b < L.grid; ++b)
                for (int j = 0; j < 128; ++j) gs[j] += hs[(size_t)b * 128 + j];
            for (size_t i = 0; i < got.size(); ++i) {
                rs[(i % 64) * 2] += got[i];
                rs[(i % 64) * 2 + 1] += got[i] * got[i];
            }
            std::string nm = std::string(name) + " [stats]";
            report(nm.c_str(), compare(gs, rs), 1e-4, gs, rs, 2);
        }
    }
    cudaFree(d_x); cudaFree(d_w); cudaFree(d_xp); cudaFree(d_wpk); cudaFree(d_out);
    if (d_stats) cudaFree(d_stats);
}

static void case_dec1(const char* name, int N, int Hl, int Wl, int cup, int cskip, int cout) {
    HostT low(N, Hl, Wl, cup);
    fill_rand_bf16(low.v, 1.0f);
    HostT skip(N, 2 * Hl, 2 * Wl, cskip > 0 ? cskip : 1);
    fill_rand_bf16(skip.v, 1.0f);
    const int cin = cup + cskip;
    std::vector<float> w((size_t)cout * cin * 9);
    const float ws = 1.0f / sqrtf((float)cin * 9);
    for (auto& x : w) x = frand() * ws;
    // reference: materialise upsample + concat, full-precision weights (the fused kernel sums weights before rounding)
    HostT cat(N, 2 * Hl, 2 * Wl, cin);
    for (int n = 0; n < N; ++n)
        for (int h = 0; h < 2 * Hl; ++h)
            for (int ww = 0; ww < 2 * Wl; ++ww) {
                for (int c = 0; c < cup; ++c) cat.at(n, h, ww, c) = low.get(n, h / 2, ww / 2, c);
                for (int c = 0; c < cskip; ++c) cat.at(n, h, ww, cup + c) = skip.get(n, h, ww, c);
            }
    HostT ref = cpu_conv(cat, w, cout, 3, 1, 1);
    for (auto& v : ref.v) v = fmaxf(v, 0.f);
    __nv_bfloat16* d_low = to_dev_bf16(low.v);
    __nv_bfloat16* d_skip = to_dev_bf16(skip.v);
    float* d_w = to_dev_f32(w);
    const long long kt = 9 * cskip + 4 * cup;
    __nv_bfloat16 *d_wpk, *d_out;
    CK(cudaMalloc(&d_wpk, 4 * cout * kt * 2));
    CK(cudaMalloc(&d_out, ref.v.size() * 2));
    CK(cudaMemset(d_out, 0xFF, ref.v.size() * 2));
    pack_dec1_w_kernel<<<64, 256>>>(d_w, d_wpk, cout, cup, cskip);
    NetSpec::Dec d;
    d.c1 = d.c2 = -1;
    d.cup = cup; d.cskip = cskip; d.cout = cout;
    EpilogueDesc ep;
    ep.relu = 1;
    for (int par = 0; par < 4; ++par) {
        IgemmLaunch L;
        std::string e = build_dec1(g_ctx, L, d, d_wpk, par, d_low, cskip ? d_skip : nullptr, N, Hl, Wl, d_out, ep);
        if (!e.empty()) {
            printf("[FAIL] %s: build: %s\n", name, e.c_str());
            g_fail++;
            return;
        }
        CK(igemm_launch(L, 0));
    }
    CK(cudaDeviceSynchronize());
    if (check_err_flag(name)) return;
    std::vector<float> got = from_dev_bf16(d_out, ref.v.size());
    report(name, compare(got, ref.v), 0.02, got, ref.v, cout);
    cudaFree(d_low); cudaFree(d_skip); cudaFree(d_w); cudaFree(d_wpk); cudaFree(d_out);
}

static void case_maxpool(const char* name, int N, int H, int W, int C) {
    HostT in(N, H, W, C);
    fill_rand_bf16(in.v, 1.0f);
    HostT ref(N, H / 2, W / 2, C);
    for (int n = 0; n < N; ++n)
        for (int ho = 0; ho < H / 2; ++ho)
            for (int wo = 0; wo < W / 2; ++wo)
                for (int c = 0; c < C; ++c) {
                    float m = -INFINITY;
                    for (int r = 0; r < 3; ++r)
                        for (int s = 0; s < 3; ++s) {
                            const int h = 2 * ho - 1 + r, w = 2 * wo - 1 + s;
                            if (h >= 0 && h < H && w >= 0 && w < W) m = fmaxf(m, in.at(n, h, w, c));
                        }
                    ref.at(n, ho, wo, c) = m;
                }
    __nv_bfloat16* d_in = to_dev_bf16(in.v);
    __nv_bfloat16* d_out;
    CK(cudaMalloc(&d_out, ref.v.size() * 2));
    maxpool3x3s2_kernel<<<ew_grid((long long)ref.v.size() / 8, 256, 148), 256>>>(d_in, d_out, N, H, W, C);
    CK(cudaDeviceSynchronize());
    std::vector<float> got = from_dev_bf16(d_out, ref.v.size());
    report(name, compare(got, ref.v), 0.0, got, ref.v, C);
    cudaFree(d_in); cudaFree(d_out);
}

static void case_head(const char* name, int N, int H, int W) {
    HostT in(N, H, W, 16);
    fill_rand_bf16(in.v, 1.0f);
    std::vector<float> w(144), wb(145);
    for (auto& x : w) x = frand() * 0.1f;
    const float bias = 0.05f;
    HostT ref = cpu_conv(in, w, 1, 3, 1, 1);
    for (auto& v : ref.v) v += bias;
    __nv_bfloat16* d_in = to_dev_bf16(in.v);
    for (int i = 0; i < 144; ++i) wb[i] = w[i];
    wb[144] = bias;
    float* d_w = to_dev_f32(wb);
    float* d_out;
    uint8_t* d_mask;
    CK(cudaMalloc(&d_out, ref.v.size() * 4));
    CK(cudaMalloc(&d_mask, ref.v.size()));
    dim3 grid((W + kHeadTile - 1) / kHeadTile, (H + kHeadTile - 1) / kHeadTile, N);
    head_conv_kernel<<<grid, 256>>>(d_in, d_w, d_w + 144, d_out, nullptr, d_mask, 0.f, N, H, W);
    CK(cudaDeviceSynchronize());
    std::vector<float> got(ref.v.size());
    CK(cudaMemcpy(got.data(), d_out, got.size() * 4, cudaMemcpyDeviceToHost));
    report(name, compare(got, ref.v), 1e-4, got, ref.v, 1);
    std::vector<uint8_t> mk(ref.v.size());
    CK(cudaMemcpy(mk.data(), d_mask, mk.size(), cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < mk.size(); ++i) bad += (mk[i] != (got[i] >= 0.f ? 255 : 0));
    printf("[%s] %s [mask] mismatches %zu\n", bad ? "FAIL" : "PASS", name, bad);
    g_fail += bad ? 1 : 0;
    cudaFree(d_in); cudaFree(d_w); cudaFree(d_out); cudaFree(d_mask);
}

// hconv: 3x3 s1 conv over cat(nearest2x(low), src) on the halo-resident kernel, vs the CPU conv of the materialised input
static void case_hconv(const char* name, int N, int H, int W, int cup, int cskip, int cout, bool residual, bool relu,
                       bool scale_shift, bool stats) {
    const int ctot = cup + cskip;
    HostT low(N, H / 2, W / 2, cup ? cup : 1), src(N, H, W, cskip ? cskip : 1), cat(N, H, W, ctot);
    fill_rand_bf16(low.v, 1.0f);
    fill_rand_bf16(src.v, 1.0f);
    for (int n = 0; n < N; ++n)
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                for (int c = 0; c < cup; ++c) cat.at(n, h, w, c) = low.at(n, h / 2, w / 2, c);
                for (int c = 0; c < cskip; ++c) cat.at(n, h, w, cup + c) = src.at(n, h, w, c);
            }
    std::vector<float> w((size_t)cout * ctot * 9);
    const float ws = 1.0f / sqrtf((float)ctot * 9);
    for (auto& x : w) x = bf16r(frand() * ws * 1.7f);
    std::vector<float> sc(cout, 1.f), sh(cout, 0.f);
    if (scale_shift)
        for (int c = 0; c < cout; ++c) { sc[c] = 0.5f + 0.5f * fabsf(frand()); sh[c] = 0.1f * frand(); }
    HostT res(N, H, W, cout);
    if (residual) fill_rand_bf16(res.v, 1.0f);
    HostT ref = cpu_conv(cat, w, cout, 3, 1, 1);
    for (size_t i = 0; i < ref.v.size(); ++i) {
        float v = ref.v[i] * sc[i % cout] + sh[i % cout];
        if (residual) v += res.v[i];
        if (relu) v = fmaxf(v, 0.f);
        ref.v[i] = v;
    }
    __nv_bfloat16* d_low = to_dev_bf16(low.v);
    __nv_bfloat16* d_src = to_dev_bf16(src.v);
    __nv_bfloat16* d_res = to_dev_bf16(res.v);
    float* d_w = to_dev_f32(w);
    float* d_sc = to_dev_f32(sc);
    float* d_sh = to_dev_f32(sh);
    __nv_bfloat16 *d_wpk, *d_out;
    CK(cudaMalloc(&d_wpk, w.size() * 2));
    CK(cudaMalloc(&d_out, ref.v.size() * 2));
    CK(cudaMemset(d_out, 0xFF, ref.v.size() * 2));
    pack_hconv_w_kernel<<<64, 256>>>(d_w, d_wpk, cout, ctot, ctot, 0, 0);
    float* d_stats = nullptr;
    EpilogueDesc ep;
    if (scale_shift) { ep.scale = d_sc; ep.shift = d_sh; }
    ep.relu = relu;
    if (residual) ep.residual = nhwc_view(d_res, N, H, W, cout);
    if (stats) {
        CK(cudaMalloc(&d_stats, (size_t)g_ctx->num_sms * cout * 2 * 4));
        CK(cudaMemset(d_stats, 0, (size_t)g_ctx->num_sms * cout * 2 * 4));
        ep.stats = d_stats;
    }
    HconvLaunch L;
    std::string e = hconv_build(L, cup ? d_low : nullptr, cup, cskip ? d_src : nullptr, cskip, d_wpk, cout, N, H, W, d_out,
                                ep, g_ctx->d_err, g_ctx->num_sms);
    if (!e.empty()) {
        printf("[FAIL] %s: %s\n", name, e.c_str());
        g_fail++;
        return;
    }
    printf("       %s: grid %d smem %u stages %d tiles %dx%dx%d\n", name, L.grid, L.smem, L.p.stages, L.p.tiles_w,
           L.p.tiles_h, N);
    CK(hconv_launch(L, 0));
    CK(cudaDeviceSynchronize());
    if (!check_err_flag(name)) {
        std::vector<float> got = from_dev_bf16(d_out, ref.v.size());
        report(name, compare(got, ref.v), 6e-3, got, ref.v, cout);
        if (stats) {
            std::vector<float> hs((size_t)g_ctx->num_sms * cout * 2);
            CK(cudaMemcpy(hs.data(), d_stats, hs.size() * 4, cudaMemcpyDeviceToHost));
            std::vector<float> gs(cout * 2, 0.f), rs(cout * 2, 0.f);
            for (int b = 0; b < L.grid; ++b)
                for (int j = 0; j < cout * 2; ++j) gs[j] += hs[(size_t)b * cout * 2 + j];
            for (size_t i = 0; i < got.size(); ++i) {
                rs[(i % cout) * 2] += got[i];
                rs[(i % cout) * 2 + 1] += got[i] * got[i];
            }
            std::string nm = std::string(name) + " [stats]";
            report(nm.c_str(), compare(gs, rs), 1e-4, gs, rs, 2);
        }
    }
    cudaFree(d_low); cudaFree(d_src); cudaFree(d_res); cudaFree(d_w); cudaFree(d_sc); cudaFree(d_sh);
    cudaFree(d_wpk); cudaFree(d_out);
    if (d_stats) cudaFree(d_stats);
}

// tconv: TMA halo conv (plain: src at output resolution; parity: nearest-2x upsample of a low-res source folded into
// 2x2 taps) vs the CPU conv of the materialised input
static void case_tconv(const char* name, int N, int H, int W, int cin, int cout, bool parity, bool residual, bool relu,
                       bool scale_shift, bool stats) {
    HostT src(N, parity ? H / 2 : H, parity ? W / 2 : W, cin), cat(N, H, W, cin);
    fill_rand_bf16(src.v, 1.0f);
    for (int n = 0; n < N; ++n)
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w)
                for (int c = 0; c < cin; ++c) cat.at(n, h, w, c) = parity ? src.at(n, h / 2, w / 2, c) : src.at(n, h, w, c);
    std::vector<float> w((size_t)cout * cin * 9);
    const float ws = 1.0f / sqrtf((float)cin * 9);
    for (auto& x : w) x = bf16r(frand() * ws * 1.7f);
    std::vector<float> sc(cout, 1.f), sh(cout, 0.f);
    if (scale_shift)
        for (int c = 0; c < cout; ++c) { sc[c] = 0.5f + 0.5f * fabsf(frand()); sh[c] = 0.1f * frand(); }
    HostT res(N, H, W, cout);
    if (residual) fill_rand_bf16(res.v, 1.0f);
    HostT ref = cpu_conv(cat, w, cout, 3, 1, 1);
    for (size_t i = 0; i < ref.v.size(); ++i) {
        float v = ref.v[i] * sc[i % cout] + sh[i % cout];
        if (residual) v += res.v[i];
        if (relu) v = fmaxf(v, 0.f);
        ref.v[i] = v;
    }
    __nv_bfloat16* d_src = to_dev_bf16(src.v);
    __nv_bfloat16* d_res = to_dev_bf16(res.v);
    float* d_w = to_dev_f32(w);
    float* d_sc = to_dev_f32(sc);
    float* d_sh = to_dev_f32(sh);
    __nv_bfloat16 *d_wpk, *d_out;
    CK(cudaMalloc(&d_wpk, (size_t)tconv_w_elems(cin, cout, parity) * 2));
    CK(cudaMalloc(&d_out, ref.v.size() * 2));
    CK(cudaMemset(d_out, 0xFF, ref.v.size() * 2));
    {
        PackTable T;
        PackEntry e = pk_entry(parity ? PK_HPAR : PK_HCONV, 0, 0, tconv_w_elems(cin, cout, parity));
        e.cout = cout; e.cin = cin; e.a = cin; e.b = 0; e.c = 0;
        T.add(e);
        CK(T.upload());
        CK(T.launch(d_w, d_wpk, 0));
        CK(cudaDeviceSynchronize());
    }
    float* d_stats = nullptr;
    EpilogueDesc ep;
    if (scale_shift) { ep.scale = d_sc; ep.shift = d_sh; }
    ep.relu = relu;
    if (residual) ep.residual = nhwc_view(d_res, N, H, W, cout);
    if (stats) {
        CK(cudaMalloc(&d_stats, (size_t)2 * g_ctx->num_sms * cout * 2 * 4));
        CK(cudaMemset(d_stats, 0, (size_t)2 * g_ctx->num_sms * cout * 2 * 4));
        ep.stats = d_stats;
    }
    TconvLaunch L;
    std::string e = tconv_build(L, d_src, cin, parity, d_wpk, cout, N, H, W, d_out, ep, g_ctx->d_err, g_ctx->num_sms);
    if (!e.empty()) {
        printf("[FAIL] %s: %s\n", name, e.c_str());
        g_fail++;
        return;
    }
    printf("       %s: grid %d occ %d iph %d smem %u stages %d nacc %d nt %d tiles %dx%dx%d\n", name, L.grid, L.occ,
           L.iph, L.smem, L.p.stages, L.p.nacc, L.p.nt, L.p.tiles_w, L.p.tiles_h, N);
    CK(tconv_launch(L, 0));
    CK(cudaDeviceSynchronize());
    if (!check_err_flag(name)) {
        std::vector<float> got = from_dev_bf16(d_out, ref.v.size());
        report(name, compare(got, ref.v), parity ? 1.2e-2 : 6e-3, got, ref.v, cout);
        if (stats) {
            std::vector<float> hs((size_t)L.grid * cout * 2);
            CK(cudaMemcpy(hs.data(), d_stats, hs.size() * 4, cudaMemcpyDeviceToHost));
            std::vector<float> gs(cout * 2, 0.f), rs(cout * 2, 0.f);
            for (int b = 0; b < L.grid; ++b)
                for (int j = 0; j < cout * 2; ++j) gs[j] += hs[(size_t)b * cout * 2 + j];
            for (size_t i = 0; i < got.size(); ++i) {
                rs[(i % cout) * 2] += got[i];
                rs[(i % cout) * 2 + 1] += got[i] * got[i];
            }
            std::string nm = std::string(name) + " [stats]";
            report(nm.c_str(), compare(gs, rs), 1e-4, gs, rs, 2);
        }
    }
    cudaFree(d_src); cudaFree(d_res); cudaFree(d_w); cudaFree(d_sc); cudaFree(d_sh);
    cudaFree(d_wpk); cudaFree(d_out);
    if (d_stats) cudaFree(d_stats);
}

static void bench_tconv(const char* name, int N, int H, int W, int cin, int cout, bool parity, bool residual, int iters) {
    const size_t src_e = (size_t)N * (parity ? H / 2 : H) * (parity ? W / 2 : W) * cin, out_e = (size_t)N * H * W * cout;
    __nv_bfloat16 *d_src, *d_out, *d_wpk, *d_res = nullptr;
    CK(cudaMalloc(&d_src, src_e * 2));
    CK(cudaMalloc(&d_out, out_e * 2));
    CK(cudaMalloc(&d_wpk, (size_t)tconv_w_elems(cin, cout, parity) * 2));
    CK(cudaMemset(d_src, 0x3C, src_e * 2));
    CK(cudaMemset(d_wpk, 0x3C, (size_t)tconv_w_elems(cin, cout, parity) * 2));
    EpilogueDesc ep;
    ep.relu = 1;
    if (residual) {
        CK(cudaMalloc(&d_res, out_e * 2));
        CK(cudaMemset(d_res, 0x3C, out_e * 2));
        ep.residual = nhwc_view(d_res, N, H, W, cout);
    }
    TconvLaunch L;
    std::string e = tconv_build(L, d_src, cin, parity, d_wpk, cout, N, H, W, d_out, ep, g_ctx->d_err, g_ctx->num_sms);
    if (!e.empty()) {
        printf("[FAIL] bench %s: %s\n", name, e.c_str());
        return;
    }
    long long* d_prof;
    CK(cudaMalloc(&d_prof, (size_t)L.grid * 16 * 8));
    L.p.prof = d_prof;
    const int modes[] = {0, 8, 9, 10, 12, 13, 15};   // 10 = MMA issue skipped: the epilogue without operand-read contention
    for (int mode : modes) {
        L.p.dbg = mode;
        CK(cudaMemset(d_prof, 0, (size_t)L.grid * 16 * 8));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        for (int i = 0; i < 2; ++i) CK(tconv_launch(L, 0));
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for (int i = 0; i < iters; ++i) CK(tconv_launch(L, 0));
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        ms /= iters;
        const double flops = 2.0 * N * H * W * (double)cout * cin * 9;
        const double bytes = (src_e + out_e * (residual ? 2 : 1)) * 2.0;
        const int tiles = L.p.tiles_w * L.p.tiles_h * N;
        printf("[BENCH-T] %-30s skip[%s%s%s] %8.1f us %7.1f TFLOP/s(3x3-eq) %7.1f GB/s %6.0f cyc/tile/CTA grid %d occ %d iph %d stages %d nacc %d nt %d smem %u\n",
               name, mode & 1 ? "L" : "-", mode & 2 ? "M" : "-", mode & 4 ? "E" : "-", ms * 1e3, flops / ms * 1e-9,
               bytes / ms * 1e-6, ms * 1e-3 * 1.965e9 / ((double)tiles / L.grid), L.grid, L.occ, L.iph, L.p.stages, L.p.nacc,
               L.p.nt, L.smem);
        if (mode & 8) {
            std::vector<long long> hp((size_t)L.grid * 16);
            CK(cudaMemcpy(hp.data(), d_prof, hp.size() * 8, cudaMemcpyDeviceToHost));
            const double per = (double)tiles / L.grid;
            // E: prefetch (residual loads) | wait_tfull | - | math+smem writes+2nd TMEM load | first TMEM load | staging-free
            // barrier | fence + store barrier + TMA store issue
            const char* nm[15] = {"P:wait_empty", "P:issue", "", "", "M:wait_tempty", "M:wait_full", "M:issue", "M:commit",
                                  "E:prefetch", "E:wait_tfull", "E:-", "E:work", "E:ld0", "E:bar_free", "E:fence+bar+store"};
            printf("          cycles/tile (CTA 0):");
            for (int k = 0; k < 15; ++k)
                if (nm[k][0]) printf(" %s=%.0f", nm[k], hp[k] / per);
            printf("\n");
        }
    }
    cudaFree(d_prof);
    check_err_flag(name);
    cudaFree(d_src); cudaFree(d_out); cudaFree(d_wpk);
    if (d_res) cudaFree(d_res);
}

// wconv: wide halo conv with streamed weights vs the CPU conv
static void case_wconv(const char* name, int N, int H, int W, int cin, int cout, bool residual, bool relu, bool scale_shift,
                       bool stats, bool pair = false) {
    HostT src(N, H, W, cin);
    fill_rand_bf16(src.v, 1.0f);
    std::vector<float> w((size_t)cout * cin * 9);
    const float ws = 1.0f / sqrtf((float)cin * 9);
    for (auto& x : w) x = bf16r(frand() * ws * 1.7f);
    std::vector<float> sc(cout, 1.f), sh(cout, 0.f);
    if (scale_shift)
        for (int c = 0; c < cout; ++c) { sc[c] = 0.5f + 0.5f * fabsf(frand()); sh[c] = 0.1f * frand(); }
    HostT res(N, H, W, cout);
    if (residual) fill_rand_bf16(res.v, 1.0f);
    HostT ref = cpu_conv(src, w, cout, 3, 1, 1);
    for (size_t i = 0; i < ref.v.size(); ++i) {
        float v = ref.v[i] * sc[i % cout] + sh[i % cout];
        if (residual) v += res.v[i];
        if (relu) v = fmaxf(v, 0.f);
        ref.v[i] = v;
    }
    __nv_bfloat16* d_src = to_dev_bf16(src.v);
    __nv_bfloat16* d_res = to_dev_bf16(res.v);
    float* d_w = to_dev_f32(w);
    float* d_sc = to_dev_f32(sc);
    float* d_sh = to_dev_f32(sh);
    __nv_bfloat16 *d_wpk, *d_out;
    CK(cudaMalloc(&d_wpk, w.size() * 2));
    CK(cudaMalloc(&d_out, ref.v.size() * 2));
    CK(cudaMemset(d_out, 0xFF, ref.v.size() * 2));
    pack_conv_w_kernel<<<256, 256>>>(d_w, d_wpk, cout, cin, 3, 3, 0);
    float* d_stats = nullptr;
    EpilogueDesc ep;
    if (scale_shift) { ep.scale = d_sc; ep.shift = d_sh; }
    ep.relu = relu;
    if (residual) ep.residual = nhwc_view(d_res, N, H, W, cout);
    if (stats) {
        CK(cudaMalloc(&d_stats, (size_t)g_ctx->num_sms * cout * 2 * 4));
        CK(cudaMemset(d_stats, 0, (size_t)g_ctx->num_sms * cout * 2 * 4));
        ep.stats = d_stats;
    }
    WconvLaunch L;
    Wconv2Launch L2;
    std::string e = pair ? wconv2_build(L2, d_src, cin, d_wpk, cout, N, H, W, d_out, ep, g_ctx->d_err, g_ctx->num_sms)
                         : wconv_build(L, d_src, cin, d_wpk, cout, N, H, W, d_out, ep, g_ctx->d_err, g_ctx->num_sms);
    if (!e.empty()) {
        printf("[FAIL] %s: %s\n", name, e.c_str());
        g_fail++;
        return;
    }
    if (pair) {
        L.grid = L2.grid;
        printf("       %s: PAIR grid %d smem %u bstages %d kN %d tiles %dx%dx%d n_tiles %d\n", name, L2.grid, L2.smem,
               L2.p.bstages, L2.kn, L2.p.tiles_w, L2.p.tiles_h, N, L2.p.n_tiles);
        CK(wconv2_launch(L2, 0));
    } else {
        printf("       %s: grid %d smem %u bstages %d tiles %dx%dx%d n_tiles %d\n", name, L.grid, L.smem, L.p.bstages,
               L.p.tiles_w, L.p.tiles_h, N, L.p.n_tiles);
        CK(wconv_launch(L, 0));
    }
    CK(cudaDeviceSynchronize());
    if (!check_err_flag(name)) {
        std::vector<float> got = from_dev_bf16(d_out, ref.v.size());
        report(name, compare(got, ref.v), 6e-3, got, ref.v, cout);
        if (stats) {
            std::vector<float> hs((size_t)L.grid * cout * 2);
            CK(cudaMemcpy(hs.data(), d_stats, hs.size() * 4, cudaMemcpyDeviceToHost));
            std::vector<float> gs(cout * 2, 0.f), rs(cout * 2, 0.f);
            for (int b = 0; b < L.grid; ++b)
                for (int j = 0; j < cout * 2; ++j) gs[j] += hs[(size_t)b * cout * 2 + j];
            for (size_t i = 0; i < got.size(); ++i) {
                rs[(i % cout) * 2] += got[i];
                rs[(i % cout) * 2 + 1] += got[i] * got[i];
            }
            std::string nm = std::string(name) + " [stats]";
            report(nm.c_str(), compare(gs, rs), 1e-4, gs, rs, 2);
        }
    }
    cudaFree(d_src); cudaFree(d_res); cudaFree(d_w); cudaFree(d_sc); cudaFree(d_sh);
    cudaFree(d_wpk); cudaFree(d_out);
    if (d_stats) cudaFree(d_stats);
}

static void bench_wconv(const char* name, int N, int H, int W, int cin, int cout, int iters, bool pair = false) {
    const size_t in_e = (size_t)N * H * W * cin, out_e = (size_t)N * H * W * cout;
    __nv_bfloat16 *d_in, *d_out, *d_wpk;
    CK(cudaMalloc(&d_in, in_e * 2));
    CK(cudaMalloc(&d_out, out_e * 2));
    CK(cudaMalloc(&d_wpk, (size_t)cout * cin * 9 * 2));
    CK(cudaMemset(d_in, 0x3C, in_e * 2));
    CK(cudaMemset(d_wpk, 0x3C, (size_t)cout * cin * 9 * 2));
    EpilogueDesc ep;
    ep.relu = 1;
    WconvLaunch L;
    Wconv2Launch L2;
    std::string e = pair ? wconv2_build(L2, d_in, cin, d_wpk, cout, N, H, W, d_out, ep, g_ctx->d_err, g_ctx->num_sms)
                         : wconv_build(L, d_in, cin, d_wpk, cout, N, H, W, d_out, ep, g_ctx->d_err, g_ctx->num_sms);
    if (!e.empty()) {
        printf("[FAIL] bench %s: %s\n", name, e.c_str());
        return;
    }
    if (pair) { L.grid = L2.grid; L.p.bstages = L2.p.bstages; }
    auto go = [&]() { return pair ? wconv2_launch(L2, 0) : wconv_launch(L, 0); };
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) CK(go());
    CK(cudaDeviceSynchronize());
    if (check_err_flag(name)) return;
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) CK(go());
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= iters;
    const double flops = 2.0 * N * H * W * (double)cout * cin * 9;
    printf("[BENCH-WC] %-34s %8.1f us  %7.1f TFLOP/s  grid %d bstages %d\n", name, ms * 1e3, flops / ms * 1e-9, L.grid,
           L.p.bstages);
    check_err_flag(name);
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_wpk);
}

// hwgrad: dW of a 3x3 s1 conv over cat(nearest2x(low), src) given dZ, vs a CPU reference; output = packed [co][9][ctot]
static void case_hwgrad(const char* name, int N, int H, int W, int cup, int cskip, int cout) {
    const int ctot = cup + cskip;
    HostT low(N, H / 2, W / 2, cup ? cup : 1), src(N, H, W, cskip ? cskip : 1), cat(N, H, W, ctot), dz(N, H, W, cout);
    fill_rand_bf16(low.v, 1.0f);
    fill_rand_bf16(src.v, 1.0f);
    fill_rand_bf16(dz.v, 1.0f);
    for (int n = 0; n < N; ++n)
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                for (int c = 0; c < cup; ++c) cat.at(n, h, w, c) = low.at(n, h / 2, w / 2, c);
                for (int c = 0; c < cskip; ++c) cat.at(n, h, w, cup + c) = src.at(n, h, w, c);
            }
    std::vector<float> ref((size_t)cout * 9 * ctot, 0.f);
#pragma omp parallel for collapse(2)
    for (int co = 0; co < cout; ++co)
        for (int tap = 0; tap < 9; ++tap) {
            const int r = tap / 3, s = tap % 3;
            for (int ci = 0; ci < ctot; ++ci) {
                double acc = 0;
                for (int n = 0; n < N; ++n)
                    for (int h = 0; h < H; ++h)
                        for (int w = 0; w < W; ++w) acc += (double)dz.at(n, h, w, co) * cat.get(n, h + r - 1, w + s - 1, ci);
                ref[((size_t)co * 9 + tap) * ctot + ci] = (float)acc;
            }
        }
    __nv_bfloat16* d_low = to_dev_bf16(low.v);
    __nv_bfloat16* d_src = to_dev_bf16(src.v);
    __nv_bfloat16* d_dz = to_dev_bf16(dz.v);
    float* d_g;
    CK(cudaMalloc(&d_g, ref.size() * 4));
    CK(cudaMemset(d_g, 0, ref.size() * 4));
    HwgradLaunch L;
    std::string e = hwgrad_build(L, cup ? d_low : nullptr, cup, cskip ? d_src : nullptr, cskip, d_dz, cout, N, H, W, d_g,
                                 g_ctx->d_err, g_ctx->num_sms);
    if (!e.empty()) {
        printf("[FAIL] %s: %s\n", name, e.c_str());
        g_fail++;
        return;
    }
    printf("       %s: grid %dx%d smem %u stages %d cw %d ndh %d\n", name, L.grid.x, L.grid.y, L.smem, L.p.stages, L.p.cw,
           L.p.ndh_max);
    CK(hwgrad_launch(L, 0));
    CK(cudaDeviceSynchronize());
    if (!check_err_flag(name)) {
        std::vector<float> got(ref.size());
        CK(cudaMemcpy(got.data(), d_g, got.size() * 4, cudaMemcpyDeviceToHost));
        report(name, compare(got, ref), 2e-3, got, ref, ctot);
    }
    cudaFree(d_low); cudaFree(d_src); cudaFree(d_dz); cudaFree(d_g);
}
// xwgrad: dW of a 3x3 s1 conv given dZ and X (TMA halo boxes, all taps per pass), written into columns [dci0, dci0+cin)
// of a packed [co][9][ctot] gradient (the other columns must stay untouched)
static void case_xwgrad(const char* name, int N, int H, int W, int cin, int cout, int cextra) {
    const int ctot = cextra + cin, dci0 = cextra;
    HostT src(N, H, W, cin), dz(N, H, W, cout);
    fill_rand_bf16(src.v, 1.0f);
    fill_rand_bf16(dz.v, 1.0f);
    std::vector<float> ref((size_t)cout * 9 * ctot, 0.f);
#pragma omp parallel for collapse(2)
    for (int co = 0; co < cout; ++co)
        for (int tap = 0; tap < 9; ++tap) {
            const int r = tap / 3, s = tap % 3;
            for (int ci = 0; ci < cin; ++ci) {
                double acc = 0;
                for (int n = 0; n < N; ++n)
                    for (int h = 0; h < H; ++h)
                        for (int w = 0; w < W; ++w) acc += (double)dz.at(n, h, w, co) * src.get(n, h + r - 1, w + s - 1, ci);
                ref[((size_t)co * 9 + tap) * ctot + dci0 + ci] = (float)acc;
            }
        }
    __nv_bfloat16* d_src = to_dev_bf16(src.v);
    __nv_bfloat16* d_dz = to_dev_bf16(dz.v);
    float* d_g;
    CK(cudaMalloc(&d_g, ref.size() * 4));
    CK(cudaMemset(d_g, 0, ref.size() * 4));
    XwgradLaunch L;
    std::string e = xwgrad_build(L, d_src, cin, d_dz, cout, N, H, W, d_g, ctot, dci0, g_ctx->d_err, g_ctx->num_sms);
    if (!e.empty()) {
        printf("[FAIL] %s: %s\n", name, e.c_str());
        g_fail++;
        return;
    }
    printf("       %s: grid %dx%d smem %u stages %d th %d co_blk %d cw %d wide %d\n", name, L.grid.x, L.grid.y, L.smem,
           L.p.stages, L.p.th, L.p.co_blk, L.p.cw, L.p.wide);
    CK(xwgrad_launch(L, 0));
    CK(cudaDeviceSynchronize());
    if (!check_err_flag(name)) {
        std::vector<float> got(ref.size());
        CK(cudaMemcpy(got.data(), d_g, got.size() * 4, cudaMemcpyDeviceToHost));
        report(name, compare(got, ref), 2e-3, got, ref, ctot);
    }
    cudaFree(d_src); cudaFree(d_dz); cudaFree(d_g);
}
static void bench_xwgrad(const char* name, int N, int H, int W, int cin, int cout, int iters) {
    __nv_bfloat16 *d_src, *d_dz;
    float* d_g;
    CK(cudaMalloc(&d_src, (size_t)N * H * W * cin * 2));
    CK(cudaMalloc(&d_dz, (size_t)N * H * W * cout * 2));
    CK(cudaMalloc(&d_g, (size_t)cout * 9 * cin * 4));
    CK(cudaMemset(d_src, 0x3C, (size_t)N * H * W * cin * 2));
    CK(cudaMemset(d_dz, 0x3C, (size_t)N * H * W * cout * 2));
    XwgradLaunch L;
    std::string e = xwgrad_build(L, d_src, cin, d_dz, cout, N, H, W, d_g, cin, 0, g_ctx->d_err, g_ctx->num_sms);
    if (!e.empty()) { printf("[FAIL] bench %s: %s\n", name, e.c_str()); return; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) CK(xwgrad_launch(L, 0));
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) CK(xwgrad_launch(L, 0));
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= iters;
    const double flops = 2.0 * N * H * W * (double)cout * cin * 9;
    const double bytes = 2.0 * N * H * W * (double)(cout + cin);
    printf("[BENCH-XW] %-30s %8.1f us %7.1f TFLOP/s %6.2f TB/s  grid %dx%d stages %d th %d smem %u\n", name, ms * 1e3,
           flops / ms * 1e-9, bytes / ms * 1e-9, L.grid.x, L.grid.y, L.p.stages, L.p.th, L.smem);
    check_err_flag(name);
    cudaFree(d_src); cudaFree(d_dz); cudaFree(d_g);
}
// xwgrad up mode: dW of a 3x3 s1 conv over nearest2x(low) given the high-resolution dZ
static void case_xwgrad_up(const char* name, int N, int H, int W, int cup, int cout, int cextra) {
    const int ctot = cup + cextra, dci0 = 0;
    HostT low(N, H / 2, W / 2, cup), upx(N, H, W, cup), dz(N, H, W, cout);
    fill_rand_bf16(low.v, 1.0f);
    fill_rand_bf16(dz.v, 1.0f);
    for (int n = 0; n < N; ++n)
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w)
                for (int c = 0; c < cup; ++c) upx.at(n, h, w, c) = low.at(n, h / 2, w / 2, c);
    std::vector<float> ref((size_t)cout * 9 * ctot, 0.f);
#pragma omp parallel for collapse(2)
    for (int co = 0; co < cout; ++co)
        for (int tap = 0; tap < 9; ++tap) {
            const int r = tap / 3, s = tap % 3;
            for (int ci = 0; ci < cup; ++ci) {
                double acc = 0;
                for (int n = 0; n < N; ++n)
                    for (int h = 0; h < H; ++h)
                        for (int w = 0; w < W; ++w) acc += (double)dz.at(n, h, w, co) * upx.get(n, h + r - 1, w + s - 1, ci);
                ref[((size_t)co * 9 + tap) * ctot + dci0 + ci] = (float)acc;
            }
        }
    __nv_bfloat16* d_low = to_dev_bf16(low.v);
    __nv_bfloat16* d_dz = to_dev_bf16(dz.v);
    float* d_g;
    CK(cudaMalloc(&d_g, ref.size() * 4));
    CK(cudaMemset(d_g, 0, ref.size() * 4));
    XwgradLaunch L;
    std::string e = xwgrad_build(L, d_low, cup, d_dz, cout, N, H / 2, W / 2, d_g, ctot, dci0, g_ctx->d_err, g_ctx->num_sms, true);
    if (!e.empty()) {
        printf("[FAIL] %s: %s\n", name, e.c_str());
        g_fail++;
        return;
    }
    printf("       %s: grid %dx%dx%d smem %u stages %d th %d co_blk %d cw %d wide %d\n", name, L.grid.x, L.grid.y, L.grid.z,
           L.smem, L.p.stages, L.p.th, L.p.co_blk, L.p.cw, L.p.wide);
    CK(xwgrad_launch(L, 0));
    CK(cudaDeviceSynchronize());
    if (!check_err_flag(name)) {
        std::vector<float> got(ref.size());
        CK(cudaMemcpy(got.data(), d_g, got.size() * 4, cudaMemcpyDeviceToHost));
        report(name, compare(got, ref), 2e-3, got, ref, ctot);
    }
    cudaFree(d_low); cudaFree(d_dz); cudaFree(d_g);
}
static void bench_xwgrad_up(const char* name, int N, int H, int W, int cup, int cout, int iters) {
    __nv_bfloat16 *d_low, *d_dz;
    float* d_g;
    CK(cudaMalloc(&d_low, (size_t)N * (H / 2) * (W / 2) * cup * 2));
    CK(cudaMalloc(&d_dz, (size_t)N * H * W * cout * 2));
    CK(cudaMalloc(&d_g, (size_t)cout * 9 * cup * 4));
    CK(cudaMemset(d_low, 0x3C, (size_t)N * (H / 2) * (W / 2) * cup * 2));
    CK(cudaMemset(d_dz, 0x3C, (size_t)N * H * W * cout * 2));
    XwgradLaunch L;
    std::string e = xwgrad_build(L, d_low, cup, d_dz, cout, N, H / 2, W / 2, d_g, cup, 0, g_ctx->d_err, g_ctx->num_sms, true);
    if (!e.empty()) { printf("[FAIL] bench %s: %s\n", name, e.c_str()); return; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) CK(xwgrad_launch(L, 0));
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) CK(xwgrad_launch(L, 0));
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= iters;
    const double bytes = 2.0 * N * H * W * (double)cout + 2.0 * N * (H / 2) * (W / 2) * (double)cup;
    printf("[BENCH-XW] %-30s %8.1f us %6.2f TB/s  grid %dx%dx%d stages %d th %d smem %u\n", name, ms * 1e3,
           bytes / ms * 1e-9, L.grid.x, L.grid.y, L.grid.z, L.p.stages, L.p.th, L.smem);
    check_err_flag(name);
    cudaFree(d_low); cudaFree(d_dz); cudaFree(d_g);
}
// dlow: gradient of a decoder conv1 w.r.t. its low-resolution input.  Host builds the PK_DLOW operand V[ci][t*cz + co]
// (3x3 weights pre-summed per high-res offset, rounded to bf16 like the device pack) and the reference
// dLow[u, ci] = sum_t dZ[2u + (oh_t, ow_t), co] * V[ci][t][co].
static void case_dlow(const char* name, int N, int Hl, int Wl, int cz, int cup, int iters = 0) {
    const int H = 2 * Hl, W = 2 * Wl;
    HostT dz(N, H, W, cz);
    if (iters == 0) fill_rand_bf16(dz.v, 1.0f);
    std::vector<float> w((size_t)cz * cup * 9);          // OIHW: [co = cz][ci = cup][3][3]
    for (auto& x : w) x = frand() * 0.2f;
    std::vector<float> V((size_t)cup * 16 * cz);
    int oh_t[16], ow_t[16];
    for (int t = 0; t < 16; ++t) {
        const int bb = t & 1, pw = (t >> 1) & 1, aa = (t >> 2) & 1, ph = (t >> 3) & 1;
        oh_t[t] = 2 - 2 * aa - ph;
        ow_t[t] = 2 - 2 * bb - pw;
        const int r0 = ph == 0 ? (aa == 0 ? 0 : 1) : (aa == 0 ? 0 : 2), r1 = ph == 0 ? (aa == 0 ? 0 : 2) : (aa == 0 ? 1 : 2);
        const int s0 = pw == 0 ? (bb == 0 ? 0 : 1) : (bb == 0 ? 0 : 2), s1 = pw == 0 ? (bb == 0 ? 0 : 2) : (bb == 0 ? 1 : 2);
        for (int c = 0; c < cup; ++c)
            for (int co = 0; co < cz; ++co) {
                float v = 0.f;
                for (int r = r0; r <= r1; ++r)
                    for (int q = s0; q <= s1; ++q) v += w[((size_t)co * cup + c) * 9 + r * 3 + q];
                V[((size_t)c * 16 + t) * cz + co] = bf16r(v);
            }
    }
    __nv_bfloat16* d_dz = to_dev_bf16(dz.v);
    __nv_bfloat16* d_v = to_dev_bf16(V);
    __nv_bfloat16* d_out;
    const size_t nout = (size_t)N * Hl * Wl * cup;
    CK(cudaMalloc(&d_out, nout * 2));
    CK(cudaMemset(d_out, 0, nout * 2));
    DlowLaunch L;
    std::string e = dlow_build(L, d_dz, cz, d_v, cup, N, Hl, Wl, d_out, g_ctx->d_err, g_ctx->num_sms);
    if (!e.empty()) {
        printf("[FAIL] %s: %s\n", name, e.c_str());
        g_fail++;
        return;
    }
    if (iters > 0) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        for (int i = 0; i < 2; ++i) CK(dlow_launch(L, 0));
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for (int i = 0; i < iters; ++i) CK(dlow_launch(L, 0));
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        ms /= iters;
        const double bytes = 2.0 * N * H * W * (double)cz + 2.0 * nout;
        printf("[BENCH-DL] %-30s %8.1f us %6.2f TB/s  grid %d stages %d smem %u\n", name, ms * 1e3, bytes / ms * 1e-9,
               L.grid, L.p.stages, L.smem);
        check_err_flag(name);
        cudaFree(d_dz); cudaFree(d_v); cudaFree(d_out);
        return;
    }
    printf("       %s: grid %d smem %u stages %d\n", name, L.grid, L.smem, L.p.stages);
    CK(dlow_launch(L, 0));
    CK(cudaDeviceSynchronize());
    if (!check_err_flag(name)) {
        std::vector<float> ref(nout, 0.f);
#pragma omp parallel for collapse(2)
        for (int n = 0; n < N; ++n)
            for (int h = 0; h < Hl; ++h)
                for (int x = 0; x < Wl; ++x)
                    for (int c = 0; c < cup; ++c) {
                        double acc = 0;
                        for (int t = 0; t < 16; ++t)
                            for (int co = 0; co < cz; ++co)
                                acc += (double)dz.get(n, 2 * h + oh_t[t], 2 * x + ow_t[t], co) * V[((size_t)c * 16 + t) * cz + co];
                        ref[(((size_t)n * Hl + h) * Wl + x) * cup + c] = (float)acc;
                    }
        std::vector<__nv_bfloat16> hb(nout);
        CK(cudaMemcpy(hb.data(), d_out, nout * 2, cudaMemcpyDeviceToHost));
        std::vector<float> got(nout);
        for (size_t i = 0; i < nout; ++i) got[i] = __bfloat162float(hb[i]);
        report(name, compare(got, ref), 6e-3, got, ref, cup);
    }
    cudaFree(d_dz); cudaFree(d_v); cudaFree(d_out);
}
// swgrad: weight gradient of the 7x7/s2 stem from the packed image and dZ, vs a CPU reference (x rounded to bf16 as packed)
static void case_swgrad(const char* name, int N, int H, int W, int iters = 0) {
    const int Ho = H / 2, Wo = W / 2;
    std::vector<float> x((size_t)N * 3 * H * W);
    for (auto& v : x) v = bf16r(frand() * 2.f);
    HostT dz(N, Ho, Wo, 64);
    if (iters == 0) fill_rand_bf16(dz.v, 1.0f);
    float* d_x = to_dev_f32(x);
    __nv_bfloat16* d_xp;
    CK(cudaMalloc(&d_xp, (size_t)N * H * (W + 8) * 4 * 2 + 128));
    CK(cudaMemset(d_xp, 0, (size_t)N * H * (W + 8) * 4 * 2 + 128));
    pack_input_kernel<<<ew_grid((long long)N * H * ((W + 8) / 2), 256, 148), 256>>>(d_x, d_xp, N, H, W);
    __nv_bfloat16* d_dz = to_dev_bf16(dz.v);
    float* d_g;
    CK(cudaMalloc(&d_g, 64 * 147 * 4));
    CK(cudaMemset(d_g, 0, 64 * 147 * 4));
    SwgradLaunch L;
    std::string e = swgrad_build(L, d_xp, d_dz, N, H, W, d_g, g_ctx->d_err, g_ctx->num_sms);
    if (!e.empty()) {
        printf("[FAIL] %s: %s\n", name, e.c_str());
        g_fail++;
        return;
    }
    if (iters > 0) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        for (int i = 0; i < 2; ++i) CK(swgrad_launch(L, 0));
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for (int i = 0; i < iters; ++i) CK(swgrad_launch(L, 0));
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        ms /= iters;
        const double bytes = 2.0 * N * Ho * Wo * 64.0 + 2.0 * N * H * (W + 8) * 4.0;
        printf("[BENCH-SW] %-30s %8.1f us %6.2f TB/s  grid %d stages %d wt %d smem %u\n", name, ms * 1e3, bytes / ms * 1e-9,
               L.grid, L.p.stages, L.p.wt, L.smem);
        check_err_flag(name);
        cudaFree(d_x); cudaFree(d_xp); cudaFree(d_dz); cudaFree(d_g);
        return;
    }
    printf("       %s: grid %d smem %u stages %d wt %d\n", name, L.grid, L.smem, L.p.stages, L.p.wt);
    CK(swgrad_launch(L, 0));
    CK(cudaDeviceSynchronize());
    if (!check_err_flag(name)) {
        std::vector<float> ref((size_t)64 * 147, 0.f);
#pragma omp parallel for collapse(2)
        for (int co = 0; co < 64; ++co)
            for (int ci = 0; ci < 3; ++ci)
                for (int r = 0; r < 7; ++r)
                    for (int q = 0; q < 7; ++q) {
                        double acc = 0;
                        for (int n = 0; n < N; ++n)
                            for (int ho = 0; ho < Ho; ++ho) {
                                const int hi = 2 * ho + r - 3;
                                if (hi < 0 || hi >= H) continue;
                                for (int wo = 0; wo < Wo; ++wo) {
                                    const int wi = 2 * wo + q - 3;
                                    if (wi < 0 || wi >= W) continue;
                                    acc += (double)dz.at(n, ho, wo, co) * x[(((size_t)n * 3 + ci) * H + hi) * W + wi];
                                }
                            }
                        ref[(size_t)co * 147 + ci * 49 + r * 7 + q] = (float)acc;
                    }
        std::vector<float> got(ref.size());
        CK(cudaMemcpy(got.data(), d_g, got.size() * 4, cudaMemcpyDeviceToHost));
        report(name, compare(got, ref), 2e-3, got, ref, 147);
    }
    cudaFree(d_x); cudaFree(d_xp); cudaFree(d_dz); cudaFree(d_g);
}
static void bench_hwgrad(const char* name, int N, int H, int W, int cup, int cskip, int cout, int iters) {
    const int ctot = cup + cskip;
    __nv_bfloat16 *d_low, *d_src, *d_dz;
    float* d_g;
    CK(cudaMalloc(&d_low, (size_t)N * (H / 2) * (W / 2) * (cup ? cup : 8) * 2));
    CK(cudaMalloc(&d_src, (size_t)N * H * W * (cskip ? cskip : 8) * 2));
    CK(cudaMalloc(&d_dz, (size_t)N * H * W * cout * 2));
    CK(cudaMalloc(&d_g, (size_t)cout * 9 * ctot * 4));
    CK(cudaMemset(d_low, 0x3C, (size_t)N * (H / 2) * (W / 2) * (cup ? cup : 8) * 2));
    CK(cudaMemset(d_src, 0x3C, (size_t)N * H * W * (cskip ? cskip : 8) * 2));
    CK(cudaMemset(d_dz, 0x3C, (size_t)N * H * W * cout * 2));
    HwgradLaunch L;
    std::string e = hwgrad_build(L, cup ? d_low : nullptr, cup, cskip ? d_src : nullptr, cskip, d_dz, cout, N, H, W, d_g,
                                 g_ctx->d_err, g_ctx->num_sms);
    if (!e.empty()) { printf("[FAIL] bench %s: %s\n", name, e.c_str()); return; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) CK(hwgrad_launch(L, 0));
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) CK(hwgrad_launch(L, 0));
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= iters;
    const double flops = 2.0 * N * H * W * (double)cout * ctot * 9;
    printf("[BENCH-W] %-30s %8.1f us %7.1f TFLOP/s  grid %dx%d stages %d smem %u\n", name, ms * 1e3, flops / ms * 1e-9,
           L.grid.x, L.grid.y, L.p.stages, L.smem);
    check_err_flag(name);
    cudaFree(d_low); cudaFree(d_src); cudaFree(d_dz); cudaFree(d_g);
}

// ------------------------------------------------------------------------------------------------ timing
static void bench_conv(const char* name, int N, int H, int W, int cin, int cout, int k, int stride, int iters) {
    const size_t in_e = (size_t)N * H * W * cin, out_e = (size_t)N * (H / stride) * (W / stride) * cout;
    __nv_bfloat16 *d_in, *d_out, *d_wpk;
    CK(cudaMalloc(&d_in, in_e * 2));
    CK(cudaMalloc(&d_out, out_e * 2));
    CK(cudaMalloc(&d_wpk, (size_t)cout * cin * k * k * 2));
    CK(cudaMemset(d_in, 0x3C, in_e * 2));  // 0x3C3C ~ 0.0115
    CK(cudaMemset(d_wpk, 0x3C, (size_t)cout * cin * k * k * 2));
    ConvRef c;
    c.cin = cin; c.cout = cout; c.k = k; c.stride = stride;
    EpilogueDesc ep;
    ep.relu = 1;
    IgemmLaunch L;
    std::string e = build_conv(g_ctx, L, c, d_wpk, d_in, N, H, W, d_out, ep);
    if (!e.empty()) {
        printf("[FAIL] bench %s: %s\n", name, e.c_str());
        return;
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) CK(igemm_launch(L, 0));
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) CK(igemm_launch(L, 0));
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= iters;
    const double flops = 2.0 * N * (H / stride) * (W / stride) * (double)cout * cin * k * k;
    const double bytes = (in_e + out_e) * 2.0;
    printf("[BENCH] %-34s %8.1f us  %7.1f TFLOP/s  %7.1f GB/s(min traffic)  grid %d stages %d\n", name, ms * 1e3,
           flops / ms * 1e-9, bytes / ms * 1e-6, L.grid, L.p.stages);
    check_err_flag(name);
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_wpk);
}

static void bench_hconv(const char* name, int N, int H, int W, int cup, int cskip, int cout, int iters) {
    const int ctot = cup + cskip;
    const size_t low_e = (size_t)N * (H / 2) * (W / 2) * (cup ? cup : 8), src_e = (size_t)N * H * W * (cskip ? cskip : 8);
    const size_t out_e = (size_t)N * H * W * cout;
    __nv_bfloat16 *d_low, *d_src, *d_out, *d_wpk;
    CK(cudaMalloc(&d_low, low_e * 2));
    CK(cudaMalloc(&d_src, src_e * 2));
    CK(cudaMalloc(&d_out, out_e * 2));
    CK(cudaMalloc(&d_wpk, (size_t)9 * ctot * cout * 2));
    CK(cudaMemset(d_low, 0x3C, low_e * 2));
    CK(cudaMemset(d_src, 0x3C, src_e * 2));
    CK(cudaMemset(d_wpk, 0x3C, (size_t)9 * ctot * cout * 2));
    EpilogueDesc ep;
    ep.relu = 1;
    HconvLaunch L;
    std::string e = hconv_build(L, cup ? d_low : nullptr, cup, cskip ? d_src : nullptr, cskip, d_wpk, cout, N, H, W, d_out,
                                ep, g_ctx->d_err, g_ctx->num_sms);
    if (!e.empty()) {
        printf("[FAIL] bench %s: %s\n", name, e.c_str());
        return;
    }
    long long* d_prof;
    CK(cudaMalloc(&d_prof, (size_t)L.grid * 16 * 8));
    L.p.prof = d_prof;
    const int modes[] = {0, 8, 13, 15, 15 + 16, 15 + 32, 15 + 48};
    for (int mode : modes) {
        L.p.dbg = mode;
        CK(cudaMemset(d_prof, 0, (size_t)L.grid * 16 * 8));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        for (int i = 0; i < 2; ++i) CK(hconv_launch(L, 0));
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for (int i = 0; i < iters; ++i) CK(hconv_launch(L, 0));
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        ms /= iters;
        const double flops = 2.0 * N * H * W * (double)cout * ctot * 9;
        const double bytes = ((cup ? low_e : 0) + (cskip ? src_e : 0) + out_e) * 2.0;
        const int tiles = L.p.tiles_w * L.p.tiles_h * N;
        printf("[BENCH-H] %-30s skip[%s%s%s] %8.1f us %7.1f TFLOP/s %7.1f GB/s  %6.0f cyc/tile  grid %d stages %d smem %u\n",
               name, mode & 1 ? "L" : "-", mode & 2 ? "M" : "-", mode & 4 ? "E" : "-", ms * 1e3, flops / ms * 1e-9,
               bytes / ms * 1e-6, ms * 1e-3 * 1.965e9 / ((double)tiles / L.grid) / L.occ, L.grid, L.p.stages, L.smem);
        if (mode & 8) {
            std::vector<long long> hp((size_t)L.grid * 16);
            CK(cudaMemcpy(hp.data(), d_prof, hp.size() * 8, cudaMemcpyDeviceToHost));
            const double per = (double)tiles / L.grid;  // counters hold the last launch only
            const char* nm[10] = {"P:wait_empty", "P:issue", "P:wait_group", "P:fence+arrive", "M:wait_tempty", "M:wait_full",
                                  "M:issue", "M:commit", "E:wait_tfull", "E:work"};
            printf("          cycles/tile (CTA 0):");
            for (int k = 0; k < 10; ++k) printf(" %s=%.0f", nm[k], hp[k] / per);
            printf("\n");
        }
    }
    cudaFree(d_prof);
    check_err_flag(name);
    cudaFree(d_low); cudaFree(d_src); cudaFree(d_out); cudaFree(d_wpk);
}

int main(int argc, char** argv) {
    const char* filt = argc > 1 ? argv[1] : "";
    auto want = [&](const char* n) { return strstr(n, filt) != nullptr; };
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device: %s sm_%d%d, %d SMs, smem/block optin %zu\n", prop.name, prop.major, prop.minor,
           prop.multiProcessorCount, prop.sharedMemPerBlockOptin);
    g_ctx = new Ctx();
    g_ctx->num_sms = prop.multiProcessorCount;
    CK(cudaMalloc(&g_ctx->d_err, 4));
    CK(cudaMemset(g_ctx->d_err, 0, 4));

    if (want("pool")) case_maxpool("maxpool 2x32x32x64", 2, 32, 32, 64);
    if (want("head")) case_head("head 2x40x48", 2, 40, 48);
    // ---- igemm basic: smallest possible pipeline first
    if (want("conv")) {
        case_conv("conv 1x1 s1 64->64 1x8x16 (1 chunk)", 1, 8, 16, 64, 64, 1, 1, false, false, false, false);
        case_conv("conv 3x3 s1 64->64 2x32x32", 2, 32, 32, 64, 64, 3, 1, false, false, false, false);
        case_conv("conv 3x3 s1 64->64 +ss+relu", 2, 32, 32, 64, 64, 3, 1, false, true, true, false);
        case_conv("conv 3x3 s1 64->64 +res+relu", 2, 32, 32, 64, 64, 3, 1, true, true, true, false);
        case_conv("conv 3x3 s1 64->64 +stats", 2, 32, 32, 64, 64, 3, 1, false, false, false, true);
        case_conv("conv 3x3 s1 128->128 1x16x16", 1, 16, 16, 128, 128, 3, 1, false, true, true, false);
        case_conv("conv 3x3 s1 256->256 2x16x16", 2, 16, 16, 256, 256, 3, 1, true, true, true, true);
        case_conv("conv 3x3 s1 512->512 2x16x16", 2, 16, 16, 512, 512, 3, 1, true, true, true, false);
        case_conv("conv 3x3 s1 32->32 1x32x32 (SW64)", 1, 32, 32, 32, 32, 3, 1, false, true, true, true);
        case_conv("conv 3x3 s1 16->16 1x32x32 (SW32)", 1, 32, 32, 16, 16, 3, 1, false, true, true, true);
        case_conv("conv 3x3 s2 64->128 2x32x32", 2, 32, 32, 64, 128, 3, 2, false, true, true, false);
        case_conv("conv 1x1 s2 64->128 2x32x32", 2, 32, 32, 64, 128, 1, 2, false, false, true, false);
        case_conv("conv 3x3 s1 64->64 3x8x8 (bn=2,partial)", 3, 8, 8, 64, 64, 3, 1, true, true, true, true);
        case_conv("conv 3x3 s1 64->64 1x24x40 (partial)", 1, 24, 40, 64, 64, 3, 1, false, true, true, true);
        case_conv("conv 3x3 s1 64->64 5x64x64 (multi-tile/CTA)", 5, 64, 64, 64, 64, 3, 1, true, true, true, true);
        case_conv("conv 3x3 s1 256->512 9x16x16", 9, 16, 16, 256, 512, 3, 1, false, true, true, false);
    }
    if (want("stem")) {
        case_stem("stem 7x7 s2 3->64 2x64x64", 2, 64, 64);
        case_stem("stem 7x7 s2 3->64 1x96x160", 1, 96, 160);
    }
    if (want("tstem")) {
        case_tstem("tstem 7x7 s2 3->64 2x64x64", 2, 64, 64, true);
        case_tstem("tstem 7x7 s2 3->64 1x96x160", 1, 96, 160, true);
        case_tstem("tstem 7x7 s2 3->64 3x512x512", 3, 512, 512, false);
    }
    if (want("sbench")) case_tstem("tstem 7x7 s2 3->64 x32 @512^2", 32, 512, 512, false, 10);
    if (want("dec1")) {
        case_dec1("dec1 up64+skip64->32 2x16x16", 2, 16, 16, 64, 64, 32);
        case_dec1("dec1 up128+skip64->64 1x16x16", 1, 16, 16, 128, 64, 64);
        case_dec1("dec1 up512+skip256->256 1x8x8", 1, 8, 8, 512, 256, 256);
        case_dec1("dec1 up32->16 (no skip) 2x16x32", 2, 16, 32, 32, 0, 16);
    }
    if (want("hconv")) {
        case_hconv("hconv 16->16 2x32x32", 2, 32, 32, 0, 16, 16, false, true, true, true);
        case_hconv("hconv 32->32 1x48x40 (partial tiles)", 1, 48, 40, 0, 32, 32, false, true, true, true);
        case_hconv("hconv 64->64 +res 3x24x24 (partial)", 3, 24, 24, 0, 64, 64, true, true, true, true);
        case_hconv("hconv 64->64 raw 5x64x64 (multi-tile/CTA)", 5, 64, 64, 0, 64, 64, false, false, false, true);
        case_hconv("hconv up32->16 2x32x64", 2, 32, 64, 32, 0, 16, false, true, true, false);
        case_hconv("hconv up64+skip64->32 2x32x32", 2, 32, 32, 64, 64, 32, false, true, true, true);
        case_hconv("hconv 32->64 (dgrad shape) 1x16x24", 1, 16, 24, 0, 32, 64, false, false, false, false);
    }
    if (want("tconv")) {
        case_tconv("tconv 16->16 2x32x32", 2, 32, 32, 16, 16, false, false, true, true, true);
        case_tconv("tconv 32->32 1x48x40 (partial tiles)", 1, 48, 40, 32, 32, false, false, true, true, true);
        case_tconv("tconv 64->64 +res 3x24x24 (partial)", 3, 24, 24, 64, 64, false, true, true, true, true);
        case_tconv("tconv 64->64 raw 5x64x64 (multi-tile/CTA)", 5, 64, 64, 64, 64, false, false, false, false, true);
        case_tconv("tconv 16->16 +res 9x64x96 (multi-tile/CTA)", 9, 64, 96, 16, 16, false, true, true, true, true);
        case_tconv("tconv 32->64 (dgrad shape) 1x16x24", 1, 16, 24, 32, 64, false, false, false, false, false);
        case_tconv("tconv 64->32 2x16x8 (narrow)", 2, 16, 8, 64, 32, false, false, true, true, true);
        case_tconv("tconv parity up32->16 2x32x64", 2, 32, 64, 32, 16, true, false, true, true, true);
        case_tconv("tconv parity up32->16 7x96x80 (multi-tile, partial)", 7, 96, 80, 32, 16, true, false, true, true, true);
        case_tconv("tconv parity up64->32 1x32x32", 1, 32, 32, 64, 32, true, false, true, true, false);
        // hi-res extents that are not multiples of the 32 x 16 staged tile: the TMA store clips, the statistics mask
        case_tconv("tconv parity up32->16 3x40x24 (clipped store)", 3, 40, 24, 32, 16, true, false, true, true, true);
        case_tconv("tconv parity up64->32 2x24x40 (clipped store)", 2, 24, 40, 64, 32, true, false, true, true, true);
    }
    if (want("split")) {
        case_dec1_split("split up128+skip128->128 1x8x8", 1, 8, 8, 128, 128, 128);
        case_dec1_split("split up512+skip256->256 2x16x16", 2, 16, 16, 512, 256, 256);
        case_dec1_split("split up256+skip128->128 3x24x40 (partial tiles)", 3, 24, 40, 256, 128, 128);
    }
    if (want("dbench")) {
        case_dec1_split("D0c1 up512+skip256->256 @32^2 x32", 32, 16, 16, 512, 256, 256, 10);
        case_dec1_split("D1c1 up256+skip128->128 @64^2 x32", 32, 32, 32, 256, 128, 128, 10);
    }
    if (want("wide")) {
        case_wconv("wide 128->128 1x16x16", 1, 16, 16, 128, 128, false, true, true, true);
        case_wconv("wide 64->128 2x24x40 (partial tiles)", 2, 24, 40, 64, 128, true, true, true, true);
        case_wconv("wide 256->256 +res 3x32x32 (multi-item/CTA)", 3, 32, 32, 256, 256, true, true, true, true);
        case_wconv("wide 256->512 raw 40x16x16 (multi-item/CTA)", 40, 16, 16, 256, 512, false, false, false, true);
        case_wconv("wide 512->128 (dgrad shape) 2x16x16", 2, 16, 16, 512, 128, false, false, false, false);
    }
    if (want("pair")) {
        case_wconv("pair 128->128 1x16x16", 1, 16, 16, 128, 128, false, true, true, true, true);
        case_wconv("pair 64->128 2x24x40 (partial tiles)", 2, 24, 40, 64, 128, true, true, true, true, true);
        case_wconv("pair 256->256 +res 3x32x32 (multi-item)", 3, 32, 32, 256, 256, true, true, true, true, true);
        case_wconv("pair 256->512 raw 40x16x16 (multi-item)", 40, 16, 16, 256, 512, false, false, false, true, true);
        case_wconv("pair 512->128 (dgrad shape) 2x16x16", 2, 16, 16, 512, 128, false, false, false, false, true);
    }
    if (want("pbench")) {
        bench_wconv("PAIR L2 3x3 128->128 @64^2 x32", 32, 64, 64, 128, 128, 20, true);
        bench_wconv("PAIR L3 3x3 256->256 @32^2 x32", 32, 32, 32, 256, 256, 20, true);
        bench_wconv("PAIR L4 3x3 512->512 @16^2 x32", 32, 16, 16, 512, 512, 20, true);
        bench_wconv("PAIR L3 3x3 256->256 @32^2 x16", 16, 32, 32, 256, 256, 20, true);
        bench_wconv("PAIR L4 3x3 512->512 @16^2 x16", 16, 16, 16, 512, 512, 20, true);
    }
    if (want("xbench")) {
        bench_wconv("L2 3x3 128->128 @64^2 x32", 32, 64, 64, 128, 128, 20);
        bench_wconv("L3 3x3 256->256 @32^2 x32", 32, 32, 32, 256, 256, 20);
        bench_wconv("L4 3x3 512->512 @16^2 x32", 32, 16, 16, 512, 512, 20);
        bench_wconv("L3 3x3 256->256 @32^2 x16", 16, 32, 32, 256, 256, 20);
        bench_wconv("L4 3x3 512->512 @16^2 x16", 16, 16, 16, 512, 512, 20);
        bench_conv("igemm L2 3x3 128->128 @64^2 x32", 32, 64, 64, 128, 128, 3, 1, 20);
        bench_conv("igemm L3 3x3 256->256 @32^2 x32", 32, 32, 32, 256, 256, 3, 1, 20);
        bench_conv("igemm L4 3x3 512->512 @16^2 x32", 32, 16, 16, 512, 512, 3, 1, 20);
        bench_conv("igemm L3 3x3 256->256 @32^2 x16", 16, 32, 32, 256, 256, 3, 1, 20);
        bench_conv("igemm L4 3x3 512->512 @16^2 x16", 16, 16, 16, 512, 512, 3, 1, 20);
    }
    if (want("tbench")) {
        bench_tconv("D4c2 16->16 @512^2 x32", 32, 512, 512, 16, 16, false, false, 5);
        bench_tconv("D3c2 32->32 @256^2 x32", 32, 256, 256, 32, 32, false, false, 5);
        bench_tconv("L1 64->64 @128^2 x32", 32, 128, 128, 64, 64, false, false, 10);
        bench_tconv("L1 64->64 +res @128^2 x32", 32, 128, 128, 64, 64, false, true, 10);
        bench_tconv("D4c1 up32->16 @512^2 x32 (parity)", 32, 512, 512, 32, 16, true, false, 5);
        bench_tconv("D3c1 up64->32 @256^2 x32 (parity)", 32, 256, 256, 64, 32, true, false, 5);
        bench_tconv("D3c1s 64->32 +res @256^2 x32", 32, 256, 256, 64, 32, false, true, 5);
    }
    if (want("bench")) {
        bench_conv("L1 3x3 64->64 @128^2 x32", 32, 128, 128, 64, 64, 3, 1, 20);
        bench_conv("L2 3x3 128->128 @64^2 x32", 32, 64, 64, 128, 128, 3, 1, 20);
        bench_conv("L3 3x3 256->256 @32^2 x32", 32, 32, 32, 256, 256, 3, 1, 20);
        bench_conv("L4 3x3 512->512 @16^2 x32", 32, 16, 16, 512, 512, 3, 1, 20);
        bench_conv("D3 3x3 32->32 @256^2 x32", 32, 256, 256, 32, 32, 3, 1, 10);
        bench_conv("D4 3x3 16->16 @512^2 x32", 32, 512, 512, 16, 16, 3, 1, 10);
    }
    if (want("hwgrad")) {
        case_hwgrad("hwgrad 16->16 2x32x32", 2, 32, 32, 0, 16, 16);
        case_hwgrad("hwgrad 32->32 1x48x40 (partial tiles)", 1, 48, 40, 0, 32, 32);
        case_hwgrad("hwgrad 64->64 3x24x24 (2 row groups)", 3, 24, 24, 0, 64, 64);
        case_hwgrad("hwgrad 16(K)->32 1x32x16", 1, 32, 16, 0, 16, 32);
        case_hwgrad("hwgrad up32->16 2x32x64", 2, 32, 64, 32, 0, 16);
        case_hwgrad("hwgrad up64+skip64->32 2x32x32 (4 groups)", 2, 32, 32, 64, 64, 32);
        case_hwgrad("hwgrad 64->64 5x64x64 (multi-tile/CTA)", 5, 64, 64, 0, 64, 64);
    }
    if (want("swgrad")) {
        case_swgrad("swgrad 2x64x128 (wt 64)", 2, 64, 128);
        case_swgrad("swgrad 1x32x96 (wt 16)", 1, 32, 96);
        case_swgrad("swgrad 3x96x64 (wt 32, several row tiles)", 3, 96, 64);
    }
    if (want("swbench")) case_swgrad("stem wgrad @512^2 x16", 16, 512, 512, 5);
    if (want("dlow")) {
        case_dlow("dlow 16->32 2x16x16 (decoder.blocks.4 shape)", 2, 16, 16, 16, 32);
        case_dlow("dlow 32->64 1x24x16 (partial tile rows)", 1, 24, 16, 32, 64);
        case_dlow("dlow 16->32 3x32x40 (multi-tile)", 3, 32, 40, 16, 32);
        case_dlow("dlow 32->64 2x8x8 (one small tile)", 2, 8, 8, 32, 64);
    }
    if (want("dlbench")) {
        case_dlow("D4c1 dlow 16->32 @256^2 x16", 16, 256, 256, 16, 32, 5);
        case_dlow("D3c1 dlow 32->64 @128^2 x16", 16, 128, 128, 32, 64, 10);
    }
    if (want("xwgrad")) {
        case_xwgrad("xwgrad 16->16 2x32x32", 2, 32, 32, 16, 16, 0);
        case_xwgrad("xwgrad 32->32 1x48x40 (th 16)", 1, 48, 40, 32, 32, 0);
        case_xwgrad("xwgrad 64->64 3x24x24 (th 8, 2 co groups)", 3, 24, 24, 64, 64, 0);
        case_xwgrad("xwgrad 16->32 1x32x16", 1, 32, 16, 16, 32, 0);
        case_xwgrad("xwgrad 64->32 2x20x16 (partial tile rows, column offset 64 of 128)", 2, 20, 16, 64, 32, 64);
        case_xwgrad("xwgrad 64->64 5x64x64 (multi-tile/CTA)", 5, 64, 64, 64, 64, 0);
        case_xwgrad("xwgrad wide 128->128 2x16x16", 2, 16, 16, 128, 128, 0);
        case_xwgrad("xwgrad wide 256->256 3x32x32", 3, 32, 32, 256, 256, 0);
        case_xwgrad("xwgrad wide 64->128 1x64x64 (column offset 128)", 1, 64, 64, 64, 128, 128);
        case_xwgrad("xwgrad wide 512->512 2x16x16", 2, 16, 16, 512, 512, 0);
        case_xwgrad_up("xwgrad up32->16 2x32x64", 2, 32, 64, 32, 16, 0);
        case_xwgrad_up("xwgrad up64->32 2x32x32 (of 128 columns)", 2, 32, 32, 64, 32, 64);
        case_xwgrad_up("xwgrad up128->64 1x48x32 (of 192 columns)", 1, 48, 32, 128, 64, 64);
        case_xwgrad_up("xwgrad wide up256->128 2x32x32 (of 384 columns)", 2, 32, 32, 256, 128, 128);
        case_xwgrad_up("xwgrad wide up512->256 1x32x32 (of 768 columns)", 1, 32, 32, 512, 256, 256);
    }
    if (want("xbench")) {
        bench_xwgrad("D4c2 16->16 @512^2 x16", 16, 512, 512, 16, 16, 5);
        bench_xwgrad("D3c2 32->32 @256^2 x16", 16, 256, 256, 32, 32, 5);
        bench_xwgrad("L1 64->64 @128^2 x16", 16, 128, 128, 64, 64, 10);
        bench_xwgrad("D3c1 skip64->32 @256^2 x16", 16, 256, 256, 64, 32, 5);
        bench_xwgrad("L2 128->128 @64^2 x16", 16, 64, 64, 128, 128, 10);
        bench_xwgrad("L3 256->256 @32^2 x16", 16, 32, 32, 256, 256, 10);
        bench_xwgrad("L4 512->512 @16^2 x16", 16, 16, 16, 512, 512, 10);
        bench_xwgrad_up("D4c1 up32->16 @512^2 x16", 16, 512, 512, 32, 16, 5);
        bench_xwgrad_up("D3c1 up64->32 @256^2 x16", 16, 256, 256, 64, 32, 5);
        bench_xwgrad_up("D2c1 up128->64 @128^2 x16", 16, 128, 128, 128, 64, 10);
        bench_xwgrad_up("D1c1 up256->128 @64^2 x16", 16, 64, 64, 256, 128, 10);
        bench_xwgrad_up("D0c1 up512->256 @32^2 x16", 16, 32, 32, 512, 256, 10);
    }
    if (want("wbench")) {
        bench_hwgrad("D4c2 16->16 @512^2 x16", 16, 512, 512, 0, 16, 16, 5);
        bench_hwgrad("D3c2 32->32 @256^2 x16", 16, 256, 256, 0, 32, 32, 5);
        bench_hwgrad("L1 64->64 @128^2 x16", 16, 128, 128, 0, 64, 64, 10);
        bench_hwgrad("D4c1 up32->16 @512^2 x16", 16, 512, 512, 32, 0, 16, 5);
        bench_hwgrad("D3c1 up64+64->32 @256^2 x16", 16, 256, 256, 64, 64, 32, 5);
    }
    if (want("hbench")) {
        bench_hconv("D4c2 16->16 @512^2 x32", 32, 512, 512, 0, 16, 16, 5);
        bench_hconv("D3c2 32->32 @256^2 x32", 32, 256, 256, 0, 32, 32, 5);
        bench_hconv("L1 64->64 @128^2 x32", 32, 128, 128, 0, 64, 64, 10);
        bench_hconv("D4c1 up32->16 @512^2 x32", 32, 512, 512, 32, 0, 16, 5);
        bench_hconv("D3c1 up64+64->32 @256^2 x32", 32, 256, 256, 64, 64, 32, 5);
    }
    printf("SELFTEST %s: %d checks, %d failures\n", g_fail ? "FAILED" : "OK", g_run, g_fail);
    return g_fail ? 1 : 0;
}
