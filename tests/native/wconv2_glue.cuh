// Host glue of the CTA-pair wide conv (wconv2.cuh): a measured negative result of round 1 (profiles/r1s4_negative_results.txt
// item 7), kept out of libunetb200.so and built only into the native self-test for its parity checks / benchmarks.
#pragma once
#include "wconv2.cuh"

namespace ub {

// ------------------------------------------------------------------------------------------------ wconv2 (CTA-pair wide conv)
struct Wconv2Launch {
    CUtensorMap a, b;
    Wconv2Params p;
    int grid = 0, kn = 0;
    uint32_t smem = 0;
};
inline bool wconv2_ok(int cin, int cout) { return wconv_ok(cin, cout); }

template <int kN>
inline cudaError_t wconv2_launch_t(const Wconv2Launch& L, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(wconv2_kernel<kN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(L.grid, 1, 1);
    cfg.blockDim = dim3(kW2Threads, 1, 1);
    cfg.dynamicSmemBytes = L.smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, wconv2_kernel<kN>, L.a, L.b, L.p);
}
inline cudaError_t wconv2_launch(const Wconv2Launch& L, cudaStream_t st) {
    return L.kn == 256 ? wconv2_launch_t<256>(L, st) : wconv2_launch_t<128>(L, st);
}

inline std::string wconv2_build(Wconv2Launch& L, const void* src, int cin, const void* wpk, int cout, int N, int H, int W,
                                void* out, const EpilogueDesc& ep, int* err, int num_sms, long long ldb = 0) {
    memset(&L.p, 0, sizeof(L.p));
    Wconv2Params& P = L.p;
    if (!wconv2_ok(cin, cout)) return "wconv2: unsupported channel configuration";
    if (ep.residual.ptr && (ep.residual.sW != cout || ep.residual.sH != (long long)W * cout ||
                            ep.residual.sN != (long long)H * W * cout))
        return "wconv2: residual must be a dense NHWC tensor of the output's shape";
    const int kn = cout % 256 == 0 ? 256 : 128;
    L.kn = kn;
    P.H = H; P.W = W; P.N = N;
    P.tiles_w = (W + 15) / 16;
    P.tiles_h = (H + 15) / 16;
    P.n_tiles = cout / kn;
    P.cin = cin; P.cout = cout;
    P.scale = ep.scale; P.shift = ep.shift; P.relu = ep.relu;
    P.out = reinterpret_cast<__nv_bfloat16*>(out);
    P.residual = reinterpret_cast<const __nv_bfloat16*>(ep.residual.ptr);
    P.stats = ep.stats;
    P.err = err;
    int bst = 9;
    while (bst > 2 && wconv2_smem(kn, bst).total + 1024 > 232448u) --bst;
    P.bstages = bst;
    L.smem = wconv2_smem(kn, bst).total + 1024;
    {
        uint64_t dims[4] = {(uint64_t)cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)cin * 2, (uint64_t)W * cin * 2, (uint64_t)H * W * cin * 2};
        uint32_t box[4] = {64, 10, 18, 1};
        uint32_t es[4] = {1, 1, 1, 1};
        std::string e = make_tmap_bf16(&L.a, src, 4, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_128B);
        if (!e.empty()) return "wconv2 A map: " + e;
    }
    {
        uint64_t dims[2] = {(uint64_t)9 * cin, (uint64_t)cout};
        uint64_t str[1] = {(uint64_t)(ldb ? ldb : 9ll * cin) * 2};
        uint32_t box[2] = {64, (uint32_t)(kn / 2)};
        uint32_t es[2] = {1, 1};
        std::string e = make_tmap_bf16(&L.b, wpk, 2, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_128B);
        if (!e.empty()) return "wconv2 B map: " + e;
    }
    const int total = P.tiles_w * P.tiles_h * N * P.n_tiles;
    const int pairs = num_sms / 2;
    const int waves = (total + pairs - 1) / pairs;
    L.grid = 2 * ((total + waves - 1) / waves);
    return "";
}


}  // namespace ub
