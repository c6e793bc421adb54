// Host-side description of one implicit-GEMM launch: tensor maps + IgemmParams + grid/smem sizing.
#pragma once
#include <string.h>

#include <string>

#include "igemm.cuh"
#include "tmap.cuh"

namespace ub {

// Strided NHWC-like bf16 view: C contiguous; strides in elements.
struct View4 {
    const void* ptr = nullptr;
    int C = 0, W = 0, H = 0, N = 0;
    long long sW = 0, sH = 0, sN = 0;
};
inline View4 nhwc_view(const void* p, int N, int H, int W, int C) {
    View4 v;
    v.ptr = p; v.C = C; v.W = W; v.H = H; v.N = N;
    v.sW = C; v.sH = (long long)W * C; v.sN = (long long)H * W * C;
    return v;
}

struct SrcDesc {
    View4 v;
    int es_w = 1, es_h = 1;  // traversal stride of the tile along w/h (== coordinate multiplier of the tile origin)
};

struct IgemmLaunch {
    CUtensorMap a0, a1, b, d;
    IgemmParams p;
    int grid = 0;
    uint32_t smem = 0;
};

inline int next_pow2(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

struct EpilogueDesc {
    const float* scale = nullptr;
    const float* shift = nullptr;
    int relu = 0;
    View4 residual;          // ptr == nullptr: none
    float* stats = nullptr;  // [m_tiles][cout][2]
};

// number of M tiles an output view is cut into (needed to size the statistics partial buffer)
inline void igemm_tile_shape(const View4& out, int& bw, int& bh, int& bn) {
    bw = next_pow2(out.W) < 16 ? next_pow2(out.W) : 16;
    bh = next_pow2(out.H) < 128 / bw ? next_pow2(out.H) : 128 / bw;
    bn = 128 / (bw * bh);
}
inline int igemm_m_tiles(const View4& out) {
    int bw, bh, bn;
    igemm_tile_shape(out, bw, bh, bn);
    return ((out.W + bw - 1) / bw) * ((out.H + bh - 1) / bh) * ((out.N + bn - 1) / bn);
}

inline std::string igemm_build(IgemmLaunch& L, const SrcDesc* src, int nsrc, const IgemmTap* taps, int ntaps,
                               int chunk_elems, const void* wpk, int ktotal, int cout, const View4& out,
                               const EpilogueDesc& ep, int* err, int num_sms) {
    memset(&L, 0, sizeof(L));
    IgemmParams& P = L.p;
    if (ntaps > kMaxTaps) return "too many taps";
    if (chunk_elems != 16 && chunk_elems != 32 && chunk_elems != 64) return "chunk_elems must be 16/32/64";
    if (cout % 16 != 0 || cout > 512) return "cout must be a multiple of 16 and <= 512";
    igemm_tile_shape(out, P.bw, P.bh, P.bn);
    P.tiles_w = (out.W + P.bw - 1) / P.bw;
    P.tiles_h = (out.H + P.bh - 1) / P.bh;
    P.tiles_n = (out.N + P.bn - 1) / P.bn;
    P.ntile = cout <= 256 ? cout : 256;
    if (cout % P.ntile != 0) return "cout not divisible by ntile";
    P.n_tiles = cout / P.ntile;
    P.chunk_elems = chunk_elems;
    P.num_taps = ntaps;
    int total = 0;
    for (int t = 0; t < ntaps; ++t) {
        P.taps[t] = taps[t];
        total += taps[t].nchunks;
        if (taps[t].src >= nsrc) return "tap source out of range";
    }
    P.total_chunks = total;
    if (total * chunk_elems != ktotal) return "tap table does not cover ktotal";
    P.Wo = out.W; P.Ho = out.H; P.Nimg = out.N; P.cout = cout;
    P.out_cblk = P.ntile < 64 ? P.ntile : 64;
    P.scale = ep.scale; P.shift = ep.shift; P.relu = ep.relu;
    P.residual = reinterpret_cast<const __nv_bfloat16*>(ep.residual.ptr);
    P.res_sw = ep.residual.sW; P.res_sh = ep.residual.sH; P.res_sn = ep.residual.sN;
    P.stats = ep.stats;
    P.err = err;

    // pipeline depth from the shared-memory budget
    int stages = 8;
    for (; stages >= 2; --stages) {
        IgemmSmem s = igemm_smem(P.ntile, chunk_elems, stages, P.out_cblk);
        if (s.total + 1024 <= 232448u) break;
    }
    if (stages < 2) return "tile does not fit in shared memory";
    if (stages > total) stages = total < 2 ? 2 : total;
    P.stages = stages;
    L.smem = igemm_smem(P.ntile, chunk_elems, stages, P.out_cblk).total + 1024;

    const CUtensorMapSwizzle swz = swizzle_for_bytes(chunk_elems * 2);
    for (int i = 0; i < 2; ++i) {
        const SrcDesc& s = src[i < nsrc ? i : 0];
        P.mulw[i] = s.es_w;
        P.mulh[i] = s.es_h;
        uint64_t dims[4] = {(uint64_t)s.v.C, (uint64_t)s.v.W, (uint64_t)s.v.H, (uint64_t)s.v.N};
        uint64_t str[3] = {(uint64_t)s.v.sW * 2, (uint64_t)s.v.sH * 2, (uint64_t)s.v.sN * 2};
        uint32_t box[4] = {(uint32_t)chunk_elems, (uint32_t)(P.bw * s.es_w), (uint32_t)(P.bh * s.es_h), (uint32_t)P.bn};
        uint32_t es[4] = {1, (uint32_t)s.es_w, (uint32_t)s.es_h, 1};
        std::string e = make_tmap_bf16(i == 0 ? &L.a0 : &L.a1, s.v.ptr, 4, dims, str, box, es, swz);
        if (!e.empty()) return "A map: " + e;
    }
    {
        uint64_t dims[2] = {(uint64_t)ktotal, (uint64_t)cout};
        uint64_t str[1] = {(uint64_t)ktotal * 2};
        uint32_t box[2] = {(uint32_t)chunk_elems, (uint32_t)P.ntile};
        uint32_t es[2] = {1, 1};
        std::string e = make_tmap_bf16(&L.b, wpk, 2, dims, str, box, es, swz);
        if (!e.empty()) return "B map: " + e;
    }
    {
        uint64_t dims[4] = {(uint64_t)out.C, (uint64_t)out.W, (uint64_t)out.H, (uint64_t)out.N};
        uint64_t str[3] = {(uint64_t)out.sW * 2, (uint64_t)out.sH * 2, (uint64_t)out.sN * 2};
        uint32_t box[4] = {(uint32_t)P.out_cblk, (uint32_t)P.bw, (uint32_t)P.bh, (uint32_t)P.bn};
        uint32_t es[4] = {1, 1, 1, 1};
        std::string e = make_tmap_bf16(&L.d, out.ptr, 4, dims, str, box, es, swizzle_for_bytes(P.out_cblk * 2));
        if (!e.empty()) return "D map: " + e;
    }
    const int total_tiles = P.tiles_w * P.tiles_h * P.tiles_n * P.n_tiles;
    const int waves = (total_tiles + num_sms - 1) / num_sms;
    L.grid = (total_tiles + waves - 1) / waves;
    return "";
}

inline cudaError_t igemm_launch(const IgemmLaunch& L, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    igemm_kernel<<<L.grid, kIgemmThreads, L.smem, st>>>(L.a0, L.a1, L.b, L.d, L.p);
    return cudaGetLastError();
}

}  // namespace ub
