// Host-side description of one implicit-GEMM launch: tensor maps + IgemmParams + grid/smem sizing.
#pragma once
#include <stdlib.h>
#include <string.h>

#include <string>

#include "hconv.cuh"
#include "hwgrad.cuh"
#include "igemm.cuh"
#include "tconv.cuh"
#include "wconv.cuh"
#include "xwgrad.cuh"
#include "dlow.cuh"
#include "swgrad.cuh"
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
    bool residual_f32 = false;  // residual.ptr addresses fp32 elements (wconv only: the training forward's partial)
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

inline cudaError_t igemm_set_attr() {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    return cudaSuccess;
}
// Weight-tile multicast across a thread-block cluster is implemented and parity-tested but OFF by default: measured on
// B200 (profiles/r1s3_igemm_cluster_ab.txt) it does not speed up the wide layers (L3 256->256: 43.9 us without, 44.8 us
// with clusters of 4) because the limit is what ONE SM can ingest (~40 B/clk: A + B tile = 48 KB per 512 MMA cycles),
// not the L2 read rate that multicast relieves.  wconv.cuh attacks the ingest instead.  UB_IGEMM_CLUSTER=1 enables it.
inline bool igemm_cluster_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* v = getenv("UB_IGEMM_CLUSTER");
        on = (v && v[0] == '1') ? 1 : 0;
    }
    return on != 0;
}
// co-resident clusters of `csize` one-CTA-per-SM blocks (GPC boundaries can strand a few SMs); queried once per size
inline int igemm_max_clusters(int csize, int num_sms) {
    if (csize == 1) return num_sms;
    static int cached[5] = {0, 0, 0, 0, 0};
    if (!cached[csize]) {
        int n = 0;
        if (igemm_set_attr() == cudaSuccess) {
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3(num_sms / csize * csize, 1, 1);
            cfg.blockDim = dim3(kIgemmThreads, 1, 1);
            cfg.dynamicSmemBytes = 200 * 1024;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = csize;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            if (cudaOccupancyMaxActiveClusters(&n, igemm_kernel, &cfg) != cudaSuccess) n = 0;
            cudaGetLastError();
        }
        if (n <= 0 || n > num_sms / csize) n = (num_sms / csize) * 3 / 4 > 0 ? (num_sms / csize) * 3 / 4 : 1;
        cached[csize] = n;
    }
    return cached[csize];
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

    // cluster multicast of the weight tile for the wide layers (L2 -> SM bandwidth bound otherwise): csize CTAs on
    // adjacent M tiles share one copy of B.  Needs 1 KB-aligned B parts (>= 8 rows of 128 B) and enough M tiles.
    const int m_tiles = P.tiles_w * P.tiles_h * P.tiles_n;
    int csize = 1;
    if (igemm_cluster_enabled() && chunk_elems == 64 && P.ntile >= 128) {
        if (m_tiles >= 8) csize = 4;
        else if (m_tiles >= 2) csize = 2;
    }
    P.csize = csize;

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
        uint32_t box[2] = {(uint32_t)chunk_elems, (uint32_t)(P.ntile / csize)};
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
    const int total_groups = ((m_tiles + csize - 1) / csize) * P.n_tiles;
    const int max_clusters = igemm_max_clusters(csize, num_sms);
    const int waves = (total_groups + max_clusters - 1) / max_clusters;
    L.grid = ((total_groups + waves - 1) / waves) * csize;
    return "";
}

inline cudaError_t igemm_launch(const IgemmLaunch& L, cudaStream_t st) {
    cudaError_t e = igemm_set_attr();
    if (e != cudaSuccess) return e;
    if (L.p.csize == 1) {
        launch_k(igemm_kernel, L.grid, kIgemmThreads, L.smem, st, L.a0, L.a1, L.b, L.d, L.p);
        return cudaGetLastError();
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(L.grid, 1, 1);
    cfg.blockDim = dim3(kIgemmThreads, 1, 1);
    cfg.dynamicSmemBytes = L.smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = L.p.csize;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, igemm_kernel, L.a0, L.a1, L.b, L.d, L.p);
}

// ------------------------------------------------------------------------------------------------ hconv (halo-resident 3x3)
struct HconvLaunch {
    HconvParams p;
    int grid = 0;
    uint32_t smem = 0;
    int occ = 1;       // CTAs per SM (1 or 2)
};

// Can the halo-resident kernel run a 3x3/s1 conv with these channel counts?  (weights of all 9 taps + >= 2 halo stages
// must fit in shared memory).  Returns the pipeline depth (0 = no).
inline int hconv_stages(int cup, int cskip, int cout) {
    const int ctot = cup + cskip;
    if (!(cout == 16 || cout == 32 || cout == 64)) return 0;
    if (!(ctot == 16 || ctot == 32 || ctot == 64 || ctot == 128) || cup % 8 || cskip % 8) return 0;
    for (int st = 6; st >= 2; --st)
        if (hconv_smem(ctot, cout, st).total + 1024 <= 232448u) return st;
    return 0;
}

inline std::string hconv_build(HconvLaunch& L, const void* low, int cup, const void* src, int cskip, const void* wpk,
                               int cout, int N, int H, int W, void* out, const EpilogueDesc& ep, int* err, int num_sms) {
    memset(&L, 0, sizeof(L));
    HconvParams& P = L.p;
    const int stages = hconv_stages(cup, cskip, cout);
    if (!stages) return "hconv: unsupported channel configuration";
    if (cup && ((H | W) & 1)) return "hconv: up-sampled source needs even H, W";
    if (ep.residual.ptr && (ep.residual.sW != cout || ep.residual.sH != (long long)W * cout ||
                            ep.residual.sN != (long long)H * W * cout))
        return "hconv: residual must be a dense NHWC tensor of the output's shape";
    P.H = H; P.W = W; P.N = N;
    P.tiles_w = (W + kHcTileW - 1) / kHcTileW;
    P.tiles_h = (H + kHcTileH - 1) / kHcTileH;
    P.cup = cup; P.cskip = cskip;
    P.low = reinterpret_cast<const __nv_bfloat16*>(low);
    P.src = reinterpret_cast<const __nv_bfloat16*>(src);
    P.wpk = reinterpret_cast<const __nv_bfloat16*>(wpk);
    P.cout = cout;
    P.stages = stages;
    P.scale = ep.scale; P.shift = ep.shift; P.relu = ep.relu;
    P.out = reinterpret_cast<__nv_bfloat16*>(out);
    P.residual = reinterpret_cast<const __nv_bfloat16*>(ep.residual.ptr);
    P.stats = ep.stats;
    P.err = err;
    L.smem = hconv_smem(cup + cskip, cout, stages).total + 1024;
    L.occ = 1;
    // two co-resident CTAs when both fit (shared memory incl. the 1 KB per-CTA reservation, TMEM 2 x 4 x cout <= 512)
    for (int st = stages < 4 ? stages : 4; st >= 2; --st) {
        const uint32_t sm2 = hconv_smem(cup + cskip, cout, st).total + 1024;
        if (2 * (sm2 + 1024) <= 232448u) {
            L.occ = 2;
            P.stages = st;
            L.smem = sm2;
            break;
        }
    }
    const int total_tiles = P.tiles_w * P.tiles_h * N;
    const int slots = num_sms * L.occ;
    const int waves = (total_tiles + slots - 1) / slots;
    L.grid = (total_tiles + waves - 1) / waves;
    return "";
}

inline cudaError_t hconv_launch(const HconvLaunch& L, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(hconv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(hconv_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 115200);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    if (L.occ == 2) launch_k(hconv_kernel<2>, L.grid, kHcThreads, L.smem, st, L.p);
    else launch_k(hconv_kernel<1>, L.grid, kHcThreads, L.smem, st, L.p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ tconv (TMA halo conv)
struct TconvLaunch {
    CUtensorMap a, d;
    TconvParams p;
    int grid = 0;
    uint32_t smem = 0;
    int occ = 1, iph = 1;
};

// plain mode: 3x3/s1 conv src[N,H,W,cin] -> out[N,H,W,cout]; parity mode (low_res != 0): src is the LOW-RES tensor
// [N,H/2,W/2,cin] and the conv is nearest-2x-upsample + 3x3 (weights = PK_HPAR parity-folded 2x2 taps).
inline bool tconv_ok(int cin, int cout, bool parity) {
    if (!(cout == 16 || cout == 32 || cout == 64)) return false;
    if (!(cin == 16 || cin == 32 || cin == 64)) return false;
    if (parity && cout > 32) return false;  // 4 accumulators x cout x >= 2 sets must fit the TMEM budget
    return true;
}
// parity mode: 18 [cout][cin] blocks (tc_issue_parity: 9 halo shifts, runs of 4 + 2 + 2 + 3 + 3 + 1 + 1 + 1 + 1 parities)
inline long long tconv_w_elems(int cin, int cout, bool parity) { return (parity ? 18ll : 9ll) * cin * cout; }

// Inference plans set this while they build: weights / folded BN constants are never written by a kernel of the
// inference stream, so tconv may copy them to shared memory BEFORE griddepcontrol.wait (PDL prologue overlap).  Training
// re-packs weights every step in the same stream and keeps the wait first.
inline bool& tconv_const_weights_flag() {
    static thread_local bool f = false;
    return f;
}
struct TconvConstWeightsScope {
    bool prev;
    TconvConstWeightsScope() : prev(tconv_const_weights_flag()) { tconv_const_weights_flag() = !getenv("UNETB200_NO_WPREFETCH"); }
    ~TconvConstWeightsScope() { tconv_const_weights_flag() = prev; }
};
inline std::string tconv_build(TconvLaunch& L, const void* src, int cin, bool parity, const void* wpk, int cout, int N,
                               int H, int W, void* out, const EpilogueDesc& ep, int* err, int num_sms) {
    memset(&L.p, 0, sizeof(L.p));
    if (tconv_const_weights_flag()) L.p.dbg |= 16;
    TconvParams& P = L.p;
    if (!tconv_ok(cin, cout, parity)) return "tconv: unsupported channel configuration";
    if (parity && ((H | W) & 1)) return "tconv: parity mode needs even H, W";
    if (ep.residual.ptr && (ep.residual.sW != cout || ep.residual.sH != (long long)W * cout ||
                            ep.residual.sN != (long long)H * W * cout))
        return "tconv: residual must be a dense NHWC tensor of the output's shape";
    const uint32_t row_bytes = cin * 2;
    P.H = H; P.W = W; P.N = N;
    P.mode = parity ? 1 : 0;
    P.cin = cin; P.cout = cout;
    // Cout <= 32: the three filter columns as N-blocks of one MMA (tc_issue_stack): a third of the A-operand reads, and
    // with lane = pixel of an image row the register stores / residual loads of the epilogue are coalesced, no staging.
    // Used where the MMA issue dominates the tile (Cin >= 64 with Cout <= 32: decoder.blocks.3.conv1 [skip], 64 -> 32 with
    // residual @256^2 x32: 155 -> 109 us).  For Cin <= 32 the nine-tap issue stays: there the one-item-per-thread epilogue
    // chain (3 TMEM loads, 32 shuffles) becomes the critical role and the stacked form is slower (16->16 @512^2 135 vs
    // 115 us, 32->32 @256^2 80 vs 65 us; profiles/r1s4_negative_results.txt #11).  UNETB200_STACK=1 forces it for every
    // Cout <= 32 launch, UNETB200_NO_STACK=1 switches it off.
    static const bool stack_all = getenv("UNETB200_STACK") != nullptr, stack_off = getenv("UNETB200_NO_STACK") != nullptr;
    const bool stack_on = !stack_off && (stack_all || cin >= 64);
    if (stack_on && !parity && cout <= 32 && W >= 32) {
        static const bool stack_occ2 = getenv("UNETB200_STACK_OCC1") == nullptr;
        const int snt = (cout == 16 ? 4 : 2) / (stack_occ2 ? 2 : 1);   // 3 * cout * snt accumulator columns, double buffered
        P.mode = 3;
        P.nt = snt;
        P.tiles_w = (W + 29) / 30;
        P.tiles_h = (H + 4 * snt - 1) / (4 * snt);
        P.halo_w = 32;
        P.tx_bytes = 32u * (uint32_t)(4 * snt + 2) * row_bytes;
        P.stage_bytes = (P.tx_bytes + 1023u) & ~1023u;
        P.w_bytes = (uint32_t)tconv_w_elems(cin, cout, false) * 2;
        P.wpk = reinterpret_cast<const __nv_bfloat16*>(wpk);
        P.scale = ep.scale; P.shift = ep.shift; P.relu = ep.relu;
        P.out = reinterpret_cast<__nv_bfloat16*>(out);
        P.residual = ep.residual_f32 ? nullptr : reinterpret_cast<const __nv_bfloat16*>(ep.residual.ptr);
        P.residual32 = ep.residual_f32 ? reinterpret_cast<const float*>(ep.residual.ptr) : nullptr;
        P.stats = ep.stats;
        P.err = err;
        L.occ = stack_occ2 ? 2 : 1;
        L.iph = 1;                                   // snt * cout / 16 = 4 (2) items over 16 (8) epilogue warps
        int sst = 0;
        for (int st = stack_occ2 ? 4 : 6; st >= 2; --st) {
            const uint32_t tot = tconv_smem(P.w_bytes, P.stage_bytes, st, 0).total + 1024;
            if (stack_occ2 ? 2 * (tot + 1024) <= 232448u : tot <= 232448u) {
                sst = st;
                break;
            }
        }
        if (sst < 2) return "tconv: does not fit in shared memory";
        P.stages = sst;
        P.nacc = 2;
        L.smem = tconv_smem(P.w_bytes, P.stage_bytes, sst, 0).total + 1024;
        L.d = CUtensorMap{};
        uint64_t dims[4] = {(uint64_t)cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)cin * 2, (uint64_t)W * cin * 2, (uint64_t)H * W * cin * 2};
        uint32_t box[4] = {(uint32_t)cin, 32, (uint32_t)(4 * snt + 2), 1};
        uint32_t es[4] = {1, 1, 1, 1};
        std::string e = make_tmap_bf16(&L.a, src, 4, dims, str, box, es, swizzle_for_bytes(row_bytes));
        if (!e.empty()) return "tconv A map: " + e;
        const int total_tiles = P.tiles_w * P.tiles_h * N;
        const int slots = num_sms * L.occ;
        const int waves = (total_tiles + slots - 1) / slots;
        L.grid = (total_tiles + waves - 1) / waves;
        return "";
    }
    int nt;
    if (parity) nt = 4;
    else {
        nt = cout == 16 ? 4 : 2;
        while (nt > 1 && 8 * (nt / 2) >= W) nt /= 2;  // narrow images: do not allocate sub-tiles that are all padding
    }
    P.nt = nt;
    const int src_h = parity ? H / 2 : H, src_w = parity ? W / 2 : W;
    P.tiles_w = parity ? (src_w + 7) / 8 : (W + 8 * nt - 1) / (8 * nt);
    P.tiles_h = (src_h + 15) / 16;
    P.halo_w = parity ? 10 : 8 * nt + 2;
    P.tx_bytes = (uint32_t)P.halo_w * kTcHaloH * row_bytes;
    P.stage_bytes = (P.tx_bytes + 1023u) & ~1023u;
    P.w_bytes = (uint32_t)tconv_w_elems(cin, cout, parity) * 2;
    P.wpk = reinterpret_cast<const __nv_bfloat16*>(wpk);
    P.scale = ep.scale; P.shift = ep.shift; P.relu = ep.relu;
    P.out = reinterpret_cast<__nv_bfloat16*>(out);
    P.residual = ep.residual_f32 ? nullptr : reinterpret_cast<const __nv_bfloat16*>(ep.residual.ptr);
    P.residual32 = ep.residual_f32 ? reinterpret_cast<const float*>(ep.residual.ptr) : nullptr;
    P.stats = ep.stats;
    P.err = err;
    // occupancy: two CTAs per SM (8 epilogue warps each) when TMEM (<= 256 columns each), the accumulator column
    // groups (<= 4: two per epilogue thread) and shared memory allow; otherwise one CTA with 16 epilogue warps
    const int items = nt * (cout / 16);
    if (items > 8) return "tconv: too many accumulator column groups";
    const int acc_cols = nt * cout;
    // cout >= 32: whole sub-tiles leave through swizzled smem staging + TMA store (see TconvParams::stage_out)
    // parity mode: with 256-bit stores the four parity sub-tiles also leave straight from registers (a lane writes the 32
    // bytes of its pixel, every second pixel of a row: full 32-byte sectors that the L2 merges with the other parity's);
    // decoder.blocks.4.conv1 96 -> 88 us, decoder.blocks.3.conv1 [up] 63 -> 57 us at batch 32.  UNETB200_PARITY_STAGE=1
    // restores the interleaved smem tile + single TMA store of round 1 for A/B runs.
    static const bool par_stage = getenv("UNETB200_PARITY_STAGE") != nullptr;
    // Plain mode stores straight from registers: ONE 256-bit store (STG.E.256) per lane and 16-channel item.  The
    // swizzled-smem + TMA-store staging it replaces for cout >= 32 was introduced when the two 128-bit stores per item
    // serialised in the LSU; with 256-bit stores the register path is as fast or faster (layer1 conv2 61.5 -> 58 us,
    // decoder.blocks.2 [skip] 61.5 -> 57.5 us at batch 32) and frees the staging buffers.  UNETB200_TC_STAGE=1 restores
    // the staged epilogue for A/B runs.
    static const bool reg_store = getenv("UNETB200_TC_STAGE") == nullptr;
    const uint32_t out_bytes = ((parity && par_stage) || (!parity && cout >= 32 && !reg_store)) ? (uint32_t)nt * 128u * cout * 2u : 0u;
    L.occ = 1;
    int stages = 0;
    if (items <= 4 && 2 * acc_cols <= 256) {
        for (int st = 4; st >= 2; --st)
            if (2 * (tconv_smem(P.w_bytes, P.stage_bytes, st, out_bytes).total + 1024 + 1024) <= 232448u) {
                L.occ = 2;
                stages = st;
                break;
            }
    }
    if (L.occ == 1)
        for (int st = 6; st >= 2; --st)
            if (tconv_smem(P.w_bytes, P.stage_bytes, st, out_bytes).total + 1024 <= 232448u) {
                stages = st;
                break;
            }
    const int groups = tc_epi_warps(L.occ) / 4;
    L.iph = (items + groups - 1) / groups;
    const int cgs = cout / 16;
    if (out_bytes && parity && items % groups == 0) {
        P.stage_out = 2;
        uint64_t dims[4] = {(uint64_t)cout, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)cout * 2, (uint64_t)W * cout * 2, (uint64_t)H * W * cout * 2};
        uint32_t box[4] = {(uint32_t)cout, 16, 32, 1};
        uint32_t es[4] = {1, 1, 1, 1};
        std::string e = make_tmap_bf16(&L.d, out, 4, dims, str, box, es, swizzle_for_bytes(cout * 2));
        if (!e.empty()) return "tconv D map: " + e;
    } else if (out_bytes && !parity && L.iph <= cgs && cgs % L.iph == 0 && items % groups == 0) {
        P.stage_out = 1;
        P.spw = (cgs / L.iph) * 128;
        uint64_t dims[4] = {(uint64_t)cout, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)cout * 2, (uint64_t)W * cout * 2, (uint64_t)H * W * cout * 2};
        uint32_t box[4] = {(uint32_t)cout, 8, 16, 1};
        uint32_t es[4] = {1, 1, 1, 1};
        std::string e = make_tmap_bf16(&L.d, out, 4, dims, str, box, es, swizzle_for_bytes(cout * 2));
        if (!e.empty()) return "tconv D map: " + e;
    }
    if (stages < 2) return "tconv: does not fit in shared memory";
    P.stages = stages;
    const int tmem_budget = L.occ == 2 ? 256 : 512;
    P.nacc = tmem_budget / acc_cols >= 4 ? 4 : (tmem_budget / acc_cols >= 2 ? 2 : 0);
    if (!P.nacc) return "tconv: accumulators do not fit in TMEM";
    L.smem = tconv_smem(P.w_bytes, P.stage_bytes, stages, P.stage_out ? out_bytes : 0u).total + 1024;
    if (!P.stage_out) L.d = CUtensorMap{};
    {
        uint64_t dims[4] = {(uint64_t)cin, (uint64_t)src_w, (uint64_t)src_h, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)cin * 2, (uint64_t)src_w * cin * 2, (uint64_t)src_h * src_w * cin * 2};
        uint32_t box[4] = {(uint32_t)cin, (uint32_t)P.halo_w, (uint32_t)kTcHaloH, 1};
        uint32_t es[4] = {1, 1, 1, 1};
        std::string e = make_tmap_bf16(&L.a, src, 4, dims, str, box, es, swizzle_for_bytes(row_bytes));
        if (!e.empty()) return "tconv A map: " + e;
    }
    const int total_tiles = P.tiles_w * P.tiles_h * N;
    const int slots = num_sms * L.occ;
    const int waves = (total_tiles + slots - 1) / slots;
    L.grid = (total_tiles + waves - 1) / waves;
    return "";
}

// 7x7/s2 stem over the packed image xp[N][H][W+8][4] (H, W = INPUT extent) -> out[N, H/2, W/2, 64]; wpk = PK_STEM2 image
inline std::string tconv_build_stem(TconvLaunch& L, const void* xp, const void* wpk, int N, int H, int W, void* out,
                                    const EpilogueDesc& ep, int* err, int num_sms) {
    memset(&L.p, 0, sizeof(L.p));
    if (tconv_const_weights_flag()) L.p.dbg |= 16;
    TconvParams& P = L.p;
    if ((H | W) & 1) return "tconv stem: H and W must be even";
    if (ep.residual.ptr) return "tconv stem: no residual";
    const int Ho = H / 2, Wo = W / 2, nt = 2;
    P.H = Ho; P.W = Wo; P.N = N;
    P.mode = 2;
    P.nt = nt;
    P.cin = 16;   // unused by the stem issue path except for K-step bookkeeping
    P.cout = 64;
    P.tiles_w = (Wo + 127) / 128;
    P.tiles_h = (Ho + nt - 1) / nt;
    P.halo_w = 136;                                  // 16-byte pixel pairs per staged input row (128 + 3 window overhang,
                                                     // rounded up to 17 x 128 B: TMA rows of 128 B, not 16 B)
    const int rows = 2 * nt + 5;
    P.tx_bytes = (uint32_t)P.halo_w * 16 * rows;
    P.stage_bytes = (P.tx_bytes + 1023u) & ~1023u;
    P.w_bytes = 64 * 224 * 2;
    P.wpk = reinterpret_cast<const __nv_bfloat16*>(wpk);
    P.scale = ep.scale; P.shift = ep.shift; P.relu = ep.relu;
    P.out = reinterpret_cast<__nv_bfloat16*>(out);
    P.stats = ep.stats;
    P.err = err;
    L.occ = 1;
    L.iph = 2;                                       // 2 sub-tiles x 4 column groups over 16 epilogue warps
    P.stages = 4;
    P.nacc = 4;
    static const bool stem_reg = getenv("UNETB200_STEM_REGSTORE") != nullptr;   // A/B: 256-bit register stores instead
    P.stage_out = stem_reg ? 0 : 1;                  // one output row segment [64 ch x 128 px] per sub-tile through TMA
    P.spw = 256;
    L.smem = tconv_smem(P.w_bytes, P.stage_bytes, P.stages, stem_reg ? 0u : (uint32_t)nt * 128u * 128u).total + 1024;
    {
        uint64_t dims[4] = {64, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)N};
        uint64_t str[3] = {128, (uint64_t)Wo * 128, (uint64_t)Ho * Wo * 128};
        uint32_t box[4] = {64, 128, 1, 1};
        uint32_t es[4] = {1, 1, 1, 1};
        std::string e = make_tmap_bf16(&L.d, out, 4, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_128B);
        if (!e.empty()) return "tconv stem D map: " + e;
    }
    {
        // view of xp as 128-byte groups of 8 pixel pairs: [64 elements][ceil((W+8)/16) groups][H rows][N].  The last group
        // of a row may run up to 120 B past the row end (into the next row; for the very last row into the 128 B of slack
        // every xp allocation carries): those pairs only feed output columns >= Wo, which are masked.
        uint64_t dims[4] = {64, (uint64_t)(W + 8 + 15) / 16, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {128, (uint64_t)(W + 8) * 8, (uint64_t)H * (W + 8) * 8};
        uint32_t box[4] = {64, 17, (uint32_t)rows, 1};
        uint32_t es[4] = {1, 1, 1, 1};
        std::string e = make_tmap_bf16(&L.a, xp, 4, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (!e.empty()) return "tconv stem A map: " + e;
    }
    const int total_tiles = P.tiles_w * P.tiles_h * N;
    const int waves = (total_tiles + num_sms - 1) / num_sms;
    L.grid = (total_tiles + waves - 1) / waves;
    return "";
}

template <int kOcc, int kIph, bool kStage, bool kStack = false>
inline cudaError_t tconv_launch_t(const TconvLaunch& L, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tconv_kernel<kOcc, kIph, kStage, false, kStack>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, kOcc == 2 ? 115200 : 232448);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    launch_k(tconv_kernel<kOcc, kIph, kStage, false, kStack>, L.grid, tc_threads(kOcc), L.smem, st, L.a, L.d, L.p);
    return cudaGetLastError();
}
// seg head Conv2d(16, 1, 3, padding=1) over src[N,H,W,16]: tconv mode 3 (three filter columns stacked as N-blocks, PK_HEAD
// weights; tc_issue_head3 in tconv.cuh).  Tile = 16 rows x 30 output columns (4 sub-tiles of 4 rows x 32 halo pixels).
inline std::string tconv_build_head(TconvLaunch& L, const void* src, const void* wpk, int N, int H, int W, int* err,
                                    int num_sms) {
    memset(&L.p, 0, sizeof(L.p));
    if (tconv_const_weights_flag()) L.p.dbg |= 16;
    TconvParams& P = L.p;
    const int nt = 4;
    const uint32_t row_bytes = 32;
    P.H = H; P.W = W; P.N = N;
    P.mode = 3;
    P.cin = 16; P.cout = 16;
    P.nt = nt;
    P.tiles_w = (W + 29) / 30;
    P.tiles_h = (H + 4 * nt - 1) / (4 * nt);
    P.halo_w = 32;
    P.tx_bytes = 32u * kTcHaloH * row_bytes;         // kTcHaloH = 4 * nt + 2 rows
    P.stage_bytes = (P.tx_bytes + 1023u) & ~1023u;
    P.w_bytes = 3 * 16 * 16 * 2;
    P.wpk = reinterpret_cast<const __nv_bfloat16*>(wpk);
    P.err = err;
    L.occ = 2;
    L.iph = 2;                                       // 4 sub-tiles over 8 epilogue warps (two per TMEM lane quadrant)
    int stages = 0;
    for (int st = 4; st >= 2; --st)
        if (2 * (tconv_smem(P.w_bytes, P.stage_bytes, st, 0).total + 1024 + 1024) <= 232448u) {
            stages = st;
            break;
        }
    if (stages < 2) return "tconv head: does not fit in shared memory";
    P.stages = stages;
    P.nacc = 4;                                      // 4 x 64 accumulator columns of the CTA's 256
    L.smem = tconv_smem(P.w_bytes, P.stage_bytes, stages, 0).total + 1024;
    L.d = CUtensorMap{};
    {
        uint64_t dims[4] = {16, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {32, (uint64_t)W * 32, (uint64_t)H * W * 32};
        uint32_t box[4] = {16, 32, (uint32_t)kTcHaloH, 1};
        uint32_t es[4] = {1, 1, 1, 1};
        std::string e = make_tmap_bf16(&L.a, src, 4, dims, str, box, es, swizzle_for_bytes(row_bytes));
        if (!e.empty()) return "tconv head A map: " + e;
    }
    const int total_tiles = P.tiles_w * P.tiles_h * N;
    const int slots = num_sms * L.occ;
    const int waves = (total_tiles + slots - 1) / slots;
    L.grid = (total_tiles + waves - 1) / waves;
    return "";
}
// seg head launch: the epilogue writes logits / prob / mask; outputs are set per call
inline cudaError_t tconv_launch_head(const TconvLaunch& L, float* logits, float* prob, uint8_t* mask, float thresh_logit,
                                     cudaStream_t st) {
    if (L.occ != 2 || L.iph != 2 || L.p.stage_out || L.p.mode != 3) return cudaErrorInvalidValue;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tconv_kernel<2, 2, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 115200);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    TconvParams p = L.p;
    p.logits = logits; p.prob = prob; p.mask = mask; p.thresh_logit = thresh_logit;
    launch_k(tconv_kernel<2, 2, false, true, true>, L.grid, tc_threads(2), L.smem, st, L.a, L.d, p);
    return cudaGetLastError();
}

inline cudaError_t tconv_launch(const TconvLaunch& L, cudaStream_t st) {
    if (L.p.mode == 3)   // stacked filter columns (tconv_build)
        return L.occ == 2 ? tconv_launch_t<2, 1, false, true>(L, st) : tconv_launch_t<1, 1, false, true>(L, st);
    const int key = (L.occ == 2 ? 4 : 0) | (L.iph == 2 ? 2 : 0) | (L.p.stage_out ? 1 : 0);
    switch (key) {
        case 0: return tconv_launch_t<1, 1, false>(L, st);
        case 1: return tconv_launch_t<1, 1, true>(L, st);
        case 2: return tconv_launch_t<1, 2, false>(L, st);
        case 3: return tconv_launch_t<1, 2, true>(L, st);
        case 4: return tconv_launch_t<2, 1, false>(L, st);
        case 5: return tconv_launch_t<2, 1, true>(L, st);
        case 6: return tconv_launch_t<2, 2, false>(L, st);
        default: return tconv_launch_t<2, 2, true>(L, st);
    }
}

// ------------------------------------------------------------------------------------------------ wconv (wide halo conv)
struct WconvLaunch {
    CUtensorMap a, b;
    WconvParams p;
    int grid = 0;
    uint32_t smem = 0;
};
inline bool wconv_ok(int cin, int cout) { return cin >= 64 && cin % 64 == 0 && cout >= 128 && cout % 128 == 0 && cout <= 512; }

// 3x3/s1/p1 conv src[N,H,W,cin] -> out[N,H,W,cout]; wpk = K-major packed weights [cout][(r*3+s)*cin + ci] (pack_conv_w)
// ldb: row pitch (elements) of the packed weight matrix when it is wider than 9*cin (a decoder conv1's PK_DEC1 matrix,
// whose first 9*cskip columns are exactly the skip-channel operand); 0 = dense.
inline std::string wconv_build(WconvLaunch& L, const void* src, int cin, const void* wpk, int cout, int N, int H, int W,
                               void* out, const EpilogueDesc& ep, int* err, int num_sms, long long ldb = 0) {
    memset(&L.p, 0, sizeof(L.p));
    L.p.const_w = tconv_const_weights_flag() ? 1 : 0;
    WconvParams& P = L.p;
    if (!wconv_ok(cin, cout)) return "wconv: unsupported channel configuration";
    if (ep.residual.ptr && (ep.residual.sW != cout || ep.residual.sH != (long long)W * cout ||
                            ep.residual.sN != (long long)H * W * cout))
        return "wconv: residual must be a dense NHWC tensor of the output's shape";
    P.H = H; P.W = W; P.N = N;
    P.tiles_w = (W + 15) / 16;
    P.tiles_h = (H + 15) / 16;
    P.n_tiles = cout / kWcN;
    P.cin = cin; P.cout = cout;
    P.scale = ep.scale; P.shift = ep.shift; P.relu = ep.relu;
    P.out = reinterpret_cast<__nv_bfloat16*>(out);
    P.residual = ep.residual_f32 ? nullptr : reinterpret_cast<const __nv_bfloat16*>(ep.residual.ptr);
    P.residual32 = ep.residual_f32 ? reinterpret_cast<const float*>(ep.residual.ptr) : nullptr;
    P.stats = ep.stats;
    P.err = err;
    int bst = 9;   // multiple of 3: the producer waits once per group of 3 weight tiles
    while (bst > 3 && wconv_smem(bst).total + 1024 > 232448u) bst -= 3;
    P.bstages = bst;
    L.smem = wconv_smem(bst).total + 1024;
    {
        uint64_t dims[4] = {(uint64_t)cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)cin * 2, (uint64_t)W * cin * 2, (uint64_t)H * W * cin * 2};
        uint32_t box[4] = {64, 18, 18, 1};
        uint32_t es[4] = {1, 1, 1, 1};
        std::string e = make_tmap_bf16(&L.a, src, 4, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_128B);
        if (!e.empty()) return "wconv A map: " + e;
    }
    {
        uint64_t dims[2] = {(uint64_t)9 * cin, (uint64_t)cout};
        uint64_t str[1] = {(uint64_t)(ldb ? ldb : 9ll * cin) * 2};
        uint32_t box[2] = {64, (uint32_t)kWcN};
        uint32_t es[2] = {1, 1};
        std::string e = make_tmap_bf16(&L.b, wpk, 2, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_128B);
        if (!e.empty()) return "wconv B map: " + e;
    }
    const int total = P.tiles_w * P.tiles_h * N * P.n_tiles;
    const int waves = (total + num_sms - 1) / num_sms;
    L.grid = (total + waves - 1) / waves;
    return "";
}
inline cudaError_t wconv_launch(const WconvLaunch& L, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(wconv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(wconv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    if (L.p.residual32) launch_k(wconv_kernel<true>, L.grid, kWcThreads, L.smem, st, L.a, L.b, L.p);
    else launch_k(wconv_kernel<false>, L.grid, kWcThreads, L.smem, st, L.a, L.b, L.p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ wpconv (wide parity conv)
struct WpconvLaunch {
    CUtensorMap a, b;
    WpconvParams p;
    int grid = 0;
    uint32_t smem = 0;
};
inline bool wpconv_ok(int cup, int cout) { return cup >= 64 && cup % 64 == 0 && cout >= 64 && cout % 64 == 0 && cout <= 512; }

// out[N, 2Hl, 2Wl, cout] = scale * conv3x3(nearest2x(low[N,Hl,Wl,cup])) by parity folding.  wpk_dec1 = the PK_DEC1 matrix
// [4 parities][cout][kt], kt = 9*cskip + 4*cup (low taps start at column koff = 9*cskip).
inline std::string wpconv_build(WpconvLaunch& L, const void* low, int cup, const void* wpk_dec1, int kt, int koff, int cout,
                                int N, int Hl, int Wl, void* out, const float* scale, int* err, int num_sms,
                                bool out_f32 = false) {
    memset(&L.p, 0, sizeof(L.p));
    WpconvParams& P = L.p;
    if (!wpconv_ok(cup, cout)) return "wpconv: unsupported channel configuration";
    P.Hl = Hl; P.Wl = Wl; P.N = N;
    P.tiles_w = (Wl + 15) / 16;
    P.tiles_h = (Hl + 15) / 16;
    P.n_tiles = cout / 64;
    P.cup = cup; P.cout = cout; P.koff = koff;
    P.scale = scale;
    P.out = out_f32 ? nullptr : reinterpret_cast<__nv_bfloat16*>(out);
    P.out32 = out_f32 ? reinterpret_cast<float*>(out) : nullptr;
    P.err = err;
    int bst = 16;  // multiple of 4: the producer waits once per group of 4 weight tiles
    while (bst > 4 && wpconv_smem(bst).total + 1024 > 232448u) bst -= 4;
    P.bstages = bst;
    L.smem = wpconv_smem(bst).total + 1024;
    {
        uint64_t dims[4] = {(uint64_t)cup, (uint64_t)Wl, (uint64_t)Hl, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)cup * 2, (uint64_t)Wl * cup * 2, (uint64_t)Hl * Wl * cup * 2};
        uint32_t box[4] = {64, 18, 18, 1};
        uint32_t es[4] = {1, 1, 1, 1};
        std::string e = make_tmap_bf16(&L.a, low, 4, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_128B);
        if (!e.empty()) return "wpconv A map: " + e;
    }
    {
        uint64_t dims[2] = {(uint64_t)kt, (uint64_t)4 * cout};
        uint64_t str[1] = {(uint64_t)kt * 2};
        uint32_t box[2] = {64, 64};
        uint32_t es[2] = {1, 1};
        std::string e = make_tmap_bf16(&L.b, wpk_dec1, 2, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_128B);
        if (!e.empty()) return "wpconv B map: " + e;
    }
    const int total = P.tiles_w * P.tiles_h * N * P.n_tiles;
    const int waves = (total + num_sms - 1) / num_sms;
    L.grid = (total + waves - 1) / waves;
    return "";
}
inline cudaError_t wpconv_launch(const WpconvLaunch& L, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(wpconv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(wpconv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    if (L.p.out32) launch_k(wpconv_kernel<true>, L.grid, kWcThreads, L.smem, st, L.a, L.b, L.p);
    else launch_k(wpconv_kernel<false>, L.grid, kWcThreads, L.smem, st, L.a, L.b, L.p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ hwgrad (halo-resident wgrad)
struct HwgradLaunch {
    HwgradParams p;
    dim3 grid;
    uint32_t smem = 0;
};
inline bool hwgrad_ok(int cup, int cskip, int cout) {
    const int ctot = cup + cskip;
    if (!(cout == 16 || cout == 32 || cout == 64)) return false;
    if (!(ctot == 16 || ctot == 32 || ctot == 64 || ctot == 128) || cup % 8 || cskip % 8) return false;
    if (ctot == 128 && cup % 64) return false;  // a 64-channel block must not straddle the two sources
    return true;
}
inline std::string hwgrad_build(HwgradLaunch& L, const void* low, int cup, const void* src, int cskip, const void* dz,
                                int cout, int N, int H, int W, float* gpk, int* err, int num_sms, int gtot = 0) {
    memset(&L.p, 0, sizeof(L.p));
    if (!hwgrad_ok(cup, cskip, cout)) return "hwgrad: unsupported channel configuration";
    if (cup && ((H | W) & 1)) return "hwgrad: up-sampled source needs even H, W";
    HwgradParams& P = L.p;
    const int ctot = cup + cskip;
    P.H = H; P.W = W; P.N = N;
    P.tiles_w = (W + kHcTileW - 1) / kHcTileW;
    P.tiles_h = (H + kHcTileH - 1) / kHcTileH;
    P.cup = cup; P.cskip = cskip;
    P.low = reinterpret_cast<const __nv_bfloat16*>(low);
    P.src = reinterpret_cast<const __nv_bfloat16*>(src);
    P.dz = reinterpret_cast<const __nv_bfloat16*>(dz);
    P.cout = cout;
    P.cw = ctot < 64 ? ctot : 64;
    P.ndh_max = 9 * P.cw <= 512 ? 3 : (6 * P.cw <= 512 ? 2 : 1);
    P.gpk = gpk;
    P.gtot = gtot ? gtot : ctot;
    P.err = err;
    int stages = 4;
    for (; stages >= 2; --stages)
        if (hwgrad_smem(P.cw, cout, stages).total + 1024 <= 232448u) break;
    if (stages < 2) return "hwgrad: does not fit in shared memory";
    P.stages = stages;
    L.smem = hwgrad_smem(P.cw, cout, stages).total + 1024;
    const int ngroups = (ctot / P.cw) * ((3 + P.ndh_max - 1) / P.ndh_max);
    const int total_tiles = P.tiles_w * P.tiles_h * N;
    int gx = num_sms / ngroups;
    if (gx < 1) gx = 1;
    if (gx > total_tiles) gx = total_tiles;
    L.grid = dim3(gx, ngroups, 1);
    return "";
}
inline cudaError_t hwgrad_launch(const HwgradLaunch& L, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(hwgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    launch_k(hwgrad_kernel, L.grid, kHcThreads, L.smem, st, L.p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ xwgrad (TMA halo wgrad)
struct XwgradLaunch {
    CUtensorMap z, x;
    XwgradParams p;
    dim3 grid;
    uint32_t smem = 0;
};
// dZ [N,H,W,cout] x X [N,H,W,cin] (both dense NHWC bf16) -> gpk[cout][9][ctot] columns [dci0, dci0 + cin)
inline bool xwgrad_ok(int cin, int cout, int H, int W) {
    if (!(cin == 16 || cin == 32 || cin % 64 == 0)) return false;
    if (!(cout == 16 || cout == 32 || cout == 64 || cout % 128 == 0)) return false;
    if (getenv("UNETB200_NO_XWGRAD")) return false;
    return W % kXwTileW == 0 && H % 2 == 0 && H >= 8;
}
// up = true: x is the LOW-resolution tensor [N,H,W,cin] of a nearest-2x up-sample and dz is [N,2H,2W,cout]
inline std::string xwgrad_build(XwgradLaunch& L, const void* x, int cin, const void* dz, int cout, int N, int H, int W,
                                float* gpk, int ctot, int dci0, int* err, int num_sms, bool up = false) {
    memset(&L.p, 0, sizeof(L.p));
    if (!xwgrad_ok(cin, cout, H, W)) return "xwgrad: unsupported configuration";
    XwgradParams& P = L.p;
    P.H = H; P.W = W; P.N = N;
    P.up = up ? 1 : 0;
    P.wide = cout >= 128;
    P.co_blk = P.wide ? 128 : (cout < 32 ? cout : 32);
    P.zc_box = P.co_blk < 64 ? P.co_blk : 64;
    P.n_zbox = P.co_blk / P.zc_box;
    P.cw = P.wide ? (cin < 32 ? cin : 32) : (cin < 64 ? cin : 64);
    P.th = (H % 32 == 0) ? 32 : (H % 16 == 0 ? 16 : 8);
    P.tiles_w = W / kXwTileW;
    P.tiles_h = (H + P.th - 1) / P.th;
    P.cout = cout; P.cin = cin;
    P.ctot = ctot; P.dci0 = dci0;
    P.gpk = gpk;
    P.err = err;
    if ((P.wide ? 3 : 1) * 3 * P.cw > 512) return "xwgrad: accumulators do not fit in TMEM";
    int stages = 6;
    for (; stages >= 2; --stages)
        if (xwgrad_smem(P.th, P.zc_box, P.n_zbox, P.cw, stages).total + 1024 <= 232448u) break;
    if (stages < 2) return "xwgrad: does not fit in shared memory";
    P.stages = stages;
    L.smem = xwgrad_smem(P.th, P.zc_box, P.n_zbox, P.cw, stages).total + 1024;
    {
        const uint64_t m = up ? 2 : 1;   // up mode: every second pixel / row of the high-resolution gradient
        uint64_t dims[4] = {(uint64_t)cout, m * W, m * H, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)cout * 2, m * W * cout * 2, m * H * m * W * cout * 2};
        uint32_t box[4] = {(uint32_t)P.zc_box, (uint32_t)(m * kXwTileW), (uint32_t)(m * (P.th + 2)), 1};
        uint32_t es[4] = {1, (uint32_t)m, (uint32_t)m, 1};
        std::string e = make_tmap_bf16(&L.z, dz, 4, dims, str, box, es, swizzle_for_bytes(P.zc_box * 2));
        if (!e.empty()) return "xwgrad dZ map: " + e;
    }
    {
        uint64_t dims[4] = {(uint64_t)cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)cin * 2, (uint64_t)W * cin * 2, (uint64_t)H * W * cin * 2};
        uint32_t box[4] = {(uint32_t)P.cw, (uint32_t)(kXwTileW + 2), (uint32_t)P.th, 1};
        uint32_t es[4] = {1, 1, 1, 1};
        std::string e = make_tmap_bf16(&L.x, x, 4, dims, str, box, es, swizzle_for_bytes(P.cw * 2));
        if (!e.empty()) return "xwgrad X map: " + e;
    }
    const int ngroups = (cout / P.co_blk) * (cin / P.cw);
    const int total_tiles = P.tiles_w * P.tiles_h * N;
    int gx = num_sms / (ngroups * (up ? 4 : 1));
    if (gx < 1) gx = 1;
    if (gx > total_tiles) gx = total_tiles;
    L.grid = dim3(gx, ngroups, up ? 4 : 1);
    return "";
}
inline cudaError_t xwgrad_launch(const XwgradLaunch& L, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(xwgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    launch_k(xwgrad_kernel, L.grid, kXwThreads, L.smem, st, L.z, L.x, L.p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ dlow (decoder conv1 -> low-res dgrad)
struct DlowLaunch {
    CUtensorMap z, w;
    DlowParams p;
    int grid = 0;
    uint32_t smem = 0;
};
inline bool dlow_ok(int cz, int cup, int Hl, int Wl) {
    if (getenv("UNETB200_NO_DLOW")) return false;
    return (cz == 16 || cz == 32) && (cup == 32 || cup == 64) && Wl % kDlTileW == 0 && Hl >= 2 && Wl >= 8;
}
// dz: [N, 2*Hl, 2*Wl, cz]; wdl: PK_DLOW operand [cup][16 * cz] bf16; out: [N, Hl, Wl, cup]
inline std::string dlow_build(DlowLaunch& L, const void* dz, int cz, const void* wdl, int cup, int N, int Hl, int Wl,
                              void* out, int* err, int num_sms) {
    memset(&L.p, 0, sizeof(L.p));
    if (!dlow_ok(cz, cup, Hl, Wl)) return "dlow: unsupported configuration";
    DlowParams& P = L.p;
    P.Hl = Hl; P.Wl = Wl; P.N = N;
    P.tiles_w = Wl / kDlTileW;
    P.tiles_h = (Hl + kDlTileH - 1) / kDlTileH;
    P.cz = cz; P.cup = cup;
    P.out = reinterpret_cast<__nv_bfloat16*>(out);
    P.err = err;
    if (const char* e = getenv("UNETB200_DLOW_DBG")) P.dbg = atoi(e);
    int stages = 6;
    for (; stages >= 2; --stages)
        if (dlow_smem(cz, cup, stages).total + 1024 <= 232448u) break;
    if (stages < 2) return "dlow: does not fit in shared memory";
    P.stages = stages;
    L.smem = dlow_smem(cz, cup, stages).total + 1024;
    {
        // pixel-pair view of the dense NHWC gradient: [2*cz, Wl pairs, 2*Hl rows, N]
        uint64_t dims[4] = {2ull * cz, (uint64_t)Wl, 2ull * Hl, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)cz * 4, 2ull * Wl * cz * 2, 4ull * Hl * Wl * cz * 2};
        uint32_t box[4] = {2u * cz, (uint32_t)kDlBoxW, (uint32_t)kDlBoxH, 1};
        uint32_t es[4] = {1, 1, 1, 1};
        std::string e = make_tmap_bf16(&L.z, dz, 4, dims, str, box, es, swizzle_for_bytes(cz * 4));
        if (!e.empty()) return "dlow dZ map: " + e;
    }
    {
        uint64_t dims[2] = {16ull * cz, (uint64_t)cup};
        uint64_t str[1] = {16ull * cz * 2};
        uint32_t box[2] = {(uint32_t)cz, (uint32_t)cup};
        uint32_t es[2] = {1, 1};
        std::string e = make_tmap_bf16(&L.w, wdl, 2, dims, str, box, es, swizzle_for_bytes(cz * 2));
        if (!e.empty()) return "dlow weight map: " + e;
    }
    const int total_tiles = P.tiles_w * P.tiles_h * N;
    L.grid = total_tiles < num_sms ? total_tiles : num_sms;
    return "";
}
inline cudaError_t dlow_launch(const DlowLaunch& L, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(dlow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    launch_k(dlow_kernel, L.grid, kDlThreads, L.smem, st, L.z, L.w, L.p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ swgrad (stem weight gradient)
struct SwgradLaunch {
    CUtensorMap z, x;
    SwgradParams p;
    int grid = 0;
    uint32_t smem = 0;
};
inline bool swgrad_ok(int H, int W) { return !getenv("UNETB200_NO_SWGRAD") && H % 16 == 0 && W % 32 == 0 && H >= 16 && W >= 32; }
// xp: packed image [N][H][W+8][4] bf16; dz: [N, H/2, W/2, 64]; grad: OIHW fp32 [64][3][7][7] (accumulated)
inline std::string swgrad_build(SwgradLaunch& L, const void* xp, const void* dz, int N, int H, int W, float* grad, int* err,
                                int num_sms) {
    memset(&L.p, 0, sizeof(L.p));
    if (!swgrad_ok(H, W)) return "swgrad: unsupported extent";
    SwgradParams& P = L.p;
    P.Ho = H / 2; P.Wo = W / 2; P.N = N;
    P.wt = P.Wo % 64 == 0 ? 64 : (P.Wo % 32 == 0 ? 32 : 16);
    P.tiles_w = P.Wo / P.wt;
    P.tiles_q = (P.Ho + kSwTQ - 1) / kSwTQ;
    P.grad = grad;
    P.err = err;
    int stages = 4;
    for (; stages >= 2; --stages)
        if (swgrad_smem(P.wt, stages).total + 1024 <= 232448u) break;
    if (stages < 2) return "swgrad: does not fit in shared memory";
    P.stages = stages;
    L.smem = swgrad_smem(P.wt, stages).total + 1024;
    {
        uint64_t dims[4] = {64, (uint64_t)P.Wo, (uint64_t)P.Ho, (uint64_t)N};
        uint64_t str[3] = {128, (uint64_t)P.Wo * 128, (uint64_t)P.Ho * P.Wo * 128};
        uint32_t box[4] = {64, (uint32_t)P.wt, (uint32_t)(kSwTQ + 3), 1};
        uint32_t es[4] = {1, 1, 1, 1};
        std::string e = make_tmap_bf16(&L.z, dz, 4, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_128B);
        if (!e.empty()) return "swgrad dZ map: " + e;
    }
    {
        // window view: "pixel" wo of an input row = 32 elements (8 input pixels x 4 channels) from padded column 2*wo
        const uint64_t Wp = (uint64_t)W + 8;
        uint64_t dims[4] = {32, (uint64_t)P.Wo, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {16, Wp * 8, (uint64_t)H * Wp * 8};
        uint32_t box[4] = {32, (uint32_t)P.wt, (uint32_t)(2 * kSwTQ), 1};
        uint32_t es[4] = {1, 1, 1, 1};
        std::string e = make_tmap_bf16(&L.x, xp, 4, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_64B);
        if (!e.empty()) return "swgrad X map: " + e;
    }
    const int total_tiles = P.tiles_w * P.tiles_q * N;
    L.grid = total_tiles < num_sms ? total_tiles : num_sms;
    return "";
}
inline cudaError_t swgrad_launch(const SwgradLaunch& L, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(swgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    launch_k(swgrad_kernel, L.grid, kSwThreads, L.smem, st, L.z, L.x, L.p);
    return cudaGetLastError();
}

}  // namespace ub
