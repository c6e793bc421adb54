// Halo-resident 3x3 / stride-1 / pad-1 convolution on tcgen05 (sm_100a) for the narrow, high-resolution layers
// (Cout <= 64: decoder blocks 2-4, encoder.layer1 and their data gradients) — SURVEY.md section 8a rows A3, A5.
//
// Why a second conv kernel: in the tap-table kernel (igemm.cuh) every filter tap re-reads its 128-pixel operand tile and
// its weight slice through L2; for Cin <= 64 that L2->SM traffic (9x the input) and the TMA's per-row request rate bound
// the kernel at 4-25 % of either roofline (profiles/r1_igemm_ncu_full_summary.csv).  Here
//   * the weights of ALL taps stay resident in shared memory for the whole persistent CTA,
//   * each 8 x 16 output tile loads its 10 x 18 input halo ONCE (cp.async, 16 B per request, coalesced along the
//     pixel's channels) as K-major rows [halo pixel][<= 64 channels] in the 32/64/128-byte swizzled layout the tensor
//     core reads at full shared-memory rate (the swizzle is an XOR of smem address bits, applied here by the loader),
//     so that a filter tap is just a whole-row shift of the UMMA descriptor's start address, and
//   * the decoder's nearest-2x upsample + concat is folded into the loader's address computation: halo pixel (h, w) of
//     the up-sampled channels is fetched from low-res pixel (h>>1, w>>1) — nothing is materialised in HBM.
// Roles: warps 0-7 cp.async producers, warp 8 tcgen05 issuer + TMEM owner, warps 9-12 epilogue.  The tiles are small
// (128 pixels x <= 64 channels), so every role's per-tile instruction stream is kept short: incremental tile
// coordinates (no integer division), fully unrolled MMA issue, and an epilogue that stores straight from registers
// (one pixel = one thread = 32..128 contiguous bytes) instead of staging through shared memory and TMA.
#pragma once
#include "ptx.cuh"

namespace ub {

constexpr int kHcProducers = 256;                              // 8 cp.async warps (2 per scheduler: latency hiding)
constexpr int kHcThreads = kHcProducers + 32 + 128;
constexpr int kHcTileW = 8, kHcTileH = 16;                    // 128 output pixels = UMMA M
constexpr int kHcHaloW = kHcTileW + 2, kHcHaloH = kHcTileH + 2;
constexpr int kHcHaloPx = kHcHaloW * kHcHaloH;                // 180

struct HconvParams {
    int H, W, N;                    // output extent
    int tiles_w, tiles_h;
    int cup, cskip;                 // channels taken from the low-res (nearest-2x) source / from the direct source
    const __nv_bfloat16* low;       // [N, H/2, W/2, cup] or nullptr
    const __nv_bfloat16* src;       // [N, H, W, cskip] or nullptr
    const __nv_bfloat16* wpk;       // [9 taps][ctot/64 blocks][cout][<= 64 ch], 16-byte chunks pre-swizzled (pack_hconv_w_kernel)
    int cout;
    int stages;
    const float* scale;             // [cout] or nullptr (identity)
    const float* shift;
    int relu;
    __nv_bfloat16* out;             // [N, H, W, cout] dense NHWC
    const __nv_bfloat16* residual;  // same shape as out, or nullptr
    float* stats;                   // [gridDim.x][cout][2] per-CTA (sum, sumsq of the bf16 output) or nullptr
    int* err;
    long long* prof;                // selftest only: [grid][32] cycle counters per role phase (dbg & 8)
    int dbg;                        // selftest only: 1 = skip halo loads, 2 = skip MMA issue, 4 = skip epilogue work
};

constexpr int kHcAcc = 4;           // TMEM accumulator buffers (4 x cout <= 256 columns)

// K-major rows of min(ctot, 64) channels; ctot = 128 uses two 64-channel blocks.
__host__ __device__ inline uint32_t hc_row_bytes(int ctot) { return (ctot < 64 ? ctot : 64) * 2; }
__host__ __device__ inline uint32_t hc_nblk(int ctot) { return ctot > 64 ? ctot / 64 : 1; }
__host__ __device__ inline uint32_t hc_blk_bytes(int ctot) { return (kHcHaloPx * hc_row_bytes(ctot) + 1023u) & ~1023u; }
// XOR swizzle of a byte offset inside a 1 KB-aligned region: 16-byte chunk index ^= 128-byte line index (masked to the row)
__host__ __device__ inline uint32_t hc_swizzle(uint32_t off, uint32_t row_bytes) {
    return off ^ (((off >> 7) & (row_bytes / 16 - 1)) << 4);
}

struct HconvSmem {
    uint32_t ss_off, cstat_off, w_off, w_bytes, halo_off, halo_bytes, bar_off, total;
};
__host__ __device__ inline HconvSmem hconv_smem(int ctot, int cout, int stages) {
    HconvSmem s;
    s.ss_off = 0;                               // scale[64], shift[64]
    s.cstat_off = s.ss_off + 2 * 64 * 4;        // per-CTA running statistics [4 epilogue warps][64][2]
    s.w_off = 4096;                              // swizzle patterns are functions of the smem address: 1 KB alignment
    s.w_bytes = 9u * ctot * cout * 2;
    s.halo_off = (s.w_off + s.w_bytes + 1023u) & ~1023u;
    s.halo_bytes = hc_nblk(ctot) * hc_blk_bytes(ctot);
    s.bar_off = (s.halo_off + stages * s.halo_bytes + 15u) & ~15u;
    s.total = s.bar_off + (2 * stages + 2 * kHcAcc) * 8 + 16;
    return s;
}

// Persistent tile walk tile = blockIdx.x, blockIdx.x + gridDim.x, ... over (tw fastest, th, tn) without divisions.
struct HcTileIter {
    int tw, th, tn, dw, dh, dn, tiles_w, tiles_h, left;
    __device__ HcTileIter(int tiles_w_, int tiles_h_, int total) : tiles_w(tiles_w_), tiles_h(tiles_h_) {
        const int t0 = blockIdx.x, g = gridDim.x;
        tw = t0 % tiles_w; th = (t0 / tiles_w) % tiles_h; tn = t0 / (tiles_w * tiles_h);
        dw = g % tiles_w; dh = (g / tiles_w) % tiles_h; dn = g / (tiles_w * tiles_h);
        left = t0 < total ? (total - t0 + g - 1) / g : 0;
    }
    __device__ __forceinline__ bool valid() const { return left > 0; }
    __device__ __forceinline__ void next() {
        --left;
        tw += dw;
        if (tw >= tiles_w) { tw -= tiles_w; ++th; }
        th += dh;
        if (th >= tiles_h) { th -= tiles_h; ++tn; }
        tn += dn;
    }
};

// Sum of v[i] over the 32 lanes, for 16 values at once: afterwards lane l (and l ^ 16) holds the total of v[l & 15].
// 16 + 8 + 4 + 2 + 1 = 31 shuffles instead of 16 x 5.
__device__ __forceinline__ float warp_reduce16(float (&v)[16], int lane) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
#pragma unroll
    for (int half = 8; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = up ? v[i] : v[i + half];
            const float keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}

// The same for 32 values (16 + 8 + 4 + 2 + 1 = 31 shuffles): afterwards lane l holds the total of v[l].  The statistics
// epilogues reduce [16 sums | 16 sums of squares] in ONE pass; two 16-value passes spent their first 16 shuffles each
// on a plain exchange (62 shuffles + 124 selects per item were two thirds of the epilogue's instructions on the
// narrow layers: the training forward of a 64-channel conv took 1.5x the inference time per image).
__device__ __forceinline__ float warp_reduce32(float (&v)[32], int lane) {
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = up ? v[i] : v[i + half];
            const float keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// All tcgen05.mma of one output tile: 9 taps x KS 16-channel K steps, fully unrolled so that every A descriptor is the
// tile's base descriptor plus a compile-time constant and every B descriptor one multiply-add: the single issuing thread
// spends a handful of instructions per MMA.  KS fixes the row width: 1 -> 32-byte rows (SWIZZLE_32B), 2 -> 64 B
// (SWIZZLE_64B), 4 -> 128 B (SWIZZLE_128B), 8 -> two 64-channel blocks of 128-byte rows.
// K-major swizzled operands: SBO = stride between 8-row groups (A: next halo row = 10 pixels; B: 8 output channels).
template <int KS>
__device__ __forceinline__ void hc_issue_tile(uint32_t d_tmem, uint32_t hbase, uint32_t wbase, uint32_t cout,
                                              uint32_t idesc) {
    constexpr uint32_t kRow = KS == 1 ? 32 : (KS == 2 ? 64 : 128);
    constexpr uint32_t kLayout = KS == 1 ? 6u : (KS == 2 ? 4u : 2u);
    constexpr uint32_t kBlk = (kHcHaloPx * kRow + 1023u) & ~1023u;
    constexpr int kNblk = KS == 8 ? 2 : 1;
    const uint64_t a_base = umma_desc(hbase, 16, kHcHaloW * kRow, kLayout);
    const uint64_t b_base = umma_desc(wbase, 16, 8 * kRow, kLayout);
    const uint32_t b_tap = (cout * kRow) >> 4;             // one (tap, block) slab of the weights, in 16-byte units
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            constexpr int dummy = 0;
            (void)dummy;
            const int blk = ks >> 2, kk = ks & 3;
            const uint32_t a_off = (uint32_t)(blk * kBlk + ((tap / 3) * kHcHaloW + (tap % 3)) * kRow + kk * 32) >> 4;
            const uint64_t ad = a_base + a_off;                               // start-address field: bits [0,14)
            const uint64_t bd = b_base + (uint32_t)(tap * kNblk + blk) * b_tap + (uint32_t)(kk * 32 >> 4);
            if (tap == 0 && ks == 0) umma_bf16_c<false>(d_tmem, ad, bd, idesc);
            else umma_bf16_c<true>(d_tmem, ad, bd, idesc);
        }
    }
}

// kOcc = CTAs per SM the register budget is sized for.  The narrow layers (small weights, small halo) run 2 CTAs per SM
// so that one CTA's barrier / TMEM / fence latencies are covered by the other's work.
template <int kOcc>
__global__ void __launch_bounds__(kHcThreads, kOcc)
hconv_kernel(const __grid_constant__ HconvParams P) {
    extern __shared__ uint8_t smem_raw[];
    griddep_launch();
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw_addr);
    const int ctot = P.cup + P.cskip;
    const HconvSmem L = hconv_smem(ctot, P.cout, P.stages);
    const uint32_t bar0 = base + L.bar_off;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (P.stages + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * P.stages + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * P.stages + kHcAcc + a); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + L.bar_off + (2 * P.stages + 2 * kHcAcc) * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = P.tiles_w * P.tiles_h * P.N;
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)kHcAcc * P.cout) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P.stages; ++s) {
            mbar_init(full_bar(s), kHcProducers / 32);  // one arrival per producer warp once its copies have landed
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < kHcAcc; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 4);
        }
        fence_mbar_init();
    }
    if (warp == kHcProducers / 32) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), tmem_cols);
        tmem_relinquish();
    }
    griddep_wait();   // PDL: nothing above touches global memory
    {
        float* ss = reinterpret_cast<float*>(sm + L.ss_off);
        float* cst = reinterpret_cast<float*>(sm + L.cstat_off);
        for (int c = threadIdx.x; c < 64; c += blockDim.x) {
            ss[c] = (P.scale && c < P.cout) ? P.scale[c] : 1.f;
            ss[64 + c] = (P.shift && c < P.cout) ? P.shift[c] : 0.f;
        }
        for (int c = threadIdx.x; c < 4 * 128; c += blockDim.x) cst[c] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const bool prof = (P.dbg & 8) != 0;
    long long t_wait = 0, t_work = 0, t_a = 0, t_b = 0, tc = prof ? clock64() : 0;
#define UB_HC_TICK(var)                         \
    if (prof) {                                 \
        const long long now_ = clock64();       \
        var += now_ - tc;                       \
        tc = now_;                              \
    }

    if (warp < kHcProducers / 32) {
        // ================================================================= cp.async producers
        const int pt = threadIdx.x;
        for (uint32_t i = pt; i < L.w_bytes / 16; i += kHcProducers)  // resident weights: one linear copy
            cp_async16(base + L.w_off + i * 16, reinterpret_cast<const uint8_t*>(P.wpk) + (size_t)i * 16, 16);
        cp_async_commit();
        const int planes = ctot / 8, up_planes = P.cup / 8;  // 16-byte chunks per pixel (all of them / up-sampled source)
        const int pmask = planes - 1, pshift = 31 - __clz(planes);
        const int chunks = (P.dbg & 1) ? 0 : planes * kHcHaloPx;
        const uint32_t row_bytes = hc_row_bytes(ctot), blk_bytes = hc_blk_bytes(ctot);
        const int Hl = P.H >> 1, Wl = P.W >> 1;
        int stage = 0, prev_stage = -1;
        uint32_t phase = 0;
        for (HcTileIter it(P.tiles_w, P.tiles_h, total_tiles); it.valid(); it.next()) {
            const int h0 = it.th * kHcTileH - 1, w0 = it.tw * kHcTileW - 1;
            if (!mbar_wait_warp(empty_bar(stage), phase ^ 1, lane)) {
                atomicExch(P.err, 21);
                goto role_done;
            }
            UB_HC_TICK(t_wait)
            const uint32_t sbase = base + L.halo_off + stage * L.halo_bytes;
            // chunk q -> (halo pixel, 16-byte chunk) with the chunk fastest: consecutive threads copy consecutive 16 B of a
            // pixel (coalesced reads, conflict-free swizzled writes).  planes is a power of two (2, 4, 8, 16) and
            // px / 10 == (px * 205) >> 11 for px < 1029: no integer division.
            const size_t low_img = (size_t)it.tn * Hl * Wl, src_img = (size_t)it.tn * P.H * P.W;
#pragma unroll 4
            for (int q = pt; q < chunks; q += kHcProducers) {
                const int plane = q & pmask;
                const int px = q >> pshift;
                const int hh = (px * 205) >> 11, ww = px - hh * kHcHaloW;
                const int gh = h0 + hh, gw = w0 + ww;
                const bool in = (unsigned)gh < (unsigned)P.H && (unsigned)gw < (unsigned)P.W;
                const __nv_bfloat16* srcp;
                if (plane < up_planes) {
                    srcp = P.low;
                    if (in) srcp += (low_img + (size_t)((gh >> 1) * Wl + (gw >> 1))) * P.cup + plane * 8;
                } else {
                    srcp = P.src;
                    if (in) srcp += (src_img + (size_t)(gh * P.W + gw)) * P.cskip + (plane - up_planes) * 8;
                }
                const uint32_t off = (plane >> 3) * blk_bytes + px * row_bytes + (plane & 7) * 16;
                cp_async16(sbase + hc_swizzle(off, row_bytes), srcp, in ? 16u : 0u);  // src size 0 => zero fill
            }
            cp_async_commit();
            UB_HC_TICK(t_work)
            if (prev_stage >= 0) {
                cp_async_wait<1>();      // everything but the tile just issued (weights included) has landed
                UB_HC_TICK(t_a)
                fence_async_smem();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(full_bar(prev_stage));
                UB_HC_TICK(t_b)
            }
            prev_stage = stage;
            if (++stage == P.stages) {
                stage = 0;
                phase ^= 1;
            }
        }
        cp_async_wait<0>();
        fence_async_smem();
        __syncwarp();
        if (prev_stage >= 0 && lane == 0) mbar_arrive(full_bar(prev_stage));
        if (prof && pt == 0) {
            P.prof[blockIdx.x * 16 + 0] = t_wait;
            P.prof[blockIdx.x * 16 + 1] = t_work;
            P.prof[blockIdx.x * 16 + 2] = t_a;
            P.prof[blockIdx.x * 16 + 3] = t_b;
        }
    } else if (warp == kHcProducers / 32) {
        // ================================================================= MMA issuer (single thread)
        if (lane == 0) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            const uint32_t idesc = umma_idesc_bf16(128, P.cout, 0, 0);
            const uint32_t wbase = base + L.w_off;
            const int ksteps = ctot / 16;
            for (int n = blockIdx.x; n < total_tiles; n += gridDim.x) {
                if (!mbar_wait(tempty_bar(acc), acc_phase ^ 1)) {
                    atomicExch(P.err, 22);
                    goto role_done;
                }
                UB_HC_TICK(t_a)
                if (!mbar_wait(full_bar(stage), phase)) {
                    atomicExch(P.err, 23);
                    goto role_done;
                }
                tc_fence_after();
                UB_HC_TICK(t_wait)
                const uint32_t hbase = base + L.halo_off + stage * L.halo_bytes;
                const uint32_t d_tmem = tmem_base + acc * P.cout;
                if (!(P.dbg & 2)) {
                    switch (ksteps) {
                        case 1: hc_issue_tile<1>(d_tmem, hbase, wbase, P.cout, idesc); break;
                        case 2: hc_issue_tile<2>(d_tmem, hbase, wbase, P.cout, idesc); break;
                        case 4: hc_issue_tile<4>(d_tmem, hbase, wbase, P.cout, idesc); break;
                        default: hc_issue_tile<8>(d_tmem, hbase, wbase, P.cout, idesc); break;
                    }
                }
                UB_HC_TICK(t_work)
                if (P.dbg & 32) {
                    mbar_arrive(empty_bar(stage));
                    mbar_arrive(tfull_bar(acc));
                } else {
                    umma_commit(empty_bar(stage));
                    umma_commit(tfull_bar(acc));
                }
                UB_HC_TICK(t_b)
                if (++stage == P.stages) {
                    stage = 0;
                    phase ^= 1;
                }
                if (++acc == kHcAcc) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
            if (prof) {
                P.prof[blockIdx.x * 16 + 4] = t_a;
                P.prof[blockIdx.x * 16 + 5] = t_wait;
                P.prof[blockIdx.x * 16 + 6] = t_work;
                P.prof[blockIdx.x * 16 + 7] = t_b;
            }
        }
    } else {
        // ================================================================= epilogue (4 warps): thread = output pixel
        const int q = warp & 3;
        const int row = q * 32 + lane;                 // TMEM lane == pixel index inside the 8 x 16 tile
        const int wl = row & (kHcTileW - 1), hl = row >> 3;
        const int et = threadIdx.x - (kHcProducers + 32);
        const float* ss = reinterpret_cast<const float*>(sm + L.ss_off);
        float* cst = reinterpret_cast<float*>(sm + L.cstat_off);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (HcTileIter it(P.tiles_w, P.tiles_h, total_tiles); it.valid(); it.next()) {
            if (!mbar_wait_warp(tfull_bar(acc), acc_phase, lane)) {
                atomicExch(P.err, 24);
                goto role_done;
            }
            tc_fence_after();
            UB_HC_TICK(t_wait)
            const int ph = it.th * kHcTileH + hl, pw = it.tw * kHcTileW + wl;
            const bool valid = ph < P.H && pw < P.W;
            const size_t pix = ((size_t)it.tn * P.H + ph) * P.W + pw;
            __nv_bfloat16* optr = P.out + pix * P.cout;
            const __nv_bfloat16* rptr = P.residual ? P.residual + pix * P.cout : nullptr;
            const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * P.cout;
            for (int c0 = 0; c0 < P.cout; c0 += 16) {
                uint32_t r[16];
                if (!(P.dbg & 16)) {
                    tmem_ld16(taddr + c0, r);
                    tmem_ld_wait();
                }
                if (c0 + 16 >= P.cout) {  // last TMEM read of this accumulator: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(acc));
                }
                if (P.dbg & 4) continue;
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) * ss[c0 + i] + ss[64 + c0 + i];
                if (rptr && valid) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint4 rv = __ldg(reinterpret_cast<const uint4*>(rptr + c0) + j);
                        v[j * 8 + 0] += bf16_lo(rv.x); v[j * 8 + 1] += bf16_hi(rv.x);
                        v[j * 8 + 2] += bf16_lo(rv.y); v[j * 8 + 3] += bf16_hi(rv.y);
                        v[j * 8 + 4] += bf16_lo(rv.z); v[j * 8 + 5] += bf16_hi(rv.z);
                        v[j * 8 + 6] += bf16_lo(rv.w); v[j * 8 + 7] += bf16_hi(rv.w);
                    }
                }
                if (P.relu) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
                }
                uint4 o[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    o[j].x = pack_bf16(v[j * 8 + 0], v[j * 8 + 1]);
                    o[j].y = pack_bf16(v[j * 8 + 2], v[j * 8 + 3]);
                    o[j].z = pack_bf16(v[j * 8 + 4], v[j * 8 + 5]);
                    o[j].w = pack_bf16(v[j * 8 + 6], v[j * 8 + 7]);
                }
                if (valid) {
                    reinterpret_cast<uint4*>(optr + c0)[0] = o[0];
                    reinterpret_cast<uint4*>(optr + c0)[1] = o[1];
                }
                if (P.stats) {
                    // statistics of the bf16 values just stored (masked pixels contribute 0)
                    float s1[16], s2[16];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint32_t w4[4] = {o[j].x, o[j].y, o[j].z, o[j].w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float lo = valid ? bf16_lo(w4[k]) : 0.f, hi = valid ? bf16_hi(w4[k]) : 0.f;
                            s1[j * 8 + 2 * k] = lo; s1[j * 8 + 2 * k + 1] = hi;
                            s2[j * 8 + 2 * k] = lo * lo; s2[j * 8 + 2 * k + 1] = hi * hi;
                        }
                    }
                    const float t1 = warp_reduce16(s1, lane), t2 = warp_reduce16(s2, lane);
                    if (lane < 16) {
                        // channel c0+lane of warp q's private slot is only ever touched by this lane: plain adds in a fixed
                        // order, i.e. the batch statistics (and with them the whole forward) are run-to-run deterministic
                        cst[q * 128 + 2 * (c0 + lane)] += t1;
                        cst[q * 128 + 2 * (c0 + lane) + 1] += t2;
                    }
                }
            }
            UB_HC_TICK(t_work)
            if (++acc == kHcAcc) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
        if (P.stats) {
            named_bar_sync(1, 128);
            float* dst = P.stats + static_cast<size_t>(blockIdx.x) * P.cout * 2;
            for (int j = et; j < 2 * P.cout; j += 128) dst[j] = (cst[j] + cst[128 + j]) + (cst[256 + j] + cst[384 + j]);
        }
        if (prof && et == 0) {
            P.prof[blockIdx.x * 16 + 8] = t_wait;
            P.prof[blockIdx.x * 16 + 9] = t_work;
        }
    }
#undef UB_HC_TICK
role_done:
    tc_fence_before();
    __syncthreads();
    if (warp == kHcProducers / 32) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

}  // namespace ub
