// Data gradient of a decoder conv1 w.r.t. its LOW-resolution input (through the nearest-2x up-sample), narrow blocks:
//     dLow[u, ci] = sum over the 4 x 4 high-res offsets o = (oh, ow) in -1..2 of  dZ[2u + o, co] * V[o][co][ci]
// (V = the 3x3 weights pre-summed per offset, PK_DLOW layout [ci][t*cz + co]).  The tap-table kernel ran this as 16
// stride-2 TMA tile loads of 128 pixels x 16..32 channels per output tile (32-byte rows: ~700 cycles per K = 16 chunk,
// 315 us for decoder.blocks.4).  Here a pipeline step is ONE dense TMA box of the high-resolution dZ under an 8 x 16
// output tile, loaded as PIXEL PAIRS: the tensor is viewed as [2*cz, Wl, 2*Hl, N] (a row = two horizontally adjacent
// pixels = 64 / 128 contiguous bytes, the swizzle span), box [2*cz, 10 pairs, 34 rows] from (w0-1, 2*h0-1).  All 16
// taps are shifted K-major UMMA descriptors into that resident box: consecutive operand rows are consecutive PAIRS
// (= every second pixel), the even / odd pixel of a pair is the K offset 0 / cz inside the row, consecutive 8-row
// groups are two box rows apart (SBO = 2 box rows = one output row).  No TMA element strides (a first version loaded
// four parity images with element strides (2, 2): the strided gather ran at ~10 cycles per pixel and was TMA-bound).
// The weights of all taps stay resident in shared memory; accumulator [128 pixels x cup] double buffered in TMEM; the
// epilogue stores bf16 rows straight from registers.
#pragma once
#include "ptx.cuh"

namespace ub {

constexpr int kDlThreads = 192;    // warp 0: TMA producer, warp 1: MMA issuer (owns TMEM), warps 2-5: epilogue
constexpr int kDlTileW = 8, kDlTileH = 16;
constexpr int kDlBoxW = kDlTileW + 2, kDlBoxH = 2 * kDlTileH + 2;   // pixel pairs x high-res rows

struct DlowParams {
    int Hl, Wl, N;                  // low-resolution extent; dZ is [N, 2*Hl, 2*Wl, cz]
    int tiles_w, tiles_h;
    int cz, cup;                    // K channels (16 / 32 / 64) and output channels (<= 64)
    int stages;
    int dbg;                        // selftest only: 1 = no dZ loads, 2 = no MMA issue, 4 = no output stores
    __nv_bfloat16* out;             // [N, Hl, Wl, cup]
    int* err;
};

struct DlowSmem {
    uint32_t box_bytes, stage_bytes, w_tap_bytes, w_off, bar_off, total;
};
__host__ __device__ inline DlowSmem dlow_smem(int cz, int cup, int stages) {
    DlowSmem s;
    s.box_bytes = ((uint32_t)kDlBoxW * kDlBoxH * cz * 4 + 1023u) & ~1023u;
    s.stage_bytes = s.box_bytes;
    s.w_tap_bytes = ((uint32_t)cup * cz * 2 + 1023u) & ~1023u;
    s.w_off = s.stage_bytes * stages;
    s.bar_off = s.w_off + 16 * s.w_tap_bytes;
    s.total = s.bar_off + (2 * stages + 5) * 8 + 16;
    return s;
}

__global__ void __launch_bounds__(kDlThreads, 1)
dlow_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmW,
            const __grid_constant__ DlowParams P) {
    extern __shared__ uint8_t smem_raw[];
    griddep_launch();
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw_addr);
    const DlowSmem L = dlow_smem(P.cz, P.cup, P.stages);
    const uint32_t bar0 = base + L.bar_off;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (P.stages + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * P.stages + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * P.stages + 2 + a); };
    const uint32_t w_bar = bar0 + 8u * (2 * P.stages + 4);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + L.bar_off + (2 * P.stages + 5) * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = P.tiles_w * P.tiles_h * P.N;
    const uint32_t Rz = P.cz * 2;                       // bytes of one pixel of a dZ box = one weight row = swizzle span
    uint32_t tmem_cols = 32;
    while (tmem_cols < 2u * P.cup) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmZ);
        tma_prefetch_desc(&tmW);
        for (int s = 0; s < P.stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 4);
        }
        mbar_init(w_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), tmem_cols);
        tmem_relinquish();
    }
    griddep_wait();   // PDL: nothing above touches global memory (the weights were re-packed earlier in this stream)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================= TMA producer (one thread)
        if (lane == 0) {
            // resident weights: one [cup x cz] tile per tap
            mbar_expect_tx(w_bar, 16u * P.cup * Rz);
            for (int t = 0; t < 16; ++t) tma_load_2d(base + L.w_off + t * L.w_tap_bytes, &tmW, w_bar, t * P.cz, 0);
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx = (uint32_t)kDlBoxW * kDlBoxH * 2u * Rz;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int tw = t % P.tiles_w, th = (t / P.tiles_w) % P.tiles_h, tn = t / (P.tiles_w * P.tiles_h);
                if (!mbar_wait(empty_bar(stage), phase ^ 1)) {
                    atomicExch(P.err, 51);
                    goto role_done;
                }
                if (P.dbg & 1) {
                    mbar_arrive(full_bar(stage));
                } else {
                    mbar_expect_tx(full_bar(stage), tx);
                    tma_load_4d(base + stage * L.stage_bytes, &tmZ, full_bar(stage), 0, tw * kDlTileW - 1,
                                2 * th * kDlTileH - 1, tn);
                }
                if (++stage == P.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer (one thread)
        if (lane == 0) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            const uint32_t idesc = umma_idesc_bf16(128, P.cup, 0, 0);
            const uint32_t lay_w = Rz == 128 ? 2u : (Rz == 64 ? 4u : 6u);   // weight rows: cz channels
            const uint32_t Rp = 2 * Rz, lay_a = Rp == 128 ? 2u : 4u;         // dZ rows: a pixel pair
            const int ksteps = P.cz / 16;
            // per tap: byte offset of the A view inside the box -- invariant over tiles
            uint32_t a_off[16];
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                const int b = t & 1, pw = (t >> 1) & 1, a = (t >> 2) & 1, ph = (t >> 3) & 1;
                const int oh = 2 - 2 * a - ph, ow = 2 - 2 * b - pw;           // high-res offset of this tap (PK_DLOW order)
                const int odd = ow & 1, pair = (ow - odd) / 2 + 1;            // pixel 2u + ow = pair u + (ow - odd)/2, box starts at -1
                a_off[t] = (uint32_t)((oh + 1) * kDlBoxW + pair) * Rp + (uint32_t)odd * Rz;
            }
            if (!mbar_wait(w_bar, 0)) {
                atomicExch(P.err, 52);
                goto role_done;
            }
            for (int n = blockIdx.x; n < total_tiles; n += gridDim.x) {
                if (!mbar_wait(tempty_bar(acc), acc_phase ^ 1) || !mbar_wait(full_bar(stage), phase)) {
                    atomicExch(P.err, 53);
                    goto role_done;
                }
                tc_fence_after();
                const uint32_t sa = base + stage * L.stage_bytes;
                const uint32_t d_tmem = tmem_base + acc * P.cup;
                uint32_t accum = 0;
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    for (int k = 0; k < ksteps && !(P.dbg & 2); ++k) {
                        // A: 8-row groups = the 8 pairs of one output row (SBO = two box rows); B: [cup x cz] tile of tap t
                        const uint64_t ad = umma_desc(sa + a_off[t] + k * 32, 16, 2 * kDlBoxW * Rp, lay_a);
                        const uint64_t bd = umma_desc(base + L.w_off + t * L.w_tap_bytes + k * 32, 16, 8 * Rz, lay_w);
                        umma_bf16(d_tmem, ad, bd, idesc, accum);
                        accum = 1;
                    }
                }
                umma_commit(empty_bar(stage));
                umma_commit(tfull_bar(acc));
                if (++stage == P.stages) {
                    stage = 0;
                    phase ^= 1;
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        // ================================================================= epilogue: thread = output pixel
        const int q = warp & 3;
        const int m = q * 32 + lane;
        const int wl = m & (kDlTileW - 1), hl = m >> 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int tw = t % P.tiles_w, th = (t / P.tiles_w) % P.tiles_h, tn = t / (P.tiles_w * P.tiles_h);
            const int h = th * kDlTileH + hl, w = tw * kDlTileW + wl;
            const bool valid = h < P.Hl && w < P.Wl;
            __nv_bfloat16* op = P.out + (((size_t)tn * P.Hl + h) * P.Wl + w) * P.cup;
            if (!mbar_wait_warp(tfull_bar(acc), acc_phase, lane)) {
                atomicExch(P.err, 54);
                goto role_done;
            }
            tc_fence_after();
            for (int c0 = 0; c0 < P.cup; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(tmem_base + (uint32_t(q * 32) << 16) + acc * P.cup + c0, v);
                tmem_ld_wait();
                if (c0 + 16 >= P.cup) {   // last TMEM read of this accumulator: hand it back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(acc));
                }
                if (valid && !(P.dbg & 4)) {
                    uint4 o0, o1;
                    o0.x = pack_bf16(__uint_as_float(v[0]), __uint_as_float(v[1]));
                    o0.y = pack_bf16(__uint_as_float(v[2]), __uint_as_float(v[3]));
                    o0.z = pack_bf16(__uint_as_float(v[4]), __uint_as_float(v[5]));
                    o0.w = pack_bf16(__uint_as_float(v[6]), __uint_as_float(v[7]));
                    o1.x = pack_bf16(__uint_as_float(v[8]), __uint_as_float(v[9]));
                    o1.y = pack_bf16(__uint_as_float(v[10]), __uint_as_float(v[11]));
                    o1.z = pack_bf16(__uint_as_float(v[12]), __uint_as_float(v[13]));
                    o1.w = pack_bf16(__uint_as_float(v[14]), __uint_as_float(v[15]));
                    st_global_256(op + c0, o0, o1);
                }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }
role_done:
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

}  // namespace ub
