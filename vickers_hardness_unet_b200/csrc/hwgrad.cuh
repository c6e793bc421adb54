// Halo-resident weight gradient of the narrow 3x3 / stride-1 convolutions (Cout <= 64, <= 128 input channels) on
// tcgen05 (sm_100a):      dW[co, (r, s), ci] = sum over pixels p of  dZ[p, co] * X[p + (r-1, s-1), ci]
//
// The general wgrad kernel (wgrad.cuh) runs one (tap, pixel range) work item at a time and therefore streams dZ and X
// through L2 nine times; for 16-64 channel tensors at 128^2..512^2 that (and the TMA's per-row request rate on 32..128
// byte rows) costs 0.4-1.5 ms per layer.  Here each persistent CTA walks 8 x 16 pixel tiles, loads the dZ tile and the
// 10 x 18 X halo ONCE (cp.async into swizzled MN-major rows [pixel][channels]) and accumulates ALL filter taps in TMEM:
//   * the three column taps s = 0..2 of a filter row are ONE MMA of N = 3 * Cin: consecutive pixels are consecutive
//     rows in shared memory, so the "next pixel" operand block is the same data 1 row further (descriptor LBO = one row);
//   * the three filter rows r are three accumulators [co x 3*Cin] that stay in TMEM for the whole kernel; the contraction
//     (K) runs over the tile's 128 pixels, 16 per MMA (two image rows of 8 pixels: descriptor SBO = one halo row);
//   * the decoder's nearest-2x upsample + concat input is formed by the loader's address arithmetic (as in hconv.cuh).
// At the end each CTA adds its partial [Cout x 9 x Cin] into the packed fp32 gradient with 16-byte vector reductions.
// blockIdx.y selects the (64-channel block, filter-row range) a CTA owns when 9 * Cin fp32 columns exceed TMEM.
#pragma once
#include "hconv.cuh"

namespace ub {

struct HwgradParams {
    int H, W, N;                    // extent of dZ (= conv output = conv input resolution)
    int tiles_w, tiles_h;
    int cup, cskip;                 // conv input = cat(nearest2x(low)[cup], src[cskip])
    const __nv_bfloat16* low;
    const __nv_bfloat16* src;
    const __nv_bfloat16* dz;        // [N, H, W, cout]
    int cout;
    int cw;                         // input channels per CTA group (min(ctot, 64))
    int ndh_max;                    // filter rows per CTA group (3, or fewer when 3 * 3 * cw > 512 TMEM columns)
    int stages;
    float* gpk;                     // packed gradient [cout][9][gtot] fp32
    int gtot;                       // gradient row length (>= ctot: further channels are another launch's)
    int* err;
};

struct HwgradSmem {
    uint32_t x_bytes, z_bytes, stage_bytes, bar_off, total;
};
__host__ __device__ inline HwgradSmem hwgrad_smem(int cw, int cout, int stages) {
    HwgradSmem s;
    s.x_bytes = (kHcHaloPx * cw * 2 + 1023u) & ~1023u;
    s.z_bytes = (128u * cout * 2 + 1023u) & ~1023u;
    s.stage_bytes = s.x_bytes + s.z_bytes;
    s.bar_off = s.stage_bytes * stages;
    s.total = s.bar_off + (2 * stages + 1) * 8 + 16;
    return s;
}

__global__ void __launch_bounds__(kHcThreads, 1)
hwgrad_kernel(const __grid_constant__ HwgradParams P) {
    extern __shared__ uint8_t smem_raw[];
    griddep_launch();
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw_addr);
    const int ctot = P.cup + P.cskip;
    const HwgradSmem L = hwgrad_smem(P.cw, P.cout, P.stages);
    const uint32_t bar0 = base + L.bar_off;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (P.stages + s); };
    const uint32_t done_bar = bar0 + 8u * (2 * P.stages);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + L.bar_off + (2 * P.stages + 1) * 8);

    // this CTA's group: 64-channel block of the input and filter-row range
    const int nblk = ctot / P.cw;
    const int ngrp_dh = (3 + P.ndh_max - 1) / P.ndh_max;
    const int blk = blockIdx.y % nblk, gdh = blockIdx.y / nblk;
    const int dh0 = gdh * P.ndh_max;
    const int ndh = (3 - dh0) < P.ndh_max ? (3 - dh0) : P.ndh_max;
    (void)ngrp_dh;
    const int c_lo = blk * P.cw;
    const int ncol = 3 * P.cw;                      // UMMA N: (s, ci)
    const uint32_t R = P.cw * 2, Rz = P.cout * 2;   // row bytes of the X halo / of the dZ tile

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = P.tiles_w * P.tiles_h * P.N;
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(ndh * ncol)) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P.stages; ++s) {
            mbar_init(full_bar(s), kHcProducers / 32);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(done_bar, 1);
        fence_mbar_init();
    }
    if (warp == kHcProducers / 32) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), tmem_cols);
        tmem_relinquish();
    }
    griddep_wait();   // PDL: nothing above touches global memory
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kHcProducers / 32) {
        // ================================================================= cp.async producers
        const int pt = threadIdx.x;
        const int xpl = P.cw / 8, zpl = P.cout / 8;          // 16-byte chunks per pixel
        const int xmask = xpl - 1, xshift = 31 - __clz(xpl), zmask = zpl - 1, zshift = 31 - __clz(zpl);
        const int xchunks = xpl * kHcHaloPx, zchunks = zpl * 128;
        const int Hl = P.H >> 1, Wl = P.W >> 1;
        int stage = 0, prev_stage = -1;
        uint32_t phase = 0;
        for (HcTileIter it(P.tiles_w, P.tiles_h, total_tiles); it.valid(); it.next()) {
            const int h0 = it.th * kHcTileH - 1, w0 = it.tw * kHcTileW - 1;
            if (!mbar_wait_warp(empty_bar(stage), phase ^ 1, lane)) {
                atomicExch(P.err, 31);
                goto role_done;
            }
            const uint32_t xb = base + stage * L.stage_bytes, zb = xb + L.x_bytes;
            const size_t low_img = (size_t)it.tn * Hl * Wl, src_img = (size_t)it.tn * P.H * P.W;
#pragma unroll 4
            for (int q = pt; q < xchunks; q += kHcProducers) {
                const int ch = q & xmask;
                const int px = q >> xshift;
                const int hh = (px * 205) >> 11, ww = px - hh * kHcHaloW;
                const int gh = h0 + hh, gw = w0 + ww;
                const bool in = (unsigned)gh < (unsigned)P.H && (unsigned)gw < (unsigned)P.W;
                const int c = c_lo + ch * 8;
                const __nv_bfloat16* srcp;
                if (c < P.cup) {
                    srcp = P.low;
                    if (in) srcp += (low_img + (size_t)((gh >> 1) * Wl + (gw >> 1))) * P.cup + c;
                } else {
                    srcp = P.src;
                    if (in) srcp += (src_img + (size_t)(gh * P.W + gw)) * P.cskip + (c - P.cup);
                }
                cp_async16(xb + hc_swizzle(px * R + ch * 16, R), srcp, in ? 16u : 0u);
            }
#pragma unroll 4
            for (int q = pt; q < zchunks; q += kHcProducers) {
                const int ch = q & zmask;
                const int px = q >> zshift;                          // 0..127: (h_local, w_local) = (px >> 3, px & 7)
                const int gh = h0 + 1 + (px >> 3), gw = w0 + 1 + (px & 7);
                const bool in = gh < P.H && gw < P.W;
                const __nv_bfloat16* srcp = P.dz;
                if (in) srcp += (src_img + (size_t)(gh * P.W + gw)) * P.cout + ch * 8;
                cp_async16(zb + hc_swizzle(px * Rz + ch * 16, Rz), srcp, in ? 16u : 0u);  // masked pixels contribute 0
            }
            cp_async_commit();
            if (prev_stage >= 0) {
                cp_async_wait<1>();
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(full_bar(prev_stage));
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
    } else if (warp == kHcProducers / 32) {
        // ================================================================= MMA issuer (single thread)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t idesc = umma_idesc_bf16(128, ncol, 1, 1);   // both operands MN-major
            const uint32_t lay_x = R == 128 ? 2u : (R == 64 ? 4u : 6u), lay_z = Rz == 128 ? 2u : (Rz == 64 ? 4u : 6u);
            uint32_t first = 0;
            for (int n = blockIdx.x; n < total_tiles; n += gridDim.x) {
                if (!mbar_wait(full_bar(stage), phase)) {
                    atomicExch(P.err, 33);
                    goto role_done;
                }
                tc_fence_after();
                const uint32_t xb = base + stage * L.stage_bytes, zb = xb + L.x_bytes;
                // A = dZ tile: M-blocks beyond the tensor's channels alias block 0 (LBO = 0): their D rows are ignored
                const uint64_t a_base = umma_desc(zb, 0, 8 * Rz, lay_z);
                // B = X halo: N-block j = filter column s = j is the same rows one pixel further (LBO = one row)
                const uint64_t b_base = umma_desc(xb, R, kHcHaloW * R, lay_x);
                for (int d = 0; d < ndh; ++d) {
                    const uint32_t d_tmem = tmem_base + d * ncol;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {   // 16 pixels = image rows 2k, 2k+1 of the tile
                        const uint64_t ad = a_base + ((uint32_t)(k * 16) * Rz >> 4);
                        const uint64_t bd = b_base + ((uint32_t)((dh0 + d + 2 * k) * kHcHaloW) * R >> 4);
                        umma_bf16(d_tmem, ad, bd, idesc, first | (uint32_t)k);
                    }
                }
                first = 1;
                umma_commit(empty_bar(stage));
                if (++stage == P.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            umma_commit(done_bar);
        }
    } else {
        // ================================================================= epilogue (4 warps): thread = output channel row
        const int q = warp & 3;
        const int co = q * 32 + lane;
        if (!mbar_wait_warp(done_bar, 0, lane)) {
            atomicExch(P.err, 34);
            goto role_done;
        }
        tc_fence_after();
        if (blockIdx.x < total_tiles) {   // CTAs without a tile hold garbage in TMEM
            for (int d = 0; d < ndh; ++d) {
                for (int c0 = 0; c0 < ncol; c0 += 16) {
                    uint32_t r[16];
                    tmem_ld16(tmem_base + (uint32_t(q * 32) << 16) + d * ncol + c0, r);
                    tmem_ld_wait();
                    if (co < P.cout) {
                        const int s = c0 / P.cw, ci = c_lo + (c0 - s * P.cw);
                        float* gp = P.gpk + ((size_t)co * 9 + (dh0 + d) * 3 + s) * P.gtot + ci;
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            red_add_v4(gp + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                       __uint_as_float(r[j + 3]));
                    }
                }
            }
        }
    }
role_done:
    tc_fence_before();
    __syncthreads();
    if (warp == kHcProducers / 32) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

}  // namespace ub
