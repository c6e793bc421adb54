// TMA-fed halo-resident weight gradient of the 3x3 / stride-1 convolutions on tcgen05 (sm_100a):
//     dW[co, (r, s), ci] = sum over pixels p of  dZ[p, co] * X[p + (r-1, s-1), ci]
//
// hwgrad.cuh (cp.async, 128-pixel tiles, one MMA per filter row) spends most of its time on tile hand-off and on M = 128
// MMAs of which 16-64 rows are used; wgrad.cuh (one tap per work item) re-reads dZ and X nine times and is bound by the
// SM's L2 ingest.  Here a pipeline step is ONE pair of TMA boxes for a tile of 8 x TH "anchor" pixels q (= X pixels):
//     X  box [cw channels, 10 pixels (columns w0-1 .. w0+8), TH rows]            -> B operand, MN-major
//     dZ box [co_blk channels, 8 pixels, TH+2 rows (rows h0-1 .. h0+TH)]          -> A operand, MN-major
// and every filter tap comes out of shifted views of those two boxes (the contraction K runs over anchors, 16 per MMA =
// two image rows of 8 pixels, descriptor SBO = one row of the box):
//   * filter COLUMN s is N-block s of the B operand: the same X rows one pixel further (LBO = one pixel), N = 3 * cw;
//   * filter ROW r (narrow mode, co_blk <= 32): M-block j of the A operand is the dZ tile one image row further
//     (LBO = one dZ row), so accumulator rows [j*co_blk, (j+1)*co_blk) hold filter row r = 2 - j: ONE MMA per 16 anchors
//     produces all nine taps (M = 3 * co_blk <= 96 of 128 rows used; hwgrad used 16-64 of 128 three times);
//   * wide mode (co_blk = 128: two 64-channel dZ boxes, LBO = box stride): three MMAs per 16 anchors whose A start
//     address moves one dZ row, into three accumulators (filter row r = 2 - d).
// Accumulators stay in TMEM for the whole kernel (the pixel range is split over the CTAs of a group); at the end each
// CTA adds its partial into the packed fp32 gradient [cout][9][ctot] with 16-byte vector reductions.  Zero fill of the
// out-of-image part of both boxes makes partial tiles and the image border exact.  blockIdx.y = (co block, ci block).
//
// Up mode (decoder conv1, channels of the nearest-2x up-sampled tensor):  X = the LOW-resolution tensor, anchors = its
// pixels u, and blockIdx.z = output parity (a, b): the dZ box is loaded with TMA element strides (2, 2) starting at
// (2*w0 + b, 2*(h0-1) + a), i.e. it is the dense image G_ab[u] = dZ[2u + (a, b)].  The main loop is unchanged; a block
// (row shift dh = 1 - j, column shift dw = s - 1 of `low` against G_ab) is added to every 3x3 tap (r, s') with
// floor((a + r - 1) / 2) = dh and floor((b + s' - 1) / 2) = dw (1, 2 or 4 taps; blocks with no tap are skipped).
#pragma once
#include "ptx.cuh"

namespace ub {

constexpr int kXwThreads = 192;   // warp 0: TMA producer, warp 1: MMA issuer (owns TMEM), warps 2-5: final epilogue
constexpr int kXwTileW = 8;

struct XwgradParams {
    int H, W, N;
    int tiles_w, tiles_h;
    int th;                 // anchor rows per tile (even)
    int cout, cin;          // channels of dZ / of X
    int co_blk, cw;         // channels per CTA group
    int zc_box, n_zbox;     // dZ box channels (<= 64) and boxes per step
    int wide;
    int up;                 // 1: up mode (H, W = extent of the low-resolution tensor; dZ is [N, 2H, 2W, cout])
    int stages;
    int ctot, dci0;         // packed gradient: row length and first column of X's channels
    float* gpk;
    int* err;
};

struct XwgradSmem {
    uint32_t zbox_bytes, z_bytes, x_bytes, stage_bytes, bar_off, total;
};
__host__ __device__ inline XwgradSmem xwgrad_smem(int th, int zc_box, int n_zbox, int cw, int stages) {
    XwgradSmem s;
    // narrow mode reads up to 128 / co_blk - 1 rows past an anchor row: the box is followed by the X region (and the
    // last stage by the barrier block + slack), so those (ignored) reads stay inside the allocation
    s.zbox_bytes = ((uint32_t)(th + 2) * kXwTileW * zc_box * 2 + 1023u) & ~1023u;
    s.z_bytes = s.zbox_bytes * n_zbox;
    s.x_bytes = ((uint32_t)th * (kXwTileW + 2) * cw * 2 + 1023u) & ~1023u;
    s.stage_bytes = s.z_bytes + s.x_bytes;
    s.bar_off = s.stage_bytes * stages;
    s.total = s.bar_off + (2 * stages + 1) * 8 + 16;
    return s;
}

__global__ void __launch_bounds__(kXwThreads, 1)
xwgrad_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmX,
              const __grid_constant__ XwgradParams P) {
    extern __shared__ uint8_t smem_raw[];
    griddep_launch();
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw_addr);
    const XwgradSmem L = xwgrad_smem(P.th, P.zc_box, P.n_zbox, P.cw, P.stages);
    const uint32_t bar0 = base + L.bar_off;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (P.stages + s); };
    const uint32_t done_bar = bar0 + 8u * (2 * P.stages);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + L.bar_off + (2 * P.stages + 1) * 8);

    const int nco = P.cout / P.co_blk;
    const int co0 = (blockIdx.y % nco) * P.co_blk, ci0 = (blockIdx.y / nco) * P.cw;
    const int ncol = 3 * P.cw;                       // UMMA N: (s, ci)
    const int nacc = P.wide ? 3 : 1;
    const uint32_t R = P.cw * 2, Rz = P.zc_box * 2;  // bytes of one pixel in the X / dZ boxes
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = P.tiles_w * P.tiles_h * P.N;
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(nacc * ncol)) tmem_cols <<= 1;
    const int pa = P.up ? (int)(blockIdx.z >> 1) : 0, pb = P.up ? (int)(blockIdx.z & 1) : 0;   // output parity
    // wide mode: accumulators (row shifts dh = 1 - d) that carry taps: all three, or two of them in up mode
    const int d_lo = (P.wide && P.up && pa == 0) ? 1 : 0, d_hi = (P.wide && P.up && pa == 1) ? 2 : nacc;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmZ);
        tma_prefetch_desc(&tmX);
        for (int s = 0; s < P.stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(done_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), tmem_cols);
        tmem_relinquish();
    }
    griddep_wait();   // PDL: nothing above touches global memory
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================= TMA producer (one thread)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx = (uint32_t)P.n_zbox * (P.th + 2) * kXwTileW * Rz + (uint32_t)P.th * (kXwTileW + 2) * R;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int tw = t % P.tiles_w, th_ = (t / P.tiles_w) % P.tiles_h, tn = t / (P.tiles_w * P.tiles_h);
                if (!mbar_wait(empty_bar(stage), phase ^ 1)) {
                    atomicExch(P.err, 41);
                    goto role_done;
                }
                const uint32_t zb = base + stage * L.stage_bytes, xb = zb + L.z_bytes;
                mbar_expect_tx(full_bar(stage), tx);
                const int zw = P.up ? 2 * tw * kXwTileW + pb : tw * kXwTileW;
                const int zh = P.up ? 2 * (th_ * P.th - 1) + pa : th_ * P.th - 1;
                for (int b = 0; b < P.n_zbox; ++b)
                    tma_load_4d(zb + b * L.zbox_bytes, &tmZ, full_bar(stage), co0 + b * P.zc_box, zw, zh, tn);
                tma_load_4d(xb, &tmX, full_bar(stage), ci0, tw * kXwTileW - 1, th_ * P.th, tn);
                if (++stage == P.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer (one thread)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t idesc = umma_idesc_bf16(128, ncol, 1, 1);   // both operands MN-major
            const uint32_t lay_x = R == 128 ? 2u : (R == 64 ? 4u : 6u), lay_z = Rz == 128 ? 2u : (Rz == 64 ? 4u : 6u);
            const uint32_t zrow = kXwTileW * Rz, xrow = (kXwTileW + 2) * R;   // one image row of each box
            const int ksteps = P.th / 2;
            uint32_t accum = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                if (!mbar_wait(full_bar(stage), phase)) {
                    atomicExch(P.err, 43);
                    goto role_done;
                }
                tc_fence_after();
                const uint32_t zb = base + stage * L.stage_bytes, xb = zb + L.z_bytes;
                // A = dZ: K groups one box row apart; M-blocks one box row apart (narrow) / one box apart (wide)
                const uint64_t a_base = umma_desc(zb, P.wide ? L.zbox_bytes : zrow, zrow, lay_z);
                // B = X: N-block s = filter column s = the same rows one pixel further
                const uint64_t b_base = umma_desc(xb, R, xrow, lay_x);
                // fully unrolled issue bodies: the single issuing thread is the critical path (a dynamic inner loop
                // over the accumulators cost 50 % on the wide layers)
                const uint32_t a_step = (2 * zrow) >> 4, b_step = (2 * xrow) >> 4, zr = zrow >> 4;
                uint64_t ad = a_base + (uint64_t)(d_lo * zr), bd = b_base;
                const uint32_t t0 = tmem_base + d_lo * ncol, t1 = t0 + ncol, t2 = t1 + ncol;
                if (!P.wide) {
#pragma unroll 4
                    for (int k = 0; k < ksteps; ++k) {
                        umma_bf16(t0, ad, bd, idesc, accum);
                        accum = 1;
                        ad += a_step;
                        bd += b_step;
                    }
                } else if (d_hi - d_lo == 3) {
#pragma unroll 4
                    for (int k = 0; k < ksteps; ++k) {
                        umma_bf16(t0, ad, bd, idesc, accum);
                        umma_bf16(t1, ad + zr, bd, idesc, accum);
                        umma_bf16(t2, ad + 2 * zr, bd, idesc, accum);
                        accum = 1;
                        ad += a_step;
                        bd += b_step;
                    }
                } else {
#pragma unroll 4
                    for (int k = 0; k < ksteps; ++k) {
                        umma_bf16(t0, ad, bd, idesc, accum);
                        umma_bf16(t1, ad + zr, bd, idesc, accum);
                        accum = 1;
                        ad += a_step;
                        bd += b_step;
                    }
                }
                umma_commit(empty_bar(stage));
                if (++stage == P.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            umma_commit(done_bar);
        }
    } else {
        // ================================================================= final epilogue: thread = accumulator row
        const int q = warp & 3;
        const int m = q * 32 + lane;
        if (!mbar_wait_warp(done_bar, 0, lane)) {
            atomicExch(P.err, 44);
            goto role_done;
        }
        tc_fence_after();
        const int j = P.wide ? 0 : m / P.co_blk;
        const int co = co0 + (P.wide ? m : m - j * P.co_blk);
        // filter taps a (shift, parity) block is added to: plain mode 1 tap; up mode 0, 1 or 2 per dimension
        auto taps = [&](int delta, int par, int plain, int (&out)[2]) -> int {
            if (!P.up) {
                out[0] = plain;
                return 1;
            }
            int n = 0;
            for (int t = 0; t < 3; ++t)
                if (((par + t + 1) >> 1) - 1 == delta) out[n++] = t;   // floor((par + t - 1) / 2) == delta
            return n;
        };
        for (int d = d_lo; d < d_hi; ++d) {
            const int jj = P.wide ? d : j;
            int rr[2] = {0, 0}, cc[2] = {0, 0};
            const int nr = (P.wide || j < 3) ? taps(1 - jj, pa, 2 - jj, rr) : 0;
            for (int c0 = 0; c0 < ncol; c0 += 16) {
                const int s = c0 / P.cw, ci = ci0 + (c0 - s * P.cw);
                const int nc = taps(s - 1, pb, s, cc);
                if (nc == 0) continue;                // warp-uniform (nr is not in narrow mode: tcgen05.ld is warp-collective)
                uint32_t v[16];
                tmem_ld16(tmem_base + (uint32_t(q * 32) << 16) + d * ncol + c0, v);
                tmem_ld_wait();
                for (int a = 0; a < nr; ++a)
                    for (int b = 0; b < nc; ++b) {
                        float* gp = P.gpk + ((size_t)co * 9 + rr[a] * 3 + cc[b]) * P.ctot + P.dci0 + ci;
#pragma unroll
                        for (int e = 0; e < 16; e += 4)
                            red_add_v4(gp + e, __uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]),
                                       __uint_as_float(v[e + 3]));
                    }
            }
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
