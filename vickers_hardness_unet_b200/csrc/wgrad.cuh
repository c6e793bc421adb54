// Weight-gradient implicit GEMM on tcgen05 (sm_100a):   dW[co, ci] (+)= sum over pixels  dZ[p, co] * X[p*mul + off, ci]
//
// The contraction runs over pixels, so both operands are "MN-major" for the tensor core: a TMA box of
// [128 pixels x 64 channels] (128-byte swizzle) is exactly the canonical MN-major SW128 layout (64 contiguous
// channels per K row, 8-row groups 1024 B apart, 64-channel blocks `LBO` apart).  Narrower tensors (16/32 channels)
// use the 32/64-byte swizzle variants.  M is always issued as 128 rows: rows >= cout read whatever follows in smem and
// are ignored by the epilogue (a D row depends only on its own A row, and UMMA M=64 costs the same time as M=128).
//
// A launch processes a table of work items (co tile, ci slice, tap shift, pixel-tile range = split-K); each CTA walks
// items persistently, accumulates one item in TMEM (double buffered) and the epilogue adds it into the fp32 OIHW
// gradient with red.global.add.f32 (fan-out to up to 4 filter taps for the parity-folded decoder weights).
#pragma once
#include "ptx.cuh"

namespace ub {

constexpr int kWgThreads = 192;

struct WgItem {
    int co0;              // first output channel (M offset)
    int ci0;              // first source channel of the B slice
    int dci0;             // first destination input-channel index in the gradient tensor
    int ncin;             // N (multiple of 16, <= 256)
    int dw, dh;           // B shift in source pixels
    int tile_begin, tile_end;  // pixel-tile range (split-K)
    int ndst;
    int dst_off[4];       // filter-tap offsets (r*S+s) the result is added to
};

struct WgParams {
    int tiles_w, tiles_h, tiles_n;
    int bw, bh, bn;
    int mulw, mulh;         // X coordinate = tile origin * mul + shift
    int cout;               // valid M rows overall
    int zc_box;             // channels per Z box (64, or cout when < 64)
    int xc_box;             // channels per X box (64, or cin slice when < 64)
    int stages;
    int num_items;
    const WgItem* items;
    float* grad;            // stem_mode: OIHW fp32 [64][3][7][7]; otherwise the PACKED gradient [co][tap][cin_total] fp32
    long long s_co, s_ci;   // stem_mode: element strides of the OIHW tensor
    int ntaps, cin_total;   // packed layout extents (ci contiguous: 16 consecutive columns = 64 B = 4 vector reductions)
    int stem_mode;          // 1: column j of the 32-wide stem window maps to (px=j/4, ch=j%4) -> ch*49 + r*7 + px-1
    int* err;
};

struct WgSmem {
    uint32_t a_bytes, b_bytes, stage_bytes, bar_off, total;
};
__host__ __device__ inline WgSmem wg_smem(int zc_box, int xc_box, int max_ncin, int stages) {
    WgSmem s;
    // A always reserves two 64-channel blocks so that the (ignored) upper M rows stay inside the stage
    s.a_bytes = 2 * 128 * 128;
    uint32_t nb = (max_ncin + xc_box - 1) / xc_box;
    s.b_bytes = nb * 128 * xc_box * 2;
    s.b_bytes = (s.b_bytes + 1023u) & ~1023u;
    s.stage_bytes = s.a_bytes + s.b_bytes;
    s.bar_off = s.stage_bytes * stages;
    s.total = s.bar_off + (2 * stages + 4) * 8 + 16;
    (void)zc_box;
    return s;
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmX,
             const __grid_constant__ WgParams P, const int max_ncin) {
    extern __shared__ uint8_t smem_raw[];
    griddep_launch();
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw_addr);
    const WgSmem L = wg_smem(P.zc_box, P.xc_box, max_ncin, P.stages);
    const uint32_t bar0 = base + L.bar_off;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (P.stages + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * P.stages + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * P.stages + 2 + a); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + L.bar_off + (2 * P.stages + 4) * 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t tmem_cols = 32;
    while (tmem_cols < 2u * max_ncin) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmZ);
        tma_prefetch_desc(&tmX);
        for (int s = 0; s < P.stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 4);
        }
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
    const int tiles_hw = P.tiles_w * P.tiles_h;
    const uint32_t zblk_bytes = 128u * P.zc_box * 2;  // one Z box
    const uint32_t xblk_bytes = 128u * P.xc_box * 2;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int it = blockIdx.x; it < P.num_items; it += gridDim.x) {
                const WgItem I = P.items[it];
                const int mrows = (P.cout - I.co0) < 128 ? (P.cout - I.co0) : 128;
                const int nzb = (mrows + P.zc_box - 1) / P.zc_box;
                const int nxb = (I.ncin + P.xc_box - 1) / P.xc_box;
                const uint32_t tx = nzb * zblk_bytes + nxb * xblk_bytes;
                for (int t = I.tile_begin; t < I.tile_end; ++t) {
                    const int tw = t % P.tiles_w, th = (t / P.tiles_w) % P.tiles_h, tn = t / tiles_hw;
                    if (!mbar_wait(empty_bar(stage), phase ^ 1)) {
                        atomicExch(P.err, 11);
                        goto role_done;
                    }
                    const uint32_t sa = base + stage * L.stage_bytes;
                    mbar_expect_tx(full_bar(stage), tx);
                    for (int b = 0; b < nzb; ++b)
                        tma_load_4d(sa + b * zblk_bytes, &tmZ, full_bar(stage), I.co0 + b * P.zc_box, tw * P.bw,
                                    th * P.bh, tn * P.bn);
                    for (int b = 0; b < nxb; ++b)
                        tma_load_4d(sa + L.a_bytes + b * xblk_bytes, &tmX, full_bar(stage), I.ci0 + b * P.xc_box,
                                    tw * P.bw * P.mulw + I.dw, th * P.bh * P.mulh + I.dh, tn * P.bn);
                    if (++stage == P.stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            // MN-major canonical layouts: SBO = 8 K-rows (pixels) * row bytes ; LBO = next channel block
            const uint32_t za_layout = P.zc_box >= 64 ? 2u : (P.zc_box == 32 ? 4u : 6u);
            const uint32_t xb_layout = P.xc_box >= 64 ? 2u : (P.xc_box == 32 ? 4u : 6u);
            const uint32_t z_row = P.zc_box * 2, x_row = P.xc_box * 2;
            for (int it = blockIdx.x; it < P.num_items; it += gridDim.x) {
                const WgItem I = P.items[it];
                const uint32_t idesc = umma_idesc_bf16(128, I.ncin, 1, 1);
                if (!mbar_wait(tempty_bar(acc), acc_phase ^ 1)) {
                    atomicExch(P.err, 12);
                    goto role_done;
                }
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * max_ncin;
                uint32_t accum = 0;
                for (int t = I.tile_begin; t < I.tile_end; ++t) {
                    if (!mbar_wait(full_bar(stage), phase)) {
                        atomicExch(P.err, 13);
                        goto role_done;
                    }
                    tc_fence_after();
                    const uint32_t sa = base + stage * L.stage_bytes;
                    const uint32_t sb = sa + L.a_bytes;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {  // 128 pixels = 8 x (UMMA_K = 16)
                        const uint64_t ad = umma_desc(sa + k * 16 * z_row, zblk_bytes, 8 * z_row, za_layout);
                        const uint64_t bd = umma_desc(sb + k * 16 * x_row, xblk_bytes, 8 * x_row, xb_layout);
                        umma_bf16(d_tmem, ad, bd, idesc, accum);
                        accum = 1;
                    }
                    umma_commit(empty_bar(stage));
                    if (++stage == P.stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(tfull_bar(acc));
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int it = blockIdx.x; it < P.num_items; it += gridDim.x) {
            const WgItem I = P.items[it];
            if (!mbar_wait(tfull_bar(acc), acc_phase)) {
                atomicExch(P.err, 14);
                goto role_done;
            }
            tc_fence_after();
            const int co = I.co0 + row;
            const bool valid = co < P.cout;
            float* grow = P.stem_mode ? P.grad + (long long)co * P.s_co
                                      : P.grad + (long long)co * P.ntaps * P.cin_total + I.dci0;
            for (int c0 = 0; c0 < I.ncin; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(tmem_base + (uint32_t(q * 32) << 16) + acc * max_ncin + c0, r);
                tmem_ld_wait();
                if (c0 + 16 >= I.ncin) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(acc));
                }
                if (!valid) continue;
                if (P.stem_mode) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int col = c0 + j;
                        const int px = col >> 2, ch = col & 3;
                        if (px >= 1 && ch < 3) atomicAdd(grow + ch * 49 + I.dst_off[0] + (px - 1), __uint_as_float(r[j]));
                    }
                } else {
                    for (int d = 0; d < I.ndst; ++d) {
                        float* gp = grow + (long long)I.dst_off[d] * P.cin_total + c0;
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            red_add_v4(gp + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                       __uint_as_float(r[j + 3]));
                    }
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
