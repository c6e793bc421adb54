// Multi-source "tap table" implicit-GEMM convolution on tcgen05 (sm_100a).
//
//   D[pixel, co] = sum over taps t, channels c of  SRC_t[pixel*mul + (dw_t, dh_t), c0_t + c] * Wpk[co, k(t, c)]
//
// One kernel covers every conv of the U-Net hot path (SURVEY.md section 8a rows A1, A3-A6 and their dgrads):
//   3x3 s1/s2, 1x1 s2, the 7x7 s2 stem (7 row-taps over an overlapped-window view), the decoder's fused
//   nearest-2x-upsample + concat conv (two sources, parity decomposition) and stride-2 dgrad (strided output view).
// Operands are NHWC bf16; A tiles are 128 output pixels (bw x bh x bn box) x chunk channels fetched by 4-D TMA
// tile loads whose out-of-bounds zero fill implements the conv padding; B tiles are [ntile x chunk] slices of the
// packed K-major weight matrix; the fp32 accumulator lives in TMEM (double buffered); the epilogue applies
// scale/shift (folded BN), residual add and ReLU, or emits raw output + per-channel batch statistics partials,
// and writes bf16 through a swizzled smem staging tile with TMA stores (which clip partial tiles).
#pragma once
#include "epilogue.cuh"
#include "ptx.cuh"

namespace ub {

constexpr int kMaxTaps = 16;
constexpr int kIgemmThreads = 192;  // warp0: TMA producer, warp1: MMA issuer + TMEM owner, warps 2-5: epilogue

struct IgemmTap {
    int16_t dw, dh;    // coordinate offset in source pixels
    int16_t c0;        // first source channel of this tap
    int16_t nchunks;   // number of chunk_elems-wide K chunks
    int32_t src;       // which A tensor map (0/1)
};

struct IgemmParams {
    int tiles_w, tiles_h, tiles_n, n_tiles;  // M-tile grid (w,h,image) and number of Cout tiles
    int bw, bh, bn;                          // pixels per tile along w, h, image: bw*bh*bn == 128
    int ntile;                               // UMMA N: 16..256, multiple of 16
    int chunk_elems;                         // K elements per pipeline stage: 16, 32 or 64 (32/64/128-byte swizzle)
    int stages;
    int csize;                               // thread-block cluster size (1, 2 or 4): the CTAs of a cluster work on csize
                                             // adjacent M tiles of the same N tile and share ONE copy of the weight tile:
                                             // each loads ntile/csize rows and multicasts them to all (L2 -> SM traffic
                                             // per MMA drops from A + B to A + B/csize)
    int num_taps, total_chunks;
    IgemmTap taps[kMaxTaps];
    int mulw[2], mulh[2];                    // source coordinate = tile origin * mul + tap offset
    int Wo, Ho, Nimg, cout;                  // logical output extent (masking) and channel count
    int out_cblk;                            // channels per staging/store block = min(64, ntile)
    const float* scale;                      // [cout] or nullptr (identity)
    const float* shift;                      // [cout] or nullptr
    int relu;
    const __nv_bfloat16* residual;           // NHWC-strided bf16 or nullptr
    long long res_sw, res_sh, res_sn;        // element strides of the residual
    float* stats;                            // [gridDim.x][cout][2] per-CTA (sum, sumsq of the bf16 output) or nullptr
    int* err;                                // device error flag (set on pipeline timeout)
};

// dynamic shared memory carve-up (all offsets relative to a 1024-aligned base)
struct IgemmSmem {
    uint32_t a_bytes, b_bytes, stage_bytes, staging_off, staging_bytes, ss_off, part_off, cstat_off, bar_off, total;
};
__host__ __device__ inline IgemmSmem igemm_smem(int ntile, int chunk_elems, int stages, int out_cblk) {
    IgemmSmem s;
    const uint32_t cb = chunk_elems * 2;
    s.a_bytes = 128 * cb;
    s.b_bytes = (ntile * cb + 1023u) & ~1023u;
    s.stage_bytes = s.a_bytes + s.b_bytes;  // a_bytes is a multiple of 1024 for every chunk size (>= 4 KB)
    s.staging_off = s.stage_bytes * stages;
    s.staging_bytes = 128 * out_cblk * 2;   // x2 buffers
    s.ss_off = s.staging_off + 2 * s.staging_bytes;
    s.part_off = s.ss_off + 2 * 512 * 4;    // scale/shift for up to 512 channels
    s.cstat_off = s.part_off + 8 * 64 * 2 * 4;  // stats partials [8 groups][64 ch][2]
    s.bar_off = s.cstat_off + 512 * 2 * 4;      // per-CTA running statistics [512 ch][2]
    s.total = s.bar_off + (2 * stages + 4) * 8 + 16;
    return s;
}

__global__ void __launch_bounds__(kIgemmThreads, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
             const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmD,
             const __grid_constant__ IgemmParams P) {
    extern __shared__ uint8_t smem_raw[];
    griddep_launch();
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw_addr);
    const IgemmSmem L = igemm_smem(P.ntile, P.chunk_elems, P.stages, P.out_cblk);

    const uint32_t bar0 = base + L.bar_off;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (P.stages + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * P.stages + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * P.stages + 2 + a); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + L.bar_off + (2 * P.stages + 4) * 8);
    volatile int* s_abort = reinterpret_cast<volatile int*>(sm + L.bar_off + (2 * P.stages + 4) * 8 + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int csize = P.csize;
    const int crank = csize > 1 ? (int)cluster_ctarank() : 0;
    const uint16_t cmask = (uint16_t)((1u << csize) - 1u);
    const int m_tiles = P.tiles_w * P.tiles_h * P.tiles_n;
    // work groups: csize adjacent M tiles x one N tile; cluster c takes groups c, c + #clusters, ...  A CTA whose M tile
    // lies beyond the tensor (last, partial group) runs on zero-filled operands and its stores are clipped.
    const int total_groups = ((m_tiles + csize - 1) / csize) * P.n_tiles;
    const int group0 = blockIdx.x / csize, group_step = gridDim.x / csize;
    uint32_t tmem_cols = 32;
    while (tmem_cols < 2u * P.ntile) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA0);
        tma_prefetch_desc(&tmA1);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmD);
        for (int s = 0; s < P.stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), csize);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 4);
        }
        *s_abort = 0;
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), tmem_cols);
        tmem_relinquish();
    }
    griddep_wait();   // PDL: nothing above touches global memory
    // per-channel scale/shift -> smem (identity when absent)
    {
        float* ss = reinterpret_cast<float*>(sm + L.ss_off);
        float* cst = reinterpret_cast<float*>(sm + L.cstat_off);
        for (int c = threadIdx.x; c < 512; c += blockDim.x) {
            ss[c] = (P.scale && c < P.cout) ? P.scale[c] : 1.f;
            ss[512 + c] = (P.shift && c < P.cout) ? P.shift[c] : 0.f;
            cst[2 * c] = 0.f;
            cst[2 * c + 1] = 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (csize > 1) cluster_sync_all();  // peers' barriers are initialised before anything is multicast into them
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================= TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx = L.a_bytes + P.ntile * P.chunk_elems * 2;
            const int b_rows = P.ntile / csize;
            const uint32_t b_part = (uint32_t)b_rows * P.chunk_elems * 2;
            for (int grp = group0; grp < total_groups; grp += group_step) {
                const int nt = grp % P.n_tiles;
                int m = (grp / P.n_tiles) * csize + crank;
                const int tw = m % P.tiles_w;
                m /= P.tiles_w;
                const int th = m % P.tiles_h;
                const int tn = m / P.tiles_h;
                int kidx = 0;
                for (int t = 0; t < P.num_taps; ++t) {
                    const IgemmTap tap = P.taps[t];
                    const CUtensorMap* mp = tap.src ? &tmA1 : &tmA0;
                    const int cw = tw * P.bw * P.mulw[tap.src] + tap.dw;
                    const int ch = th * P.bh * P.mulh[tap.src] + tap.dh;
                    for (int cc = 0; cc < tap.nchunks; ++cc, ++kidx) {
                        if (!mbar_wait(empty_bar(stage), phase ^ 1)) {
                            *s_abort = 1;
                            atomicExch(P.err, 1);
                            goto role_done;
                        }
                        const uint32_t sa = base + stage * L.stage_bytes;
                        mbar_expect_tx(full_bar(stage), tx);
                        tma_load_4d(sa, mp, full_bar(stage), tap.c0 + cc * P.chunk_elems, cw, ch, tn * P.bn);
                        if (csize > 1)
                            tma_load_2d_mc(sa + L.a_bytes + crank * b_part, &tmB, full_bar(stage), kidx * P.chunk_elems,
                                           nt * P.ntile + crank * b_rows, cmask);
                        else
                            tma_load_2d(sa + L.a_bytes, &tmB, full_bar(stage), kidx * P.chunk_elems, nt * P.ntile);
                        if (++stage == P.stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer (single thread)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            const uint32_t idesc = umma_idesc_bf16(128, P.ntile, 0, 0);
            const uint32_t sbo = 8u * P.chunk_elems * 2;
            const uint32_t layout = (P.chunk_elems == 64) ? 2u : (P.chunk_elems == 32 ? 4u : 6u);
            const int ksteps = P.chunk_elems / 16;
            for (int grp = group0; grp < total_groups; grp += group_step) {
                if (!mbar_wait(tempty_bar(acc), acc_phase ^ 1)) {
                    *s_abort = 1;
                    atomicExch(P.err, 2);
                    goto role_done;
                }
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * P.ntile;
                uint32_t accum = 0;
                for (int kc = 0; kc < P.total_chunks; ++kc) {
                    if (!mbar_wait(full_bar(stage), phase)) {
                        *s_abort = 1;
                        atomicExch(P.err, 3);
                        goto role_done;
                    }
                    tc_fence_after();
                    const uint32_t sa = base + stage * L.stage_bytes;
                    const uint32_t sb = sa + L.a_bytes;
                    for (int k = 0; k < ksteps; ++k) {
                        const uint64_t ad = umma_desc(sa + k * 32, 16, sbo, layout);
                        const uint64_t bd = umma_desc(sb + k * 32, 16, sbo, layout);
                        umma_bf16(d_tmem, ad, bd, idesc, accum);
                        accum = 1;
                    }
                    // frees this smem stage (in every CTA that multicasts into it) once the MMAs above have read it
                    if (csize > 1) umma_commit_mc(empty_bar(stage), cmask);
                    else umma_commit(empty_bar(stage));
                    if (++stage == P.stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        // ================================================================= epilogue (4 warps, one TMEM lane quadrant each)
        const int q = warp & 3;
        const int et = threadIdx.x - 64;        // 0..127
        EpiParams E;
        E.ntile = P.ntile; E.out_cblk = P.out_cblk; E.relu = P.relu; E.cout = P.cout;
        E.bw = P.bw; E.bh = P.bh; E.bn = P.bn; E.Wo = P.Wo; E.Ho = P.Ho; E.Nimg = P.Nimg;
        E.residual = P.residual; E.res_sw = P.res_sw; E.res_sh = P.res_sh; E.res_sn = P.res_sn;
        E.stats = P.stats;
        EpiSmem ES;
        ES.staging = sm + L.staging_off; ES.staging_bytes = L.staging_bytes;
        ES.ss = reinterpret_cast<const float*>(sm + L.ss_off);
        ES.part = reinterpret_cast<float*>(sm + L.part_off);
        ES.cst = reinterpret_cast<float*>(sm + L.cstat_off);
        int acc = 0;
        uint32_t acc_phase = 0;
        uint32_t blk_counter = 0;
        for (int grp = group0; grp < total_groups; grp += group_step) {
            const int nt = grp % P.n_tiles;
            int m = (grp / P.n_tiles) * csize + crank;
            const int tw = m % P.tiles_w;
            m /= P.tiles_w;
            const int th = m % P.tiles_h;
            const int tn = m / P.tiles_h;
            if (!mbar_wait(tfull_bar(acc), acc_phase)) {
                *s_abort = 1;
                atomicExch(P.err, 4);
                goto role_done;
            }
            tc_fence_after();
            epilogue_tile(E, ES, &tmD, tmem_base + acc * P.ntile, tempty_bar(acc), tw, th, tn, nt, q, lane, et, blk_counter);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        epilogue_finish(E, ES, q, lane, et);
    }
role_done:
    tc_fence_before();
    __syncthreads();
    if (csize > 1) cluster_sync_all();  // no CTA exits while a peer may still multicast into it / arrive on its barriers
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

}  // namespace ub
