// TMA-fed halo-resident 3x3 / stride-1 / pad-1 convolution on tcgen05 (sm_100a), second generation of hconv.cuh for
// the narrow, high-resolution layers (Cin, Cout <= 64): encoder.layer1, decoder blocks 2-4 conv2, their data gradients,
// and — in "parity" mode — decoder.blocks.4.conv1, whose nearest-2x upsample is folded into four 2x2-tap convolutions
// on the low-resolution tensor.
//
// What the role-cycle counters of hconv.cuh showed on a B200 (profiles/r1s3_hconv_role_cycles.txt): with 128-pixel
// tiles every producer -> MMA -> epilogue hand-off costs ~1000 cycles per tile per CTA (mbarrier round trips, tcgen05
// fences, per-tile address arithmetic of single warps) while the tile's MMAs need 300-1700 and its HBM traffic 350;
// and the cp.async halo loader is issue-bound for wide rows.  Hence here:
//   * one pipeline step covers NT = 2..4 sub-tiles of 128 output pixels (a 16..32 x 16 output rectangle, or the four
//     output parities of an 8 x 16 low-res tile): one hand-off per 256..512 pixels, halo overhead 1.2x instead of 1.4x,
//   * the halo is ONE 4-D TMA box load [C, 8*NT+2, 18, 1] issued by one thread: out-of-bounds zero fill is the conv
//     padding, the hardware swizzle writes the K-major operand layout, completion is an mbarrier transaction count
//     (no proxy fence, no producer warps),
//   * a tap is a whole-row shift of the A descriptor's start address inside the resident halo (fully unrolled issue
//     loop on the uniform datapath); the weights of all taps stay resident in smem,
//   * 8 or 16 epilogue warps (two / four per TMEM lane quadrant) split the sub-tiles / channel groups, batch their TMEM loads,
//     prefetch the residual before waiting for the accumulator and store straight from registers.
// Roles: warp 0 = TMA producer, warp 1 = tcgen05 issuer + TMEM owner, warps 2.. = epilogue.
#pragma once
#include "hconv.cuh"
#include "ptx.cuh"

namespace ub {

// epilogue warps per CTA: 8 when two CTAs share an SM, 16 (four per TMEM lane quadrant) when one CTA owns it
__host__ __device__ constexpr int tc_epi_warps(int occ) { return occ == 2 ? 8 : 16; }
__host__ __device__ constexpr int tc_threads(int occ) { return 64 + 32 * tc_epi_warps(occ); }
constexpr int kTcHaloH = 18;        // 16 output rows + 2

struct TconvParams {
    int H, W, N;                    // OUTPUT extent
    int mode;                       // 0: plain 3x3 over src[N,H,W,cin]; 1: parity (src = low-res [N,H/2,W/2,cin], 2x2 taps);
                                    // 2: 7x7/s2 stem over the packed image xp[N][2H][2W+8][4] (see tc_issue_stem)
                                    // 3: seg head with the three filter columns stacked as N-blocks (see tc_issue_head3)
    int nt;                         // accumulators (sub-tiles) per pipeline step
    int tiles_w, tiles_h;           // tile grid per image (plain: 8*nt x 16 output px; parity: 8 x 16 low-res px)
    int cin, cout;
    int halo_w;                     // halo pixels per row
    uint32_t stage_bytes, tx_bytes; // smem per stage (1 KB multiple) / bytes one halo box delivers
    int stages, nacc;
    uint32_t w_bytes;
    const __nv_bfloat16* wpk;       // resident weights, pre-swizzled rows of cin channels (pack_hconv_w / PK_HPAR)
    const float* scale;
    const float* shift;
    int relu;
    __nv_bfloat16* out;             // [N, H, W, cout]
    const __nv_bfloat16* residual;  // same shape or nullptr
    const float* residual32;        // fp32 residual of the same shape (training forward of decoder.blocks.2.conv1: the
                                    // partial over the up-sampled channels stays fp32), or nullptr
    float* stats;                   // [gridDim.x][cout][2] or nullptr
    int stage_out;                  // 1: the epilogue transposes through swizzled smem and stores whole sub-tiles with TMA
                                    // (cout >= 32: a register store of 16 channels per pixel touches 32 different 128-byte
                                    // lines per warp instruction and the LSU serialises them: measured 2100 of 5900
                                    // cycles per tile on the stem); 0: straight from registers (cout = 16);
                                    // 2: parity mode, one interleaved 32 x 16 hi-res tile per pipeline step
    int spw;                        // epilogue threads per sub-tile (named-barrier population), stage_out only
    // segmentation-head epilogue (kHead kernels: Conv2d(16,1,3,padding=1) + bias, the conv has cout = 16 accumulator
    // columns of which column 0 holds the bf16 high part and column 1 the bf16 low part of the fp32 weights)
    const float* head_bias;         // [1]
    float* logits;                  // fp32 [N,1,H,W] or nullptr
    float* prob;                    // sigmoid(logits) or nullptr
    uint8_t* mask;                  // {0,255} = logits >= thresh_logit, or nullptr
    float thresh_logit;
    int* err;
    long long* prof;                // selftest only: [grid][16] cycle counters per role phase (dbg & 8)
    int dbg;                        // selftest only: 1 = skip halo loads, 2 = skip MMA issue, 4 = skip epilogue math + stores; 16 = constant weights (inference)
};

struct TconvSmem {
    uint32_t ss_off, cstat_off, bar_off, w_off, halo_off, out_off, total;
};
// out_bytes: epilogue staging (nt sub-tiles x 128 pixels x cout x 2 B) or 0
__host__ __device__ inline TconvSmem tconv_smem(uint32_t w_bytes, uint32_t stage_bytes, int stages, uint32_t out_bytes = 0) {
    TconvSmem s;
    s.ss_off = 0;                                  // scale[64], shift[64]
    s.cstat_off = 512;                             // [<= 16 epilogue warps][64 ch][2]
    s.bar_off = s.cstat_off + 16 * 128 * 4;        // 8704
    s.w_off = 9216;
    s.halo_off = (s.w_off + w_bytes + 1023u) & ~1023u;
    s.out_off = s.halo_off + stages * stage_bytes;
    s.total = s.out_off + out_bytes;
    return s;
}

// All tcgen05.mma of one pipeline step, fully unrolled over taps and 16-channel K steps so that every descriptor is a
// uniform base plus a compile-time multiple of the (uniform) row pitch: the issuing thread's instruction stream stays on
// the uniform datapath (an offset table read from shared memory costs an R2UR round trip per MMA: 130 instead of 55
// cycles per MMA measured).  KS = cin / 16; rows are cin*2 bytes; pitch16 = halo row pitch in 16-byte units.
template <int KS>
__device__ __forceinline__ void tc_issue_plain(uint32_t d_tmem, uint64_t a_base, uint64_t b_base, uint32_t pitch16,
                                               uint32_t cout, uint32_t idesc, int nt) {
    constexpr uint32_t kRow16 = KS * 2;                      // row bytes / 16
    const uint32_t b_tap = cout * kRow16;                    // one tap's weight slab in 16-byte units
    for (int s = 0; s < nt; ++s) {
        const uint64_t a_s = a_base + (uint32_t)(8 * s) * kRow16;
        const uint32_t d = d_tmem + s * cout;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
                const uint64_t ad = a_s + (uint32_t)(tap / 3) * pitch16 + (uint32_t)((tap % 3) * kRow16 + kk * 2);
                const uint64_t bd = b_base + (uint32_t)tap * b_tap + (uint32_t)(kk * 2);
                if (tap == 0 && kk == 0) umma_bf16_c<false>(d, ad, bd, idesc);
                else umma_bf16_c<true>(d, ad, bd, idesc);
            }
        }
    }
}
// parity mode: accumulator s = (ph, pw); low-res neighbour (a, b) sits at halo pixel (a + ph, b + pw), so a halo shift
// (r, c) of the 3x3 neighbourhood serves every parity with r - ph and c - pw in {0, 1}: the centre all four, an edge shift
// two, a corner one.  ONE MMA per shift covers a contiguous run of the parity accumulators (N = 4, 2, 3 or 1 x cout; the
// N = 3 runs of the shifts (1,0) / (1,2) carry a zero weight block for the parity in the middle): 9 reads of the 4 KB A
// operand per K step instead of 16 — that read is what binds these layers (profiles/r1s3_umma_smem_bandwidth.txt).
// Weight slabs (PK_HPAR, pack.cuh hpar_block): 18 blocks of [cout][cin] in the issue order below.
template <int KS>
__device__ __forceinline__ void tc_issue_parity(uint32_t d_tmem, uint64_t a_base, uint64_t b_base, uint32_t pitch16,
                                                uint32_t cout, uint32_t idesc) {
    constexpr uint32_t kRow16 = KS * 2;
    const uint32_t b_blk = cout * kRow16;                    // one [cout][cin] weight block in 16-byte units
    const uint32_t n_step = ((cout >> 3) & 0x3Fu) << 17;     // idesc n_dim increment per extra parity
    //                          centre  top  bottom  left  right  corners
    constexpr int kR[9]   = {1, 0, 2, 1, 1, 0, 0, 2, 2};
    constexpr int kC[9]   = {1, 1, 1, 0, 2, 0, 2, 0, 2};
    constexpr int kLo[9]  = {0, 0, 2, 0, 1, 0, 1, 2, 3};     // first parity accumulator of the run
    constexpr int kNp[9]  = {4, 2, 2, 3, 3, 1, 1, 1, 1};     // parities in the run
    constexpr int kBlk[9] = {0, 4, 6, 8, 11, 14, 15, 16, 17};
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        const uint32_t d = d_tmem + kLo[j] * cout;
        const uint32_t id = idesc + (uint32_t)(kNp[j] - 1) * n_step;
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
            const uint64_t ad = a_base + (uint32_t)kR[j] * pitch16 + (uint32_t)(kC[j] * kRow16 + kk * 2);
            const uint64_t bd = b_base + (uint32_t)kBlk[j] * b_blk + (uint32_t)(kk * 2);
            if (j == 0 && kk == 0) umma_bf16_c<false>(d, ad, bd, id);
            else umma_bf16_c<true>(d, ad, bd, id);
        }
    }
}

// Stem mode (encoder.conv1, 7x7 stride 2, 3(+1 zero) input channels): out(ho, wo) = sum_r sum_{px<8} xp[2ho+r-3][2wo+px][4ch]
// * w[r][px][ch] — 7 row taps with K = 32 each.  The operand row of output pixel wo is the 64 contiguous bytes of input
// row 2ho+r-3 starting at byte 16*wo: consecutive rows OVERLAP (pitch 16 B, length 64 B).  The non-swizzled K-major
// descriptor expresses exactly that: a core matrix (8 rows x 16 B) is 128 contiguous bytes, core matrices adjacent in K
// are LBO = 16 B apart, adjacent in M SBO = 128 B apart — so the raw input rows in shared memory ARE the im2col matrix
// and nothing is expanded.  Sub-tile s = output row, 128 consecutive wo.  Weights: [K chunk of 8][cout group of 8][8][8]
// (LBO = 1024 B between K chunks, SBO = 128 B between groups of 8 output channels), PK_STEM2.
__device__ __forceinline__ void tc_issue_stem(uint32_t d_tmem, uint32_t stage_addr, uint32_t w_addr, uint32_t row_pitch,
                                              uint32_t idesc, int nt) {
    const uint64_t a_base = umma_desc(stage_addr, 16, 128, 0u);
    const uint64_t b_base = umma_desc(w_addr, 1024, 128, 0u);
    for (int s = 0; s < nt; ++s) {
        const uint32_t d = d_tmem + s * 64;
#pragma unroll
        for (int r = 0; r < 7; ++r) {
            const uint64_t a_r = a_base + (((uint32_t)(2 * s + r) * row_pitch) >> 4);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                const uint64_t ad = a_r + (uint32_t)(kk * 2);
                const uint64_t bd = b_base + (uint32_t)(((r * 4 + kk * 2) * 1024) >> 4);
                if (r == 0 && kk == 0) umma_bf16_c<false>(d, ad, bd, idesc);
                else umma_bf16_c<true>(d, ad, bd, idesc);
            }
        }
    }
}

// Head mode 3 (Conv2d(16, 1, 3, padding=1)): the A operand (4 KB per MMA) is what a 16-channel conv pays for, so the three
// filter COLUMNS become N-blocks of one MMA instead of three shifted A reads: a sub-tile is 4 image rows x 32 consecutive
// pixels of a halo box that is exactly 32 pixels wide (image rows are contiguous in smem, so SBO = 8 pixels), accumulator
// column 2c / 2c + 1 of pixel x holds (hi / lo weight parts of filter column c) . input(x) summed over the three filter
// rows (row taps = whole-row shifts of the A start address): 3 MMAs per sub-tile instead of 9.  The epilogue adds the
// neighbours' columns with two warp shuffles (lane = pixel): logit(x) = D_0(x-1) + D_1(x) + D_2(x+1); pixels 0 and 31 of a
// row are halo, a tile yields 30 output columns.
__device__ __forceinline__ void tc_issue_head3(uint32_t d_tmem, uint64_t a_base, uint64_t b_base, uint32_t idesc, int nt) {
    for (int s = 0; s < nt; ++s) {
        const uint32_t d = d_tmem + s * 16;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const uint64_t ad = a_base + (uint32_t)((4 * s + r) * 32 * 2);   // 32 pixels x 32 B per image row, 16-byte units
            const uint64_t bd = b_base + (uint32_t)(r * 16 * 2);            // 16 rows x 32 B per filter row
            if (r == 0) umma_bf16_c<false>(d, ad, bd, idesc);
            else umma_bf16_c<true>(d, ad, bd, idesc);
        }
    }
}

// Mode 3 for a whole conv (Cout <= 32): same sub-tile geometry as tc_issue_head3, B = the three filter-column slabs of
// filter row r (contiguous in the PK_HCONV layout: N = 3 * cout), accumulator block c of pixel x holds W[.][c] . input(x)
// summed over filter rows and input channels; the epilogue combines D_0(x-1) + D_1(x) + D_2(x+1) with warp shuffles.
// 3 * KS reads of the 4 KB A operand per sub-tile instead of 9 * KS.
template <int KS>
__device__ __forceinline__ void tc_issue_stack(uint32_t d_tmem, uint64_t a_base, uint64_t b_base, uint32_t cout,
                                               uint32_t idesc, int nt) {
    constexpr uint32_t kRow16 = KS * 2;
    const uint32_t b_row = 3 * cout * kRow16;                // three tap slabs = one filter row, 16-byte units
    for (int s = 0; s < nt; ++s) {
        const uint32_t d = d_tmem + s * 3 * cout;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
                const uint64_t ad = a_base + (uint32_t)((4 * s + r) * 32) * kRow16 + (uint32_t)(kk * 2);
                const uint64_t bd = b_base + (uint32_t)r * b_row + (uint32_t)(kk * 2);
                if (r == 0 && kk == 0) umma_bf16_c<false>(d, ad, bd, idesc);
                else umma_bf16_c<true>(d, ad, bd, idesc);
            }
        }
    }
}

// kIph = accumulator column groups (16 channels of one sub-tile) each epilogue thread handles per pipeline step
template <int kOcc, int kIph, bool kStage, bool kHead = false, bool kStack = false>
__global__ void __launch_bounds__(tc_threads(kOcc), kOcc)
tconv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmD,
             const __grid_constant__ TconvParams P) {
    constexpr int kEw = tc_epi_warps(kOcc), kTcThreads = tc_threads(kOcc);
    extern __shared__ uint8_t smem_raw[];
    griddep_launch();
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw_addr);
    const TconvSmem L = tconv_smem(P.w_bytes, P.stage_bytes, P.stages, kStage ? (uint32_t)P.nt * 128u * P.cout * 2u : 0u);
    const uint32_t bar0 = base + L.bar_off;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (P.stages + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * P.stages + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * P.stages + P.nacc + a); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + L.bar_off + (2 * P.stages + 2 * P.nacc) * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = P.tiles_w * P.tiles_h * P.N;
    const int sub_cols = (kStack && !kHead) ? 3 * P.cout : P.cout;   // accumulator columns per sub-tile
    const int acc_cols = P.nt * sub_cols;
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(P.nacc * acc_cols)) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        if (kStage) tma_prefetch_desc(&tmD);
        for (int s = 0; s < P.stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < P.nacc; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), kEw);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), tmem_cols);
        tmem_relinquish();
    }
    const bool const_w = (P.dbg & 16) != 0;   // inference: weights / scale / shift are constants of the stream
    if (!const_w) griddep_wait();             // PDL: nothing above touches global memory
    {
        float* ss = reinterpret_cast<float*>(sm + L.ss_off);
        float* cst = reinterpret_cast<float*>(sm + L.cstat_off);
        for (int c = threadIdx.x; c < 64; c += kTcThreads) {
            ss[c] = (P.scale && c < P.cout) ? P.scale[c] : 1.f;
            ss[64 + c] = (P.shift && c < P.cout) ? P.shift[c] : 0.f;
        }
        for (int c = threadIdx.x; c < kEw * 128; c += kTcThreads) cst[c] = 0.f;
        // resident weights: linear copy (the array is pre-swizzled relative to a 1 KB-aligned origin)
        const uint4* wsrc = reinterpret_cast<const uint4*>(P.wpk);
        uint4* wdst = reinterpret_cast<uint4*>(sm + L.w_off);
        for (uint32_t i = threadIdx.x; i < P.w_bytes / 16; i += kTcThreads) wdst[i] = __ldg(wsrc + i);
        fence_async_smem();  // generic-proxy writes of the weights -> visible to the tensor core (async proxy)
    }
    if (const_w) griddep_wait();   // the weight copy above overlapped the previous kernel's tail
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t row_bytes = (uint32_t)P.cin * 2;
    // Role-cycle counters exist only in builds with -DUB_TC_PROF (tests/native/selftest): seven live 64-bit counters cost
    // the 96-register kernel 16-40 bytes of local-memory spill per thread, measurable on every launch.
#ifdef UB_TC_PROF
    const bool prof = (P.dbg & 8) != 0;
    long long t_wait = 0, t_work = 0, t_a = 0, t_b = 0, t_c = 0, t_d = 0, t_e = 0, tc = prof ? clock64() : 0;
#define UB_TC_TICK(var)                         \
    if (prof) {                                 \
        const long long now_ = clock64();       \
        var += now_ - tc;                       \
        tc = now_;                              \
    }
#define UB_TC_PROF_ONLY(...) __VA_ARGS__
#else
#define UB_TC_TICK(var)
#define UB_TC_PROF_ONLY(...)
#endif

    if (warp == 0) {
        // ================================================================= TMA producer (one thread)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (HcTileIter it(P.tiles_w, P.tiles_h, total_tiles); it.valid(); it.next()) {
                if (!mbar_wait(empty_bar(stage), phase ^ 1)) {
                    atomicExch(P.err, 31);
                    break;
                }
                UB_TC_TICK(t_wait)
                int x0 = (P.mode ? it.tw * 8 : it.tw * 8 * P.nt) - 1, y0 = it.th * 16 - 1;
                if (P.mode == 2) {  // 128-byte groups of 8 pixel pairs of the packed image: output column wo <-> pair wo
                    x0 = it.tw * 16;
                    y0 = 2 * it.th * P.nt - 3;
                } else if (kStack) {
                    x0 = it.tw * 30 - 1;
                    y0 = it.th * 4 * P.nt - 1;
                }
                if (P.dbg & 1) {
                    mbar_arrive(full_bar(stage));
                } else {
                    mbar_expect_tx(full_bar(stage), P.tx_bytes);
                    tma_load_4d(base + L.halo_off + stage * P.stage_bytes, &tmA, full_bar(stage), 0, x0, y0, it.tn);
                }
                UB_TC_TICK(t_work)
                if (++stage == P.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            UB_TC_PROF_ONLY(if (prof) {
                P.prof[blockIdx.x * 16 + 0] = t_wait;
                P.prof[blockIdx.x * 16 + 1] = t_work;
            })
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer (one thread)
        if (lane == 0) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            const uint32_t idesc = umma_idesc_bf16(128, (kStack && !kHead) ? 3 * P.cout : P.cout, 0, 0);
            const uint32_t layout = row_bytes == 32 ? 6u : (row_bytes == 64 ? 4u : 2u);
            const uint64_t b_base = umma_desc(base + L.w_off, 16, 8 * row_bytes, layout);
            const uint64_t a_base0 = umma_desc(base + L.halo_off, 16, ((kStack) ? 8 : P.halo_w) * row_bytes, layout);
            const uint32_t pitch16 = ((uint32_t)P.halo_w * row_bytes) >> 4;
            const int ks = P.cin >> 4;
            for (int n = blockIdx.x; n < total_tiles; n += gridDim.x) {
                if (!mbar_wait(tempty_bar(acc), acc_phase ^ 1)) {
                    atomicExch(P.err, 32);
                    break;
                }
                UB_TC_TICK(t_a)
                if (!mbar_wait(full_bar(stage), phase)) {
                    atomicExch(P.err, 33);
                    break;
                }
                tc_fence_after();
                UB_TC_TICK(t_wait)
                const uint64_t a_base = a_base0 + (uint64_t)((stage * P.stage_bytes) >> 4);
                const uint32_t d_tmem = tmem_base + acc * acc_cols;
                if (!(P.dbg & 2)) {
                    if (P.mode == 2) {
                        tc_issue_stem(d_tmem, base + L.halo_off + stage * P.stage_bytes, base + L.w_off,
                                      (uint32_t)P.halo_w * 16u, idesc, P.nt);
                    } else if (kStack && kHead) {
                        tc_issue_head3(d_tmem, a_base, b_base, idesc, P.nt);
                    } else if (kStack) {
                        switch (ks) {
                            case 1: tc_issue_stack<1>(d_tmem, a_base, b_base, P.cout, idesc, P.nt); break;
                            case 2: tc_issue_stack<2>(d_tmem, a_base, b_base, P.cout, idesc, P.nt); break;
                            default: tc_issue_stack<4>(d_tmem, a_base, b_base, P.cout, idesc, P.nt); break;
                        }
                    } else if (P.mode) {
                        switch (ks) {
                            case 1: tc_issue_parity<1>(d_tmem, a_base, b_base, pitch16, P.cout, idesc); break;
                            case 2: tc_issue_parity<2>(d_tmem, a_base, b_base, pitch16, P.cout, idesc); break;
                            default: tc_issue_parity<4>(d_tmem, a_base, b_base, pitch16, P.cout, idesc); break;
                        }
                    } else {
                        switch (ks) {
                            case 1: tc_issue_plain<1>(d_tmem, a_base, b_base, pitch16, P.cout, idesc, P.nt); break;
                            case 2: tc_issue_plain<2>(d_tmem, a_base, b_base, pitch16, P.cout, idesc, P.nt); break;
                            default: tc_issue_plain<4>(d_tmem, a_base, b_base, pitch16, P.cout, idesc, P.nt); break;
                        }
                    }
                }
                UB_TC_TICK(t_work)
                umma_commit(empty_bar(stage));
                umma_commit(tfull_bar(acc));
                UB_TC_TICK(t_b)
                if (++stage == P.stages) {
                    stage = 0;
                    phase ^= 1;
                }
                if (++acc == P.nacc) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
            UB_TC_PROF_ONLY(if (prof) {
                P.prof[blockIdx.x * 16 + 4] = t_a;
                P.prof[blockIdx.x * 16 + 5] = t_wait;
                P.prof[blockIdx.x * 16 + 6] = t_work;
                P.prof[blockIdx.x * 16 + 7] = t_b;
            })
        }
    } else {
        // ================================================================= epilogue (kEw warps): thread = pixel of a sub-tile
        const int e = warp - 2;
        const int q = warp & 3;                        // TMEM lane quadrant this warp may read
        const int half = e >> 2;                       // which share of the accumulator column groups
        const int row = q * 32 + lane;
        const int wl = row & 7, hl = row >> 3;
        const int cgs = P.cout >> 4;                   // 16-channel groups per sub-tile
        const int items = P.nt * cgs;
        const int i0 = half * kIph;                    // this thread's items: [i0, min(i0 + kIph, items))
        const float* ss = reinterpret_cast<const float*>(sm + L.ss_off);
        float* cst = reinterpret_cast<float*>(sm + L.cstat_off) + e * 128;
        // per-item constants: sub-tile, channel group, pixel offset inside the tile
        int it_s[kIph], it_c0[kIph], it_dh[kIph], it_dw[kIph];
#pragma unroll
        for (int k = 0; k < kIph; ++k) {
            const int i = i0 + k < items ? i0 + k : items - 1;
            it_s[k] = i / cgs;
            it_c0[k] = (i - it_s[k] * cgs) * 16;
            if (P.mode == 2) {
                it_dh[k] = it_s[k];
                it_dw[k] = row;
            } else if (kStack) {
                it_dh[k] = 4 * it_s[k] + q;
                it_dw[k] = lane - 1;
            } else if (P.mode) {
                it_dh[k] = 2 * hl + (it_s[k] >> 1);
                it_dw[k] = 2 * wl + (it_s[k] & 1);
            } else {
                it_dh[k] = hl;
                it_dw[k] = 8 * it_s[k] + wl;
            }
        }
        const int n_mine = items - i0 < kIph ? (items - i0 > 0 ? items - i0 : 0) : kIph;
        const float head_bias = kHead ? __ldg(P.head_bias) : 0.f;
        // staged stores: all items of a thread belong to ONE sub-tile (host guarantees kIph divides cout/16)
        const int my_sub = it_s[0];
        const uint32_t out_row_bytes = (uint32_t)P.cout * 2u, out_swz = out_row_bytes / 16u - 1u;
        // stage_out == 2 (parity mode): the four parity sub-tiles interleave into ONE dense 32 x 16 hi-res staging tile
        // (a register store would write cout*2 bytes every second pixel: half-used 128-byte lines from every warp store)
        const bool par_stage = kStage && P.stage_out == 2;
        const bool store_issuer = par_stage ? (e == 0 && lane == 0)
                                            : (kStage && n_mine > 0 && q == 0 && lane == 0 && (i0 % cgs) == 0);
        const int th_px = P.mode == 2 ? P.nt : ((kStack) ? 4 * P.nt : (P.mode ? 32 : 16));
        const int tw_px = P.mode == 2 ? 128 : ((kStack) ? 30 : (P.mode ? 16 : 8 * P.nt));
        int acc = 0;
        uint32_t acc_phase = 0;
        for (HcTileIter it(P.tiles_w, P.tiles_h, total_tiles); it.valid(); it.next()) {
            const int h0 = it.th * th_px, w0 = it.tw * tw_px;
            const int pix0 = (it.tn * P.H + h0) * P.W + w0;      // pixel index of the tile origin (fits 32 bits)
            bool valid[kIph];
            uint4 rv[kIph][2];
#pragma unroll
            for (int k = 0; k < kIph; ++k) {
                valid[k] = k < n_mine && h0 + it_dh[k] < P.H && w0 + it_dw[k] < P.W;
                if (kStack) valid[k] = valid[k] && lane >= 1 && lane <= 30;
                if (P.residual && valid[k]) {  // issued before the accumulator wait: the loads overlap the tile's MMAs
                    ld_global_nc_256(P.residual + (size_t)(pix0 + it_dh[k] * P.W + it_dw[k]) * P.cout + it_c0[k],
                                     rv[k][0], rv[k][1]);
                }
            }
            UB_TC_TICK(t_a)
            if (!mbar_wait_warp(tfull_bar(acc), acc_phase, lane)) {
                atomicExch(P.err, 34);
                break;
            }
            tc_fence_after();
            UB_TC_TICK(t_wait)
            UB_TC_TICK(t_b)
            const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * acc_cols;
#pragma unroll
            for (int k = 0; k < kIph; ++k) {
                uint32_t r[16];
                if (kStack && !kHead) {
                    // D_0 of the left neighbour + own D_1 + D_2 of the right neighbour (lane = pixel of the image row)
                    uint32_t t[16];
                    const uint32_t ta = taddr + it_s[k] * sub_cols + it_c0[k];
                    tmem_ld16(ta, r);
                    tmem_ld16(ta + P.cout, t);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        r[j] = __float_as_uint(__shfl_up_sync(0xffffffffu, __uint_as_float(r[j]), 1) + __uint_as_float(t[j]));
                    tmem_ld16(ta + 2 * P.cout, t);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        r[j] = __float_as_uint(__uint_as_float(r[j]) + __shfl_down_sync(0xffffffffu, __uint_as_float(t[j]), 1));
                } else {
                    tmem_ld16(taddr + it_s[k] * P.cout + it_c0[k], r);
                    tmem_ld_wait();
                }
                if (k == 0) { UB_TC_TICK(t_c) }
                if (k == kIph - 1) {  // last TMEM read of this accumulator: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(acc));
                }
                if (k >= n_mine || (P.dbg & 4)) continue;
                if (kHead) {
                    // fp32 logit = (w_hi + w_lo) . x + bias; sigmoid(x) >= t  <=>  x >= logit(t)
                    float lg = __uint_as_float(r[0]) + __uint_as_float(r[1]);
                    if (kStack) {   // stacked filter columns: left neighbour's column 0, own column 1, right neighbour's 2
                        const float c1 = __uint_as_float(r[2]) + __uint_as_float(r[3]);
                        const float c2 = __uint_as_float(r[4]) + __uint_as_float(r[5]);
                        lg = __shfl_up_sync(0xffffffffu, lg, 1) + c1 + __shfl_down_sync(0xffffffffu, c2, 1);
                    }
                    lg += head_bias;
                    if (valid[k]) {
                        const size_t o = (size_t)(pix0 + it_dh[k] * P.W + it_dw[k]);
                        if (P.logits) P.logits[o] = lg;
                        if (P.prob) P.prob[o] = 1.f / (1.f + __expf(-lg));
                        if (P.mask) P.mask[o] = lg >= P.thresh_logit ? 255 : 0;
                    }
                    continue;
                }
                const int c0 = it_c0[k];
                float v[16];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 sc = *reinterpret_cast<const float4*>(ss + c0 + 4 * j);
                    const float4 sh = *reinterpret_cast<const float4*>(ss + 64 + c0 + 4 * j);
                    v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) * sc.x + sh.x;
                    v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) * sc.y + sh.y;
                    v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) * sc.z + sh.z;
                    v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) * sc.w + sh.w;
                }
                if (P.residual32 && valid[k]) {   // 16 fp32 channels = two 256-bit loads (one launch per train step: not prefetched)
                    const float* rp = P.residual32 + (size_t)(pix0 + it_dh[k] * P.W + it_dw[k]) * P.cout + c0;
                    uint4 f[4];
                    ld_global_nc_256(rp, f[0], f[1]);
                    ld_global_nc_256(rp + 8, f[2], f[3]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        v[4 * j + 0] += __uint_as_float(f[j].x); v[4 * j + 1] += __uint_as_float(f[j].y);
                        v[4 * j + 2] += __uint_as_float(f[j].z); v[4 * j + 3] += __uint_as_float(f[j].w);
                    }
                }
                if (P.residual && valid[k]) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        v[j * 8 + 0] += bf16_lo(rv[k][j].x); v[j * 8 + 1] += bf16_hi(rv[k][j].x);
                        v[j * 8 + 2] += bf16_lo(rv[k][j].y); v[j * 8 + 3] += bf16_hi(rv[k][j].y);
                        v[j * 8 + 4] += bf16_lo(rv[k][j].z); v[j * 8 + 5] += bf16_hi(rv[k][j].z);
                        v[j * 8 + 6] += bf16_lo(rv[k][j].w); v[j * 8 + 7] += bf16_hi(rv[k][j].w);
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
                if (kStage && k == 0) {
                    // the previous tile's TMA store has finished READING the staging buffer before anyone rewrites it; the
                    // first item's TMEM load and epilogue math above overlapped that store
                    if (store_issuer) tma_wait_read<0>();
                    UB_TC_TICK(t_work)
                    if (par_stage) named_bar_sync(2, 32 * kEw);
                    else named_bar_sync(2 + 2 * my_sub, P.spw);
                    UB_TC_TICK(t_d)
                }
                if (kStage) {
                    // row = TMEM lane = pixel in TMA box order; 16-byte chunks XOR-swizzled like the output tensor map
                    uint32_t so = (uint32_t)(par_stage ? it_dh[k] * 16 + it_dw[k] : row) * out_row_bytes + (uint32_t)c0 * 2u;
                    uint8_t* sp = sm + L.out_off + (par_stage ? 0u : it_s[k] * 128u * out_row_bytes);
                    *reinterpret_cast<uint4*>(sp + (so ^ (((so >> 7) & out_swz) << 4))) = o[0];
                    so += 16;
                    *reinterpret_cast<uint4*>(sp + (so ^ (((so >> 7) & out_swz) << 4))) = o[1];
                } else if (valid[k]) {
                    st_global_256(P.out + (size_t)(pix0 + it_dh[k] * P.W + it_dw[k]) * P.cout + c0, o[0], o[1]);
                }
                if (P.stats) {
                    // statistics of the bf16 values just stored (masked pixels contribute 0); fixed summation order;
                    // sums and sums of squares of the 16 channels share ONE 32-value reduction (hconv.cuh warp_reduce32)
                    float sv[32];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint32_t w4[4] = {o[j].x, o[j].y, o[j].z, o[j].w};
#pragma unroll
                        for (int m = 0; m < 4; ++m) {
                            const float lo = valid[k] ? bf16_lo(w4[m]) : 0.f, hi = valid[k] ? bf16_hi(w4[m]) : 0.f;
                            sv[j * 8 + 2 * m] = lo; sv[j * 8 + 2 * m + 1] = hi;
                            sv[16 + j * 8 + 2 * m] = lo * lo; sv[16 + j * 8 + 2 * m + 1] = hi * hi;
                        }
                    }
                    const float t = warp_reduce32(sv, lane);   // lane l: sum (l < 16) / sum of squares (l >= 16) of channel l & 15
                    cst[2 * (c0 + (lane & 15)) + (lane >> 4)] += t;
                }
            }
            UB_TC_TICK(t_work)
            if (kStage && n_mine > 0 && !(P.dbg & 4)) {
                fence_async_smem();                       // staging writes -> visible to the TMA store (async proxy)
                if (par_stage) {
                    named_bar_sync(3, 32 * kEw);
                    if (store_issuer) {
                        tma_store_4d(&tmD, base + L.out_off, 0, w0, h0, it.tn);   // [cout, 16, 32, 1], clipped by the TMA
                        tma_commit();
                    }
                } else {
                named_bar_sync(3 + 2 * my_sub, P.spw);
                if (store_issuer) {
                    const uint32_t src = base + L.out_off + my_sub * 128u * out_row_bytes;
                    if (P.mode == 2) tma_store_4d(&tmD, src, 0, w0, h0 + my_sub, it.tn);
                    else tma_store_4d(&tmD, src, 0, w0 + 8 * my_sub, h0, it.tn);   // partial tiles are clipped by the TMA
                    tma_commit();
                }
                }
            }
            UB_TC_TICK(t_e)
            if (++acc == P.nacc) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
        if (kStage && store_issuer) tma_wait_all<0>();  // all output stores complete before the CTA exits
        UB_TC_PROF_ONLY(if (prof && threadIdx.x == 64) {
            P.prof[blockIdx.x * 16 + 8] = t_a;
            P.prof[blockIdx.x * 16 + 9] = t_wait;
            P.prof[blockIdx.x * 16 + 10] = t_b;
            P.prof[blockIdx.x * 16 + 11] = t_work;
            P.prof[blockIdx.x * 16 + 12] = t_c;
            P.prof[blockIdx.x * 16 + 13] = t_d;
            P.prof[blockIdx.x * 16 + 14] = t_e;
        })
        if (P.stats) {
            named_bar_sync(1, 32 * kEw);
            const float* call = reinterpret_cast<const float*>(sm + L.cstat_off);
            float* dst = P.stats + static_cast<size_t>(blockIdx.x) * P.cout * 2;
            for (int j = threadIdx.x - 64; j < 2 * P.cout; j += 32 * kEw) {
                float a = 0.f;
#pragma unroll
                for (int w8 = 0; w8 < kEw; ++w8) a += call[w8 * 128 + j];
                dst[j] = a;
            }
        }
    }
#undef UB_TC_TICK
#undef UB_TC_PROF_ONLY
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

}  // namespace ub
