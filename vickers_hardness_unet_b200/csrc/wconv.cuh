// Halo-resident 3x3 / stride-1 / pad-1 convolution with STREAMED weights on tcgen05 (sm_100a) for the wide layers
// (Cin, Cout multiples of 64/128: encoder.layer2-4, decoder conv2 of blocks 0-1 and all their stride-1 data gradients)
// — SURVEY.md section 8a rows A3, A5, A11.
//
// Why: the tap-table kernel (igemm.cuh) re-loads a 128-pixel A tile for every filter tap and a 256-row B tile for every
// K chunk; measured (profiles/r1s3_igemm_cluster_ab.txt) one SM ingests ~40 B/clk into shared memory, so A + B = 48 KB
// per 512 MMA cycles caps those layers at ~40 % of the tensor pipe whatever the L2 does (multicast does not help).
// Here, per 64-channel K chunk:
//   * ONE 4-D TMA box [64 ch, 18, 18, 1] brings the halo of a 16 x 16 output tile (two 128-pixel sub-tiles); all nine
//     taps of both sub-tiles read it through shifted UMMA descriptors (A ingest / 6),
//   * nine [128 cout x 64 ch] weight tiles stream through a ring; each is used by BOTH sub-tiles (B ingest / 2),
//   => 41.5 KB + 9 x 16 KB = 185 KB per 4608 MMA cycles = 40 B/clk: at the ingest cap instead of 2.3x above it.
// Accumulators: 2 sub-tiles x 128 columns, double buffered (512 TMEM columns).  Roles: warp 0 TMA producer, warp 1
// tcgen05 issuer + TMEM owner, warps 2-17 epilogue (four per TMEM lane quadrant; a thread owns one pixel x 64 channels).
// Weights are the tap-table kernel's packed matrix [cout][(r*3+s)*cin + ci] (same tensor map geometry).
#pragma once
#include "hconv.cuh"
#include "ptx.cuh"

namespace ub {

constexpr int kWcThreads = 64 + 16 * 32;
constexpr int kWcN = 128;                                  // UMMA N (output channels per work item)
constexpr uint32_t kWcHaloBytes = 18 * 18 * 128;           // 41472
constexpr uint32_t kWcHaloStage = (kWcHaloBytes + 1023u) & ~1023u;
constexpr uint32_t kWcBStage = kWcN * 128;                 // 16 KB
constexpr int kWcHaloStages = 2;

struct WconvParams {
    int H, W, N;
    int tiles_w, tiles_h, n_tiles;     // 16 x 16 output tiles per image; cout / 128
    int cin, cout;
    int bstages;
    int const_w;                       // inference: weights are constants of the stream -> first tiles fetched before the PDL wait
    const float* scale;                // [cout] or nullptr
    const float* shift;
    int relu;
    __nv_bfloat16* out;                // [N, H, W, cout]
    const __nv_bfloat16* residual;     // dense, same shape, or nullptr
    const float* residual32;           // fp32 residual (kRes32 instantiation: the training forward's up-sampled partial)
    float* stats;                      // [gridDim.x][cout][2] or nullptr
    int* err;
};

struct WconvSmem {
    uint32_t ss_off, cstat_off, bar_off, halo_off, b_off, total;
};
__host__ __device__ inline WconvSmem wconv_smem(int bstages) {
    WconvSmem s;
    s.ss_off = 0;                                  // scale[512], shift[512]
    s.cstat_off = 4096;                            // [4 quadrants x 2 sub-tiles][512 ch][2]: one writer warp per slot
    s.bar_off = s.cstat_off + 8 * 512 * 2 * 4;     // 36864
    s.halo_off = 37888;                            // 37 KB, 1 KB aligned
    s.b_off = s.halo_off + kWcHaloStages * kWcHaloStage;
    s.total = s.b_off + bstages * kWcBStage;
    return s;
}

template <bool kRes32>
__global__ void __launch_bounds__(kWcThreads, 1)
wconv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ WconvParams P) {
    extern __shared__ uint8_t smem_raw[];
    griddep_launch();
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw_addr);
    const WconvSmem L = wconv_smem(P.bstages);
    const uint32_t bar0 = base + L.bar_off;
    auto hfull = [&](int s) { return bar0 + 8u * s; };
    auto hempty = [&](int s) { return bar0 + 8u * (2 + s); };
    auto tfull = [&](int a) { return bar0 + 8u * (4 + a); };
    auto tempty = [&](int a) { return bar0 + 8u * (6 + a); };
    auto bfull = [&](int s) { return bar0 + 8u * (8 + s); };
    auto bempty = [&](int s) { return bar0 + 8u * (8 + P.bstages + s); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + L.bar_off + (8 + 2 * P.bstages) * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_img = P.tiles_w * P.tiles_h;
    const int total_items = tiles_img * P.N * P.n_tiles;
    const int chunks = P.cin >> 6;
    const int n_pre = P.const_w ? (P.bstages < 9 ? P.bstages : 9) : 0;   // weight tiles issued before the PDL wait

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < 2; ++s) {
            mbar_init(hfull(s), 1);
            mbar_init(hempty(s), 1);
            mbar_init(tfull(s), 1);
            mbar_init(tempty(s), 16);
        }
        for (int s = 0; s < P.bstages; ++s) {
            mbar_init(bfull(s), 1);
            mbar_init(bempty(s), 1);
        }
        fence_mbar_init();
        // PDL prologue: the first weight tiles of this CTA's first item (taps 0.. of chunk 0) do not depend on the
        // previous kernel when the weights are constants of the stream (inference)
        if (P.const_w && (int)blockIdx.x < total_items) {
            const int nt0 = blockIdx.x % P.n_tiles;
            for (int t = 0; t < n_pre; ++t) {
                mbar_expect_tx(bfull(t), kWcBStage);
                tma_load_2d(base + L.b_off + t * kWcBStage, &tmB, bfull(t), t * P.cin, nt0 * kWcN);
            }
        }
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
        tmem_relinquish();
    }
    griddep_wait();   // PDL: nothing above touches global memory that a kernel of this stream writes
    {
        float* ss = reinterpret_cast<float*>(sm + L.ss_off);
        float* cst = reinterpret_cast<float*>(sm + L.cstat_off);
        for (int c = threadIdx.x; c < 512; c += kWcThreads) {
            ss[c] = (P.scale && c < P.cout) ? P.scale[c] : 1.f;
            ss[512 + c] = (P.shift && c < P.cout) ? P.shift[c] : 0.f;
        }
        for (int c = threadIdx.x; c < 8 * 1024; c += kWcThreads) cst[c] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // work item i -> (n tile fastest, then tile column, row, image)
    auto decode = [&](int item, int& nt, int& tw, int& th, int& tn) {
        nt = item % P.n_tiles;
        int t = item / P.n_tiles;
        tw = t % P.tiles_w;
        t /= P.tiles_w;
        th = t % P.tiles_h;
        tn = t / P.tiles_h;
    };

    if (warp == 0) {
        // ================================================================= TMA producer (one thread)
        if (lane == 0) {
            int hs = 0, bs = 0, pre = n_pre;
            uint32_t hph = 0, bph = 0;
            for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
                int nt, tw, th, tn;
                decode(item, nt, tw, th, tn);
                for (int c = 0; c < chunks; ++c) {
                    if (!mbar_wait(hempty(hs), hph ^ 1)) {
                        atomicExch(P.err, 41);
                        goto done;
                    }
                    mbar_expect_tx(hfull(hs), kWcHaloBytes);
                    tma_load_4d(base + L.halo_off + hs * kWcHaloStage, &tmA, hfull(hs), c * 64, tw * 16 - 1, th * 16 - 1, tn);
                    if (++hs == kWcHaloStages) {
                        hs = 0;
                        hph ^= 1;
                    }
                    for (int tap = 0; tap < 9; ++tap) {
                        // one wait per 3 weight tiles: ring slots free in order, so the LAST slot of the group being free
                        // implies the first two are (bstages is a multiple of 3: a group never straddles the wrap).  A
                        // wait + expect_tx + TMA issue per tile made the single producer thread the bottleneck
                        // (~900 cycles per tile whatever its size).
                        if (tap % 3 == 0 && !mbar_wait(bempty(bs + 2), bph ^ 1)) {
                            atomicExch(P.err, 42);
                            goto done;
                        }
                        if (pre > 0) {
                            --pre;   // issued in the prologue
                        } else {
                            mbar_expect_tx(bfull(bs), kWcBStage);
                            tma_load_2d(base + L.b_off + bs * kWcBStage, &tmB, bfull(bs), tap * P.cin + c * 64, nt * kWcN);
                        }
                        if (++bs == P.bstages) {
                            bs = 0;
                            bph ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer (one thread)
        if (lane == 0) {
            int hs = 0, bs = 0, acc = 0;
            uint32_t hph = 0, bph = 0, aph = 0;
            const uint32_t idesc = umma_idesc_bf16(128, kWcN, 0, 0);
            const uint64_t a_desc0 = umma_desc(base + L.halo_off, 16, 18 * 128, 2u);
            const uint64_t b_desc0 = umma_desc(base + L.b_off, 16, 8 * 128, 2u);
            for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
                if (!mbar_wait(tempty(acc), aph ^ 1)) {
                    atomicExch(P.err, 43);
                    goto done;
                }
                tc_fence_after();
                const uint32_t d0 = tmem_base + acc * (2 * kWcN);
                for (int c = 0; c < chunks; ++c) {
                    if (!mbar_wait(hfull(hs), hph)) {
                        atomicExch(P.err, 44);
                        goto done;
                    }
                    tc_fence_after();
                    const uint64_t a_base = a_desc0 + (uint64_t)((hs * kWcHaloStage) >> 4);
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        if (!mbar_wait(bfull(bs), bph)) {
                            atomicExch(P.err, 45);
                            goto done;
                        }
                        tc_fence_after();
                        const uint64_t bd = b_desc0 + (uint64_t)((bs * kWcBStage) >> 4);
                        const uint64_t ad = a_base + (uint32_t)((((tap / 3) * 18 + tap % 3) * 128) >> 4);
#pragma unroll
                        for (int s = 0; s < 2; ++s) {
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                const uint64_t a = ad + (uint32_t)((s * 8 * 128 + kk * 32) >> 4);
                                const uint64_t b = bd + (uint32_t)((kk * 32) >> 4);
                                if (tap == 0 && kk == 0) umma_bf16(d0 + s * kWcN, a, b, idesc, c > 0 ? 1u : 0u);
                                else umma_bf16_c<true>(d0 + s * kWcN, a, b, idesc);
                            }
                        }
                        umma_commit(bempty(bs));
                        if (++bs == P.bstages) {
                            bs = 0;
                            bph ^= 1;
                        }
                    }
                    umma_commit(hempty(hs));
                    if (++hs == kWcHaloStages) {
                        hs = 0;
                        hph ^= 1;
                    }
                }
                umma_commit(tfull(acc));
                acc ^= 1;
                if (acc == 0) aph ^= 1;
            }
        }
    } else {
        // ================================================================= epilogue (16 warps)
        const int e = warp - 2;
        const int q = warp & 3;               // TMEM lane quadrant
        const int g = e >> 2;                 // 0..3: sub-tile g >> 1, channels (g & 1) * 64 .. + 63 of the N tile
        const int sub = g >> 1, ch0 = (g & 1) * 64;
        const int row = q * 32 + lane;
        const int wl = row & 7, hl = row >> 3;
        const float* ss = reinterpret_cast<const float*>(sm + L.ss_off);
        float* cst = reinterpret_cast<float*>(sm + L.cstat_off) + (q * 2 + sub) * 1024;
        int acc = 0;
        uint32_t aph = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
            int nt, tw, th, tn;
            decode(item, nt, tw, th, tn);
            const int ph = th * 16 + hl, pw = tw * 16 + sub * 8 + wl;
            const bool valid = ph < P.H && pw < P.W;
            const int cbase = nt * kWcN + ch0;
            const size_t off = (((size_t)tn * P.H + ph) * P.W + pw) * P.cout + cbase;
            if (!mbar_wait_warp(tfull(acc), aph, lane)) {
                atomicExch(P.err, 46);
                goto done;
            }
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * (2 * kWcN) + sub * kWcN + ch0;
#pragma unroll
            for (int half = 0; half < 2; ++half) {   // 2 x 32 channels
                uint32_t r[32];
                uint4 rv[4];
                if (!kRes32 && P.residual && valid) {
                    ld_global_nc_256(P.residual + off + half * 32, rv[0], rv[1]);
                    ld_global_nc_256(P.residual + off + half * 32 + 16, rv[2], rv[3]);
                }
                tmem_ld32(taddr + half * 32, r);
                tmem_ld_wait();
                if (half == 1) {  // last TMEM read of this accumulator: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty(acc));
                }
                uint4 o_prev = make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int j = 0; j < 4; ++j) {  // 4 x 8 channels
                    const int c = cbase + half * 32 + j * 8;
                    float v[8];
                    const float4 sc0 = *reinterpret_cast<const float4*>(ss + c);
                    const float4 sc1 = *reinterpret_cast<const float4*>(ss + c + 4);
                    const float4 sh0 = *reinterpret_cast<const float4*>(ss + 512 + c);
                    const float4 sh1 = *reinterpret_cast<const float4*>(ss + 512 + c + 4);
                    v[0] = __uint_as_float(r[j * 8 + 0]) * sc0.x + sh0.x;
                    v[1] = __uint_as_float(r[j * 8 + 1]) * sc0.y + sh0.y;
                    v[2] = __uint_as_float(r[j * 8 + 2]) * sc0.z + sh0.z;
                    v[3] = __uint_as_float(r[j * 8 + 3]) * sc0.w + sh0.w;
                    v[4] = __uint_as_float(r[j * 8 + 4]) * sc1.x + sh1.x;
                    v[5] = __uint_as_float(r[j * 8 + 5]) * sc1.y + sh1.y;
                    v[6] = __uint_as_float(r[j * 8 + 6]) * sc1.z + sh1.z;
                    v[7] = __uint_as_float(r[j * 8 + 7]) * sc1.w + sh1.w;
                    if (kRes32) {
                        if (valid) {
                            uint4 r0, r1;
                            ld_global_nc_256(P.residual32 + off + half * 32 + j * 8, r0, r1);
                            v[0] += __uint_as_float(r0.x); v[1] += __uint_as_float(r0.y); v[2] += __uint_as_float(r0.z);
                            v[3] += __uint_as_float(r0.w); v[4] += __uint_as_float(r1.x); v[5] += __uint_as_float(r1.y);
                            v[6] += __uint_as_float(r1.z); v[7] += __uint_as_float(r1.w);
                        }
                    } else if (P.residual && valid) {
                        v[0] += bf16_lo(rv[j].x); v[1] += bf16_hi(rv[j].x);
                        v[2] += bf16_lo(rv[j].y); v[3] += bf16_hi(rv[j].y);
                        v[4] += bf16_lo(rv[j].z); v[5] += bf16_hi(rv[j].z);
                        v[6] += bf16_lo(rv[j].w); v[7] += bf16_hi(rv[j].w);
                    }
                    if (P.relu) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
                    }
                    uint4 o;
                    o.x = pack_bf16(v[0], v[1]);
                    o.y = pack_bf16(v[2], v[3]);
                    o.z = pack_bf16(v[4], v[5]);
                    o.w = pack_bf16(v[6], v[7]);
                    if (j & 1) {   // two 8-channel groups = 32 bytes: one 256-bit store per pair
                        if (valid) st_global_256(P.out + off + half * 32 + (j - 1) * 8, o_prev, o);
                    } else {
                        o_prev = o;
                    }
                    if (P.stats && (j & 1)) {
                        // statistics of the bf16 values just stored, 16 channels (this group and the previous one) at a
                        // time over the warp's 32 pixels: [16 sums | 16 sums of squares] in ONE 32-value reduction
                        const uint32_t w8[8] = {o_prev.x, o_prev.y, o_prev.z, o_prev.w, o.x, o.y, o.z, o.w};
                        float sv[32];
#pragma unroll
                        for (int m = 0; m < 8; ++m) {
                            const float lo = valid ? bf16_lo(w8[m]) : 0.f, hi = valid ? bf16_hi(w8[m]) : 0.f;
                            sv[2 * m] = lo; sv[2 * m + 1] = hi;
                            sv[16 + 2 * m] = lo * lo; sv[16 + 2 * m + 1] = hi * hi;
                        }
                        const float t = warp_reduce32(sv, lane);  // lane l: sum (l < 16) / sum of squares (l >= 16) of channel l & 15
                        cst[2 * (c - 8 + (lane & 15)) + (lane >> 4)] += t;
                    }
                }
            }
            acc ^= 1;
            if (acc == 0) aph ^= 1;
        }
        if (P.stats) {
            named_bar_sync(1, 512);
            const float* call = reinterpret_cast<const float*>(sm + L.cstat_off);
            float* dst = P.stats + static_cast<size_t>(blockIdx.x) * P.cout * 2;
            for (int j = threadIdx.x - 64; j < 2 * P.cout; j += 512) {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) a += call[k * 1024 + j];
                dst[j] = a;
            }
        }
    }
done:
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}


// ------------------------------------------------------------------------------------------------ wpconv (parity mode)
// The up-sampled half of a wide decoder conv1 (decoder.blocks.0-2): out = scale * conv3x3(nearest2x(low)) computed on the
// LOW-RES tensor as four 2x2-tap parity convolutions (weights summed per parity: the PK_DEC1 packed matrix the tap-table
// kernel uses, columns [koff, koff + 4*cup)), written as a bf16 partial that the skip-channel conv (wconv / tconv) then
// adds through its residual path — nothing up-sampled or concatenated ever exists in memory, and the 4 parity launches
// + two-source tap table of the igemm path (23 % of the inference step) become two halo-resident launches.
// Per 64-channel K chunk: ONE TMA halo box [64 ch, 18, 18, 1] of a 16 x 16 low-res tile (two 128-pixel sub-tiles) and
// sixteen [64 cout x 64 ch] weight tiles (parity x tap), each used by both sub-tiles.  Accumulators: 4 parities x 2
// sub-tiles x 64 columns = all 512 TMEM columns (single buffered: the epilogue of an item is ~10 % of its MMA time).
struct WpconvParams {
    int Hl, Wl, N;                     // LOW-RES extent; the output is [N, 2*Hl, 2*Wl, cout]
    int tiles_w, tiles_h, n_tiles;     // 16 x 16 low-res tiles per image; cout / 64
    int cup, cout, koff;               // K chunks = cup / 64; koff = first low-tap column of the packed weight matrix
    int bstages;
    const float* scale;                // [cout] or nullptr
    __nv_bfloat16* out;
    float* out32;                      // kOut32 instantiation: fp32 partial for the training forward
    int* err;
};
constexpr uint32_t kWpBStage = 64 * 128;   // 8 KB

struct WpconvSmem {
    uint32_t ss_off, bar_off, halo_off, b_off, total;
};
__host__ __device__ inline WpconvSmem wpconv_smem(int bstages) {
    WpconvSmem s;
    s.ss_off = 0;                      // scale[512]
    s.bar_off = 2048;
    s.halo_off = 3072;
    s.b_off = s.halo_off + kWcHaloStages * kWcHaloStage;
    s.total = s.b_off + bstages * kWpBStage;
    return s;
}

template <bool kOut32>
__global__ void __launch_bounds__(kWcThreads, 1)
wpconv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
              const __grid_constant__ WpconvParams P) {
    extern __shared__ uint8_t smem_raw[];
    griddep_launch();
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw_addr);
    const WpconvSmem L = wpconv_smem(P.bstages);
    const uint32_t bar0 = base + L.bar_off;
    auto hfull = [&](int s) { return bar0 + 8u * s; };
    auto hempty = [&](int s) { return bar0 + 8u * (2 + s); };
    const uint32_t tfull = bar0 + 8u * 4, tempty = bar0 + 8u * 5;
    auto bfull = [&](int s) { return bar0 + 8u * (6 + s); };
    auto bempty = [&](int s) { return bar0 + 8u * (6 + P.bstages + s); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + L.bar_off + (6 + 2 * P.bstages) * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_items = P.tiles_w * P.tiles_h * P.N * P.n_tiles;
    const int chunks = P.cup >> 6;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < 2; ++s) {
            mbar_init(hfull(s), 1);
            mbar_init(hempty(s), 1);
        }
        mbar_init(tfull, 1);
        mbar_init(tempty, 16);
        for (int s = 0; s < P.bstages; ++s) {
            mbar_init(bfull(s), 1);
            mbar_init(bempty(s), 1);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
        tmem_relinquish();
    }
    griddep_wait();   // PDL: nothing above touches global memory
    {
        float* ss = reinterpret_cast<float*>(sm + L.ss_off);
        for (int c = threadIdx.x; c < 512; c += kWcThreads) ss[c] = (P.scale && c < P.cout) ? P.scale[c] : 1.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto decode = [&](int item, int& nt, int& tw, int& th, int& tn) {
        nt = item % P.n_tiles;
        int t = item / P.n_tiles;
        tw = t % P.tiles_w;
        t /= P.tiles_w;
        th = t % P.tiles_h;
        tn = t / P.tiles_h;
    };

    if (warp == 0) {
        // ================================================================= TMA producer (one thread)
        if (lane == 0) {
            int hs = 0, bs = 0;
            uint32_t hph = 0, bph = 0;
            for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
                int nt, tw, th, tn;
                decode(item, nt, tw, th, tn);
                for (int c = 0; c < chunks; ++c) {
                    if (!mbar_wait(hempty(hs), hph ^ 1)) {
                        atomicExch(P.err, 51);
                        goto done;
                    }
                    mbar_expect_tx(hfull(hs), kWcHaloBytes);
                    tma_load_4d(base + L.halo_off + hs * kWcHaloStage, &tmA, hfull(hs), c * 64, tw * 16 - 1, th * 16 - 1, tn);
                    if (++hs == kWcHaloStages) {
                        hs = 0;
                        hph ^= 1;
                    }
                    for (int pt = 0; pt < 16; ++pt) {  // parity (pt >> 2) x low-res neighbour (pt & 3)
                        // one wait per 4 weight tiles (see wconv_kernel; bstages is a multiple of 4)
                        if ((pt & 3) == 0 && !mbar_wait(bempty(bs + 3), bph ^ 1)) {
                            atomicExch(P.err, 52);
                            goto done;
                        }
                        mbar_expect_tx(bfull(bs), kWpBStage);
                        tma_load_2d(base + L.b_off + bs * kWpBStage, &tmB, bfull(bs), P.koff + (pt & 3) * P.cup + c * 64,
                                    (pt >> 2) * P.cout + nt * 64);
                        if (++bs == P.bstages) {
                            bs = 0;
                            bph ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer (one thread)
        if (lane == 0) {
            int hs = 0, bs = 0;
            uint32_t hph = 0, bph = 0, aph = 0;
            const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
            const uint64_t a_desc0 = umma_desc(base + L.halo_off, 16, 18 * 128, 2u);
            const uint64_t b_desc0 = umma_desc(base + L.b_off, 16, 8 * 128, 2u);
            for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
                if (!mbar_wait(tempty, aph ^ 1)) {
                    atomicExch(P.err, 53);
                    goto done;
                }
                tc_fence_after();
                for (int c = 0; c < chunks; ++c) {
                    if (!mbar_wait(hfull(hs), hph)) {
                        atomicExch(P.err, 54);
                        goto done;
                    }
                    tc_fence_after();
                    const uint64_t a_base = a_desc0 + (uint64_t)((hs * kWcHaloStage) >> 4);
#pragma unroll
                    for (int pt = 0; pt < 16; ++pt) {
                        if (!mbar_wait(bfull(bs), bph)) {
                            atomicExch(P.err, 55);
                            goto done;
                        }
                        tc_fence_after();
                        const int par = pt >> 2, ab = pt & 3;
                        const uint64_t bd = b_desc0 + (uint64_t)((bs * kWpBStage) >> 4);
                        // low-res neighbour (a, b) of output parity (ph, pw) sits at halo pixel (a + ph, b + pw)
                        const uint64_t ad = a_base + (uint32_t)(((((ab >> 1) + (par >> 1)) * 18 + (ab & 1) + (par & 1)) * 128) >> 4);
#pragma unroll
                        for (int s = 0; s < 2; ++s) {
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                const uint64_t a = ad + (uint32_t)((s * 8 * 128 + kk * 32) >> 4);
                                const uint64_t b = bd + (uint32_t)((kk * 32) >> 4);
                                const uint32_t d = tmem_base + (par * 2 + s) * 64;
                                if (ab == 0 && kk == 0) umma_bf16(d, a, b, idesc, c > 0 ? 1u : 0u);
                                else umma_bf16_c<true>(d, a, b, idesc);
                            }
                        }
                        umma_commit(bempty(bs));
                        if (++bs == P.bstages) {
                            bs = 0;
                            bph ^= 1;
                        }
                    }
                    umma_commit(hempty(hs));
                    if (++hs == kWcHaloStages) {
                        hs = 0;
                        hph ^= 1;
                    }
                }
                umma_commit(tfull);
                aph ^= 1;
            }
        }
    } else {
        // ================================================================= epilogue (16 warps): warp group g = output parity
        const int e = warp - 2;
        const int q = warp & 3;
        const int par = e >> 2, ph = par >> 1, pw = par & 1;
        const int row = q * 32 + lane;
        const int wl = row & 7, hl = row >> 3;
        const float* ss = reinterpret_cast<const float*>(sm + L.ss_off);
        const int Ho = 2 * P.Hl, Wo = 2 * P.Wl;
        uint32_t aph = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
            int nt, tw, th, tn;
            decode(item, nt, tw, th, tn);
            if (!mbar_wait_warp(tfull, aph, lane)) {
                atomicExch(P.err, 56);
                goto done;
            }
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int lh = th * 16 + hl, lw = tw * 16 + s * 8 + wl;
                const bool valid = lh < P.Hl && lw < P.Wl;
                const size_t ooff = (((size_t)tn * Ho + 2 * lh + ph) * Wo + 2 * lw + pw) * P.cout + nt * 64;
                __nv_bfloat16* op = P.out + ooff;
                const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + (par * 2 + s) * 64;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t r[32];
                    tmem_ld32(taddr + half * 32, r);
                    tmem_ld_wait();
                    if (s == 1 && half == 1) {  // last TMEM read of this item: hand the accumulators back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty);
                    }
                    uint4 o_prev = make_uint4(0, 0, 0, 0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = nt * 64 + half * 32 + j * 8;
                        const float4 sc0 = *reinterpret_cast<const float4*>(ss + c);
                        const float4 sc1 = *reinterpret_cast<const float4*>(ss + c + 4);
                        if (kOut32) {
                            if (valid) {
                                uint4 lo, hi;   // 8 fp32 channels = one 256-bit store
                                lo.x = __float_as_uint(__uint_as_float(r[j * 8 + 0]) * sc0.x);
                                lo.y = __float_as_uint(__uint_as_float(r[j * 8 + 1]) * sc0.y);
                                lo.z = __float_as_uint(__uint_as_float(r[j * 8 + 2]) * sc0.z);
                                lo.w = __float_as_uint(__uint_as_float(r[j * 8 + 3]) * sc0.w);
                                hi.x = __float_as_uint(__uint_as_float(r[j * 8 + 4]) * sc1.x);
                                hi.y = __float_as_uint(__uint_as_float(r[j * 8 + 5]) * sc1.y);
                                hi.z = __float_as_uint(__uint_as_float(r[j * 8 + 6]) * sc1.z);
                                hi.w = __float_as_uint(__uint_as_float(r[j * 8 + 7]) * sc1.w);
                                st_global_256(P.out32 + ooff + half * 32 + j * 8, lo, hi);
                            }
                            continue;
                        }
                        uint4 o;
                        o.x = pack_bf16(__uint_as_float(r[j * 8 + 0]) * sc0.x, __uint_as_float(r[j * 8 + 1]) * sc0.y);
                        o.y = pack_bf16(__uint_as_float(r[j * 8 + 2]) * sc0.z, __uint_as_float(r[j * 8 + 3]) * sc0.w);
                        o.z = pack_bf16(__uint_as_float(r[j * 8 + 4]) * sc1.x, __uint_as_float(r[j * 8 + 5]) * sc1.y);
                        o.w = pack_bf16(__uint_as_float(r[j * 8 + 6]) * sc1.z, __uint_as_float(r[j * 8 + 7]) * sc1.w);
                        if (j & 1) {   // 16 channels = one 256-bit store
                            if (valid) st_global_256(op + half * 32 + (j - 1) * 8, o_prev, o);
                        } else {
                            o_prev = o;
                        }
                    }
                }
            }
            aph ^= 1;
        }
    }
done:
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace ub
