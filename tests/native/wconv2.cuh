// CTA-pair (tcgen05 cta_group::2) version of the wide halo conv (wconv.cuh) — 3x3 / stride-1 / pad-1, Cin multiple of
// 64, Cout multiple of 128.
//
// Why: every single-CTA MMA of this network runs at what one SM can read from its shared memory for the tensor core,
// measured ~72 B/clk (profiles/r1s3_umma_smem_bandwidth.txt): a 128 x N x 16 MMA reads 4 KB of A + 32*N B of B, so N = 128
// takes 113 cycles instead of its 64-cycle compute floor (57 % of the tensor pipe, what ncu shows for wconv_kernel).
// In a CTA pair one MMA covers M = 256 pixels (128 rows of A from each CTA's own halo) and each CTA holds only HALF of
// the B tile's rows: per CTA 4 KB + 16*N B per MMA -> N = 256 needs 8 KB per 128-cycle MMA = 64 B/clk: under the limit.
// Work item of a pair = one 16 x 16 output tile (CTA rank r owns the 8-pixel-wide half r: its own [64 ch,10,18,1] TMA
// halo box per K chunk) x kN output channels.  Per tap each CTA streams kN/2 weight rows.  Only the leader (rank 0)
// issues tcgen05.mma.cta_group::2; its commits are multicast to both CTAs' barriers.  The peer's "its operands have
// landed" reaches the leader through a forwarder thread (peer's otherwise idle MMA warp: wait own barrier -> remote
// mbarrier.arrive on the leader's peer-full barrier); both CTAs' epilogue warps arrive on the leader's tempty barrier.
#pragma once
#include "../../vickers_hardness_unet_b200/csrc/hconv.cuh"
#include "../../vickers_hardness_unet_b200/csrc/ptx.cuh"

namespace ub {

constexpr int kW2Threads = 64 + 16 * 32;
constexpr uint32_t kW2HaloBytes = 10 * 18 * 128;                       // 23040
constexpr uint32_t kW2HaloStage = (kW2HaloBytes + 1023u) & ~1023u;     // 23552
constexpr int kW2HaloStages = 2;

struct Wconv2Params {
    int H, W, N;
    int tiles_w, tiles_h, n_tiles;     // 16 x 16 output tiles per image; cout / kN
    int cin, cout;
    int bstages;
    const float* scale;
    const float* shift;
    int relu;
    __nv_bfloat16* out;
    const __nv_bfloat16* residual;
    float* stats;                      // [gridDim.x][cout][2] or nullptr
    int* err;
};

struct Wconv2Smem {
    uint32_t ss_off, cstat_off, bar_off, halo_off, b_off, total;
};
__host__ __device__ inline Wconv2Smem wconv2_smem(int kN, int bstages) {
    Wconv2Smem s;
    s.ss_off = 0;                                  // scale[512], shift[512]
    s.cstat_off = 4096;                            // [4 quadrants][512 ch][2]
    s.bar_off = s.cstat_off + 4 * 512 * 2 * 4;     // 20480
    s.halo_off = 21504;
    s.b_off = s.halo_off + kW2HaloStages * kW2HaloStage;   // 68608, 1 KB aligned
    s.total = s.b_off + bstages * (kN / 2) * 128;
    return s;
}

template <int kN>
__global__ void __launch_bounds__(kW2Threads, 1)
wconv2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
              const __grid_constant__ Wconv2Params P) {
    constexpr uint32_t kBStage = (kN / 2) * 128;
    constexpr int kCg = kN / 4;                    // channels per epilogue warp group
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw_addr);
    const Wconv2Smem L = wconv2_smem(kN, P.bstages);
    const uint32_t bar0 = base + L.bar_off;
    auto hfull = [&](int s) { return bar0 + 8u * s; };
    auto hempty = [&](int s) { return bar0 + 8u * (2 + s); };
    auto phfull = [&](int s) { return bar0 + 8u * (4 + s); };     // leader only: the peer's halo stage s has landed
    auto tfull = [&](int a) { return bar0 + 8u * (6 + a); };
    auto tempty = [&](int a) { return bar0 + 8u * (8 + a); };     // leader only: 32 epilogue warps of the pair
    auto bfull = [&](int s) { return bar0 + 8u * (10 + s); };
    auto bempty = [&](int s) { return bar0 + 8u * (10 + P.bstages + s); };
    auto pbfull = [&](int s) { return bar0 + 8u * (10 + 2 * P.bstages + s); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + L.bar_off + (10 + 3 * P.bstages) * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int total_items = P.tiles_w * P.tiles_h * P.N * P.n_tiles;
    const int pair0 = blockIdx.x >> 1, pair_step = gridDim.x >> 1;
    const int chunks = P.cin >> 6;
    constexpr uint32_t kTmemCols = 2 * kN;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < 2; ++s) {
            mbar_init(hfull(s), 1);
            mbar_init(hempty(s), 1);
            mbar_init(phfull(s), 1);
            mbar_init(tfull(s), 1);
            mbar_init(tempty(s), 32);
        }
        for (int s = 0; s < P.bstages; ++s) {
            mbar_init(bfull(s), 1);
            mbar_init(bempty(s), 1);
            mbar_init(pbfull(s), 1);
        }
        fence_mbar_init();
    }
    {
        float* ss = reinterpret_cast<float*>(sm + L.ss_off);
        float* cst = reinterpret_cast<float*>(sm + L.cstat_off);
        for (int c = threadIdx.x; c < 512; c += kW2Threads) {
            ss[c] = (P.scale && c < P.cout) ? P.scale[c] : 1.f;
            ss[512 + c] = (P.shift && c < P.cout) ? P.shift[c] : 0.f;
        }
        for (int c = threadIdx.x; c < 4 * 1024; c += kW2Threads) cst[c] = 0.f;
    }
    __syncthreads();
    cluster_sync_all();                 // both CTAs' barriers exist before any remote arrive / multicast commit
    if (warp == 1) {
        tmem_alloc2(smem_u32(const_cast<uint32_t*>(tmem_slot)), kTmemCols);
        tmem_relinquish2();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
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
        // ================================================================= TMA producer (one thread, both CTAs)
        if (lane == 0) {
            int hs = 0, bs = 0;
            uint32_t hph = 0, bph = 0;
            for (int item = pair0; item < total_items; item += pair_step) {
                int nt, tw, th, tn;
                decode(item, nt, tw, th, tn);
                for (int c = 0; c < chunks; ++c) {
                    if (!mbar_wait(hempty(hs), hph ^ 1)) {
                        atomicExch(P.err, 61);
                        goto done;
                    }
                    mbar_expect_tx(hfull(hs), kW2HaloBytes);
                    tma_load_4d(base + L.halo_off + hs * kW2HaloStage, &tmA, hfull(hs), c * 64, tw * 16 + 8 * (int)rank - 1,
                                th * 16 - 1, tn);
                    if (++hs == kW2HaloStages) {
                        hs = 0;
                        hph ^= 1;
                    }
                    for (int tap = 0; tap < 9; ++tap) {
                        if (!mbar_wait(bempty(bs), bph ^ 1)) {
                            atomicExch(P.err, 62);
                            goto done;
                        }
                        mbar_expect_tx(bfull(bs), kBStage);
                        tma_load_2d(base + L.b_off + bs * kBStage, &tmB, bfull(bs), tap * P.cin + c * 64,
                                    nt * kN + (int)rank * (kN / 2));
                        if (++bs == P.bstages) {
                            bs = 0;
                            bph ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            // ============================================================= MMA issuer (leader CTA, one thread)
            int hs = 0, bs = 0, acc = 0;
            uint32_t hph = 0, bph = 0, aph = 0;
            const uint32_t idesc = umma_idesc_bf16(256, kN, 0, 0);
            const uint64_t a_desc0 = umma_desc(base + L.halo_off, 16, 10 * 128, 2u);
            const uint64_t b_desc0 = umma_desc(base + L.b_off, 16, 8 * 128, 2u);
            for (int item = pair0; item < total_items; item += pair_step) {
                if (!mbar_wait(tempty(acc), aph ^ 1)) {
                    atomicExch(P.err, 63);
                    goto done;
                }
                tc_fence_after();
                const uint32_t d0 = tmem_base + acc * kN;
                for (int c = 0; c < chunks; ++c) {
                    if (!mbar_wait(hfull(hs), hph) || !mbar_wait(phfull(hs), hph)) {
                        atomicExch(P.err, 64);
                        goto done;
                    }
                    tc_fence_after();
                    const uint64_t a_base = a_desc0 + (uint64_t)((hs * kW2HaloStage) >> 4);
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        if (!mbar_wait(bfull(bs), bph) || !mbar_wait(pbfull(bs), bph)) {
                            atomicExch(P.err, 65);
                            goto done;
                        }
                        tc_fence_after();
                        const uint64_t bd = b_desc0 + (uint64_t)((bs * kBStage) >> 4);
                        const uint64_t ad = a_base + (uint32_t)((((tap / 3) * 10 + tap % 3) * 128) >> 4);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma2_bf16(d0, ad + (uint32_t)((kk * 32) >> 4), bd + (uint32_t)((kk * 32) >> 4), idesc,
                                       (c > 0 || tap > 0 || kk > 0) ? 1u : 0u);
                        if (tap % 3 == 2) {             // frees the row's three weight stages in BOTH CTAs (bstages % 3 == 0)
                            umma2_commit(bempty(bs - 2));
                            umma2_commit(bempty(bs - 1));
                            umma2_commit(bempty(bs));
                        }
                        if (++bs == P.bstages) {
                            bs = 0;
                            bph ^= 1;
                        }
                    }
                    umma2_commit(hempty(hs));
                    if (++hs == kW2HaloStages) {
                        hs = 0;
                        hph ^= 1;
                    }
                }
                umma2_commit(tfull(acc));               // accumulators of both CTAs complete
                acc ^= 1;
                if (acc == 0) aph ^= 1;
            }
        } else if (lane == 0) {
            // ============================================================= forwarder (peer CTA): "my operands have landed"
            int hs = 0, bs = 0;
            uint32_t hph = 0, bph = 0;
            for (int item = pair0; item < total_items; item += pair_step) {
                for (int c = 0; c < chunks; ++c) {
                    if (!mbar_wait(hfull(hs), hph)) {
                        atomicExch(P.err, 66);
                        goto done;
                    }
                    mbar_arrive_remote(mapa_shared(phfull(hs), 0));
                    if (++hs == kW2HaloStages) {
                        hs = 0;
                        hph ^= 1;
                    }
                    for (int tap = 0; tap < 9; ++tap) {
                        if (!mbar_wait(bfull(bs), bph)) {
                            atomicExch(P.err, 67);
                            goto done;
                        }
                        mbar_arrive_remote(mapa_shared(pbfull(bs), 0));
                        if (++bs == P.bstages) {
                            bs = 0;
                            bph ^= 1;
                        }
                    }
                }
            }
        }
    } else {
        // ================================================================= epilogue (16 warps, both CTAs)
        const int e = warp - 2;
        const int q = warp & 3;               // TMEM lane quadrant
        const int g = e >> 2;                 // channel group: kCg channels of the N tile
        const int row = q * 32 + lane;
        const int wl = row & 7, hl = row >> 3;
        const float* ss = reinterpret_cast<const float*>(sm + L.ss_off);
        float* cst = reinterpret_cast<float*>(sm + L.cstat_off) + q * 1024;
        int acc = 0;
        uint32_t aph = 0;
        for (int item = pair0; item < total_items; item += pair_step) {
            int nt, tw, th, tn;
            decode(item, nt, tw, th, tn);
            const int ph = th * 16 + hl, pw = tw * 16 + 8 * (int)rank + wl;
            const bool valid = ph < P.H && pw < P.W;
            const int cbase = nt * kN + g * kCg;
            const size_t off = (((size_t)tn * P.H + ph) * P.W + pw) * P.cout + cbase;
            if (!mbar_wait_warp(tfull(acc), aph, lane)) {
                atomicExch(P.err, 68);
                goto done;
            }
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * kN + g * kCg;
#pragma unroll
            for (int half = 0; half < kCg / 32; ++half) {
                uint32_t r[32];
                uint4 rv[4];
                if (P.residual && valid) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        rv[j] = __ldg(reinterpret_cast<const uint4*>(P.residual + off + half * 32) + j);
                }
                tmem_ld32(taddr + half * 32, r);
                tmem_ld_wait();
                if (half == kCg / 32 - 1) {  // last TMEM read of this accumulator: tell the leader's MMA thread
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (rank == 0) mbar_arrive(tempty(acc));
                        else mbar_arrive_remote(mapa_shared(tempty(acc), 0));
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
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
                    if (P.residual && valid) {
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
                    if (valid) *reinterpret_cast<uint4*>(P.out + off + half * 32 + j * 8) = o;
                    if (P.stats) {
                        const uint32_t w4[4] = {o.x, o.y, o.z, o.w};
                        float s1[16];
#pragma unroll
                        for (int m = 0; m < 4; ++m) {
                            const float lo = valid ? bf16_lo(w4[m]) : 0.f, hi = valid ? bf16_hi(w4[m]) : 0.f;
                            s1[2 * m] = lo; s1[2 * m + 1] = hi;
                            s1[8 + 2 * m] = lo * lo; s1[8 + 2 * m + 1] = hi * hi;
                        }
                        const float t = warp_reduce16(s1, lane);  // lane l < 8: sum of channel l; 8 <= l < 16: sum of squares
                        if (lane < 16) cst[2 * (c + (lane & 7)) + (lane >> 3)] += t;
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
            for (int j = threadIdx.x - 64; j < 2 * P.cout; j += 512)
                dst[j] = (call[j] + call[1024 + j]) + (call[2048 + j] + call[3072 + j]);
        }
    }
done:
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                 // the peer may still arrive on / multicast into this CTA's barriers until here
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, kTmemCols);
    }
}

}  // namespace ub
