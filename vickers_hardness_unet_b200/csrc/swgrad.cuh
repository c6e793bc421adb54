// Weight gradient of the 7x7 / stride-2 stem on tcgen05 (sm_100a):
//     dW[co, ci, r, s] = sum over (n, ho, wo) of  dZ[n, ho, wo, co] * x[n, 2*ho + r - 3, 2*wo + s - 3, ci]
// The tap-table kernel (wgrad.cuh, stem mode) ran one filter row r per work item and so streamed the 134 MB dZ seven
// times (0.30 ms, HBM-bound).  Here the anchors are INPUT rows: a pipeline step loads, for TQ pairs of input rows and WT
// output columns,
//     dZ box [64 co, WT, TQ+3 rows] (rows q0-1 .. q0+TQ+1)      -> A operand, MN-major (a pixel = 128 bytes, SW128)
//     X  box [32, WT, 2*TQ rows] of the window view of the packed image xp[N][H][W+8][4]: "pixel" wo of a row is the
//            64-byte window of 8 input pixels x 4 channels starting at padded column 2*wo (16-byte pitch, overlapping:
//            the TMA materialises the windows)                   -> B operand, MN-major, N = 32
// and every filter row comes from pairing an input row with the right dZ rows: input row 2q + par meets dZ row ho with
// r = 2q + par - 2*ho + 3, i.e. ho in {q-1, q, q+1, q+2}.  The A operand carries TWO dZ rows as its two 64-channel
// M-blocks (LBO = one dZ box row), so per 16 output columns an input row costs two MMAs (views v = 0: rows q-1, q;
// v = 1: rows q+1, q+2) into the accumulator (par, v); accumulator rows [64 j, 64 j + 64) hold filter row
// r = (par ? 6 : 5) - 2 * (2 v + j)  (r = -1: unused).  Four accumulators x 32 columns stay in TMEM for the whole
// kernel; at the end each CTA adds its partial into the OIHW fp32 gradient (column = (window pixel, channel) ->
// kernel column s = pixel - 1, pixel 0 and channel 3 are padding).
#pragma once
#include "ptx.cuh"

namespace ub {

constexpr int kSwThreads = 192;   // warp 0: TMA producer, warp 1: MMA issuer (owns TMEM), warps 2-5: final epilogue
constexpr int kSwTQ = 4;          // input-row pairs per step

struct SwgradParams {
    int Ho, Wo, N;                // extent of dZ (input is 2*Ho x 2*Wo)
    int wt;                       // output columns per step (16 / 32 / 64)
    int tiles_w, tiles_q;
    int stages;
    float* grad;                  // OIHW fp32 [64][3][7][7], accumulated with atomics (zeroed by the caller)
    int* err;
};

struct SwgradSmem {
    uint32_t z_bytes, x_bytes, stage_bytes, bar_off, total;
};
__host__ __device__ inline SwgradSmem swgrad_smem(int wt, int stages) {
    SwgradSmem s;
    s.z_bytes = (uint32_t)(kSwTQ + 3) * wt * 128;        // multiples of 1 KB for wt >= 8
    s.x_bytes = (uint32_t)(2 * kSwTQ) * wt * 64;
    s.stage_bytes = s.z_bytes + s.x_bytes;
    s.bar_off = s.stage_bytes * stages;
    s.total = s.bar_off + (2 * stages + 1) * 8 + 16;
    return s;
}

__global__ void __launch_bounds__(kSwThreads, 1)
swgrad_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmX,
              const __grid_constant__ SwgradParams P) {
    extern __shared__ uint8_t smem_raw[];
    griddep_launch();
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw_addr);
    const SwgradSmem L = swgrad_smem(P.wt, P.stages);
    const uint32_t bar0 = base + L.bar_off;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (P.stages + s); };
    const uint32_t done_bar = bar0 + 8u * (2 * P.stages);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + L.bar_off + (2 * P.stages + 1) * 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = P.tiles_w * P.tiles_q * P.N;
    constexpr uint32_t kCols = 128;   // 4 accumulators x 32 columns

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
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), kCols);
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
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int tw = t % P.tiles_w, tq = (t / P.tiles_w) % P.tiles_q, tn = t / (P.tiles_w * P.tiles_q);
                if (!mbar_wait(empty_bar(stage), phase ^ 1)) {
                    atomicExch(P.err, 61);
                    goto role_done;
                }
                const uint32_t zb = base + stage * L.stage_bytes, xb = zb + L.z_bytes;
                mbar_expect_tx(full_bar(stage), L.stage_bytes);
                tma_load_4d(zb, &tmZ, full_bar(stage), 0, tw * P.wt, tq * kSwTQ - 1, tn);
                tma_load_4d(xb, &tmX, full_bar(stage), 0, tw * P.wt, 2 * tq * kSwTQ, tn);
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
            bool first_tile = true;
            const uint32_t idesc = umma_idesc_bf16(128, 32, 1, 1);   // both operands MN-major
            const uint32_t zrow = (uint32_t)P.wt * 128, xrow = (uint32_t)P.wt * 64;
            const int ksteps = P.wt / 16;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                if (!mbar_wait(full_bar(stage), phase)) {
                    atomicExch(P.err, 63);
                    goto role_done;
                }
                tc_fence_after();
                const uint32_t zb = base + stage * L.stage_bytes, xb = zb + L.z_bytes;
                // A: K groups of 8 output columns (1 KB), the two M-blocks are two consecutive dZ rows
                const uint64_t a_base = umma_desc(zb, zrow, 1024, 2u);
                // B: K groups of 8 windows (512 B), one 32-element N block
                const uint64_t b_base = umma_desc(xb, 0, 512, 4u);
#pragma unroll
                for (int ql = 0; ql < kSwTQ; ++ql) {
#pragma unroll
                    for (int par = 0; par < 2; ++par) {
                        for (int k = 0; k < ksteps; ++k) {
                            const uint64_t bd = b_base + (((uint32_t)(2 * ql + par) * xrow + (uint32_t)k * 1024) >> 4);
                            // the first MMA into each accumulator (par, v) of the kernel overwrites, all others accumulate
                            const uint32_t accum = (first_tile && ql == 0 && k == 0) ? 0u : 1u;
#pragma unroll
                            for (int v = 0; v < 2; ++v)
                                umma_bf16(tmem_base + (par * 2 + v) * 32,
                                          a_base + (((uint32_t)(ql + 2 * v) * zrow + (uint32_t)k * 2048) >> 4), bd, idesc, accum);
                        }
                    }
                }
                first_tile = false;
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
        const int j = m >> 6, co = m & 63;
        if (!mbar_wait_warp(done_bar, 0, lane)) {
            atomicExch(P.err, 64);
            goto role_done;
        }
        tc_fence_after();
        if ((int)blockIdx.x < total_tiles) {
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int par = a >> 1, v = a & 1;
                const int r = (par ? 6 : 5) - 2 * (2 * v + j);
                uint32_t c[32];
                tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + a * 32, c);
                tmem_ld_wait();
                if (r < 0) continue;
                float* g = P.grad + (size_t)co * 147 + r * 7;
#pragma unroll
                for (int col = 4; col < 32; ++col) {          // window pixel 0 is padding
                    const int px = col >> 2, ch = col & 3;
                    if (ch < 3) atomicAdd(g + ch * 49 + (px - 1), __uint_as_float(c[col]));
                }
            }
        }
    }
role_done:
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kCols);
    }
}

}  // namespace ub
