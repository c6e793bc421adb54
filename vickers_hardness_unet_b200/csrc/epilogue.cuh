// Shared epilogue of the tcgen05 conv kernels (igemm.cuh, hconv.cuh): TMEM accumulator -> scale/shift (folded BN)
// (+ residual) (+ ReLU) -> bf16 -> swizzled smem staging -> TMA store (clips partial tiles), or raw output + per-channel
// batch-statistics partials.  Executed by the 4 epilogue warps (128 threads: thread `et`, warp quadrant q, lane).
// Named barriers 1..3 are reserved for these 128 threads.
#pragma once
#include "ptx.cuh"

namespace ub {

struct EpiParams {
    int ntile, out_cblk, relu, cout;
    int bw, bh, bn, Wo, Ho, Nimg;            // tile shape (bw*bh*bn == 128 rows) and logical output extent
    const __nv_bfloat16* residual;           // NHWC-strided bf16 or nullptr
    long long res_sw, res_sh, res_sn;
    float* stats;                            // [gridDim.x][cout][2] or nullptr
};
struct EpiSmem {
    uint8_t* staging;                        // 2 x staging_bytes
    uint32_t staging_bytes;
    const float* ss;                         // [512] scale, [512] shift
    float* part;                             // [8][64][2]
    float* cst;                              // [512][2] per-CTA running statistics
};

// One finished accumulator: 128 rows x ntile fp32 columns at tmem_acc (column base of this accumulator buffer).
// Arrives on tempty_bar (count 4: one per epilogue warp) right after the last TMEM read.
__device__ __forceinline__ void epilogue_tile(const EpiParams& E, const EpiSmem& S, const CUtensorMap* tmD,
                                              uint32_t tmem_acc, uint32_t tempty_bar, int tw, int th, int tn, int nt,
                                              int q, int lane, int et, uint32_t& blk_counter,
                                              long long* tm = nullptr) {
    long long tc = tm ? clock64() : 0;
#define UB_EPI_TICK(k)                          \
    if (tm) {                                   \
        const long long now_ = clock64();       \
        tm[k] += now_ - tc;                     \
        tc = now_;                              \
    }
    const int row = q * 32 + lane;          // tile row == TMEM lane == pixel index inside the tile
    const float* ss = S.ss;
    float* part = S.part;
    float* cst = S.cst;
    const int cblk = E.out_cblk;
    const int nblk = E.ntile / cblk;
    const uint32_t row_bytes = cblk * 2;
    const uint32_t swz_mask = (cblk >= 64) ? 7u : (cblk == 32 ? 3u : 1u);
    const int pw = tw * E.bw + row % E.bw;
    const int ph = th * E.bh + (row / E.bw) % E.bh;
    const int pn = tn * E.bn + row / (E.bw * E.bh);
    const bool valid = (pw < E.Wo) && (ph < E.Ho) && (pn < E.Nimg);
        for (int cb = 0; cb < nblk; ++cb, ++blk_counter) {
            const int cbase = nt * E.ntile + cb * cblk;  // first output channel of this block
            uint8_t* sbuf = S.staging + (blk_counter & 1) * S.staging_bytes;
            if (et == 0) tma_wait_read<1>();  // the store issued two blocks ago has finished reading this buffer
            UB_EPI_TICK(0)
            named_bar_sync(1, 128);
            UB_EPI_TICK(1)
            for (int h0 = 0; h0 < cblk; h0 += 32) {
                const int ncol = (cblk - h0) < 32 ? (cblk - h0) : 32;  // 16 or 32
                uint32_t r[32];
                const uint32_t taddr = tmem_acc + (uint32_t(q * 32) << 16) + cb * cblk + h0;
                if (ncol == 32) {
                    tmem_ld32(taddr, r);
                } else {
                    uint32_t r16[16];
                    tmem_ld16(taddr, r16);
#pragma unroll
                    for (int i = 0; i < 16; ++i) r[i] = r16[i];
#pragma unroll
                    for (int i = 16; i < 32; ++i) r[i] = 0;
                }
                tmem_ld_wait();
                UB_EPI_TICK(2)
                if (cb == nblk - 1 && h0 + 32 >= cblk) {
                    // last TMEM read of this accumulator: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar);
                }
                const __nv_bfloat16* resp = nullptr;
                if (E.residual && valid)
                    resp = E.residual + pn * E.res_sn + ph * E.res_sh + pw * E.res_sw + cbase + h0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {  // 4 x (8 channels = 16 B)
                    if (j * 8 >= ncol) break;
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int c = cbase + h0 + j * 8 + i;
                        v[i] = __uint_as_float(r[j * 8 + i]) * ss[c] + ss[512 + c];
                    }
                    if (resp) {
                        const uint4 rv = __ldg(reinterpret_cast<const uint4*>(resp + j * 8));
                        v[0] += bf16_lo(rv.x); v[1] += bf16_hi(rv.x);
                        v[2] += bf16_lo(rv.y); v[3] += bf16_hi(rv.y);
                        v[4] += bf16_lo(rv.z); v[5] += bf16_hi(rv.z);
                        v[6] += bf16_lo(rv.w); v[7] += bf16_hi(rv.w);
                    }
                    if (E.relu) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
                    }
                    if (!valid) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = 0.f;
                    }
                    uint4 o;
                    o.x = pack_bf16(v[0], v[1]);
                    o.y = pack_bf16(v[2], v[3]);
                    o.z = pack_bf16(v[4], v[5]);
                    o.w = pack_bf16(v[6], v[7]);
                    uint32_t off = row * row_bytes + (h0 * 2 + j * 16);
                    off ^= ((off >> 7) & swz_mask) << 4;
                    *reinterpret_cast<uint4*>(sbuf + off) = o;
                }
            }
            UB_EPI_TICK(3)
            fence_async_smem();
            UB_EPI_TICK(4)
            named_bar_sync(2, 128);
            UB_EPI_TICK(5)
            if (et == 0) {
                tma_store_4d(tmD, smem_u32(sbuf), cbase, tw * E.bw, th * E.bh, tn * E.bn);
                tma_commit();
            }
            UB_EPI_TICK(6)
            if (E.stats) {
                // per-channel sum / sum of squares over the tile's 128 pixels, from the bf16 values just staged.  A thread
                // reads 16-byte chunks (8 channels) of cblk/8 rows; row groups living in the same warp are combined with
                // shuffles, the four warps through `part`.  (One bf16 scalar per load and 64 loads per thread made this
                // tail 2-3x the cost of the rest of the epilogue: the stride-2 / 1x1 convs of the training forward took
                // 2.5x their inference time per image.)  Fixed summation order: deterministic.
                const int nch = cblk >> 3;                 // 16-byte chunks per staged row: 2, 4 or 8
                const int cc = et % nch, rg = et / nch, ngrp = 128 / nch;
                float s1[8], s2[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) s1[k] = s2[k] = 0.f;
                for (int rr = rg; rr < 128; rr += ngrp) {
                    uint32_t off = rr * row_bytes + cc * 16;
                    off ^= ((off >> 7) & swz_mask) << 4;
                    const uint4 v = *reinterpret_cast<const uint4*>(sbuf + off);
                    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const float lo = bf16_lo(w4[m]), hi = bf16_hi(w4[m]);
                        s1[2 * m] += lo; s1[2 * m + 1] += hi;
                        s2[2 * m] += lo * lo; s2[2 * m + 1] += hi * hi;
                    }
                }
                for (int o = nch; o < 32; o <<= 1) {       // lanes with equal lane % nch hold the same channel chunk
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], o);
                        s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], o);
                    }
                }
                if (lane < nch) {                          // part[q][channel][2], channel = lane * 8 + k
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        part[(q * 64 + lane * 8 + k) * 2 + 0] = s1[k];
                        part[(q * 64 + lane * 8 + k) * 2 + 1] = s2[k];
                    }
                }
                named_bar_sync(3, 128);
                if (et < cblk) {
                    float t1 = 0.f, t2 = 0.f;
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        t1 += part[(qq * 64 + et) * 2 + 0];
                        t2 += part[(qq * 64 + et) * 2 + 1];
                    }
                    // channel cbase+et is always owned by this thread (et == channel % cblk): no race
                    cst[2 * (cbase + et)] += t1;
                    cst[2 * (cbase + et) + 1] += t2;
                }
                // `part` is rewritten only after the next block's named_bar_sync(1), which orders it after these reads
            }
        }
#undef UB_EPI_TICK
}

// After the last tile: publish this CTA's running statistics (one row of the partials table per CTA).
__device__ __forceinline__ void epilogue_finish(const EpiParams& E, const EpiSmem& S, int q, int lane, int et) {
    (void)q; (void)lane;
    if (E.stats) {
        named_bar_sync(3, 128);
        float* dst = E.stats + static_cast<size_t>(blockIdx.x) * E.cout * 2;
        for (int j = et; j < 2 * E.cout; j += 128) dst[j] = S.cst[j];
    }
    if (et == 0) tma_wait_all<0>();  // all output stores complete before the CTA exits
}

}  // namespace ub
