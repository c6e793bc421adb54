// Bandwidth-bound kernels of the training step: BatchNorm (batch statistics) forward/backward, max-pool backward,
// seg-head backward, fused BCE+Dice loss, fused AdamW, dgrad weight packs.  NHWC bf16 activations, fp32 statistics.
#pragma once
#include "ptx.cuh"

namespace ub {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
    f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 o;
    o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]); o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
    return o;
}

// ------------------------------------------------------------------------------------------------ BN forward (train)
// Second stage of the batch statistics: the conv epilogues left per-CTA partial (sum, sumsq) rows; up to 4 segments
// (the 4 parity launches of a decoder conv1).  Produces scale/shift for the apply pass, saves mean/invstd for the
// backward and updates the running statistics like nn.BatchNorm2d (momentum 0.1, unbiased running variance,
// num_batches_tracked += 1)  [torch semantics used at /root/reference/train.py:413,436].
struct StatSegs {
    const float* ptr[4];
    int rows[4];
    int n;
};
// one WARP per channel: lanes stride over the partial rows (blockDim = 256 -> 8 channels per block)
__global__ void __launch_bounds__(256)
bn_finalize_kernel(StatSegs segs, int C, double count, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* __restrict__ running_mean,
                   float* __restrict__ running_var, long long* __restrict__ counter, float momentum,
                   float eps, float* __restrict__ scale, float* __restrict__ shift,
                   float* __restrict__ mean_out, float* __restrict__ invstd_out) {
    griddep_launch();
    griddep_wait();
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0 && counter) *counter += 1;
    if (c >= C) return;
    double s1 = 0, s2 = 0;
    for (int s = 0; s < segs.n; ++s)
        for (int r = lane; r < segs.rows[s]; r += 32) {
            const float2 v = *reinterpret_cast<const float2*>(segs.ptr[s] + ((size_t)r * C + c) * 2);
            s1 += v.x;
            s2 += v.y;
        }
    s1 = warp_sum_d(s1);
    s2 = warp_sum_d(s2);
    if (lane) return;
    const double mean = s1 / count;
    double var = s2 / count - mean * mean;
    if (var < 0) var = 0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * invstd;
    scale[c] = sc;
    shift[c] = beta[c] - (float)mean * sc;
    mean_out[c] = (float)mean;
    invstd_out[c] = invstd;
    if (running_mean) {
        const double unbiased = count > 1 ? var * count / (count - 1) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

// a = [relu]( z*scale + shift [+ residual] ), 8 channels per thread.  blockDim * gridDim is a multiple of C/8, so a
// thread's channel group never changes inside the grid-stride loop: its 16 constants are loaded once.
__global__ void __launch_bounds__(256)
bn_apply_kernel(const __nv_bfloat16* __restrict__ z, const float* __restrict__ scale,
                const float* __restrict__ shift, const __nv_bfloat16* __restrict__ residual, int relu,
                __nv_bfloat16* __restrict__ out, long long npix, int C) {
    griddep_launch();
    griddep_wait();
    const int C8 = C / 8;
    const long long total = npix * C8;
    const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int c0 = int(i0 % C8) * 8;
    float sc[8], sh[8];
    {
        const float4 a = __ldg(reinterpret_cast<const float4*>(scale + c0)), b = __ldg(reinterpret_cast<const float4*>(scale + c0) + 1);
        const float4 c = __ldg(reinterpret_cast<const float4*>(shift + c0)), d = __ldg(reinterpret_cast<const float4*>(shift + c0) + 1);
        sc[0] = a.x; sc[1] = a.y; sc[2] = a.z; sc[3] = a.w; sc[4] = b.x; sc[5] = b.y; sc[6] = b.z; sc[7] = b.w;
        sh[0] = c.x; sh[1] = c.y; sh[2] = c.z; sh[3] = c.w; sh[4] = d.x; sh[5] = d.y; sh[6] = d.z; sh[7] = d.w;
    }
    const uint4* z4 = reinterpret_cast<const uint4*>(z);
    const uint4* r4 = reinterpret_cast<const uint4*>(residual);
    auto one = [&](const uint4& zv, const uint4& rv) {
        float v[8], r[8];
        unpack8(zv, v);
        if (residual) unpack8(rv, r);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float y = v[k] * sc[k] + sh[k];
            if (residual) y += r[k];
            if (relu) y = fmaxf(y, 0.f);
            v[k] = y;
        }
        return pack8(v);
    };
    const long long step = (long long)gridDim.x * blockDim.x;
    long long i = i0;
    for (; i + step < total; i += 2 * step) {   // two independent elements in flight per thread
        const long long j = i + step;
        const uint4 za = __ldg(z4 + i), zb = __ldg(z4 + j);
        uint4 ra = za, rb = zb;
        if (residual) {
            ra = __ldg(r4 + i);
            rb = __ldg(r4 + j);
        }
        reinterpret_cast<uint4*>(out)[i] = one(za, ra);
        reinterpret_cast<uint4*>(out)[j] = one(zb, rb);
    }
    if (i < total) {
        const uint4 za = __ldg(z4 + i);
        uint4 ra = za;
        if (residual) ra = __ldg(r4 + i);
        reinterpret_cast<uint4*>(out)[i] = one(za, ra);
    }
}

// ------------------------------------------------------------------------------------------------ BN backward
// g = dA * [a > 0] ; partial sums per block of (sum g, sum g*zhat), zhat = (z - mean) * invstd.
// The ReLU mask comes from the stored activation `a_mask` (units with a residual input) or, when a_mask == nullptr and
// scale != nullptr, is recomputed from z (a = relu(z*scale + shift) > 0  <=>  z*scale + shift > 0): one tensor less to read.
// blockDim = 256; thread t owns channel group t % C8 for the pixels t / C8 + k * (256 / C8).  Two pixels per loop
// iteration keep 4-6 independent 16-byte loads in flight per thread; the host sizes the grid so that a block runs
// >= 8 iterations (few partial rows for the small, wide tensors of layer3/4: the row traffic was as large as the tensor).
__device__ __forceinline__ void bn_bwd_mask8(float (&g)[8], const float (&zz)[8], const uint4* a_mask, long long idx,
                                             const float (&sc)[8], const float (&sh)[8]) {
    if (a_mask) {
        float m[8];
        unpack8(__ldg(a_mask + idx), m);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = m[k] > 0.f ? g[k] : 0.f;
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = (zz[k] * sc[k] + sh[k]) > 0.f ? g[k] : 0.f;
    }
}
// Block tail shared by every kernel that accumulates the BatchNorm-backward sums (sum g, sum g*zhat): thread t owns the
// channel group t % (C/8) for its pixels; the block's per-channel totals go to partial[blockIdx.x][C][2].
// red: 256 * 17 floats of shared memory (row pitch 17 words: conflict-free column walks).
__device__ __forceinline__ void bn_partials_store(const float (&s1)[8], const float (&s2)[8], int C,
                                                  float* __restrict__ partial, float* red) {
    const int C8 = C / 8, ppb = 256 / C8;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        red[threadIdx.x * 17 + k] = s1[k];
        red[threadIdx.x * 17 + 8 + k] = s2[k];
    }
    __syncthreads();
    // thread t < C*2 sums its (channel, which) over the ppb pixel lanes
    for (int j = threadIdx.x; j < C * 2; j += 256) {
        const int c = j >> 1, which = j & 1;
        const int g8 = c / 8, k = c % 8;
        float acc = 0.f;
        for (int l = 0; l < ppb; ++l) acc += red[(l * C8 + g8) * 17 + which * 8 + k];
        partial[((size_t)blockIdx.x * C + c) * 2 + which] = acc;
    }
}
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dA, const __nv_bfloat16* __restrict__ a_mask,
                     const float* __restrict__ scale, const float* __restrict__ shift,
                     const __nv_bfloat16* __restrict__ z, const float* __restrict__ mean,
                     const float* __restrict__ invstd, float* __restrict__ partial, long long npix, int C) {
    griddep_launch();
    griddep_wait();
    __shared__ float red[256 * 17];
    const int C8 = C / 8;
    const int cg = threadIdx.x % C8;
    const int lane_p = threadIdx.x / C8;
    const int ppb = 256 / C8;  // pixels per block iteration
    float s1[8], s2[8], mu[8], is[8], sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        s1[k] = s2[k] = 0.f;
        mu[k] = mean[cg * 8 + k];
        is[k] = invstd[cg * 8 + k];
        sc[k] = scale ? scale[cg * 8 + k] : 0.f;
        sh[k] = scale ? shift[cg * 8 + k] : 1.f;   // no mask: 0*z + 1 > 0 always
    }
    const uint4* dA4 = reinterpret_cast<const uint4*>(dA);
    const uint4* z4 = reinterpret_cast<const uint4*>(z);
    const uint4* m4 = reinterpret_cast<const uint4*>(a_mask);
    if (lane_p < ppb) {
        const long long step = (long long)gridDim.x * ppb;
        long long p = (long long)blockIdx.x * ppb + lane_p;
        for (; p + step < npix; p += 2 * step) {
            const long long i0 = p * C8 + cg, i1 = (p + step) * C8 + cg;
            const uint4 ga = __ldg(dA4 + i0), gb = __ldg(dA4 + i1);
            const uint4 za = __ldg(z4 + i0), zb = __ldg(z4 + i1);
            float g[8], zz[8], h[8], zy[8];
            unpack8(ga, g);
            unpack8(za, zz);
            unpack8(gb, h);
            unpack8(zb, zy);
            bn_bwd_mask8(g, zz, m4, i0, sc, sh);
            bn_bwd_mask8(h, zy, m4, i1, sc, sh);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                s1[k] += g[k];
                s2[k] += g[k] * (zz[k] - mu[k]) * is[k];
                s1[k] += h[k];
                s2[k] += h[k] * (zy[k] - mu[k]) * is[k];
            }
        }
        if (p < npix) {
            const long long idx = p * C8 + cg;
            float g[8], zz[8];
            unpack8(__ldg(dA4 + idx), g);
            unpack8(__ldg(z4 + idx), zz);
            bn_bwd_mask8(g, zz, m4, idx, sc, sh);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                s1[k] += g[k];
                s2[k] += g[k] * (zz[k] - mu[k]) * is[k];
            }
        }
    }
    bn_partials_store(s1, s2, C, partial, red);
}

// sums the per-block partials; writes dgamma / dbeta and the affine apply coefficients
//   dz = k1 * (g - s1/n - zhat * s2/n),  zhat = (z - mean) * invstd,  k1 = gamma * invstd
//      = cA * g + cB * z + cC   with  cA = k1,  cB = -k1 * invstd * s2/n,  cC = -k1 * s1/n - cB * mean
// stored as coef[0..C) = cA, coef[C..2C) = cB, coef[2C..3C) = cC (16-byte loads in the apply pass).
// one WARP per channel (blockDim = 256 -> 8 channels per block)
__global__ void __launch_bounds__(256)
bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblocks, int C, double count,
                       const float* __restrict__ gamma, const float* __restrict__ mean,
                       const float* __restrict__ invstd, float* __restrict__ dgamma, float* __restrict__ dbeta,
                       float* __restrict__ coef) {
    griddep_launch();
    griddep_wait();
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= C) return;
    double s1 = 0, s2 = 0;
    for (int b = lane; b < nblocks; b += 32) {
        const float2 v = *reinterpret_cast<const float2*>(partial + ((size_t)b * C + c) * 2);
        s1 += v.x;
        s2 += v.y;
    }
    s1 = warp_sum_d(s1);
    s2 = warp_sum_d(s2);
    if (lane) return;
    dbeta[c] = (float)s1;
    dgamma[c] = (float)s2;
    const double k1 = (double)gamma[c] * (double)invstd[c];
    const double cB = -k1 * (double)invstd[c] * (s2 / count);
    coef[c] = (float)k1;
    coef[C + c] = (float)cB;
    coef[2 * C + c] = (float)(-k1 * (s1 / count) - cB * (double)mean[c]);
}

// dz = cA * g + cB * z + cC;  optionally also writes g (the masked gradient, for the residual identity path).
// Mask as in bn_bwd_reduce_kernel; per-thread channel constants are loop invariant (see bn_apply_kernel) and arrive as
// 16-byte loads (56 scalar loads per thread were the larger part of this kernel on the <= 16 MB tensors).
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dA, const __nv_bfloat16* __restrict__ a_mask,
                    const float* __restrict__ scale, const float* __restrict__ shift,
                    const __nv_bfloat16* __restrict__ z, const float* __restrict__ coef,
                    __nv_bfloat16* __restrict__ dz, __nv_bfloat16* __restrict__ g_out, long long npix, int C) {
    griddep_launch();
    griddep_wait();
    const int C8 = C / 8;
    const long long total = npix * C8;
    const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int c0 = int(i0 % C8) * 8;
    float cA[8], cB[8], cC[8], sc[8], sh[8];
    auto ld8 = [](const float* p, float (&o)[8]) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
    };
    ld8(coef + c0, cA);
    ld8(coef + C + c0, cB);
    ld8(coef + 2 * C + c0, cC);
    if (scale) {
        ld8(scale + c0, sc);
        ld8(shift + c0, sh);
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) sc[k] = 0.f, sh[k] = 1.f;
    }
    const uint4* dA4 = reinterpret_cast<const uint4*>(dA);
    const uint4* z4 = reinterpret_cast<const uint4*>(z);
    const uint4* m4 = reinterpret_cast<const uint4*>(a_mask);
    const long long step = (long long)gridDim.x * blockDim.x;
    long long i = i0;
    for (; i + step < total; i += 2 * step) {
        const long long j = i + step;
        const uint4 ga = __ldg(dA4 + i), gb = __ldg(dA4 + j);
        const uint4 za = __ldg(z4 + i), zb = __ldg(z4 + j);
        float g[8], zz[8], h[8], zy[8], o[8], q[8];
        unpack8(ga, g);
        unpack8(za, zz);
        unpack8(gb, h);
        unpack8(zb, zy);
        bn_bwd_mask8(g, zz, m4, i, sc, sh);
        bn_bwd_mask8(h, zy, m4, j, sc, sh);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            o[k] = cA[k] * g[k] + (cB[k] * zz[k] + cC[k]);
            q[k] = cA[k] * h[k] + (cB[k] * zy[k] + cC[k]);
        }
        reinterpret_cast<uint4*>(dz)[i] = pack8(o);
        reinterpret_cast<uint4*>(dz)[j] = pack8(q);
        if (g_out) {
            reinterpret_cast<uint4*>(g_out)[i] = pack8(g);
            reinterpret_cast<uint4*>(g_out)[j] = pack8(h);
        }
    }
    if (i < total) {
        float g[8], zz[8], o[8];
        unpack8(__ldg(dA4 + i), g);
        unpack8(__ldg(z4 + i), zz);
        bn_bwd_mask8(g, zz, m4, i, sc, sh);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = cA[k] * g[k] + (cB[k] * zz[k] + cC[k]);
        reinterpret_cast<uint4*>(dz)[i] = pack8(o);
        if (g_out) reinterpret_cast<uint4*>(g_out)[i] = pack8(g);
    }
}

// ------------------------------------------------------------------------------------------------ max-pool (train)
// Forward that also records, per output element, WHICH of the 9 window positions (r*3+s, first maximum in scan order —
// the element PyTorch's max_pool2d backward picks) won; the backward is then a pure gather.
__global__ void maxpool3x3s2_idx_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                        uint8_t* __restrict__ idx, int N, int H, int W, int C) {
    griddep_launch();
    griddep_wait();
    const int Ho = H / 2, Wo = W / 2, C8 = C / 8;
    const long long total = (long long)N * Ho * Wo * C8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        // 32-bit index arithmetic (the host keeps element counts below 2^32): 64-bit div/mod per element is ~100 instructions
        const unsigned u = (unsigned)i, t1 = u / (unsigned)C8, t2 = t1 / (unsigned)Wo;
        const int c8 = int(u - t1 * (unsigned)C8);
        const int wo = int(t1 - t2 * (unsigned)Wo);
        const int n = int(t2 / (unsigned)Ho);
        const int ho = int(t2 - (unsigned)n * (unsigned)Ho);
        float best[8];
        uint32_t bi[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            best[k] = -INFINITY;
            bi[k] = 0;
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int h = ho * 2 - 1 + r;
            if (h < 0 || h >= H) continue;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int w = wo * 2 - 1 + s;
                if (w < 0 || w >= W) continue;
                float v[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(in + (((long long)n * H + h) * W + w) * C) + c8), v);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (v[k] > best[k]) {
                        best[k] = v[k];
                        bi[k] = r * 3 + s;
                    }
            }
        }
        reinterpret_cast<uint4*>(out)[i] = pack8(best);
        uint2 o;
        o.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
        o.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
        reinterpret_cast<uint2*>(idx)[i] = o;
    }
}

// dF[n,h,w,c] = dSkip[n,h,w,c] + sum over the (<= 4) windows containing (h,w) whose recorded winner is (h,w) of dP[window]
__global__ void maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dP, const uint8_t* __restrict__ idx,
                                   const __nv_bfloat16* __restrict__ dSkip, __nv_bfloat16* __restrict__ dF, int N,
                                   int H, int W, int C) {
    griddep_launch();
    griddep_wait();
    const int Ho = H / 2, Wo = W / 2, C8 = C / 8;
    const long long total = (long long)N * H * W * C8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const unsigned u = (unsigned)i, t1 = u / (unsigned)C8, t2 = t1 / (unsigned)W;   // 32-bit: see maxpool3x3s2_idx_kernel
        const int c8 = int(u - t1 * (unsigned)C8);
        const int w = int(t1 - t2 * (unsigned)W);
        const int n = int(t2 / (unsigned)H);
        const int h = int(t2 - (unsigned)n * (unsigned)H);
        float acc[8];
        if (dSkip) unpack8(__ldg(reinterpret_cast<const uint4*>(dSkip) + i), acc);
        else {
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = 0.f;
        }
        // windows (ho,wo) with 2*ho-1 <= h <= 2*ho+1
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int ho = h / 2 + a;
            if (ho >= Ho || 2 * ho - 1 > h) continue;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int wo = w / 2 + b;
                if (wo >= Wo || 2 * wo - 1 > w) continue;
                const uint32_t pos = (h - (2 * ho - 1)) * 3 + (w - (2 * wo - 1));
                const long long o = (((long long)n * Ho + ho) * Wo + wo) * C8 + c8;
                const uint2 id = __ldg(reinterpret_cast<const uint2*>(idx) + o);
                float d[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(dP) + o), d);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (((id.x >> (8 * k)) & 0xFF) == pos) acc[k] += d[k];
                    if (((id.y >> (8 * k)) & 0xFF) == pos) acc[4 + k] += d[4 + k];
                }
            }
        }
        reinterpret_cast<uint4*>(dF)[i] = pack8(acc);
    }
}

// ------------------------------------------------------------------------------------------------ seg head backward
// dA[n,h,w,c] = sum_{r,s} dL[n, h+1-r, w+1-s] * w[c][r][s]   (bf16 out, 16 channels per pixel)
__global__ void __launch_bounds__(256)
head_bwd_data_kernel(const float* __restrict__ dL, const float* __restrict__ w,
                     __nv_bfloat16* __restrict__ dA, int N, int H, int W) {
    griddep_launch();
    griddep_wait();
    // thread = (pixel, 8-channel half): its 72 weights live in registers for the whole grid-stride loop (the first
    // version read 144 weights per pixel from shared memory: LDS-bound)
    const int half = threadIdx.x & 1;
    float wr[72];
#pragma unroll
    for (int k = 0; k < 72; ++k) wr[k] = __ldg(w + half * 72 + k);   // [c][r][s], c = half*8 + k/9
    const long long total = (long long)N * H * W;
    for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 1; i < total;
         i += ((long long)gridDim.x * blockDim.x) >> 1) {
        const unsigned ui = (unsigned)i, row = ui / (unsigned)W;   // N*H*W < 2^32 (checked by the host): 32-bit div/mod
        const int x = int(ui - row * (unsigned)W);
        const int y = int(row % (unsigned)H);
        const long long nb = (long long)(row / (unsigned)H) * H * W;
        float d[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int yy = y + 1 - r, xx = x + 1 - s;
                d[r * 3 + s] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(dL + nb + (long long)yy * W + xx) : 0.f;
            }
        float o[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 9; ++k) a += d[k] * wr[c * 9 + k];
            o[c] = a;
        }
        reinterpret_cast<uint4*>(dA)[i * 2 + half] = pack8(o);
    }
}

// per-block partials of dW[c][r][s] = sum_p A[p + (r-1, s-1)][c] * dL[p] and dbias = sum dL   -> partial[block][145]
// Indexed by the ACTIVATION pixel q = p + (r-1, s-1): a thread owns 8 channels of q (one 16-byte load of the 134 MB
// tensor, the only HBM stream), needs the nine dL values around q (fp32, L1 / L2 hits: neighbouring lanes read the same
// lines) and keeps all 72 (channel, tap) partial sums in registers.  Lanes 2k / 2k+1 hold the two channel halves of one
// pixel, so a warp reads 512 contiguous bytes of A.  The first version walked the 9 taps with 9 warps, each re-reading A
// (L1-bound, 188 us at batch 16 / 512^2 against an HBM floor of ~25 us).
__global__ void __launch_bounds__(256, 2)
head_bwd_weight_kernel(const __nv_bfloat16* __restrict__ A, const float* __restrict__ dL, float* __restrict__ partial,
                       int N, int H, int W) {
    griddep_launch();
    griddep_wait();
    const int half = threadIdx.x & 1, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc[72], bsum = 0.f;
#pragma unroll
    for (int k = 0; k < 72; ++k) acc[k] = 0.f;
    const unsigned total = (unsigned)N * H * W;   // < 2^31 (checked by the host)
    const unsigned stride = gridDim.x * 128u;
    const uint4* A4 = reinterpret_cast<const uint4*>(A);
    for (unsigned q = blockIdx.x * 128u + (threadIdx.x >> 1); q < total; q += stride) {
        const unsigned row = q / (unsigned)W;
        const int x = int(q - row * (unsigned)W), y = int(row % (unsigned)H);
        const uint4 av = __ldg(A4 + (size_t)q * 2 + half);
        float d[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s2 = 0; s2 < 3; ++s2) {
                const int yy = y + 1 - r, xx = x + 1 - s2;
                d[r * 3 + s2] = (yy >= 0 && yy < H && xx >= 0 && xx < W)
                                    ? __ldg(dL + (long long)q + (long long)(1 - r) * W + (1 - s2)) : 0.f;
            }
        float v[8];
        unpack8(av, v);
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
            for (int k = 0; k < 9; ++k) acc[c * 9 + k] += v[c] * d[k];
        if (half == 0) bsum += d[4];
    }
    __shared__ float red[8][2][72];
    __shared__ float redb[8];
#pragma unroll
    for (int k = 0; k < 72; ++k) {
        float t = acc[k];
#pragma unroll
        for (int o = 16; o >= 2; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);   // lanes of equal parity
        if (lane < 2) red[warp][lane][k] = t;
    }
    bsum = warp_sum(bsum);
    if (lane == 0) redb[warp] = bsum;
    __syncthreads();
    float* out = partial + (size_t)blockIdx.x * 145;
    if (threadIdx.x < 144) {
        const int hf = threadIdx.x / 72, k = threadIdx.x % 72;   // out index = (hf*8 + c)*9 + tap = threadIdx.x
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][hf][k];
        out[threadIdx.x] = t;
    } else if (threadIdx.x == 144) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += redb[w];
        out[144] = t;
    }
}
// out[j] = sum_b partial[b][j]; one WARP per column j (blockDim = 256 -> 8 columns per block)
__global__ void __launch_bounds__(256)
sum_rows_kernel(const float* __restrict__ partial, int nrows, int width, float* __restrict__ out) {
    griddep_launch();
    griddep_wait();
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (j >= width) return;
    double s = 0;
    for (int b = lane; b < nrows; b += 32) s += partial[(size_t)b * width + j];
    s = warp_sum_d(s);
    if (lane == 0) out[j] = (float)s;
}

// ------------------------------------------------------------------------------------------------ BCE + Dice loss
// nn.BCEWithLogitsLoss() (mean) + smp DiceLoss(mode="binary") over the whole batch (/root/reference/train.py:438,600-601)
// pass 1: per-block partials of [sum bce_i, sum p*t, sum p, sum t]
__global__ void __launch_bounds__(256)
loss_partial_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ partial,
                    long long n) {
    griddep_launch();
    griddep_wait();
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const float xv = __ldg(x + i), t = __ldg(y + i);
        const float e = __expf(-fabsf(xv));
        const float l1p = log1pf(e);
        s[0] += fmaxf(xv, 0.f) - xv * t + l1p;      // BCE with logits, stable form
        const float p = __expf(fminf(xv, 0.f) - l1p);  // exp(logsigmoid(x))
        s[1] += p * t;
        s[2] += p;
        s[3] += t;
    }
    __shared__ float red[8][4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float v = warp_sum(s[k]);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        float v = 0.f;
        for (int wv = 0; wv < 8; ++wv) v += red[wv][threadIdx.x];
        partial[blockIdx.x * 4 + threadIdx.x] = v;
    }
}
// result[0..2] = bce, dice, bce+dice ; result[3..5] = I, P, T (saved for backward); one warp
__global__ void loss_finalize_kernel(const float* __restrict__ partial, int nblocks, double n, float eps,
                                     float* __restrict__ result) {
    griddep_launch();
    griddep_wait();
    if (blockIdx.x || threadIdx.x >= 32) return;   // one warp: lanes stride over the partial rows (fixed order: deterministic)
    double s[4] = {0, 0, 0, 0};
    for (int b = threadIdx.x; b < nblocks; b += 32) {
        const float4 v = *reinterpret_cast<const float4*>(partial + b * 4);
        s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) s[k] = warp_sum_d(s[k]);
    if (threadIdx.x) return;
    const double card = s[2] + s[3];
    double dice = 1.0 - 2.0 * s[1] / (card > eps ? card : (double)eps);
    if (!(s[3] > 0)) dice = 0.0;
    result[0] = (float)(s[0] / n);
    result[1] = (float)dice;
    result[2] = (float)(s[0] / n + dice);
    result[3] = (float)s[1];
    result[4] = (float)s[2];
    result[5] = (float)s[3];
}
// dx_i = g_bce * (sigma_i - y_i)/n  +  g_dice * ( -(2 t_i C - 2 I) / C^2 ) * sigma_i (1 - sigma_i)   (0 if T == 0)
// g points to device scalars (upstream gradients), so no host sync is needed.
__global__ void loss_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                const float* __restrict__ result, const float* __restrict__ g_bce,
                                const float* __restrict__ g_dice, float gscale, float eps, float* __restrict__ dx,
                                long long n) {
    griddep_launch();
    griddep_wait();
    const float I = result[3], P = result[4], T = result[5];
    const float gb = (g_bce ? *g_bce : 0.f) * gscale / (float)n;
    float gd = (g_dice ? *g_dice : 0.f) * gscale;
    const float Cc = P + T;
    const bool clamped = !(Cc > eps);
    if (!(T > 0.f)) gd = 0.f;
    const float invC = clamped ? 1.f / eps : 1.f / Cc;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const float xv = __ldg(x + i), t = __ldg(y + i);
        const float sg = 1.f / (1.f + __expf(-xv));
        // d(2I/C)/dp = 2t/C - 2I/C^2 (C not clamped) or 2t/eps (clamped)
        const float dscore = clamped ? 2.f * t * invC : (2.f * t * invC - 2.f * I * invC * invC);
        dx[i] = gb * (sg - t) - gd * dscore * sg * (1.f - sg);
    }
}

// ------------------------------------------------------------------------------------------------ validation metrics
// dice_coef / iou_coef of /root/reference/train.py:230-281 on the device: pred = [p > thresh] (thresh 0.5 for
// probabilities, 0 for logits since sigmoid(x) > 0.5 <=> x > 0); per image inter = sum pred*t, sp = sum pred, st = sum t;
// dice_n = (2*inter + eps) / (sp + st + eps); iou_n = (inter + eps) / (sp + st - inter + eps); mean over the batch.
constexpr int kMetricBlocks = 64;   // blocks per image (fixed: deterministic partial layout [N][64][3])
__global__ void __launch_bounds__(256)
seg_metrics_partial_kernel(const float* __restrict__ p, const float* __restrict__ t, long long hw, float thresh,
                           float* __restrict__ partial) {
    griddep_launch();
    griddep_wait();
    const int n = blockIdx.y;
    const float* pn = p + (long long)n * hw;
    const float* tn = t + (long long)n * hw;
    float inter = 0.f, sp = 0.f, st = 0.f;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < hw; i += 256ll * kMetricBlocks) {
        const float pred = __ldg(pn + i) > thresh ? 1.f : 0.f;
        const float tv = __ldg(tn + i);
        inter += pred * tv;
        sp += pred;
        st += tv;
    }
    __shared__ float red[8][3];
    inter = warp_sum(inter); sp = warp_sum(sp); st = warp_sum(st);
    if ((threadIdx.x & 31) == 0) {
        red[threadIdx.x >> 5][0] = inter; red[threadIdx.x >> 5][1] = sp; red[threadIdx.x >> 5][2] = st;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float a = 0.f;
        for (int w = 0; w < 8; ++w) a += red[w][threadIdx.x];
        partial[((long long)n * kMetricBlocks + blockIdx.x) * 3 + threadIdx.x] = a;
    }
}
// one block: thread n < N finishes image n (fp64), then the batch mean -> out[0] = dice, out[1] = iou
__global__ void __launch_bounds__(256)
seg_metrics_finalize_kernel(const float* __restrict__ partial, int N, float eps, float* __restrict__ out) {
    griddep_launch();
    griddep_wait();
    double dsum = 0, isum = 0;
    for (int n = threadIdx.x; n < N; n += 256) {
        double inter = 0, sp = 0, st = 0;
        for (int b = 0; b < kMetricBlocks; ++b) {
            const float* q = partial + ((long long)n * kMetricBlocks + b) * 3;
            inter += q[0]; sp += q[1]; st += q[2];
        }
        dsum += (2 * inter + eps) / (sp + st + eps);
        isum += (inter + eps) / (sp + st - inter + eps);
    }
    __shared__ double rd[8], ri[8];
    dsum = warp_sum_d(dsum); isum = warp_sum_d(isum);
    if ((threadIdx.x & 31) == 0) { rd[threadIdx.x >> 5] = dsum; ri[threadIdx.x >> 5] = isum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int w = 0; w < 8; ++w) { a += rd[w]; b += ri[w]; }
        out[0] = (float)(a / N);
        out[1] = (float)(b / N);
    }
}

// ------------------------------------------------------------------------------------------------ fused AdamW
// torch.optim.AdamW semantics (/root/reference/train.py:606): decoupled decay on every tensor, bias-corrected moments.
// Optionally scales the gradient (1/world_size, 1/loss_scale) and zeroes it afterwards.
__global__ void adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             long long n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                             float gscale, int zero_grad) {
    griddep_launch();
    griddep_wait();
    const long long n4 = n / 4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4 + (n & 3);
         i += (long long)gridDim.x * blockDim.x) {
        if (i < n4) {
            float4 pp = reinterpret_cast<float4*>(p)[i], gg = reinterpret_cast<float4*>(g)[i];
            float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
            float* P4 = &pp.x; float* G4 = &gg.x; float* M4 = &mm.x; float* V4 = &vv.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float gr = G4[k] * gscale;
                P4[k] *= (1.f - lr * wd);
                M4[k] = b1 * M4[k] + (1.f - b1) * gr;
                V4[k] = b2 * V4[k] + (1.f - b2) * gr * gr;
                const float denom = sqrtf(V4[k]) / bc2_sqrt + eps;
                P4[k] -= (lr / bc1) * (M4[k] / denom);
            }
            reinterpret_cast<float4*>(p)[i] = pp;
            reinterpret_cast<float4*>(m)[i] = mm;
            reinterpret_cast<float4*>(v)[i] = vv;
            if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            const long long j = n4 * 4 + (i - n4);
            const float gr = g[j] * gscale;
            float pv = p[j] * (1.f - lr * wd);
            const float mv = b1 * m[j] + (1.f - b1) * gr;
            const float vv2 = b2 * v[j] + (1.f - b2) * gr * gr;
            pv -= (lr / bc1) * (mv / (sqrtf(vv2) / bc2_sqrt + eps));
            p[j] = pv; m[j] = mv; v[j] = vv2;
            if (zero_grad) g[j] = 0.f;
        }
    }
}

// ------------------------------------------------------------------------------------------------ packed gradient unpack
// The wgrad kernels accumulate into a packed fp32 buffer [co][tap][ci] (ci contiguous -> vector reductions); this turns
// it into the OIHW layout of the parameter / .grad tensors:  grad[co][ci][tap] = gpk[co][tap][ci].  One launch handles a
// table of convs (a backward stage); block b belongs to the conv whose block range contains it.
struct UnpackEntry {
    long long src_off, dst_off;   // element offsets into the packed buffer / the flat gradient array
    int cout, cin, ntaps, block_begin;
};
struct UnpackTable {
    int n;
    UnpackEntry e[24];
};
__global__ void __launch_bounds__(256)
unpack_grads_kernel(UnpackTable T, const float* __restrict__ gpk, float* __restrict__ grads) {
    griddep_launch();
    griddep_wait();
    int k = 0;
    while (k + 1 < T.n && (int)blockIdx.x >= T.e[k + 1].block_begin) ++k;
    const UnpackEntry E = T.e[k];
    const long long total = (long long)E.cout * E.cin * E.ntaps;
    const int per = E.cin * E.ntaps;
    for (long long i = (long long)(blockIdx.x - E.block_begin) * 2048 + threadIdx.x, it = 0; it < 8 && i < total;
         i += 256, ++it) {
        const int co = int(i / per);
        const int rem = int(i - (long long)co * per);
        const int ci = rem / E.ntaps, tap = rem - ci * E.ntaps;
        grads[E.dst_off + i] = gpk[E.src_off + ((long long)co * E.ntaps + tap) * E.cin + ci];
    }
}

// ------------------------------------------------------------------------------------------------ dgrad weight packs
// Generic tap-list pack of a K-major dgrad operand:  out[ci][col0 + t*cout + co] = w[co][ci0+ci][r_t][s_t]
// (w: OIHW fp32 [cout][cin_total][R][S]); `ld` = row length of out.
struct TapList {
    int n;
    int r[16], s[16];
};
__global__ void pack_dgrad_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout,
                                    int cin_total, int ci0, int cin, int R, int S, int ld, int col0, TapList taps) {
    griddep_launch();
    griddep_wait();
    const long long total = (long long)cin * taps.n * cout;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int co = int(i % cout);
        const int t = int((i / cout) % taps.n);
        const int ci = int(i / ((long long)cout * taps.n));
        out[(long long)ci * ld + col0 + t * cout + co] =
            ub_f2s(w[(((long long)co * cin_total + ci0 + ci) * R + taps.r[t]) * S + taps.s[t]]);
    }
}
// decoder conv1 dLow operand: 16 taps (ph,a,pw,b) -> out[c][t*cout + co] = sum_{r in R(ph,a), s in S(pw,b)} w[co][c][r][s]
// tap index t = ((ph*2 + a)*2 + pw)*2 + b
__global__ void pack_dec1_dlow_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout, int cup,
                                        int cin_total) {
    griddep_launch();
    griddep_wait();
    const long long total = (long long)cup * 16 * cout;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int co = int(i % cout);
        const int t = int((i / cout) % 16);
        const int c = int(i / ((long long)cout * 16));
        const int b = t & 1, pw = (t >> 1) & 1, a = (t >> 2) & 1, ph = (t >> 3) & 1;
        const int r0 = ph == 0 ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2);
        const int r1 = ph == 0 ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2);
        const int s0 = pw == 0 ? (b == 0 ? 0 : 1) : (b == 0 ? 0 : 2);
        const int s1 = pw == 0 ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : 2);
        float v = 0.f;
        for (int r = r0; r <= r1; ++r)
            for (int s = s0; s <= s1; ++s) v += w[((long long)co * cin_total + c) * 9 + r * 3 + s];
        out[i] = ub_f2s(v);
    }
}

}  // namespace ub
