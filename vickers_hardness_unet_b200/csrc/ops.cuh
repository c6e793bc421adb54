// Bandwidth-bound kernels around the implicit-GEMM convs: input pack, weight re-pack, BN fold, max-pool, seg head.
// All activations are NHWC bf16; vector width is 16 B (8 channels) wherever the layout allows.
#pragma once
#include "ptx.cuh"

namespace ub {

// ------------------------------------------------------------------------------------------------ input pack
// x fp32 NCHW [N,3,H,W]  ->  xp bf16 [N][H][W+8][4]  (4 zero pixels of padding left and right, channel 3 = 0).
// The stem reads 8-pixel x 4-channel windows (64 B) starting at padded pixel 2*ow, so every window start is
// 16 B aligned and no window leaves the row (SURVEY.md section 7 "7x7 stem with Cin=3").
__global__ void pack_input_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xp, int N, int H, int W) {
    griddep_launch();
    griddep_wait();
    const int Wp = W + 8;
    const long long total = (long long)N * H * (Wp / 2);  // two pixels (16 B) per thread
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const unsigned u = (unsigned)i, nh = u / (unsigned)(Wp / 2);   // 32-bit index arithmetic
        const int wp2 = int(u - nh * (unsigned)(Wp / 2));
        const int n = int(nh / (unsigned)H);
        const int h = int(nh - (unsigned)n * (unsigned)H);
        float v[2][3];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int w = wp2 * 2 + k - 4;
            const bool in = (w >= 0 && w < W);
#pragma unroll
            for (int c = 0; c < 3; ++c)
                v[k][c] = in ? __ldg(x + (((long long)n * 3 + c) * H + h) * W + w) : 0.f;
        }
        uint4 o;
        o.x = pack_bf16(v[0][0], v[0][1]);
        o.y = pack_bf16(v[0][2], 0.f);
        o.z = pack_bf16(v[1][0], v[1][1]);
        o.w = pack_bf16(v[1][2], 0.f);
        *reinterpret_cast<uint4*>(xp + ((long long)(n * (long long)H + h) * Wp + wp2 * 2) * 4) = o;
    }
}

// Same packed layout straight from camera frames: img uint8 HWC [N,H,W,3] (bgr != 0: channel order B,G,R as cv2.imread
// delivers it) -> ((v / 255) - mean[c]) / std[c] per RGB channel, i.e. the reference's host-side pre-processing
// (/root/reference/infer_pth_gui.py:46-48, ui_infer_rectangle.py:530-533) fused into the pack: 0.75 MB instead of
// 3 MB per 512x512 image cross the PCIe bus and the host never touches the pixels.
struct NormParams {
    float mean[3], inv_std[3];
};
__global__ void pack_input_u8_kernel(const uint8_t* __restrict__ img, __nv_bfloat16* __restrict__ xp, int N, int H, int W,
                                     int bgr, NormParams np) {
    griddep_launch();
    griddep_wait();
    const int Wp = W + 8;
    const long long total = (long long)N * H * (Wp / 2);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int wp2 = int(i % (Wp / 2));
        const long long nh = i / (Wp / 2);
        float v[2][3];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int w = wp2 * 2 + k - 4;
            const bool in = (w >= 0 && w < W);
            const uint8_t* px = img + (nh * W + (in ? w : 0)) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float raw = (float)__ldg(px + (bgr ? 2 - c : c)) * (1.f / 255.f);
                v[k][c] = in ? (raw - np.mean[c]) * np.inv_std[c] : 0.f;
            }
        }
        uint4 o;
        o.x = pack_bf16(v[0][0], v[0][1]);
        o.y = pack_bf16(v[0][2], 0.f);
        o.z = pack_bf16(v[1][0], v[1][1]);
        o.w = pack_bf16(v[1][2], 0.f);
        *reinterpret_cast<uint4*>(xp + (nh * Wp + wp2 * 2) * 4) = o;
    }
}

// ------------------------------------------------------------------------------------------------ weight re-pack
// OIHW fp32 [co][ci][R][S] -> K-major bf16 [co][(r*S+s)*cin + ci]; flip=1 additionally mirrors r,s and swaps the
// roles of co/ci (the dgrad operand: [ci][( (R-1-r)*S + (S-1-s) )*cout + co]).
__global__ void pack_conv_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout, int cin,
                                   int R, int S, int flip) {
    griddep_launch();
    griddep_wait();
    const long long total = (long long)cout * cin * R * S;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        // i indexes the OUTPUT so that writes are coalesced
        if (!flip) {
            const int ci = int(i % cin);
            long long t = i / cin;
            const int rs = int(t % (R * S));
            const int co = int(t / (R * S));
            out[i] = ub_f2s(w[((long long)co * cin + ci) * R * S + rs]);
        } else {
            const int co = int(i % cout);
            long long t = i / cout;
            const int rs = int(t % (R * S));
            const int ci = int(t / (R * S));
            const int r = R - 1 - rs / S, s = S - 1 - rs % S;
            out[i] = ub_f2s(w[(((long long)co * cin + ci) * R + r) * S + s]);
        }
    }
}

// hconv operand: OIHW fp32 -> K-major rows [tap][block = c/64][co][c % 64 (or all ctot < 64 channels)] bf16 for the channel
// slice [ci0, ci0+ctot), with the 16-byte chunks XOR-swizzled by the 128-byte line index of the byte offset inside the
// whole array (the kernel copies the array linearly to a 1 KB-aligned shared-memory region: csrc/hconv.cuh).
// transposed=1 builds the data-gradient operand instead: output rows are the conv's INPUT channels (cout_ := cin slice),
// K runs over the conv's output channels and the taps are mirrored:  value = w[c][ci0 + co][2-r][2-s].
__global__ void pack_hconv_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout, int ctot,
                                    int dim1_total, int ci0, int transposed) {
    griddep_launch();
    griddep_wait();
    const long long total = 9ll * ctot * cout;
    const int cpr = ctot < 64 ? ctot : 64;          // channels per row
    const int nblk = ctot > 64 ? ctot / 64 : 1;
    const unsigned row_bytes = cpr * 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        // i enumerates the LOGICAL (unswizzled) layout
        const int cc = int(i % cpr);
        long long t = i / cpr;
        const int co = int(t % cout);
        t /= cout;
        const int blk = int(t % nblk);
        const int tap = int(t / nblk);
        const int c = blk * 64 + cc;
        const int r = tap / 3, s = tap - 3 * r;
        float v;
        if (!transposed) v = w[(((long long)co * dim1_total + ci0 + c) * 3 + r) * 3 + s];
        else v = w[(((long long)c * dim1_total + ci0 + co) * 3 + (2 - r)) * 3 + (2 - s)];
        const unsigned off = (unsigned)(i * 2);
        const unsigned phys = off ^ (((off >> 7) & (row_bytes / 16 - 1)) << 4);
        out[phys / 2] = ub_f2s(v);
    }
}

// Stem: [64][3][7][7] -> [64][r*32 + px*4 + ch], px 0..7 <-> kernel column s = px-1 (px 0 and ch 3 are zero).
__global__ void pack_stem_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out) {
    griddep_launch();
    griddep_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 64 * 224) return;
    const int ch = i % 4, px = (i / 4) % 8, r = (i / 32) % 7, co = i / 224;
    float v = 0.f;
    if (ch < 3 && px >= 1) v = w[((co * 3 + ch) * 7 + r) * 7 + (px - 1)];
    out[i] = ub_f2s(v);
}

// Decoder conv1 (nearest-2x upsample + concat fused away): for output parity (ph,pw) the 3x3 taps over the
// up-sampled channels collapse onto a 2x2 low-resolution neighbourhood with summed weights.
//   w: [cout][cup+cskip][3][3] (cat order: up-sampled channels first)  ->  out[parity][cout][9*cskip + 4*cup]
//   K order: skip taps (r*3+s)*cskip + c, then low taps (a*2+b)*cup + c.
//   row sets: ph=0: a=0<-{r=0}, a=1<-{1,2};  ph=1: a=0<-{0,1}, a=1<-{2}   (same for columns).
__global__ void pack_dec1_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout, int cup,
                                   int cskip) {
    griddep_launch();
    griddep_wait();
    const int kt = 9 * cskip + 4 * cup;
    const long long total = 4ll * cout * kt;
    const int cin = cup + cskip;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int k = int(i % kt);
        const int co = int((i / kt) % cout);
        const int par = int(i / ((long long)kt * cout));
        const int ph = par >> 1, pw = par & 1;
        float v = 0.f;
        if (k < 9 * cskip) {
            const int c = k % cskip, rs = k / cskip;
            v = w[((long long)co * cin + cup + c) * 9 + rs];
        } else {
            const int kk = k - 9 * cskip;
            const int c = kk % cup, ab = kk / cup;
            const int a = ab >> 1, b = ab & 1;
            const int r0 = ph == 0 ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2);
            const int r1 = ph == 0 ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2);
            const int s0 = pw == 0 ? (b == 0 ? 0 : 1) : (b == 0 ? 0 : 2);
            const int s1 = pw == 0 ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : 2);
            for (int r = r0; r <= r1; ++r)
                for (int s = s0; s <= s1; ++s) v += w[((long long)co * cin + c) * 9 + r * 3 + s];
        }
        out[i] = ub_f2s(v);
    }
}

// eval-mode BatchNorm folded to per-channel scale/shift: y = x*scale + shift.
__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var, float eps,
                               float* __restrict__ scale, float* __restrict__ shift, int C) {
    griddep_launch();
    griddep_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float s = gamma[c] * rsqrtf(var[c] + eps);
    scale[c] = s;
    shift[c] = beta[c] - mean[c] * s;
}

// ------------------------------------------------------------------------------------------------ max-pool 3x3 s2 p1
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) { return ub_max2(a, b); }
__global__ void maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int N,
                                    int H, int W, int C) {
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
        uint4 m = make_uint4(kNegInfPair, kNegInfPair, kNegInfPair, kNegInfPair);  // -inf pairs
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int h = ho * 2 - 1 + r;
            if (h < 0 || h >= H) continue;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int w = wo * 2 - 1 + s;
                if (w < 0 || w >= W) continue;
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + (((long long)n * H + h) * W + w) * C) + c8);
                m.x = bf16x2_max(m.x, v.x);
                m.y = bf16x2_max(m.y, v.y);
                m.z = bf16x2_max(m.z, v.z);
                m.w = bf16x2_max(m.w, v.w);
            }
        }
        reinterpret_cast<uint4*>(out + (((long long)n * Ho + ho) * Wo + wo) * C)[c8] = m;
    }
}

// ------------------------------------------------------------------------------------------------ segmentation head
// Conv2d(16,1,3,padding=1) with bias on NHWC bf16 -> fp32 logits [N,1,H,W]; optionally also sigmoid probabilities
// and/or a uint8 {0,255} mask at `thresh_logit` (sigmoid(x) >= t  <=>  x >= logit(t)).
// One thread per output pixel; a 16x16 pixel tile (+halo) is staged in shared memory.
constexpr int kHeadTile = 16;
__global__ void __launch_bounds__(256)
head_conv_kernel(const __nv_bfloat16* __restrict__ in, const float* __restrict__ w /*[16][3][3]*/,
                 const float* __restrict__ bias, float* __restrict__ logits, float* __restrict__ prob,
                 uint8_t* __restrict__ mask, float thresh_logit, int N, int H, int W) {
    griddep_launch();
    griddep_wait();
    __shared__ uint4 tile[(kHeadTile + 2) * (kHeadTile + 2) * 2];  // [18][18][16 ch bf16 = 2 x uint4]
    __shared__ float sw[9 * 16];
    const int tx = threadIdx.x % kHeadTile, ty = threadIdx.x / kHeadTile;
    const int w0 = blockIdx.x * kHeadTile, h0 = blockIdx.y * kHeadTile, n = blockIdx.z;
    if (threadIdx.x < 144) {
        const int c = threadIdx.x % 16, rs = threadIdx.x / 16;  // sw[rs][c]
        sw[threadIdx.x] = w[c * 9 + rs];
    }
    for (int i = threadIdx.x; i < (kHeadTile + 2) * (kHeadTile + 2) * 2; i += 256) {
        const int half = i & 1, p = i >> 1;
        const int ww = w0 - 1 + p % (kHeadTile + 2), hh = h0 - 1 + p / (kHeadTile + 2);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (ww >= 0 && ww < W && hh >= 0 && hh < H)
            v = __ldg(reinterpret_cast<const uint4*>(in + (((long long)n * H + hh) * W + ww) * 16) + half);
        tile[i] = v;
    }
    __syncthreads();
    const int ow = w0 + tx, oh = h0 + ty;
    if (ow >= W || oh >= H) return;
    float acc = bias[0];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const int p = (ty + r) * (kHeadTile + 2) + (tx + s);
            const float* wr = sw + (r * 3 + s) * 16;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const uint4 v = tile[p * 2 + half];
                const float* wh = wr + half * 8;
                acc += bf16_lo(v.x) * wh[0] + bf16_hi(v.x) * wh[1] + bf16_lo(v.y) * wh[2] + bf16_hi(v.y) * wh[3] +
                       bf16_lo(v.z) * wh[4] + bf16_hi(v.z) * wh[5] + bf16_lo(v.w) * wh[6] + bf16_hi(v.w) * wh[7];
            }
        }
    const long long o = ((long long)n * H + oh) * W + ow;
    if (logits) logits[o] = acc;
    if (prob) prob[o] = 1.f / (1.f + __expf(-acc));
    if (mask) mask[o] = acc >= thresh_logit ? 255 : 0;
}

// out[j] = sum_t part[t][j]  (deterministic second stage of the per-tile statistics reduction)
__global__ void reduce_partials_kernel(const float* __restrict__ part, float* __restrict__ out, int ntiles, int width) {
    griddep_launch();
    griddep_wait();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= width) return;
    float s = 0.f;
    for (int t = 0; t < ntiles; ++t) s += part[(size_t)t * width + j];
    out[j] = s;
}

inline int ew_grid(long long total, int block, int num_sms) {
    long long g = (total + block - 1) / block;
    const long long cap = (long long)num_sms * 16;
    return int(g < cap ? (g > 0 ? g : 1) : cap);
}

// grid for kernels that process two elements per loop iteration and whose threads keep a fixed channel group:
// half the blocks of ew_grid for small tensors, and blockDim * grid stays a multiple of the channel-group count c8.
inline int ew_grid2(long long total, int block, int num_sms, int c8) {
    long long g = (total + 2LL * block - 1) / (2LL * block);
    const long long cap = (long long)num_sms * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    if (c8 > block) {   // blockDim * g must be a multiple of c8
        const long long m = c8 / block;
        g = (g + m - 1) / m * m;
    }
    return int(g);
}

}  // namespace ub
