// Table-driven weight re-pack: ONE launch converts every fp32 OIHW master tensor of a table into its bf16 operand layout
// (instead of ~50 small launches per optimizer step).  An entry describes one destination matrix; the element functions
// are the same layouts as the stand-alone kernels in ops.cuh / train_ops.cuh (which remain for the unit-test entry points).
#pragma once
#include <vector>

#include "ops.cuh"

namespace ub {

enum PackType { PK_CONV = 0, PK_STEM = 1, PK_DEC1 = 2, PK_DLOW = 3, PK_TAPS = 4, PK_HCONV = 5, PK_HPAR = 6, PK_STEM2 = 7, PK_HEAD = 8 };

struct PackEntry {
    int type;
    int cout, cin;            // PK_CONV: OIHW dims; PK_DEC1/PK_DLOW: cout, cup; PK_HCONV: rows (cout_), ctot; PK_TAPS: cout, cin
    int a, b, c, d;           // type-specific (see pack_elem)
    int ntaps;                // PK_TAPS
    unsigned long long taps;  // PK_TAPS: 4 bits per tap = (r << 2) | s
    long long src_off;        // element offset into the flat fp32 parameter array
    long long dst_off;        // element offset into the bf16 destination arena
    long long total;          // logical elements of this entry
    int block_begin;          // first block of this entry in the launch
    int pad;
    // iteration order (set by PackTable::add): a thread owns `jper` columns j and loops over the entry's `nloop` filter
    // taps, logical element i = (j / tap_stride) * tap_stride * nloop + j % tap_stride + t * tap_stride.  All taps of
    // one (co, ci) pair are 36 contiguous source bytes: read by ONE thread they hit L1 after the first load (the
    // element-per-thread order pulled every 32-byte sector of the fp32 weights nine times through L2).
    int nloop, jper;
    long long tap_stride;
};

// value of logical element i of entry E and the destination index (relative to E.dst_off) it is stored at
// IdxT = int for every entry of this network (largest: 512*512*9 elements): 64-bit div/mod per element made the pack
// kernels instruction-bound.
template <typename IdxT>
__device__ __forceinline__ float pack_elem(const PackEntry& E, const float* __restrict__ w, IdxT i, IdxT& dst) {
    dst = i;
    switch (E.type) {
        case PK_CONV: {  // a = R, b = S, c = flip: [co][(r*S+s)*cin + ci]   (flip: [ci][((R-1-r)*S + (S-1-s))*cout + co])
            const int R = E.a, S = E.b;
            if (!E.c) {
                const int ci = int(i % E.cin);
                const IdxT t = i / E.cin;
                const int rs = int(t % (R * S)), co = int(t / (R * S));
                return w[((IdxT)co * E.cin + ci) * R * S + rs];
            }
            const int co = int(i % E.cout);
            const IdxT t = i / E.cout;
            const int rs = int(t % (R * S)), ci = int(t / (R * S));
            const int r = R - 1 - rs / S, s = S - 1 - rs % S;
            return w[(((IdxT)co * E.cin + ci) * R + r) * S + s];
        }
        case PK_STEM: {  // [64][r*32 + px*4 + ch], px 0..7 <-> kernel column px-1 (px 0 and ch 3 are zero)
            const int ch = int(i % 4), px = int((i / 4) % 8), r = int((i / 32) % 7), co = int(i / 224);
            return (ch < 3 && px >= 1) ? w[((co * 3 + ch) * 7 + r) * 7 + (px - 1)] : 0.f;
        }
        case PK_DEC1: {  // cin = cup, a = cskip: out[parity][cout][9*cskip + 4*cup] (parity-folded 2x2 taps on the low-res x)
            const int cup = E.cin, cskip = E.a, kt = 9 * cskip + 4 * cup, cint = cup + cskip;
            const int k = int(i % kt), co = int((i / kt) % E.cout), par = int(i / ((IdxT)kt * E.cout));
            const int ph = par >> 1, pw = par & 1;
            if (k < 9 * cskip) {
                const int c = k % cskip, rs = k / cskip;
                return w[((IdxT)co * cint + cup + c) * 9 + rs];
            }
            const int kk = k - 9 * cskip, c = kk % cup, ab = kk / cup, aa = ab >> 1, bb = ab & 1;
            const int r0 = ph == 0 ? (aa == 0 ? 0 : 1) : (aa == 0 ? 0 : 2), r1 = ph == 0 ? (aa == 0 ? 0 : 2) : (aa == 0 ? 1 : 2);
            const int s0 = pw == 0 ? (bb == 0 ? 0 : 1) : (bb == 0 ? 0 : 2), s1 = pw == 0 ? (bb == 0 ? 0 : 2) : (bb == 0 ? 1 : 2);
            float v = 0.f;
            for (int r = r0; r <= r1; ++r)
                for (int s = s0; s <= s1; ++s) v += w[((IdxT)co * cint + c) * 9 + r * 3 + s];
            return v;
        }
        case PK_DLOW: {  // cin = cup, a = cin_total: out[c][t*cout + co], t = ((ph*2 + a)*2 + pw)*2 + b
            const int co = int(i % E.cout), t = int((i / E.cout) % 16), c = int(i / ((IdxT)E.cout * 16));
            const int bb = t & 1, pw = (t >> 1) & 1, aa = (t >> 2) & 1, ph = (t >> 3) & 1;
            const int r0 = ph == 0 ? (aa == 0 ? 0 : 1) : (aa == 0 ? 0 : 2), r1 = ph == 0 ? (aa == 0 ? 0 : 2) : (aa == 0 ? 1 : 2);
            const int s0 = pw == 0 ? (bb == 0 ? 0 : 1) : (bb == 0 ? 0 : 2), s1 = pw == 0 ? (bb == 0 ? 0 : 2) : (bb == 0 ? 1 : 2);
            float v = 0.f;
            for (int r = r0; r <= r1; ++r)
                for (int s = s0; s <= s1; ++s) v += w[((IdxT)co * E.a + c) * 9 + r * 3 + s];
            return v;
        }
        case PK_TAPS: {  // a = cin_total, b = ci0, c = (R << 8) | S, d = (ld << 0); pad = col0:  out[ci][col0 + t*cout + co]
            const int co = int(i % E.cout), t = int((i / E.cout) % E.ntaps), ci = int(i / ((IdxT)E.cout * E.ntaps));
            const int R = E.c >> 8, S = E.c & 0xFF;
            const int rs = int((E.taps >> (4 * t)) & 0xF), r = rs >> 2, s = rs & 3;
            dst = (IdxT)ci * E.d + E.pad + t * E.cout + co;
            return w[(((IdxT)co * E.a + E.b + ci) * R + r) * S + s];
        }
        case PK_STEM2: {  // tconv stem operand: [K chunk j = r*4 + q][cout group g][co % 8][8 elements], element e of chunk
            // (r, q) = input pixel px = 2q + e/4 (kernel column px - 1; px 0 is padding), channel e % 4 (3 is padding)
            const int e8 = int(i % 8), co8 = int((i / 8) % 8), g = int((i / 64) % 8), j = int(i / 512);
            const int r = j >> 2, q = j & 3, px = 2 * q + (e8 >> 2), ch = e8 & 3, co = g * 8 + co8;
            return (ch < 3 && px >= 1) ? w[((co * 3 + ch) * 7 + r) * 7 + (px - 1)] : 0.f;
        }
        case PK_HEAD: {  // seg head [1][16][3][3] as the stacked-column tconv operand (tconv mode 3): [filter row r][16 rows]
            // [16 ci], row 2c = bf16(w[r][c]), row 2c + 1 = the bf16 of the rounding residual w - bf16(w) for filter column
            // c = 0..2 (the epilogue adds the two accumulator columns: fp32-accurate weights on the bf16 tensor core),
            // rows 6..15 = 0; 32-byte rows, SWIZZLE_32B like PK_HCONV.  Elements past the 3 x 16 x 16 block are zero.
            const int ci = int(i % 16), row = int((i / 16) % 16), r = int(i / 256);
            const unsigned off = (unsigned)(i * 2);
            dst = (off ^ (((off >> 7) & 1u) << 4)) / 2;
            if (r > 2 || row > 5) return 0.f;
            const float wv = w[ci * 9 + r * 3 + (row >> 1)];
            const float hi = ub_s2f(ub_f2s(wv));
            return (row & 1) == 0 ? hi : wv - hi;
        }
        case PK_HPAR: {  // tconv parity operand: cout = rows, cin = cup (<= 64), a = cin_total of the OIHW tensor.
            // 18 blocks [blk][co][c] in the issue order of tc_issue_parity (tconv.cuh): block = (halo shift (r, s), output
            // parity (ph, pw)); value = sum of the 3x3 taps that land on low-res neighbour (a, b) = (r - ph, s - pw) for
            // that parity (same row/column sets as PK_DEC1), ZERO when (a, b) is not a neighbour (the filler blocks of
            // the N = 3 runs); 16-byte chunks swizzled like PK_HCONV
            const int cup = E.cin;
            const unsigned row_bytes = cup * 2;
            const int c = int(i % cup);
            IdxT t = i / cup;
            const int co = int(t % E.cout);
            const int blk = int(t / E.cout);
            int hr, hs, par;
            if (blk < 4) hr = 1, hs = 1, par = blk;
            else if (blk < 6) hr = 0, hs = 1, par = blk - 4;
            else if (blk < 8) hr = 2, hs = 1, par = blk - 4;
            else if (blk < 11) hr = 1, hs = 0, par = blk - 8;
            else if (blk < 14) hr = 1, hs = 2, par = blk - 10;
            else hr = ((blk - 14) >> 1) * 2, hs = ((blk - 14) & 1) * 2, par = blk - 14;
            const int ph = par >> 1, pw = par & 1, aa = hr - ph, bb = hs - pw;
            const unsigned off = (unsigned)(i * 2);
            dst = (off ^ (((off >> 7) & (row_bytes / 16 - 1)) << 4)) / 2;
            if (aa < 0 || aa > 1 || bb < 0 || bb > 1) return 0.f;
            const int r0 = ph == 0 ? (aa == 0 ? 0 : 1) : (aa == 0 ? 0 : 2), r1 = ph == 0 ? (aa == 0 ? 0 : 2) : (aa == 0 ? 1 : 2);
            const int s0 = pw == 0 ? (bb == 0 ? 0 : 1) : (bb == 0 ? 0 : 2), s1 = pw == 0 ? (bb == 0 ? 0 : 2) : (bb == 0 ? 1 : 2);
            float v = 0.f;
            for (int r = r0; r <= r1; ++r)
                for (int s = s0; s <= s1; ++s) v += w[((IdxT)co * E.a + c) * 9 + r * 3 + s];
            return v;
        }
        default: {  // PK_HCONV: cout = rows, cin = ctot, a = dim1_total, b = ci0, c = transposed (see pack_hconv_w_kernel)
            const int ctot = E.cin, cpr = ctot < 64 ? ctot : 64, nblk = ctot > 64 ? ctot / 64 : 1;
            const unsigned row_bytes = cpr * 2;
            const int cc = int(i % cpr);
            IdxT t = i / cpr;
            const int co = int(t % E.cout);
            t /= E.cout;
            const int blk = int(t % nblk), tap = int(t / nblk);
            const int c = blk * 64 + cc, r = tap / 3, s = tap - 3 * r;
            const unsigned off = (unsigned)(i * 2);
            dst = (off ^ (((off >> 7) & (row_bytes / 16 - 1)) << 4)) / 2;
            if (!E.c) return w[(((IdxT)co * E.a + E.b + c) * 3 + r) * 3 + s];
            return w[(((IdxT)c * E.a + E.b + co) * 3 + (2 - r)) * 3 + (2 - s)];
        }
    }
}

// The layouts that hold ~22 of the 24.4 M weights — every 3x3 conv of the wide layers, plain [co][(r,s)][ci] for the
// forward and flipped / transposed [ci][(2-r,2-s)][co] for the data gradient — as tiled transposes through shared memory:
// 128-byte coalesced fp32 loads, 64-byte coalesced bf16 stores.  (One thread per (co, ci) pair reading its nine taps
// straight from global memory issued 32 different lines per load instruction: LSU-bound, ~5x above the HBM time.)
// A block owns 256 (co, ci) pairs either way, so the table's block ranges are unchanged.
__device__ __forceinline__ bool pack_conv3x3_tiled(const PackEntry& E, const float* __restrict__ w,
                                                   __nv_bfloat16* __restrict__ out, float* tile /* [32 * 73] */) {
    if (E.type != PK_CONV || E.a != 3 || E.b != 3 || E.total >= (1ll << 30)) return false;
    const int tid = threadIdx.x, b = (int)blockIdx.x - E.block_begin;
    const int npairs = E.cout * E.cin;
    if (!E.c) {
        // plain: pair j = co * cin + ci; the block's 256 pairs are 2304 contiguous source floats
        const int j0 = b * 256, n = min(256, npairs - j0);
        if (n <= 0) return true;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int idx = k * 256 + tid;
            if (idx < n * 9) tile[idx] = __ldg(w + (long long)j0 * 9 + idx);
        }
        __syncthreads();
        if (tid < n) {
            const int j = j0 + tid, co = j / E.cin, ci = j - co * E.cin;
            __nv_bfloat16* dst = out + (long long)co * 9 * E.cin + ci;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) dst[tap * E.cin] = ub_f2s(tile[tid * 9 + tap]);   // smem stride 9: no conflicts
        }
        return true;
    }
    if ((E.cin & 7) || (E.cout & 31)) return false;
    // flipped / transposed: tile = 8 ci x 32 co; per co the 8 ci are 72 contiguous source floats
    const int tiles_co = E.cout / 32, ci0 = (b / tiles_co) * 8, co0 = (b % tiles_co) * 32;
    if (ci0 >= E.cin) return true;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int idx = k * 256 + tid, co_l = idx / 72, off = idx - co_l * 72;
        tile[co_l * 73 + off] = __ldg(w + ((long long)(co0 + co_l) * E.cin + ci0) * 9 + off);
    }
    __syncthreads();
    const int co_l = tid & 31;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int r = k * 8 + (tid >> 5), ci_l = r / 9, tap = r - ci_l * 9;
        out[((long long)(ci0 + ci_l) * 9 + tap) * E.cout + co0 + co_l] = ub_f2s(tile[co_l * 73 + ci_l * 9 + (8 - tap)]);
    }
    return true;
}

template <typename IdxT>
__device__ __forceinline__ void pack_entry_block(const PackEntry& E, const float* __restrict__ w, __nv_bfloat16* __restrict__ out) {
    const IdxT nloop = E.nloop, ts = (IdxT)E.tap_stride;
    const IdxT jtotal = (IdxT)(E.total / E.nloop);
    IdxT j = (IdxT)(blockIdx.x - E.block_begin) * (IdxT)(256 * E.jper) + (IdxT)threadIdx.x;
    for (int it = 0; it < E.jper; ++it, j += 256) {
        if (j >= jtotal) break;
        // fast paths of the two layouts that hold almost all elements: one index decomposition per (co, ci) pair, the
        // nine taps are 9 consecutive source floats (the generic path spent ~185 instructions per element on div/mod)
        if (E.type == PK_HCONV) {
            const int ctot = E.cin, cpr = ctot < 64 ? ctot : 64;
            const unsigned row_bytes = cpr * 2, swz = row_bytes / 16 - 1;
            const IdxT t0 = j / cpr;
            const int cc = int(j - t0 * cpr);
            const IdxT blk = t0 / E.cout;
            const int co = int(t0 - blk * E.cout), c = int(blk) * 64 + cc;
            const float* src = E.c ? w + ((IdxT)c * E.a + E.b + co) * 9 : w + ((IdxT)co * E.a + E.b + c) * 9;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const unsigned off = (unsigned)((j + (IdxT)tap * ts) * 2);
                out[(off ^ (((off >> 7) & swz) << 4)) / 2] = ub_f2s(__ldg(src + (E.c ? 8 - tap : tap)));
            }
            continue;
        }
        if (E.type == PK_CONV && E.a == 3 && E.b == 3) {
            const IdxT hi = j / ts;                 // co (plain) / ci (flip)
            const int lo = int(j - hi * ts);        // ci (plain) / co (flip)
            const float* src = E.c ? w + ((IdxT)lo * E.cin + hi) * 9 : w + j * 9;
            __nv_bfloat16* dst = out + hi * 9 * ts + lo;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) dst[(IdxT)tap * ts] = ub_f2s(__ldg(src + (E.c ? 8 - tap : tap)));
            continue;
        }
        const IdxT grp = j / ts;
        const IdxT i0 = grp * ts * nloop + (j - grp * ts);
        for (IdxT t = 0; t < nloop; ++t) {
            IdxT d;
            const float v = pack_elem<IdxT>(E, w, i0 + t * ts, d);
            out[d] = ub_f2s(v);
        }
    }
}

__global__ void __launch_bounds__(256)
pack_table_kernel(const PackEntry* __restrict__ tab, int n, const float* __restrict__ params,
                  __nv_bfloat16* __restrict__ dst_base) {
    griddep_launch();
    griddep_wait();
    int lo = 0, hi = n - 1;  // last entry whose block_begin <= blockIdx.x
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tab[mid].block_begin <= (int)blockIdx.x) lo = mid;
        else hi = mid - 1;
    }
    const PackEntry E = tab[lo];
    const float* w = params + E.src_off;
    __nv_bfloat16* out = dst_base + E.dst_off;
    __shared__ float tile[32 * 73];
    if (pack_conv3x3_tiled(E, w, out, tile)) return;
    if (E.total < (1ll << 30)) pack_entry_block<int>(E, w, out);
    else pack_entry_block<long long>(E, w, out);
}

// host side: a table under construction, uploaded once (the layouts only depend on the network description)
struct PackTable {
    std::vector<PackEntry> host;
    PackEntry* dev = nullptr;
    int nblocks = 0;
    void add(PackEntry e) {
        e.block_begin = nblocks;
        e.nloop = 1;
        e.tap_stride = e.total > 0 ? e.total : 1;
        if (e.type == PK_CONV && e.a * e.b > 1) {
            e.nloop = e.a * e.b;
            e.tap_stride = e.c ? e.cout : e.cin;
        } else if (e.type == PK_HCONV) {
            e.nloop = 9;
            e.tap_stride = e.total / 9;
        } else if (e.type == PK_TAPS && e.ntaps > 1) {
            e.nloop = e.ntaps;
            e.tap_stride = e.cout;
        }
        if (e.total % e.nloop) {   // not expected: fall back to one element per loop
            e.nloop = 1;
            e.tap_stride = e.total > 0 ? e.total : 1;
        }
        e.jper = e.nloop >= 8 ? 1 : 8 / e.nloop;
        const long long jtotal = e.total / e.nloop, per_block = 256ll * e.jper;
        nblocks += (int)((jtotal + per_block - 1) / per_block);
        host.push_back(e);
    }
    cudaError_t upload() {
        cudaFree(dev);
        dev = nullptr;
        cudaError_t e = cudaMalloc(&dev, host.size() * sizeof(PackEntry));
        if (e != cudaSuccess) return e;
        return cudaMemcpy(dev, host.data(), host.size() * sizeof(PackEntry), cudaMemcpyHostToDevice);
    }
    cudaError_t launch(const float* params, __nv_bfloat16* dst_base, cudaStream_t st) const {
        if (!nblocks) return cudaSuccess;
        launch_k(pack_table_kernel, nblocks, 256, 0, st, dev, (int)host.size(), params, dst_base);
        return cudaGetLastError();
    }
    ~PackTable() { cudaFree(dev); }
};

inline PackEntry pk_entry(int type, long long src_off, long long dst_off, long long total) {
    PackEntry e;
    memset(&e, 0, sizeof(e));
    e.type = type; e.src_off = src_off; e.dst_off = dst_off; e.total = total;
    return e;
}

}  // namespace ub
