// radix.cu -- hand-written stable LSD radix sort for the k-mer tuple lists (8-byte keys) and the
// seed-hit lists (16-byte records), 8-bit digits.  Replaces daligner's threaded byte-radix sort
// of k-mer tuples (what `daligner`/`damapper` run behind dazzler.d:6131-6170).
//
// Per pass: (1) per-CTA digit histogram over a contiguous range, (2) one-CTA exclusive scan of the
// [digit][cta] table, (3) stable scatter.  Grid = 4 CTAs per SM, each CTA walks its range in tiles.
// Algorithmic HBM traffic per pass: 2 reads + 1 write of the array.
#include "common.cuh"
#include <stdlib.h>

namespace dn {
namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;                       // per thread per tile
constexpr int RS_WTILE = 32 * RS_ITEMS;           // items per warp per tile (contiguous)
constexpr int RS_TILE = RS_WTILE * RS_WARPS;      // 2048

template <typename Item, int FIELD> __device__ __forceinline__ u32 digit_of(const Item &it, int shift);
template <> __device__ __forceinline__ u32 digit_of<u64, 0>(const u64 &it, int shift) { return (u32)(it >> shift) & 255u; }
template <> __device__ __forceinline__ u32 digit_of<ulonglong2, 0>(const ulonglong2 &it, int shift) { return (u32)(it.x >> shift) & 255u; }
template <> __device__ __forceinline__ u32 digit_of<ulonglong2, 1>(const ulonglong2 &it, int shift) { return (u32)((it.y & 0xffffffffull) >> shift) & 255u; }

__host__ __device__ inline size_t cta_range(size_t n, int G) {
    size_t per = (n + G - 1) / G;
    return (per + RS_TILE - 1) / RS_TILE * RS_TILE;
}

template <typename Item, int FIELD>
__global__ void __launch_bounds__(RS_THREADS) k_radix_hist(const Item *__restrict__ in, size_t n, int shift, u32 *__restrict__ hist /* [G][256] */) {
    __shared__ u32 h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const size_t per = cta_range(n, gridDim.x);
    size_t beg = per * blockIdx.x, end = beg + per; if (end > n) end = n;
    for (size_t i = beg + threadIdx.x; i < end; i += RS_THREADS) atomicAdd(&h[digit_of<Item, FIELD>(in[i], shift)], 1u);
    __syncthreads();
    hist[(size_t)blockIdx.x * 256 + threadIdx.x] = h[threadIdx.x];        // [cta][digit]: coalesced here and in the scan
}

// one CTA, 1024 threads: exclusive scan of hist in digit-major order; 4 threads share a digit's column (quarter each)
__global__ void __launch_bounds__(1024) k_radix_scan(u32 *hist, int G) {
    __shared__ u32 part[4][256];
    __shared__ u32 tot[256];
    const int dgt = threadIdx.x & 255, q = threadIdx.x >> 8;
    const int per = (G + 3) / 4, c0 = q * per, c1 = min(G, c0 + per);
    u32 *col = hist + dgt;
    u32 s = 0;
#pragma unroll 8
    for (int c = c0; c < c1; c++) s += col[(size_t)c * 256];
    part[q][dgt] = s;
    __syncthreads();
    if (threadIdx.x < 256) tot[threadIdx.x] = part[0][threadIdx.x] + part[1][threadIdx.x] + part[2][threadIdx.x] + part[3][threadIdx.x];
    __syncthreads();
    if (threadIdx.x < 32) {                              // exclusive scan of the 256 digit totals by one warp
        u32 v[8], t = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { v[k] = tot[threadIdx.x * 8 + k]; t += v[k]; }
        u32 inc = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u32 x = __shfl_up_sync(0xffffffffu, inc, o); if ((int)threadIdx.x >= o) inc += x; }
        u32 ex = inc - t;
#pragma unroll
        for (int k = 0; k < 8; k++) { tot[threadIdx.x * 8 + k] = ex; ex += v[k]; }
    }
    __syncthreads();
    u32 a = tot[dgt];
    for (int k = 0; k < q; k++) a += part[k][dgt];
#pragma unroll 8
    for (int c = c0; c < c1; c++) { u32 t = col[(size_t)c * 256]; col[(size_t)c * 256] = a; a += t; }
}

// Stable scatter of one pass.  Ranks come from warp-level match_any; the tile is then staged in shared memory in its
// sorted order, so that consecutive threads store consecutive items of a digit run: full sectors per run instead of one
// partial sector per thread (the direct register scatter ran at 1.4 TB/s on 16-byte tuples).
template <typename Item, int FIELD>
__global__ void __launch_bounds__(RS_THREADS, 4) k_radix_scatter(const Item *__restrict__ in, Item *__restrict__ out, size_t n, int shift,
                                                              const u32 *__restrict__ hist) {
    __shared__ u32 whist[RS_WARPS][256];
    __shared__ u32 base[256];                 // global position of the next item of each digit for this CTA
    __shared__ u32 tstart[256];               // first tile-local slot of each digit
    __shared__ Item stage[RS_TILE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    base[threadIdx.x] = hist[(size_t)blockIdx.x * 256 + threadIdx.x];
    const size_t per = cta_range(n, gridDim.x);
    size_t beg = per * blockIdx.x, end = beg + per; if (end > n) end = n;
    for (size_t t0 = beg; t0 < end; t0 += RS_TILE) {
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) whist[w][threadIdx.x] = 0;
        __syncthreads();
        Item it[RS_ITEMS]; u32 rk[RS_ITEMS]; u32 dg[RS_ITEMS];
        const size_t wbase = t0 + (size_t)warp * RS_WTILE;
#pragma unroll
        for (int s = 0; s < RS_ITEMS; s++) {
            size_t idx = wbase + s * 32 + lane;
            bool valid = idx < end;
            if (valid) it[s] = in[idx];
            u32 vmask = __ballot_sync(0xffffffffu, valid);
            if (valid) {
                u32 d = digit_of<Item, FIELD>(it[s], shift);
                dg[s] = d;
                u32 peers = __match_any_sync(vmask, d);
                u32 before = whist[warp][d];
                rk[s] = before + __popc(peers & ((1u << lane) - 1u));
                __syncwarp(vmask);
                if ((peers & ((1u << lane) - 1u)) == 0) whist[warp][d] = before + __popc(peers);
            }
            __syncwarp();
        }
        __syncthreads();
        u32 dtot;
        {   // digit = threadIdx.x: exclusive scan over warps (tile-local), digit total
            u32 a = 0;
#pragma unroll
            for (int w = 0; w < RS_WARPS; w++) { u32 t = whist[w][threadIdx.x]; whist[w][threadIdx.x] = a; a += t; }
            dtot = a;
        }
        {   // exclusive scan of the 256 digit totals -> tstart (warp shuffles + one smem hop)
            u32 inc = dtot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { u32 x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
            __shared__ u32 wsum[RS_WARPS];
            if (lane == 31) wsum[warp] = inc;
            __syncthreads();
            u32 add = 0;
#pragma unroll
            for (int w = 0; w < RS_WARPS; w++) if (w < warp) add += wsum[w];
            tstart[threadIdx.x] = add + inc - dtot;
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < RS_ITEMS; s++) {
            size_t idx = wbase + s * 32 + lane;
            if (idx < end) stage[tstart[dg[s]] + whist[warp][dg[s]] + rk[s]] = it[s];
        }
        __syncthreads();
        const u32 tile_n = (u32)((end - t0) < (size_t)RS_TILE ? (end - t0) : (size_t)RS_TILE);
#pragma unroll
        for (int s = 0; s < RS_ITEMS; s++) {
            const u32 p = (u32)s * RS_THREADS + threadIdx.x;
            if (p < tile_n) {
                const Item v = stage[p];
                const u32 d = digit_of<Item, FIELD>(v, shift);
                out[base[d] + (p - tstart[d])] = v;
            }
        }
        __syncthreads();
        base[threadIdx.x] += dtot;
        __syncthreads();
    }
}

// ---- one-sweep variant: digit histograms of ALL passes in one read of the input, then one kernel per pass in which every
// tile publishes its digit counts and finds its global offsets by looking back over its predecessors (decoupled look-back,
// one status word per (tile, digit): 2 flag bits + 30 bits).  Tiles are numbered by a ticket counter, so a tile only waits
// for tiles that already run.  Per pass 1R + 1W and one launch, instead of 2R + 1W and three launches.
constexpr int RS_MAXPASS = 8;

template <typename Item, int FIELD>
__global__ void __launch_bounds__(RS_THREADS) k_radix_hist_all(const Item *__restrict__ in, size_t n, int bit_lo, int npass,
                                                               u32 *__restrict__ ghist /* [npass][256] */) {
    __shared__ u32 h[RS_MAXPASS][256];
    for (int p = 0; p < npass; p++) h[p][threadIdx.x] = 0;
    __syncthreads();
    const size_t per = cta_range(n, gridDim.x);
    size_t beg = per * blockIdx.x, end = beg + per; if (end > n) end = n;
    for (size_t i = beg + threadIdx.x; i < end; i += RS_THREADS) {
        const Item it = in[i];
        for (int p = 0; p < npass; p++) atomicAdd(&h[p][digit_of<Item, FIELD>(it, bit_lo + 8 * p)], 1u);
    }
    __syncthreads();
    for (int p = 0; p < npass; p++) if (h[p][threadIdx.x]) atomicAdd(&ghist[p * 256 + threadIdx.x], h[p][threadIdx.x]);
}

template <typename Item, int FIELD>
__global__ void __launch_bounds__(RS_THREADS, 4) k_radix_onesweep(const Item *__restrict__ in, Item *__restrict__ out, size_t n, int shift,
                                                               const u32 *__restrict__ ghist /* [256] of this pass */,
                                                               u32 *status /* [0] ticket, [256 + tile * 256 + digit] */) {
    constexpr u32 FLAG_A = 1u << 30, FLAG_P = 2u << 30, VAL = (1u << 30) - 1u;
    __shared__ u32 whist[RS_WARPS][256];
    __shared__ u32 base[256];                 // global position of the first item of each digit of this tile
    __shared__ u32 tstart[256];               // first tile-local slot of each digit
    __shared__ Item stage[RS_TILE];
    __shared__ u32 wsum[RS_WARPS];
    __shared__ u32 s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(status, 1u);
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) whist[w][threadIdx.x] = 0;
    __syncthreads();
    const u32 tile = s_tile;
    const size_t t0 = (size_t)tile * RS_TILE, end = n;
    Item it[RS_ITEMS]; u32 rk[RS_ITEMS]; u32 dg[RS_ITEMS];
    const size_t wbase = t0 + (size_t)warp * RS_WTILE;
#pragma unroll
    for (int s = 0; s < RS_ITEMS; s++) {
        size_t idx = wbase + s * 32 + lane;
        bool valid = idx < end;
        if (valid) it[s] = in[idx];
        u32 vmask = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            u32 d = digit_of<Item, FIELD>(it[s], shift);
            dg[s] = d;
            u32 peers = __match_any_sync(vmask, d);
            u32 before = whist[warp][d];
            rk[s] = before + __popc(peers & ((1u << lane) - 1u));
            __syncwarp(vmask);
            if ((peers & ((1u << lane) - 1u)) == 0) whist[warp][d] = before + __popc(peers);
        }
        __syncwarp();
    }
    __syncthreads();
    u32 dtot;
    {   // digit = threadIdx.x: exclusive scan over warps (tile-local), digit total
        u32 a = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) { u32 t = whist[w][threadIdx.x]; whist[w][threadIdx.x] = a; a += t; }
        dtot = a;
    }
    // publish this tile's count of digit threadIdx.x; look back for the number of items of that digit in earlier tiles
    volatile u32 *st = status + 256;
    u32 excl = 0;
    if (tile == 0) st[threadIdx.x] = FLAG_P | dtot;
    else {
        st[(size_t)tile * 256 + threadIdx.x] = FLAG_A | dtot;
        for (long long j = (long long)tile - 1; j >= 0; j--) {
            u32 w;
            do { w = st[(size_t)j * 256 + threadIdx.x]; } while ((w >> 30) == 0u);
            excl += w & VAL;
            if (w & FLAG_P) break;
        }
        st[(size_t)tile * 256 + threadIdx.x] = FLAG_P | (excl + dtot);
    }
    u32 gex;
    {   // exclusive scans over the 256 digits: tile-local starts (tstart) and the pass's global digit starts (gex)
        const u32 gv = ghist[threadIdx.x];
        u32 inc = dtot, ginc = gv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 x = __shfl_up_sync(0xffffffffu, inc, o), y = __shfl_up_sync(0xffffffffu, ginc, o);
            if (lane >= o) { inc += x; ginc += y; }
        }
        __shared__ u32 gsum[RS_WARPS];
        if (lane == 31) { wsum[warp] = inc; gsum[warp] = ginc; }
        __syncthreads();
        u32 add = 0, gadd = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) if (w < warp) { add += wsum[w]; gadd += gsum[w]; }
        tstart[threadIdx.x] = add + inc - dtot;
        gex = gadd + ginc - gv;
    }
    base[threadIdx.x] = gex + excl;
    __syncthreads();
#pragma unroll
    for (int s = 0; s < RS_ITEMS; s++) {
        size_t idx = wbase + s * 32 + lane;
        if (idx < end) stage[tstart[dg[s]] + whist[warp][dg[s]] + rk[s]] = it[s];
    }
    __syncthreads();
    const u32 tile_n = (u32)((end - t0) < (size_t)RS_TILE ? (end - t0) : (size_t)RS_TILE);
#pragma unroll
    for (int s = 0; s < RS_ITEMS; s++) {
        const u32 p = (u32)s * RS_THREADS + threadIdx.x;
        if (p < tile_n) {
            const Item v = stage[p];
            const u32 d = digit_of<Item, FIELD>(v, shift);
            out[base[d] + (p - tstart[d])] = v;
        }
    }
}

template <typename Item, int FIELD>
Item *radix_sort_onesweep(Item *a, Item *b, size_t n, int bit_lo, int bit_hi, cudaStream_t s) {
    const int npass = (bit_hi - bit_lo + 7) / 8;
    const size_t tiles = (n + RS_TILE - 1) / RS_TILE;
    int G = sm_count() * 4; if ((size_t)G > tiles) G = (int)tiles;
    // one zeroed workspace: [npass][256] digit histograms, then per pass {256 words (ticket), tiles * 256 status words}
    const size_t per_pass = 256 + tiles * 256;
    DBuf<u32> ws((size_t)npass * 256 + (size_t)npass * per_pass); ws.zero(s);
    DN_LAUNCH((k_radix_hist_all<Item, FIELD>), G, RS_THREADS, 0, s, (const Item *)a, n, bit_lo, npass, ws.p);
    Item *src = a, *dst = b;
    for (int p = 0; p < npass; p++) {
        DN_LAUNCH((k_radix_onesweep<Item, FIELD>), (unsigned)tiles, RS_THREADS, 0, s, (const Item *)src, dst, n, bit_lo + 8 * p,
                  (const u32 *)(ws.p + p * 256), ws.p + (size_t)npass * 256 + (size_t)p * per_pass);
        Item *t = src; src = dst; dst = t;
    }
    return src;
}

template <typename Item, int FIELD>
Item *radix_sort_impl(Item *a, Item *b, size_t n, int bit_lo, int bit_hi, cudaStream_t s) {
    if (n == 0 || bit_hi <= bit_lo) return a;
    if (n >= (1ull << 32)) throw Error("radix sort: more than 2^32 items");
    static const bool three_launch = getenv("DN_RADIX3") != nullptr;
    if (!three_launch && n < (1ull << 30) && (bit_hi - bit_lo + 7) / 8 <= RS_MAXPASS) return radix_sort_onesweep<Item, FIELD>(a, b, n, bit_lo, bit_hi, s);
    static const int per_sm = getenv("DN_RADIX_CTAS") ? atoi(getenv("DN_RADIX_CTAS")) : 4;
    int G = sm_count() * (per_sm > 0 ? per_sm : 4);
    size_t need = (n + RS_TILE - 1) / RS_TILE;
    if ((size_t)G > need) G = (int)need;
    DBuf<u32> hist((size_t)256 * G);
    Item *src = a, *dst = b;
    for (int shift = bit_lo; shift < bit_hi; shift += 8) {
        DN_LAUNCH((k_radix_hist<Item, FIELD>), G, RS_THREADS, 0, s, (const Item *)src, n, shift, hist.p);
        DN_LAUNCH(k_radix_scan, 1, 1024, 0, s, hist.p, G);
        DN_LAUNCH((k_radix_scatter<Item, FIELD>), G, RS_THREADS, 0, s, (const Item *)src, dst, n, shift, (const u32 *)hist.p);
        Item *t = src; src = dst; dst = t;
    }
    return src;
}

}  // namespace

u64 *radix_sort_u64(u64 *keys, u64 *tmp, size_t n, int bit_lo, int bit_hi, cudaStream_t s) {
    return radix_sort_impl<u64, 0>(keys, tmp, n, bit_lo, bit_hi, s);
}
ulonglong2 *radix_sort_rec16(ulonglong2 *recs, ulonglong2 *tmp, size_t n, int field, int bit_lo, int bit_hi, cudaStream_t s) {
    if (field == 0) return radix_sort_impl<ulonglong2, 0>(recs, tmp, n, bit_lo, bit_hi, s);
    return radix_sort_impl<ulonglong2, 1>(recs, tmp, n, bit_lo, bit_hi, s);
}

}  // namespace dn
