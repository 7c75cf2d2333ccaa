// scan.cuh -- the single-pass (decoupled look-back) exclusive scan as a template over a load / store functor, so that a
// producer can be fused into the load and a consumer into the store (seed.cu: hit cover + band flags -> band index + cover sums).
#pragma once
#include "common.cuh"

namespace dn {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename Tout>
__device__ __forceinline__ Tout block_exclusive(Tout v, Tout *total, Tout *smem /* 33 */) {
    // exclusive scan of one value per thread across the block
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Tout inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { Tout t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        Tout w = lane < (SCAN_THREADS / 32) ? smem[lane] : Tout(0);
        Tout winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { Tout t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
        smem[lane] = winc - w;               // exclusive warp offsets
        if (lane == 31) smem[32] = winc;     // block total
    }
    __syncthreads();
    Tout res = smem[warp] + inc - v;
    if (total) *total = smem[32];
    __syncthreads();
    return res;
}

// Every tile publishes its sum, then the inclusive prefix of everything up to it, in one 64-bit status word (2 flag bits +
// 62 value bits); a tile adds up its predecessors' words until it meets a prefix.  Tiles are numbered by a ticket counter,
// so a tile only ever waits for tiles that are already running.  One launch and 1R + 1W of the array instead of three
// launches and 2R + 1W.  status[0] = ticket, status[1 + t] = tile t; zeroed by the host.  Values must stay below 2^62.
// Op: Tout load(size_t i) const; void store(size_t i, Tout exclusive_prefix) const.
template <typename Tout, typename Op>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_chained(Op op, size_t n, unsigned long long *status, Tout *total_out) {
    constexpr unsigned long long FLAG_A = 1ull << 62, FLAG_P = 2ull << 62, VAL = (1ull << 62) - 1ull;
    __shared__ Tout sm[33];
    __shared__ Tout tile[SCAN_TILE + SCAN_TILE / 32];
    __shared__ unsigned s_tile; __shared__ Tout s_excl;
    auto at = [](int x) { return x + (x >> 5); };
    if (threadIdx.x == 0) s_tile = (unsigned)atomicAdd(status, 1ull);
    __syncthreads();
    const unsigned t = s_tile;
    const size_t base = (size_t)t * SCAN_TILE;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        size_t idx = base + (size_t)i * SCAN_THREADS + threadIdx.x;
        tile[at(i * SCAN_THREADS + threadIdx.x)] = idx < n ? op.load(idx) : Tout(0);
    }
    __syncthreads();
    Tout v[SCAN_ITEMS]; Tout acc = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) { v[i] = tile[at(threadIdx.x * SCAN_ITEMS + i)]; acc += v[i]; }
    Tout total;
    Tout ex = block_exclusive<Tout>(acc, &total, sm);
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        volatile unsigned long long *st = status + 1;
        Tout excl = 0;
        if (t == 0) { if (lane == 0) st[0] = FLAG_P | (unsigned long long)total; }
        else {
            if (lane == 0) st[t] = FLAG_A | (unsigned long long)total;
            long long j = (long long)t - 1;
            for (;;) {
                const long long idx = j - lane;
                const unsigned long long w = idx >= 0 ? st[idx] : FLAG_P;          // before the first tile: prefix 0
                const unsigned f = (unsigned)(w >> 62);
                const unsigned nr = __ballot_sync(0xffffffffu, f == 0u), pm = __ballot_sync(0xffffffffu, f == 2u);
                const int fp = pm ? __ffs(pm) - 1 : 32;                             // closest predecessor holding a prefix
                const unsigned need = fp >= 31 ? 0xffffffffu : ((2u << fp) - 1u);
                if (nr & need) continue;                                            // one of the words we need is not published yet
                Tout x = lane <= fp ? (Tout)(w & VAL) : Tout(0);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
                excl += x;
                if (fp < 32) break;
                j -= 32;
            }
            if (lane == 0) st[t] = FLAG_P | (unsigned long long)(excl + total);
        }
        if (lane == 0) { s_excl = excl; if (total_out && t == gridDim.x - 1) *total_out = excl + total; }
    }
    __syncthreads();
    ex += s_excl;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) { tile[at(threadIdx.x * SCAN_ITEMS + i)] = ex; ex += v[i]; }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        size_t idx = base + (size_t)i * SCAN_THREADS + threadIdx.x;
        if (idx < n) op.store(idx, tile[at(i * SCAN_THREADS + threadIdx.x)]);
    }
}

template <typename Tout, typename Op>
void scan_chained(Op op, size_t n, Tout *d_total, cudaStream_t s) {
    if (n == 0) { if (d_total) DN_CUDA(cudaMemsetAsync(d_total, 0, sizeof(Tout), s)); return; }
    const size_t nbt = (n + SCAN_TILE - 1) / SCAN_TILE;
    DBuf<unsigned long long> st(nbt + 1); st.zero(s);
    DN_LAUNCH((k_scan_chained<Tout, Op>), (unsigned)nbt, SCAN_THREADS, 0, s, op, n, st.p, d_total);
}

template <typename Tin, typename Tout> struct ScanPlain {
    const Tin *in; Tout *out;
    __device__ __forceinline__ Tout load(size_t i) const { return (Tout)in[i]; }
    __device__ __forceinline__ void store(size_t i, Tout v) const { out[i] = v; }
};

}  // namespace dn
