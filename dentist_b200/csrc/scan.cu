// scan.cu -- hand-written exclusive prefix sums (three-phase: tile reduce, recursive scan of the
// tile sums, tile scan + offset).  HBM-bound: reads the input twice, writes the output once.
#include "common.cuh"
#include "scan.cuh"
#include <mutex>
#include <stdlib.h>

namespace dn {

std::atomic<unsigned long long> g_launches{0};

Arena &arena() { static Arena a; return a; }

void *Arena::alloc(size_t bytes) {
    static const bool no_arena = getenv("DN_NO_ARENA") != nullptr;
    if (no_arena) { void *p = nullptr; DN_CUDA(cudaMalloc(&p, bytes ? bytes : 1)); loose.push_back(p); return p; }
    bytes = (bytes + 511) & ~(size_t)511;
    cur += bytes; if (cur > high) high = cur;
    if (!chunks.empty()) {
        Chunk &c = chunks.back();
        if (c.used + bytes <= c.cap) { void *r = c.p + c.used; c.used += bytes; return r; }
    }
    size_t cap = bytes > (1ull << 30) ? bytes : (1ull << 30);
    Chunk c{nullptr, cap, bytes};
    DN_CUDA(cudaMalloc((void **)&c.p, cap));
    chunks.push_back(c);
    return c.p;
}

void Arena::reset() {
    if (!loose.empty()) { cudaDeviceSynchronize(); for (void *p : loose) cudaFree(p); loose.clear(); }
    cur = 0;
    if (chunks.size() > 1 || (chunks.size() == 1 && chunks[0].cap < high)) {     // coalesce into one slab of the high-water size
        cudaDeviceSynchronize();
        for (Chunk &c : chunks) cudaFree(c.p);
        chunks.clear();
        size_t cap = high + (high >> 3) + (64ull << 20);
        Chunk c{nullptr, cap, 0};
        if (cudaMalloc((void **)&c.p, cap) == cudaSuccess) chunks.push_back(c); else cudaGetLastError();
    }
    for (Chunk &c : chunks) c.used = 0;
}

void Arena::destroy() { for (Chunk &c : chunks) cudaFree(c.p); chunks.clear(); high = cur = 0; }

namespace {
struct DBlock { void *p; size_t cap; };
std::mutex g_dmu; std::vector<DBlock> g_dfree; std::vector<DBlock> g_dlive; size_t g_dfree_bytes = 0;
}
void *dcache_alloc(size_t bytes) {
    bytes = (bytes + 511) & ~(size_t)511;
    std::lock_guard<std::mutex> lk(g_dmu);
    int best = -1;
    for (int i = 0; i < (int)g_dfree.size(); i++)
        if (g_dfree[i].cap >= bytes && g_dfree[i].cap <= bytes + (bytes >> 1) + (1 << 20) && (best < 0 || g_dfree[i].cap < g_dfree[best].cap)) best = i;
    DBlock b;
    if (best >= 0) { b = g_dfree[best]; g_dfree.erase(g_dfree.begin() + best); g_dfree_bytes -= b.cap; }
    else { b.cap = bytes; DN_CUDA(cudaMalloc(&b.p, bytes)); }
    g_dlive.push_back(b);
    return b.p;
}
void dcache_free(void *p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_dmu);
    for (size_t i = 0; i < g_dlive.size(); i++)
        if (g_dlive[i].p == p) {
            DBlock b = g_dlive[i]; g_dlive.erase(g_dlive.begin() + i);
            if (g_dfree.size() < 64 && g_dfree_bytes + b.cap <= (8ull << 30)) { g_dfree.push_back(b); g_dfree_bytes += b.cap; }
            else cudaFree(b.p);
            return;
        }
    cudaFree(p);
}
void dcache_destroy() {
    std::lock_guard<std::mutex> lk(g_dmu);
    for (auto &b : g_dfree) cudaFree(b.p);
    g_dfree.clear(); g_dfree_bytes = 0;
}

void *PinnedBuf::get(size_t bytes) {
    if (bytes > cap) {
        if (p) cudaFreeHost(p);
        p = nullptr; cap = bytes + (bytes >> 2) + (1 << 20);
        DN_CUDA(cudaHostAlloc(&p, cap, cudaHostAllocDefault));
    }
    return p;
}

// The L2 set-aside for persisting accesses comes OUT of the L2 every other access shares (measured: the device maximum as a
// standing limit costs the lookup join 1.5 ms per step), so it is sized by the window that is about to use it and changed
// only when that size changes.  Returns the bytes a window of `want` bytes can count on.
size_t l2_persisting_bytes(size_t want) {
    static size_t maxb = [] {
        int dev = 0, mx = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&mx, cudaDevAttrMaxPersistingL2CacheSize, dev);
        return (size_t)(mx > 0 ? mx : 0);
    }();
    static size_t cur = (size_t)-1;
    size_t lim = want + (want >> 3); if (lim > maxb) lim = maxb;
    if (lim != cur) {
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, lim) != cudaSuccess) { cudaGetLastError(); return 0; }
        cur = lim;
    }
    return want < lim ? want : lim;
}

int sm_count() {
    static int n = 0;
    if (!n) { int dev; DN_CUDA(cudaGetDevice(&dev)); DN_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev)); }
    return n;
}

namespace {
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const Tin *__restrict__ in, Tout *__restrict__ sums, size_t n) {
    __shared__ Tout sm[33];
    size_t base = (size_t)blockIdx.x * SCAN_TILE;
    Tout acc = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        size_t idx = base + (size_t)i * SCAN_THREADS + threadIdx.x;     // coalesced
        if (idx < n) acc += (Tout)in[idx];
    }
    Tout total;
    block_exclusive<Tout>(acc, &total, sm);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const Tin *__restrict__ in, Tout *__restrict__ out,
                                                             const Tout *__restrict__ offs, size_t n, Tout *total_out) {
    __shared__ Tout sm[33];
    // one pad slot per 32 elements: the transposed accesses (thread t owns slots 16t .. 16t+15) would otherwise hit the
    // same bank from every second thread (16-way conflicts)
    __shared__ Tout tile[SCAN_TILE + SCAN_TILE / 32];
    auto at = [](int x) { return x + (x >> 5); };
    size_t base = (size_t)blockIdx.x * SCAN_TILE;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        size_t idx = base + (size_t)i * SCAN_THREADS + threadIdx.x;
        tile[at(i * SCAN_THREADS + threadIdx.x)] = idx < n ? (Tout)in[idx] : Tout(0);
    }
    __syncthreads();
    Tout v[SCAN_ITEMS]; Tout acc = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) { v[i] = tile[at(threadIdx.x * SCAN_ITEMS + i)]; acc += v[i]; }
    Tout total;
    Tout ex = block_exclusive<Tout>(acc, &total, sm) + (offs ? offs[blockIdx.x] : Tout(0));
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) { tile[at(threadIdx.x * SCAN_ITEMS + i)] = ex; ex += v[i]; }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        size_t idx = base + (size_t)i * SCAN_THREADS + threadIdx.x;
        if (idx < n) out[idx] = tile[at(i * SCAN_THREADS + threadIdx.x)];
    }
    if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0)
        *total_out = (offs ? offs[blockIdx.x] : Tout(0)) + total;
}

template <typename Tin, typename Tout>
void scan_rec(const Tin *in, Tout *out, size_t n, Tout *d_total, cudaStream_t s) {
    if (n == 0) { if (d_total) DN_CUDA(cudaMemsetAsync(d_total, 0, sizeof(Tout), s)); return; }
    static const bool three_phase = getenv("DN_SCAN3") != nullptr;
    if (!three_phase) { scan_chained<Tout>(ScanPlain<Tin, Tout>{in, out}, n, d_total, s); return; }
    size_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (nb == 1) {
        DN_LAUNCH((k_scan_apply<Tin, Tout>), 1, SCAN_THREADS, 0, s, in, out, (const Tout *)nullptr, n, d_total);
        return;
    }
    DBuf<Tout> sums(nb), offs(nb);
    DN_LAUNCH((k_scan_reduce<Tin, Tout>), (unsigned)nb, SCAN_THREADS, 0, s, in, sums.p, n);
    scan_rec<Tout, Tout>(sums.p, offs.p, nb, nullptr, s);
    DN_LAUNCH((k_scan_apply<Tin, Tout>), (unsigned)nb, SCAN_THREADS, 0, s, in, out, (const Tout *)offs.p, n, d_total);
}
}  // namespace

void exclusive_scan_u32_to_i64(const u32 *in, int64_t *out, size_t n, int64_t *d_total, cudaStream_t s) {
    scan_rec<u32, long long>(in, (long long *)out, n, (long long *)d_total, s);
}
void exclusive_scan_i32(const int32_t *in, int32_t *out, size_t n, int32_t *d_total, cudaStream_t s) {
    scan_rec<int32_t, int32_t>(in, out, n, d_total, s);
}

}  // namespace dn
