// common.cuh -- shared device/host plumbing for the dentist_b200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdexcept>
#include <vector>
#include <atomic>

namespace dn {

typedef unsigned long long u64;
typedef unsigned int u32;

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

#define DN_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    char b_[512]; snprintf(b_, sizeof b_, "%s:%d: CUDA error %s (%s)", __FILE__, __LINE__, cudaGetErrorName(e_), cudaGetErrorString(e_)); \
    throw dn::Error(b_); } } while (0)

extern std::atomic<unsigned long long> g_launches;   // kernels launched by this library
#define DN_LAUNCH(kernel, grid, block, smem, stream, ...) do { \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); \
    dn::g_launches.fetch_add(1, std::memory_order_relaxed); \
    DN_CUDA(cudaGetLastError()); } while (0)

// Workspace arena: one grow-only slab of HBM per process, bump-allocated during a call and reset at
// the start of the next one, so a steady-state step performs no cudaMalloc / cudaFree at all
// (driver allocation calls cost milliseconds and serialise the device).
struct Arena {
    struct Chunk { char *p; size_t cap, used; };
    std::vector<Chunk> chunks;
    std::vector<void *> loose;                // DN_NO_ARENA=1: one cudaMalloc per buffer so compute-sanitizer sees every bound
    size_t high = 0, cur = 0;                 // bytes handed out in this call / high-water mark
    void *alloc(size_t bytes);
    void reset();                             // start of a call: reclaim everything, coalesce chunks
    void destroy();
};
Arena &arena();

// Recycling allocator for buffers that outlive a call (resident blocks): cudaMalloc / cudaFree cost
// milliseconds and synchronise the device, so freed buffers are kept and handed out again.
void *dcache_alloc(size_t bytes);
void dcache_free(void *p);
void dcache_destroy();

// RAII device buffer.  Default: carved from the arena (freed wholesale at the next reset).
// persistent(): an owned cudaMalloc allocation that outlives the call (resident blocks).
template <typename T> struct DBuf {
    T *p = nullptr; size_t n = 0; bool owned = false;
    DBuf() {}
    explicit DBuf(size_t n_) { alloc(n_); }
    DBuf(const DBuf &) = delete; DBuf &operator=(const DBuf &) = delete;
    DBuf(DBuf &&o) noexcept : p(o.p), n(o.n), owned(o.owned) { o.p = nullptr; o.n = 0; }
    DBuf &operator=(DBuf &&o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; owned = o.owned; o.p = nullptr; o.n = 0; } return *this; }
    ~DBuf() { release(); }
    void alloc(size_t n_) { release(); n = n_; owned = false; if (n) p = (T *)arena().alloc(n * sizeof(T)); }
    void persistent(size_t n_) { release(); n = n_; owned = true; if (n) p = (T *)dcache_alloc(n * sizeof(T)); }
    void release() { if (p && owned) dcache_free(p); p = nullptr; n = 0; }
    void zero(cudaStream_t s) { if (n) DN_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
    size_t bytes() const { return n * sizeof(T); }
};

// Pinned host staging buffer (grow-only, reused across calls) for device->host downloads.
struct PinnedBuf {
    void *p = nullptr; size_t cap = 0;
    void *get(size_t bytes);
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
};

int sm_count();
size_t l2_persisting_bytes(size_t want);     // scan.cu: sizes the device's persisting-L2 set-aside for a window of `want` bytes; returns the bytes it can count on

// ---- device primitives implemented in scan.cu / radix.cu ---------------------------------
// exclusive prefix sum: out[i] = sum_{j<i} in[j]; returns total through *d_total (device, 1 elem) if non-null
void exclusive_scan_u32_to_i64(const u32 *in, int64_t *out, size_t n, int64_t *d_total, cudaStream_t s);
void exclusive_scan_i32(const int32_t *in, int32_t *out, size_t n, int32_t *d_total, cudaStream_t s);

// Stable LSD radix sorts (hand-written; 8-bit digits).  `tmp` must hold n items.  Result is left in
// the buffer returned (either `keys` or `tmp`).
u64 *radix_sort_u64(u64 *keys, u64 *tmp, size_t n, int bit_lo, int bit_hi, cudaStream_t s);
// 16-byte records {u64 key, u64 val}.  field 0 sorts on key bits, field 1 on the low 32 bits of val.
ulonglong2 *radix_sort_rec16(ulonglong2 *recs, ulonglong2 *tmp, size_t n, int field, int bit_lo, int bit_hi, cudaStream_t s);

}  // namespace dn
