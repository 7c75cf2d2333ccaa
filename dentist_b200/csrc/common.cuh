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

// Stream the engine is currently enqueuing on (thread-local; set by the C-ABI entry points).
cudaStream_t &cur_stream();

// RAII device buffer on the stream-ordered allocator (cudaMallocAsync): allocations and frees are
// enqueued on the engine stream and served from a cached pool, so a step does no cudaMalloc/cudaFree.
template <typename T> struct DBuf {
    T *p = nullptr; size_t n = 0;
    DBuf() {}
    explicit DBuf(size_t n_) { alloc(n_); }
    DBuf(const DBuf &) = delete; DBuf &operator=(const DBuf &) = delete;
    DBuf(DBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DBuf &operator=(DBuf &&o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
    ~DBuf() { release(); }
    void alloc(size_t n_) { release(); n = n_; if (n) DN_CUDA(cudaMallocAsync((void **)&p, n * sizeof(T), cur_stream())); }
    void release() { if (p) cudaFreeAsync(p, cur_stream()); p = nullptr; n = 0; }
    void zero(cudaStream_t s) { if (n) DN_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
    size_t bytes() const { return n * sizeof(T); }
};

int sm_count();

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
