// bucket.cu -- the A-side k-mer index built WITHOUT a full sort.
//
// The lookup join needs the 8-byte tuples (kmer << shift | position) ordered by (kmer, position) plus the direct-address
// prefix table over the leading `tbits` bits of the k-mer.  The table is sized so that a bucket holds one or two tuples on
// average, so instead of five LSD radix passes (what daligner's threaded byte-radix sort of the k-mer tuples does behind
// dazzler.d:6131-6170) the tuples are counted per bucket, the counts are scanned -- which IS the prefix table --, the tuples
// are scattered to their bucket and every bucket is put in order on its own (two tuples: one compare; up to 32: insertion
// sort by its thread; up to 4096: a CTA's bitonic sort in shared memory).  The tuples are unique (the position is part of
// them), so the result is the same array the stable radix sort produces.  Blocks with a larger bucket (a k-mer prefix
// repeated thousands of times) take the radix path: build_index_u64 reports that and the caller falls back.
#include "seed.cuh"
#include "scan.cuh"

namespace dn {
namespace {

struct BucketOf {
    int key_shift, sh; u32 nq;
    __device__ __forceinline__ u32 operator()(u64 t) const { const u64 q = (t >> key_shift) >> sh; return q > nq ? nq : (u32)q; }
};

// C[q + 2] counts bucket q (q = nq: the invalid tuples of the k <= 15 form)
__global__ void __launch_bounds__(256) k_bucket_hist(const u64 *__restrict__ t, int64_t n, BucketOf bk, u32 *__restrict__ C) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&C[bk(__ldcs(t + i)) + 2], 1u);
}

// exclusive scan of C in place; side effect of the load: the largest bucket and the number of buckets beyond 32 tuples
struct BucketScan {
    u32 *C; u32 nq; u32 *stat /* [0] max, [1] buckets > 32 */;
    __device__ __forceinline__ u32 load(size_t i) const {
        const u32 v = C[i];
        if (v > 32u && i >= 2 && i - 2 < nq) { atomicMax(&stat[0], v); atomicAdd(&stat[1], 1u); }
        return v;
    }
    __device__ __forceinline__ void store(size_t i, u32 v) const { C[i] = v; }
};

// after the scan C[q + 2] = first slot of bucket q; every tuple takes the next slot of its bucket, which leaves
// C[q + 2] = first slot of bucket q + 1, i.e. T = C + 1 is the prefix table: T[q] = begin, T[q + 1] = end of bucket q
__global__ void __launch_bounds__(256) k_bucket_scatter(const u64 *__restrict__ t, int64_t n, BucketOf bk, u32 *__restrict__ C, u64 *__restrict__ out) {
    // four tuples per thread: the slot comes back from L2 (an atomic with a return value), so the four round trips overlap
    const int64_t base = (int64_t)blockIdx.x * (blockDim.x * 4) + threadIdx.x;
    u64 v[4]; u32 p[4];
#pragma unroll
    for (int x = 0; x < 4; x++) { const int64_t i = base + x * (int64_t)blockDim.x; if (i < n) v[x] = __ldcs(t + i); }
#pragma unroll
    for (int x = 0; x < 4; x++) { const int64_t i = base + x * (int64_t)blockDim.x; if (i < n) p[x] = atomicAdd(&C[bk(v[x]) + 2], 1u); }
#pragma unroll
    for (int x = 0; x < 4; x++) { const int64_t i = base + x * (int64_t)blockDim.x; if (i < n) out[p[x]] = v[x]; }
}

__global__ void __launch_bounds__(256) k_bucket_order(u64 *__restrict__ a, const u32 *__restrict__ T, u32 nq, u32 *__restrict__ biglist, u32 *__restrict__ nbig) {
    const u32 q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const u32 s = T[q], n = T[q + 1] - s;
    if (n < 2) return;
    u64 *b = a + s;
    if (n == 2) { const u64 x = b[0], y = b[1]; if (y < x) { b[0] = y; b[1] = x; } return; }
    if (n > 32) { biglist[atomicAdd(nbig, 1u)] = q; return; }
    // the bucket is read once (independent loads), ordered in thread-local storage (L1) and written back: an insertion sort
    // on the global array itself is a chain of dependent L2 round trips
    u64 v[32];
    for (u32 i = 0; i < n; i++) v[i] = b[i];
    for (u32 i = 1; i < n; i++) {
        const u64 x = v[i]; u32 j = i;
        while (j > 0 && v[j - 1] > x) { v[j] = v[j - 1]; j--; }
        v[j] = x;
    }
    for (u32 i = 0; i < n; i++) b[i] = v[i];
}

// buckets of 33 .. 4096 tuples: one CTA each, bitonic sort in shared memory
__global__ void __launch_bounds__(256) k_bucket_order_big(u64 *__restrict__ a, const u32 *__restrict__ T, const u32 *__restrict__ biglist) {
    __shared__ u64 sm[4096];
    const u32 q = biglist[blockIdx.x];
    const u32 s = T[q], n = T[q + 1] - s;
    u32 m = 64; while (m < n) m <<= 1;
    for (u32 i = threadIdx.x; i < m; i += blockDim.x) sm[i] = i < n ? a[s + i] : ~0ull;
    __syncthreads();
    for (u32 k = 2; k <= m; k <<= 1)
        for (u32 j = k >> 1; j > 0; j >>= 1) {
            for (u32 i = threadIdx.x; i < m; i += blockDim.x) {
                const u32 l = i ^ j;
                if (l > i) {
                    const u64 x = sm[i], y = sm[l];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { sm[i] = y; sm[l] = x; }
                }
            }
            __syncthreads();
        }
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) a[s + i] = sm[i];
}

}  // namespace

// tuples in `t` (n of them), out = the ordered index, C = nq + 3 words: the prefix table is C + 1 afterwards.
// Returns false (nothing usable written) when a bucket exceeds 4096 tuples: the caller sorts with the radix sort.
// One host synchronisation (the bucket statistics).
bool build_index_u64(const u64 *t, u64 *out, int64_t n, int key_shift, int sh, u32 nq, u32 *C, cudaStream_t s) {
    static const bool off = getenv("DN_INDEX_RADIX") != nullptr;
    if (off || n == 0 || n >= (1ll << 31)) return false;
    DN_CUDA(cudaMemsetAsync(C, 0, sizeof(u32) * ((size_t)nq + 3), s));
    DBuf<u32> stat(2); stat.zero(s);
    const BucketOf bk{key_shift, sh, nq};
    DN_LAUNCH(k_bucket_hist, (unsigned)((n + 255) / 256), 256, 0, s, t, n, bk, C);
    scan_chained<u32>(BucketScan{C, nq, stat.p}, (size_t)nq + 3, (u32 *)nullptr, s);
    u32 hs[2] = {0, 0};
    DN_CUDA(cudaMemcpyAsync(hs, stat.p, sizeof hs, cudaMemcpyDeviceToHost, s));
    DN_CUDA(cudaStreamSynchronize(s));
    if (hs[0] > 4096u) return false;
    // (the scatter completes a random 32-byte sector per 8-byte store in DRAM: 397 MB written + 220 MB read for an 80 MB
    //  array.  Holding the array in L2 through a persisting window was measured: the set-aside has to be re-sized around it,
    //  and that costs more than it saves -- 10.97 vs 10.78 ms per step.)
    DN_LAUNCH(k_bucket_scatter, (unsigned)((n + 1023) / 1024), 256, 0, s, t, n, bk, C, out);
    DBuf<u32> biglist((size_t)hs[1] + 1), nbig(1); nbig.zero(s);
    DN_LAUNCH(k_bucket_order, (nq + 255) / 256, 256, 0, s, out, (const u32 *)(C + 1), nq, biglist.p, nbig.p);
    if (hs[1] > 0) DN_LAUNCH(k_bucket_order_big, hs[1], 256, 0, s, out, (const u32 *)(C + 1), (const u32 *)biglist.p);
    return true;
}

}  // namespace dn
