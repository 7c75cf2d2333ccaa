// segsort.cu -- segmented sort of the seed hits.  The index-lookup join emits hits grouped by
// (strand, B read) in ascending order, so only the order INSIDE each (strand, read) segment is missing:
// one CTA per segment sorts (gd << aposbits | apos) keys with a shared-memory bitonic network and
// writes the segment back in place.  Replaces 9 global radix passes over the whole hit list
// (13 ms of a 41 ms step on configs[1]) by one pass whose traffic is 2 x 16 B per hit.
#include "seed.cuh"

namespace dn {
namespace {

__global__ void __launch_bounds__(256) k_seg_offsets(const int64_t *__restrict__ woff, const int64_t *__restrict__ b_off, int nr,
                                                     int64_t nwB, int64_t H, int64_t *__restrict__ seg_off) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > 2 * nr) return;
    if (i == 2 * nr) { seg_off[i] = H; return; }
    const int st = i >= nr, r = i - st * nr;
    seg_off[i] = woff[st * nwB + (b_off[r] >> 4)];
}

__global__ void __launch_bounds__(1024) k_segsort(ulonglong2 *__restrict__ hits, const int64_t *__restrict__ seg_off,
                                                 const int32_t *__restrict__ seglist, int gdbits, int aposbits) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int seg = seglist[blockIdx.x];
    const int64_t beg = seg_off[seg];
    const int n = (int)(seg_off[seg + 1] - beg);
    int Np = 2; while (Np < n) Np <<= 1;
    u64 *key = reinterpret_cast<u64 *>(smem_raw);
    u32 *val = reinterpret_cast<u32 *>(key + Np);
    const u64 gdmask = (1ull << gdbits) - 1ull, amask = (1ull << aposbits) - 1ull;
    u64 hi = 0;
    for (int i = threadIdx.x; i < Np; i += blockDim.x) {
        if (i < n) {
            const ulonglong2 h = hits[beg + i];
            key[i] = ((h.x & gdmask) << aposbits) | (h.y & 0xffffffffull);
            val[i] = (u32)(h.y >> 32);
            hi = h.x & ~gdmask;
        } else { key[i] = ~0ull; val[i] = 0; }
    }
    __syncthreads();
    for (int k = 2; k <= Np; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (Np >> 1); t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const u64 a = key[i], b = key[i + j];
                const bool up = (i & k) == 0;
                if ((a > b) == up) {
                    key[i] = b; key[i + j] = a;
                    const u32 va = val[i]; val[i] = val[i + j]; val[i + j] = va;
                }
            }
            __syncthreads();
        }
    }
    // every thread needs the segment's high bits: broadcast from whoever loaded a record
    __shared__ u64 s_hi;
    if (threadIdx.x == 0) s_hi = hits[beg].x & ~gdmask;
    __syncthreads();
    (void)hi;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const u64 kk = key[i];
        hits[beg + i] = make_ulonglong2(s_hi | (kk >> aposbits), (kk & amask) | ((u64)val[i] << 32));
    }
}

}  // namespace

void launch_seg_offsets(const int64_t *woff, const int64_t *b_off, int nr, int64_t nwB, int64_t H, int64_t *seg_off, cudaStream_t s) {
    DN_LAUNCH(k_seg_offsets, (2 * nr + 1 + 255) / 256, 256, 0, s, woff, b_off, nr, nwB, H, seg_off);
}

// cap = per-segment capacity class (power of two); segments in `seglist` have at most cap hits
void launch_segsort(ulonglong2 *hits, const int64_t *seg_off, const int32_t *seglist, int nseg, int cap, int gdbits, int aposbits,
                    cudaStream_t s) {
    if (nseg == 0) return;
    const size_t smem = (size_t)cap * 12;
    static size_t attr_set = 0;
    if (smem > attr_set) {
        DN_CUDA(cudaFuncSetAttribute(k_segsort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = smem;
    }
    // big segments need many warps to hide shared-memory latency: 1024 threads for the 8k/16k classes
    DN_LAUNCH(k_segsort, nseg, cap > 2048 ? 1024 : 256, smem, s, hits, seg_off, seglist, gdbits, aposbits);
}

}  // namespace dn
