// segsort.cu -- segmented sort of the seed hits.  The index-lookup join emits hits grouped by
// (strand, B read) in ascending order, so only the order INSIDE each (strand, read) segment is missing:
// one CTA per segment sorts (gd << aposbits | apos) keys with a shared-memory bitonic network and
// writes the segment back in place.  Replaces 9 global radix passes over the whole hit list
// (13 ms of a 41 ms step on configs[1]) by one pass whose traffic is 2 x 16 B per hit.
#include "seed.cuh"

namespace dn {
namespace {

// Segment extents from the scanned word offsets (H = total hits, still on the device), and -- in the same pass -- the
// segments dealt out by size class: lists[c] collects the segments of class c (caps[0..2]; class 3 = larger than any
// shared-memory class), counts[c] their number.  The host reads the four counts together with H: no second round trip.
__global__ void __launch_bounds__(256) k_seg_offsets(const int64_t *__restrict__ woff, const int64_t *__restrict__ b_off, int nr,
                                                     int64_t nwB, const int64_t *__restrict__ dH, int64_t *__restrict__ seg_beg,
                                                     int32_t *__restrict__ seg_len, int minlen, int cap0, int cap1, int cap2,
                                                     int32_t *__restrict__ lists /* [4][2 * nr] */, u32 *__restrict__ counts /* [4] */) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * nr) return;
    const int64_t H = *dH;
    // a block that ends with empty reads has b_off[r] == padded total for them: their strand-1 segment starts at H, not past the array
    auto at = [&](int q) -> int64_t {
        if (q == 2 * nr) return H;
        const int st = q >= nr, r = q - st * nr;
        const int64_t x = (int64_t)st * nwB + (b_off[r] >> 4);
        return x >= 2 * nwB ? H : woff[x];
    };
    const int64_t b = at(i);
    const int64_t len = at(i + 1) - b;
    seg_beg[i] = b; seg_len[i] = (int32_t)len;
    if (len < minlen) return;
    const int c = len <= cap0 ? 0 : len <= cap1 ? 1 : len <= cap2 ? 2 : 3;
    lists[(int64_t)c * 2 * nr + atomicAdd(&counts[c], 1u)] = i;
}

__global__ void __launch_bounds__(1024) k_segsort(ulonglong2 *__restrict__ hits, const int64_t *__restrict__ seg_beg,
                                                 const int32_t *__restrict__ seg_len,
                                                 const int32_t *__restrict__ seglist, int gdbits, int aposbits) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int seg = seglist[blockIdx.x];
    const int64_t beg = seg_beg[seg];
    const int n = seg_len[seg];
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

// Stable in-CTA LSD radix sort of one segment by diagonal coordinate gd (<= 32 bits), 8-bit digits.
// Hits of one (strand, read) segment are emitted in ascending bpos, so on a fixed diagonal they are
// already in ascending apos: a STABLE sort by gd alone yields the (gd, apos) order -- half the key bits
// of the bitonic kernel and O(n) work per pass.  Keys (u32 gd) and original indices (u16) ping-pong in
// shared memory; the 16-byte records are gathered once at the end into the output buffer.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_segsort_radix(const ulonglong2 *__restrict__ in, ulonglong2 *__restrict__ out,
                                                              const int64_t *__restrict__ seg_beg, const int32_t *__restrict__ seg_len,
                                                              const int32_t *__restrict__ seglist, int cap, int gdbits) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u32 *key0 = reinterpret_cast<u32 *>(smem_raw), *key1 = key0 + cap;
    unsigned short *idx0 = reinterpret_cast<unsigned short *>(key1 + cap), *idx1 = idx0 + cap;
    u32 *whist = reinterpret_cast<u32 *>(idx1 + cap);          // [WARPS][256]
    u32 *dtot = whist + WARPS * 256;                            // [256] digit totals / bases
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int seg = seglist[blockIdx.x];
    const int64_t beg = seg_beg[seg];
    const int n = seg_len[seg];
    const u64 gdmask = (1ull << gdbits) - 1ull;
    for (int i = threadIdx.x; i < n; i += WARPS * 32) { key0[i] = (u32)(in[beg + i].x & gdmask); idx0[i] = (unsigned short)i; }
    const int per = ((n + WARPS - 1) / WARPS + 31) & ~31;       // contiguous, 32-aligned sub-range per warp
    const int wb = min(n, warp * per), we = min(n, wb + per);
    __syncthreads();
    u32 *kin = key0, *kout = key1; unsigned short *iin = idx0, *iout = idx1;
    for (int shift = 0; shift < gdbits; shift += 8) {
        for (int t = threadIdx.x; t < WARPS * 256; t += WARPS * 32) whist[t] = 0;
        __syncthreads();
        u32 *wh = whist + warp * 256;
        for (int i0 = wb; i0 < we; i0 += 32) {                  // per-warp digit histogram
            const int i = i0 + lane; const bool v = i < we;
            const u32 vm = __ballot_sync(0xffffffffu, v);
            if (v) {
                const u32 dg = (kin[i] >> shift) & 255u;
                const u32 peers = __match_any_sync(vm, dg);
                if ((peers & ((1u << lane) - 1u)) == 0) wh[dg] += __popc(peers);
            }
            __syncwarp();
        }
        __syncthreads();
        if (threadIdx.x < 256) {                                // digit totals, per-warp exclusive offsets
            u32 a = 0;
#pragma unroll
            for (int w = 0; w < WARPS; w++) { u32 t = whist[w * 256 + threadIdx.x]; whist[w * 256 + threadIdx.x] = a; a += t; }
            dtot[threadIdx.x] = a;
        }
        __syncthreads();
        if (warp == 0) {                                        // exclusive scan of the 256 digit totals
            u32 v[8], s = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) { v[q] = dtot[lane * 8 + q]; s += v[q]; }
            u32 inc = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            u32 ex = inc - s;
#pragma unroll
            for (int q = 0; q < 8; q++) { dtot[lane * 8 + q] = ex; ex += v[q]; }
        }
        __syncthreads();
        for (int i0 = wb; i0 < we; i0 += 32) {                  // stable scatter
            const int i = i0 + lane; const bool v = i < we;
            const u32 vm = __ballot_sync(0xffffffffu, v);
            if (v) {
                const u32 kk = kin[i];
                const u32 dg = (kk >> shift) & 255u;
                const u32 peers = __match_any_sync(vm, dg);
                const u32 before = wh[dg];
                const u32 pos = dtot[dg] + before + __popc(peers & ((1u << lane) - 1u));
                kout[pos] = kk; iout[pos] = iin[i];
                __syncwarp(vm);
                if ((peers & ((1u << lane) - 1u)) == 0) wh[dg] = before + __popc(peers);
            }
            __syncwarp();
        }
        __syncthreads();
        { u32 *t = kin; kin = kout; kout = t; unsigned short *u = iin; iin = iout; iout = u; }
    }
    const u64 hi = n > 0 ? (in[beg].x & ~gdmask) : 0ull;
    for (int i = threadIdx.x; i < n; i += WARPS * 32)
        out[beg + i] = make_ulonglong2(hi | kin[i], in[beg + iin[i]].y);
}

}  // namespace

void launch_seg_offsets(const int64_t *woff, const int64_t *b_off, int nr, int64_t nwB, const int64_t *dH, int64_t *seg_beg, int32_t *seg_len,
                        int minlen, const int *caps, int32_t *lists, u32 *counts, cudaStream_t s) {
    DN_LAUNCH(k_seg_offsets, (2 * nr + 255) / 256, 256, 0, s, woff, b_off, nr, nwB, dH, seg_beg, seg_len, minlen, caps[0], caps[1], caps[2],
              lists, counts);
}

// cap = per-segment capacity class (power of two); segments in `seglist` have at most cap hits
void launch_segsort(ulonglong2 *hits, const int64_t *seg_beg, const int32_t *seg_len, const int32_t *seglist, int nseg, int cap, int gdbits, int aposbits,
                    cudaStream_t s) {
    if (nseg == 0) return;
    const size_t smem = (size_t)cap * 12;
    static size_t attr_set = 0;
    if (smem > attr_set) {
        DN_CUDA(cudaFuncSetAttribute(k_segsort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = smem;
    }
    // big segments need many warps to hide shared-memory latency: 1024 threads for the 8k/16k classes
    DN_LAUNCH(k_segsort, nseg, cap > 2048 ? 1024 : 256, smem, s, hits, seg_beg, seg_len, seglist, gdbits, aposbits);
}

// radix variant (gdbits <= 32): reads `in`, writes `out`; segments not listed must be copied by the caller
void launch_segsort_radix(const ulonglong2 *in, ulonglong2 *out, const int64_t *seg_beg, const int32_t *seg_len, const int32_t *seglist, int nseg, int cap,
                          int gdbits, cudaStream_t s) {
    if (nseg == 0) return;
    if (cap <= 2048) {
        const size_t smem = (size_t)cap * 12 + 8 * 256 * 4 + 1024;
        DN_LAUNCH((k_segsort_radix<8>), nseg, 256, smem, s, in, out, seg_beg, seg_len, seglist, cap, gdbits);
    } else {
        const size_t smem = (size_t)cap * 12 + 16 * 256 * 4 + 1024;
        static size_t attr_set = 0;
        if (smem > attr_set) {
            DN_CUDA(cudaFuncSetAttribute(k_segsort_radix<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = smem;
        }
        DN_LAUNCH((k_segsort_radix<16>), nseg, 512, smem, s, in, out, seg_beg, seg_len, seglist, cap, gdbits);
    }
}

}  // namespace dn
