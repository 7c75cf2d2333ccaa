// dust.cu -- low-complexity (DUST) mask of a resident block: what `DBdust` writes into the `dust` track
// behind dbdust() (dazzler.d:3815-3818; calls processPileUps/package.d:476, 655).
// Specification (oracle/dust.py; DAZZ_DB's DBdust is absent -> parity unpinned): windows of `w` bases at
// stride w/2 over every read; a window with l = len-2 >= 14 triplets is low-complexity when
// 10 * sum_t c_t*(c_t-1)/2 > T10 * (l-1)  (c_t = count of triplet t, T10 = round(10*threshold));
// masked intervals = union of low-complexity windows, merged, intervals shorter than `minlen` dropped.
#include "api_internal.hpp"
#include <string.h>

namespace dn {
namespace {

// thread per window over ALL windows of the block (woff = first window of each read): a block of few long reads (flanking
// contigs) fills the GPU like one of many short reads
__global__ void __launch_bounds__(256) k_dust_windows(const u32 *__restrict__ seq, const int64_t *__restrict__ off,
                                                      const int32_t *__restrict__ len, const int64_t *__restrict__ woff,
                                                      int nreads, int64_t nwin_total, int w, int t10, uint8_t *__restrict__ flags) {
    const int64_t gw = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gw >= nwin_total) return;
    int lo = 0, hi = nreads;                                   // last read r with woff[r] <= gw (reads without windows share their successor's offset)
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (woff[mid] <= gw) lo = mid; else hi = mid; }
    const int r = lo;
    const int L = len[r], stride = w >> 1;
    const int q = (int)(gw - woff[r]);
    const int start = q * stride, n = min(w, L - start), l = n - 2;
    uint8_t f = 0;
    if (l >= 14) {
        unsigned char c[64];
#pragma unroll
        for (int i = 0; i < 64; i++) c[i] = 0;
        const int64_t g = off[r] + start;
        int sum = 0, t = 0;
        for (int i = 0; i < n; i++) {
            const int64_t p = g + i;
            const int b = (int)((seq[p >> 4] >> ((p & 15) << 1)) & 3u);
            t = ((t << 2) | b) & 63;
            if (i >= 2) { sum += c[t]; c[t]++; }           // sum_t c_t*(c_t-1)/2 accumulated incrementally
        }
        f = (10 * sum > t10 * (l - 1)) ? 1 : 0;
    }
    flags[gw] = f;
}

}  // namespace
}  // namespace dn

using namespace dn;
using namespace dnapi;

// Device window flags -> merged (begin, end) intervals per read in the track layout of dazzler.d:4943-5052.
// Caller holds g_mu and has reset the arena.
static void dust_intervals(const DevBlock &B, int window, double threshold, int minlen, std::vector<int64_t> &anno, std::vector<int32_t> &iv) {
    const int stride = window / 2, t10 = (int)(threshold * 10.0 + 0.5);
    std::vector<int64_t> woff(B.nreads + 1, 0);
    for (int r = 0; r < B.nreads; r++) woff[r + 1] = woff[r] + (B.h_len[r] > 0 ? (B.h_len[r] + stride - 1) / stride : 0);
    const int64_t nw = woff[B.nreads];
    std::vector<uint8_t> flags(nw + 1);
    if (nw > 0) {
        DBuf<int64_t> dw(B.nreads + 1); DBuf<uint8_t> df(nw);
        DN_CUDA(cudaMemcpyAsync(dw.p, woff.data(), sizeof(int64_t) * (B.nreads + 1), cudaMemcpyHostToDevice, g_stream));
        DN_LAUNCH(k_dust_windows, (unsigned)((nw + 255) / 256), 256, 0, g_stream, (const u32 *)B.fwd.p, (const int64_t *)B.off.p, (const int32_t *)B.len.p,
                  (const int64_t *)dw.p, B.nreads, nw, window, t10, df.p);
        DN_CUDA(cudaMemcpyAsync(flags.data(), df.p, nw, cudaMemcpyDeviceToHost, g_stream));
        DN_CUDA(cudaStreamSynchronize(g_stream));
    }
    anno.assign(B.nreads + 1, 0); iv.clear();
    for (int r = 0; r < B.nreads; r++) {
        anno[r] = 4 * (int64_t)iv.size();
        const int64_t n = woff[r + 1] - woff[r]; const int L = B.h_len[r];
        int64_t q = 0;
        while (q < n) {
            if (!flags[woff[r] + q]) { q++; continue; }
            int64_t e = q;
            while (e + 1 < n && flags[woff[r] + e + 1]) e++;
            int b0 = (int)(q * stride), e0 = (int)std::min<int64_t>(e * stride + window, L);
            if (e0 - b0 >= minlen) { iv.push_back(b0); iv.push_back(e0); }
            q = e + 1;
        }
    }
    anno[B.nreads] = 4 * (int64_t)iv.size();
}

static int check_dust_args(int32_t window) {
    if (window < 16 || window > 64 || (window & 1)) return fail(DN_ERR_INVALID, "DUST window must be even and in [16,64]");
    return DN_OK;
}

extern "C" {

int dn_dust_block(const dn_block *blk, int32_t window, double threshold, int32_t minlen, int64_t **anno, int32_t **data) {
    if (!blk || !anno || !data) return fail(DN_ERR_INVALID, "null argument");
    if (int rc = check_dust_args(window)) return rc;
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = ensure_device()) return rc;
    return guarded([&] {
        cudaSetDevice(g_device); arena().reset();
        const DevBlock &B = blk->b;
        std::vector<int64_t> a; std::vector<int32_t> iv;
        dust_intervals(B, window, threshold, minlen, a, iv);
        int64_t *ha = (int64_t *)hcache_alloc(sizeof(int64_t) * (B.nreads + 1));
        memcpy(ha, a.data(), sizeof(int64_t) * (B.nreads + 1));
        int32_t *hd = (int32_t *)hcache_alloc(sizeof(int32_t) * (iv.size() + 2));
        if (!iv.empty()) memcpy(hd, iv.data(), sizeof(int32_t) * iv.size());
        *anno = ha; *data = hd;
        return DN_OK;
    });
}

int dn_block_mask_dust(dn_block *blk, int32_t window, double threshold, int32_t minlen, int64_t *masked_bases) {
    if (!blk) return fail(DN_ERR_INVALID, "null argument");
    if (int rc = check_dust_args(window)) return rc;
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = ensure_device()) return rc;
    return guarded([&] {
        cudaSetDevice(g_device); arena().reset();
        DevBlock &B = blk->b;
        std::vector<int64_t> a; std::vector<int32_t> iv;
        dust_intervals(B, window, threshold, minlen, a, iv);
        int64_t m = 0;
        for (size_t i = 0; i + 1 < iv.size(); i += 2) m += iv[i + 1] - iv[i];
        if (masked_bases) *masked_bases = m;
        iv.push_back(0); iv.push_back(0);                                   // keeps data() valid for an empty track
        block_add_mask(B, a.data(), iv.data(), g_stream);
        return DN_OK;
    });
}

}  // extern "C"
