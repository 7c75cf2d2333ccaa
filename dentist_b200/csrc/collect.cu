// collect.cu -- the alignment filters of collectPileUps on the device:
// LQ, Improper, WeaklyAnchored, Contained, Ambiguous, Redundant (commands/collectPileUps/filter.d:122-356, applied
// in the order of collectPileUps/package.d:129-141) over the AlignmentChains of a chained ref-vs-reads LAS.
// Specification = oracle/collect_filters.py.  Output: one status byte per chain (0 kept, 1 LQ, 2 improper,
// 3 weakly anchored, 4 contained, 5 ambiguous read, 6 redundant read, 7 disabled on input) + "read used" flags.
#include "api_internal.hpp"
#include <string.h>
#include <vector>

namespace dn {
namespace {

struct ChainSum { int32_t first, a, b, comp, fab, fbb, lae, lbe, cov_a, diffs, uniq, disabled; };

__device__ __forceinline__ bool is_start(const dn_las_record &r) { return !(r.flags & DN_LAS_NEXT); }

__global__ void __launch_bounds__(256) k_chain_starts(const dn_las_record *__restrict__ rec, int64_t n, int32_t *__restrict__ sflag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sflag[i] = is_start(rec[i]) ? 1 : 0;
}

// one thread per chain start: walk the NEXT records, union of the A intervals minus the repeat mask
__global__ void __launch_bounds__(128) k_chain_summary(const dn_las_record *__restrict__ rec, int64_t n, const int32_t *__restrict__ sflag,
                                                       const int32_t *__restrict__ sidx, const int64_t *__restrict__ manno,
                                                       const int32_t *__restrict__ mdata, ChainSum *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !sflag[i]) return;
    ChainSum c;
    const dn_las_record f = rec[i];
    c.first = (int32_t)i; c.a = f.aread; c.b = f.bread; c.comp = f.flags & DN_LAS_COMP; c.fab = f.abpos; c.fbb = f.bbpos;
    c.disabled = (f.flags & DN_LAS_ELIM) ? 1 : 0;
    int cov = 0, diffs = 0, uniq = 0;
    int us = f.abpos, ue = f.abpos;                       // current run of the union of A intervals (abpos ascends along a chain)
    int64_t j = i;
    dn_las_record l = f;
    auto close_run = [&]() {
        int covd = 0;
        if (manno) {
            const int64_t m0 = manno[c.a] / 8, m1 = manno[c.a + 1] / 8;
            int64_t lo = m0, hi = m1;                     // first mask interval with end > us
            while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (mdata[2 * mid + 1] <= us) lo = mid + 1; else hi = mid; }
            for (int64_t q = lo; q < m1 && mdata[2 * q] < ue; q++) covd += max(0, min(ue, mdata[2 * q + 1]) - max(us, mdata[2 * q]));
        }
        uniq += (ue - us) - covd;
    };
    for (;;) {
        cov += l.aepos - l.abpos; diffs += l.diffs;
        if (l.abpos <= ue) ue = max(ue, l.aepos); else { close_run(); us = l.abpos; ue = l.aepos; }
        if (j + 1 < n && (f.flags & (DN_LAS_START | DN_LAS_BEST)) && (rec[j + 1].flags & DN_LAS_NEXT)) { j++; l = rec[j]; } else break;
    }
    close_run();
    c.lae = l.aepos; c.lbe = l.bepos; c.cov_a = cov; c.diffs = diffs; c.uniq = uniq;
    out[sidx[i]] = c;
}

__global__ void __launch_bounds__(256) k_collect_basic(const ChainSum *__restrict__ ch, int nc, const int32_t *__restrict__ alen,
                                                       const int32_t *__restrict__ blen, double max_err, int allowance, int min_anchor,
                                                       uint8_t *__restrict__ st) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nc) return;
    const ChainSum c = ch[i];
    uint8_t s = 0;
    if (c.disabled) s = 7;
    else if ((double)c.diffs / (double)c.cov_a > max_err) s = 1;
    else if (!((c.fab <= allowance || c.fbb <= allowance) && (c.lae + allowance >= alen[c.a] || c.lbe + allowance >= blen[c.b]))) s = 2;
    else if (c.uniq <= min_anchor) s = 3;
    st[i] = s;
}

__global__ void __launch_bounds__(256) k_collect_setkey(const ChainSum *__restrict__ ch, ulonglong2 *__restrict__ items, int nc, int field) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nc) return;
    const u64 idx = field == 0 ? (u64)i : items[i].y;
    const ChainSum c = ch[idx];
    u64 key;
    switch (field) {
        case 0: key = (u32)c.lbe; break;
        case 1: key = (u32)c.lae; break;
        case 2: key = (u32)c.fbb; break;
        case 3: key = (u32)c.fab; break;
        case 4: key = ((u64)(u32)c.a << 32) | (u32)c.b; break;
        default: key = (u32)c.b; break;                    // field 5: by read only (stable on the index order)
    }
    items[i] = make_ulonglong2(key, idx);
}

__device__ __forceinline__ void b_interval(const ChainSum &c, int bl, int &s, int &e) {
    if (c.comp) { s = bl - c.lbe; e = bl - c.fbb; } else { s = c.fbb; e = c.lbe; }
}

// one thread per chain (as the containing one), `order` = chains in AlignmentChain.opCmp order
__global__ void __launch_bounds__(256) k_collect_contained(const ChainSum *__restrict__ ch, const ulonglong2 *__restrict__ order, int nc,
                                                           const int32_t *__restrict__ blen, const uint8_t *__restrict__ st_in,
                                                           uint8_t *__restrict__ contained) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nc) return;
    const int i = (int)order[p].y;
    if (st_in[i]) return;
    const ChainSum c1 = ch[i];
    int s1, e1; b_interval(c1, blen[c1.b], s1, e1);
    for (int q = p + 1; q < nc; q++) {
        const int j = (int)order[q].y;
        const ChainSum c2 = ch[j];
        if (!(c2.a == c1.a && c1.fab <= c2.fab && c2.lae <= c1.lae)) break;
        if (c2.comp != c1.comp || c2.b != c1.b) continue;
        int s2, e2; b_interval(c2, blen[c2.b], s2, e2);
        if (s1 <= s2 && e2 <= e1) contained[j] = 1;
    }
}

// one thread per read segment of `byread` (chains sorted by read): ambiguous, then redundant
__global__ void __launch_bounds__(128) k_collect_reads(const ChainSum *__restrict__ ch, const ulonglong2 *__restrict__ byread, int nc,
                                                       const int32_t *__restrict__ alen, const int32_t *__restrict__ blen,
                                                       uint8_t *__restrict__ st, uint8_t *__restrict__ used) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nc) return;
    const int b = (int)byread[p].x;
    if (p > 0 && (int)byread[p - 1].x == b) return;       // not the first chain of this read
    int e = p; while (e < nc && (int)byread[e].x == b) e++;
    const int bl = blen[b];
    bool amb = false;
    for (int x = p; x < e && !amb; x++) {
        const int i = (int)byread[x].y; if (st[i]) continue;
        int s1, e1; b_interval(ch[i], bl, s1, e1);
        for (int y = x + 1; y < e; y++) {
            const int j = (int)byread[y].y; if (st[j]) continue;
            int s2, e2; b_interval(ch[j], bl, s2, e2);
            if (max(s1, s2) < min(e1, e2)) { amb = true; break; }
        }
    }
    if (amb) { used[b] = 1; for (int x = p; x < e; x++) { const int i = (int)byread[x].y; if (!st[i]) st[i] = 5; } return; }
    bool red = false;
    for (int x = p; x < e; x++) {
        const int i = (int)byread[x].y; if (st[i]) continue;
        const ChainSum c = ch[i];
        if (c.fbb <= c.fab && c.lae + bl - c.lbe < alen[c.a]) { red = true; break; }
    }
    if (red) { used[b] = 1; for (int x = p; x < e; x++) { const int i = (int)byread[x].y; if (!st[i]) st[i] = 6; } }
}

__global__ void __launch_bounds__(256) k_collect_merge(const uint8_t *__restrict__ contained, ChainSum *__restrict__ ch, int nc,
                                                       uint8_t *__restrict__ st, int32_t *__restrict__ first) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nc) return;
    if (!st[i] && contained[i]) st[i] = 4;
    first[i] = ch[i].first;
}

int bitsof(uint64_t v) { int b = 0; while (v) { b++; v >>= 1; } return b; }

}  // namespace
}  // namespace dn

using namespace dn;
using namespace dnapi;

extern "C" int dn_collect_filter(const dn_las_buf *las, const int32_t *alen, int32_t na, const int32_t *blen, int32_t nb,
                                 const int64_t *mask_anno, const int32_t *mask_data, double max_err, int32_t allowance,
                                 int32_t min_anchor, int64_t *nchains, int32_t **chain_first, uint8_t **chain_status, uint8_t **read_used) {
    if (!las || !alen || !blen || !nchains || !chain_first || !chain_status || !read_used) return fail(DN_ERR_INVALID, "null argument");
    int64_t maxc = 1;
    for (int64_t i = 0; i < las->nrec; i++) {
        const dn_las_record &r = las->rec[i];
        if (r.aread < 0 || r.aread >= na || r.bread < 0 || r.bread >= nb) return fail(DN_ERR_INVALID, "contig id out of bounds");
        if (i == 0 && (r.flags & DN_LAS_NEXT)) return fail(DN_ERR_INVALID, "chain is missing a start");         // dazzler.d:739
        maxc = std::max<int64_t>(maxc, std::max(r.aepos, r.bepos));
    }
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = ensure_device()) return rc;
    return guarded([&]() -> int {
        cudaSetDevice(g_device); arena().reset();
        cudaStream_t s = g_stream;
        const int64_t n = las->nrec;
        uint8_t *hused = (uint8_t *)hcache_alloc((size_t)nb + 1); memset(hused, 0, (size_t)nb + 1);
        if (n == 0) { *nchains = 0; *chain_first = (int32_t *)hcache_alloc(64); *chain_status = (uint8_t *)hcache_alloc(64); *read_used = hused; return DN_OK; }
        DBuf<dn_las_record> drec(n); DBuf<int32_t> sflag(n), sidx(n), tot(1), dal(na), dbl(nb);
        DN_CUDA(cudaMemcpyAsync(drec.p, las->rec, sizeof(dn_las_record) * n, cudaMemcpyHostToDevice, s));
        DN_CUDA(cudaMemcpyAsync(dal.p, alen, sizeof(int32_t) * na, cudaMemcpyHostToDevice, s));
        DN_CUDA(cudaMemcpyAsync(dbl.p, blen, sizeof(int32_t) * nb, cudaMemcpyHostToDevice, s));
        DBuf<int64_t> dman; DBuf<int32_t> dmd;
        const bool masked = mask_anno && mask_data && mask_anno[na] > 0;
        if (masked) {
            dman.alloc(na + 1); dmd.alloc(mask_anno[na] / 4 + 2);
            DN_CUDA(cudaMemcpyAsync(dman.p, mask_anno, sizeof(int64_t) * (na + 1), cudaMemcpyHostToDevice, s));
            DN_CUDA(cudaMemcpyAsync(dmd.p, mask_data, mask_anno[na], cudaMemcpyHostToDevice, s));
        }
        DN_LAUNCH(k_chain_starts, (unsigned)((n + 255) / 256), 256, 0, s, (const dn_las_record *)drec.p, n, sflag.p);
        exclusive_scan_i32(sflag.p, sidx.p, n, tot.p, s);
        int32_t nc = 0; DN_CUDA(cudaMemcpyAsync(&nc, tot.p, 4, cudaMemcpyDeviceToHost, s)); DN_CUDA(cudaStreamSynchronize(s));
        DBuf<ChainSum> ch(nc); DBuf<uint8_t> st(nc), cont(nc), used((size_t)nb + 1); DBuf<int32_t> first(nc);
        cont.zero(s); used.zero(s);
        DN_LAUNCH(k_chain_summary, (unsigned)((n + 127) / 128), 128, 0, s, (const dn_las_record *)drec.p, n, (const int32_t *)sflag.p,
                  (const int32_t *)sidx.p, masked ? (const int64_t *)dman.p : nullptr, masked ? (const int32_t *)dmd.p : nullptr, ch.p);
        DN_LAUNCH(k_collect_basic, (nc + 255) / 256, 256, 0, s, (const ChainSum *)ch.p, nc, (const int32_t *)dal.p, (const int32_t *)dbl.p,
                  max_err, allowance, min_anchor, st.p);
        // AlignmentChain.opCmp order (base.d:766-777) by five stable LSD sorts; then the by-read order
        DBuf<ulonglong2> i1(nc), i2(nc);
        ulonglong2 *cur = i1.p, *oth = i2.p;
        const int cb = bitsof((uint64_t)maxc), fb[5] = {cb, cb, cb, cb, 32 + bitsof((uint64_t)na)};
        for (int f = 0; f < 5; f++) {
            DN_LAUNCH(k_collect_setkey, (nc + 255) / 256, 256, 0, s, (const ChainSum *)ch.p, cur, nc, f);
            ulonglong2 *r = radix_sort_rec16(cur, oth, nc, 0, 0, fb[f], s);
            if (r != cur) { oth = cur; cur = r; }
        }
        DN_LAUNCH(k_collect_contained, (nc + 255) / 256, 256, 0, s, (const ChainSum *)ch.p, (const ulonglong2 *)cur, nc,
                  (const int32_t *)dbl.p, (const uint8_t *)st.p, cont.p);
        DN_LAUNCH(k_collect_merge, (nc + 255) / 256, 256, 0, s, (const uint8_t *)cont.p, ch.p, nc, st.p, first.p);
        DN_LAUNCH(k_collect_setkey, (nc + 255) / 256, 256, 0, s, (const ChainSum *)ch.p, oth, nc, 0);      // fresh identity order ...
        DN_LAUNCH(k_collect_setkey, (nc + 255) / 256, 256, 0, s, (const ChainSum *)ch.p, oth, nc, 5);      // ... keyed by read
        ulonglong2 *byread = radix_sort_rec16(oth, cur, nc, 0, 0, bitsof((uint64_t)nb), s);
        DN_LAUNCH(k_collect_reads, (nc + 127) / 128, 128, 0, s, (const ChainSum *)ch.p, (const ulonglong2 *)byread, nc,
                  (const int32_t *)dal.p, (const int32_t *)dbl.p, st.p, used.p);
        int32_t *hf = (int32_t *)hcache_alloc(sizeof(int32_t) * ((size_t)nc + 1));
        uint8_t *hs = (uint8_t *)hcache_alloc((size_t)nc + 1);
        DN_CUDA(cudaMemcpyAsync(hf, first.p, sizeof(int32_t) * nc, cudaMemcpyDeviceToHost, s));
        DN_CUDA(cudaMemcpyAsync(hs, st.p, nc, cudaMemcpyDeviceToHost, s));
        DN_CUDA(cudaMemcpyAsync(hused, used.p, nb, cudaMemcpyDeviceToHost, s));
        DN_CUDA(cudaStreamSynchronize(s));
        *nchains = nc; *chain_first = hf; *chain_status = hs; *read_used = hused;
        return DN_OK;
    });
}
