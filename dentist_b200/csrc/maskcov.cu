// maskcov.cu -- `dentist mask-repetitive-regions` on the device (mask-stage consumer of the mapping LAS):
// BadAlignmentCoverageAssessor (commands/maskRepetitiveRegions.d:258-420) over the A intervals of the alignment chains
// (:186-205).  The reference sorts 2 events per chain and walks them; coverage is piecewise constant between events, so
// the same mask falls out of a dense formulation that needs no sort:
//   diff[contig slot + abpos] += 1, diff[contig slot + aepos] -= 1     (one thread per chain, atomics)
//   cov = inclusive prefix sum of diff                                    (every contig's diffs cancel -> one global scan)
//   a mask interval starts where the coverage zone turns bad (< lower or > upper), ends where it turns ok or the contig ends
// Specification = oracle/mask_oracle.py, pinned by the reference's own vectors (maskRepetitiveRegions.d:395-411, 582-617).
// HBM-bound streaming: 4 B read + 4 B written per reference base and pass.
#include "api_internal.hpp"
#include <string.h>
#include <algorithm>
#include <vector>

namespace dn {
namespace {

// slot layout: contig c owns positions 0..len at slots coff[c] .. coff[c] + len
__global__ void __launch_bounds__(256) k_cov_edges(const int64_t *__restrict__ coff, int na, uint8_t *__restrict__ edge) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= na) return;
    edge[coff[c]] |= 1; edge[coff[c + 1] - 1] |= 2;                // a zero-length contig has both on its only slot
}

__global__ void __launch_bounds__(256) k_cov_events(const dn_las_record *__restrict__ rec, int64_t n, const int32_t *__restrict__ alen,
                                                    const int32_t *__restrict__ blen, const int64_t *__restrict__ coff,
                                                    int improper_only, int allowance, int32_t *__restrict__ diff, int32_t *__restrict__ nsel) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const dn_las_record f = rec[i];
    if (f.flags & DN_LAS_NEXT) return;                             // chains are walked from their first record (dazzler.d:708-743)
    int64_t j = i;
    if (f.flags & (DN_LAS_START | DN_LAS_BEST)) while (j + 1 < n && (rec[j + 1].flags & DN_LAS_NEXT)) j++;
    const dn_las_record l = rec[j];
    if (improper_only) {                                           // !isProper(allowance), base.d:537-557
        const bool proper = (f.abpos <= allowance || f.bbpos <= allowance) &&
                            (l.aepos + allowance >= alen[f.aread] || l.bepos + allowance >= blen[f.bread]);
        if (proper) return;
    }
    atomicAdd(nsel, 1);
    atomicAdd(&diff[coff[f.aread] + f.abpos], 1);
    atomicAdd(&diff[coff[f.aread] + l.aepos], -1);
}

__device__ __forceinline__ bool bad_zone(int cov, double lo, double hi) { return (double)cov < lo || (double)cov > hi; }

// excl[g] = coverage left of position p (= on [p-1, p)), excl[g] + diff[g] = coverage on [p, p+1)
__global__ void __launch_bounds__(256) k_cov_flags(const int32_t *__restrict__ diff, const int32_t *__restrict__ excl,
                                                   const uint8_t *__restrict__ edge, int64_t nslots, double lo, double hi,
                                                   int32_t *__restrict__ sflag, int32_t *__restrict__ eflag) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nslots) return;
    const uint8_t e = edge[g];
    const bool first = e & 1, last = e & 2;
    const bool before = !first && bad_zone(excl[g], lo, hi);
    const bool here = !last && bad_zone(excl[g] + diff[g], lo, hi);
    sflag[g] = (here && !before) ? 1 : 0;
    eflag[g] = (before && !here) ? 1 : 0;
}

__global__ void __launch_bounds__(256) k_cov_scatter(const int32_t *__restrict__ sflag, const int32_t *__restrict__ sidx,
                                                     const int32_t *__restrict__ eflag, const int32_t *__restrict__ eidx,
                                                     int64_t nslots, int64_t *__restrict__ starts, int64_t *__restrict__ ends) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nslots) return;
    if (sflag[g]) starts[sidx[g]] = g;
    if (eflag[g]) ends[eidx[g]] = g;
}

// Trace.tracePointsUpTo!"contigA" (base.d:211-242) + the prefix sum of translateTracePoint (:185-203): the B coordinate
// the trace assigns to A position `pos` rounded down (ceil = 0) or up (ceil = 1) to a tile boundary.
__device__ int translate_b(const dn_las_record &r, const uint16_t *__restrict__ tr, int ts, int pos, int ceil_mode) {
    const int nt = r.tlen >> 1;
    const int second = (r.abpos / ts) * ts + ts;
    int idx;
    if (!ceil_mode) idx = pos < second ? 0 : (pos < r.aepos ? 1 + (pos - second) / ts : nt);
    else {
        const int second_last = ((r.aepos - 1) / ts) * ts;
        idx = pos == r.abpos ? 0 : (pos <= second ? 1 : (pos <= second_last ? 1 + (pos - second + ts - 1) / ts : nt));
    }
    int b = r.bbpos;
    for (int t = 0; t < idx; t++) b += tr[2 * t + 1];
    return b;
}

// one thread per local alignment: the mask intervals of its A contig that it overlaps, carried over to the B read
// (propagateMaskPerContig / propagateIntervals, commands/propagateMask.d:147-292) as +1/-1 coverage events
__global__ void __launch_bounds__(128) k_propagate(const dn_las_record *__restrict__ rec, const int64_t *__restrict__ toff,
                                                   const uint16_t *__restrict__ trace, int64_t n, int ts,
                                                   const int64_t *__restrict__ manno, const int32_t *__restrict__ mdata,
                                                   const int32_t *__restrict__ blen, const int64_t *__restrict__ boff,
                                                   int32_t *__restrict__ diff, int32_t *__restrict__ nsel) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const dn_las_record r = rec[i];
    const int64_t m0 = manno[r.aread] >> 3, m1 = manno[r.aread + 1] >> 3;      // (begin, end) pairs of this contig
    int64_t lo = m0, hi = m1;
    while (lo < hi) { const int64_t m = (lo + hi) >> 1; if (mdata[2 * m + 1] <= r.abpos) lo = m + 1; else hi = m; }   // first mask end > abpos
    const uint16_t *tr = trace + toff[i];
    const int L = blen[r.bread];
    for (int64_t m = lo; m < m1 && mdata[2 * m] < r.aepos; m++) {
        const int mb = max(mdata[2 * m], r.abpos), me = min(mdata[2 * m + 1], r.aepos);   // only the first / last interval can stick out
        int pb = translate_b(r, tr, ts, mb, 0), pe = translate_b(r, tr, ts, me, 1);
        if (r.flags & DN_LAS_COMP) { const int t = L - pe; pe = L - pb; pb = t; }
        if (pb < 0 || pe > L || pb >= pe) continue;
        atomicAdd(nsel, 1);
        atomicAdd(&diff[boff[r.bread] + pb], 1);
        atomicAdd(&diff[boff[r.bread] + pe], -1);
    }
}

}  // namespace
}  // namespace dn

using namespace dn;
using namespace dnapi;

// Shared tail: coverage events in `diff` (slot layout coff, len+1 slots per sequence) -> intervals whose coverage is
// < lower or > upper, in the mask-track layout.  Caller holds g_mu; `nsel` = device counter of contributing events.
static void dense_mask(DBuf<int32_t> &diff, const std::vector<int64_t> &coff, const int64_t *dcoff, int nseq, double lower, double upper,
                       const int32_t *nsel_dev, int64_t *ha, int32_t **data, cudaStream_t s) {
    const int64_t nslots = coff[nseq];
    DBuf<int32_t> excl(nslots), sflag(nslots), eflag(nslots), sidx(nslots), eidx(nslots), tot(2);
    DBuf<uint8_t> edge(nslots);
    edge.zero(s);
    DN_LAUNCH(k_cov_edges, (nseq + 255) / 256, 256, 0, s, dcoff, nseq, edge.p);
    exclusive_scan_i32(diff.p, excl.p, nslots, tot.p, s);
    DN_LAUNCH(k_cov_flags, (unsigned)((nslots + 255) / 256), 256, 0, s, (const int32_t *)diff.p, (const int32_t *)excl.p,
              (const uint8_t *)edge.p, nslots, lower, upper, sflag.p, eflag.p);
    exclusive_scan_i32(sflag.p, sidx.p, nslots, tot.p, s);
    exclusive_scan_i32(eflag.p, eidx.p, nslots, tot.p + 1, s);
    int32_t cnt[2] = {0, 0}, nsel = 0;
    DN_CUDA(cudaMemcpyAsync(cnt, tot.p, 8, cudaMemcpyDeviceToHost, s));
    DN_CUDA(cudaMemcpyAsync(&nsel, nsel_dev, 4, cudaMemcpyDeviceToHost, s));
    DN_CUDA(cudaStreamSynchronize(s));
    if (cnt[0] != cnt[1]) throw Error("coverage mask: unbalanced interval edges");
    const int32_t m = nsel == 0 ? 0 : cnt[0];                      // nothing selected: empty region (maskRepetitiveRegions.d:349-350)
    std::vector<int64_t> hs(m), he(m);
    if (m > 0) {
        DBuf<int64_t> ds(m), de(m);
        DN_LAUNCH(k_cov_scatter, (unsigned)((nslots + 255) / 256), 256, 0, s, (const int32_t *)sflag.p, (const int32_t *)sidx.p,
                  (const int32_t *)eflag.p, (const int32_t *)eidx.p, nslots, ds.p, de.p);
        DN_CUDA(cudaMemcpyAsync(hs.data(), ds.p, sizeof(int64_t) * m, cudaMemcpyDeviceToHost, s));
        DN_CUDA(cudaMemcpyAsync(he.data(), de.p, sizeof(int64_t) * m, cudaMemcpyDeviceToHost, s));
        DN_CUDA(cudaStreamSynchronize(s));
    }
    // slots -> (sequence, begin, end) in the mask-track layout (dazzler.d:4943-5052); intervals arrive sorted
    int32_t *hd = (int32_t *)hcache_alloc(sizeof(int32_t) * (2 * (size_t)m + 2));
    int c = 0;
    ha[0] = 0;
    for (int32_t k = 0; k < m; k++) {
        while (coff[c + 1] <= hs[k]) { c++; ha[c] = 8 * (int64_t)k; }
        hd[2 * k] = (int32_t)(hs[k] - coff[c]); hd[2 * k + 1] = (int32_t)(he[k] - coff[c]);
    }
    while (c < nseq) { c++; ha[c] = 8 * (int64_t)m; }
    *data = hd;
}

extern "C" int dn_mask_coverage(const dn_las_buf *las, const int32_t *alen, int32_t na, const int32_t *blen, int32_t nb,
                                double lower, double upper, int32_t improper_only, int32_t allowance,
                                int64_t **anno, int32_t **data) {
    if (!las || !alen || !blen || !anno || !data || na < 0 || nb < 0) return fail(DN_ERR_INVALID, "null argument");
    for (int64_t i = 0; i < las->nrec; i++) {
        const dn_las_record &r = las->rec[i];
        if (r.aread < 0 || r.aread >= na || r.bread < 0 || r.bread >= nb) return fail(DN_ERR_INVALID, "contig id out of bounds");
        if (r.abpos < 0 || r.aepos > alen[r.aread] || r.abpos > r.aepos) return fail(DN_ERR_INVALID, "alignment outside its contig");
    }
    std::vector<int64_t> coff(na + 1, 0);
    for (int c = 0; c < na; c++) { if (alen[c] < 0) return fail(DN_ERR_INVALID, "negative contig length"); coff[c + 1] = coff[c] + alen[c] + 1; }
    const int64_t nslots = coff[na], n = las->nrec;
    if (nslots >= (int64_t)1 << 31) return fail(DN_ERR_INVALID, "reference block too large for one coverage pass (split it)");
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = ensure_device()) return rc;
    return guarded([&]() -> int {
        cudaSetDevice(g_device); arena().reset();
        cudaStream_t s = g_stream;
        int64_t *ha = (int64_t *)hcache_alloc(sizeof(int64_t) * ((size_t)na + 1));
        memset(ha, 0, sizeof(int64_t) * ((size_t)na + 1));
        if (n == 0 || nslots == 0) {                               // "if (alignmentIntervals.empty) return ReferenceRegion()" :349-350
            *anno = ha; *data = (int32_t *)hcache_alloc(64); return DN_OK;
        }
        DBuf<dn_las_record> drec(n); DBuf<int32_t> dal(na), dbl(nb), diff(nslots), nsel(1);
        DBuf<int64_t> dcoff(na + 1);
        DN_CUDA(cudaMemcpyAsync(drec.p, las->rec, sizeof(dn_las_record) * n, cudaMemcpyHostToDevice, s));
        DN_CUDA(cudaMemcpyAsync(dal.p, alen, sizeof(int32_t) * na, cudaMemcpyHostToDevice, s));
        DN_CUDA(cudaMemcpyAsync(dbl.p, blen, sizeof(int32_t) * nb, cudaMemcpyHostToDevice, s));
        DN_CUDA(cudaMemcpyAsync(dcoff.p, coff.data(), sizeof(int64_t) * (na + 1), cudaMemcpyHostToDevice, s));
        diff.zero(s); nsel.zero(s);
        DN_LAUNCH(k_cov_events, (unsigned)((n + 255) / 256), 256, 0, s, (const dn_las_record *)drec.p, n, (const int32_t *)dal.p,
                  (const int32_t *)dbl.p, (const int64_t *)dcoff.p, improper_only, allowance, diff.p, nsel.p);
        int32_t *hd = nullptr;
        dense_mask(diff, coff, dcoff.p, na, lower, upper, nsel.p, ha, &hd, s);
        *anno = ha; *data = hd;
        return DN_OK;
    });
}

extern "C" int dn_propagate_mask(const dn_las_buf *las, int32_t na, const int64_t *mask_anno, const int32_t *mask_data,
                                 const int32_t *blen, int32_t nb, int64_t **anno, int32_t **data) {
    if (!las || !mask_anno || !mask_data || !blen || !anno || !data || na < 0 || nb < 0) return fail(DN_ERR_INVALID, "null argument");
    if (las->tspace < 1) return fail(DN_ERR_INVALID, "LAS without trace spacing");
    for (int64_t i = 0; i < las->nrec; i++) {
        const dn_las_record &r = las->rec[i];
        if (r.aread < 0 || r.aread >= na || r.bread < 0 || r.bread >= nb) return fail(DN_ERR_INVALID, "contig id out of bounds");
        if (r.bbpos < 0 || r.bepos > blen[r.bread] || r.abpos > r.aepos) return fail(DN_ERR_INVALID, "alignment outside its read");
        const int nt = r.aepos > r.abpos ? (r.aepos - 1) / las->tspace - r.abpos / las->tspace + 1 : 0;
        if (r.tlen != 2 * nt) return fail(DN_ERR_INVALID, "trace length does not match the alignment (propagation needs trace points)");
    }
    std::vector<int64_t> boff(nb + 1, 0);
    for (int c = 0; c < nb; c++) { if (blen[c] < 0) return fail(DN_ERR_INVALID, "negative read length"); boff[c + 1] = boff[c] + blen[c] + 1; }
    const int64_t nslots = boff[nb], n = las->nrec;
    if (nslots >= (int64_t)1 << 31) return fail(DN_ERR_INVALID, "read block too large for one propagation pass (split it)");
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = ensure_device()) return rc;
    return guarded([&]() -> int {
        cudaSetDevice(g_device); arena().reset();
        cudaStream_t s = g_stream;
        int64_t *ha = (int64_t *)hcache_alloc(sizeof(int64_t) * ((size_t)nb + 1));
        memset(ha, 0, sizeof(int64_t) * ((size_t)nb + 1));
        const int64_t mbytes = mask_anno[na];
        if (n == 0 || nslots == 0 || mbytes <= 0) { *anno = ha; *data = (int32_t *)hcache_alloc(64); return DN_OK; }
        DBuf<dn_las_record> drec(n); DBuf<int64_t> dtoff(n), dman(na + 1), dboff(nb + 1); DBuf<uint16_t> dtr(las->ntrace + 2);
        DBuf<int32_t> dmd(mbytes / 4 + 2), dbl(nb), diff(nslots), nsel(1);
        DN_CUDA(cudaMemcpyAsync(drec.p, las->rec, sizeof(dn_las_record) * n, cudaMemcpyHostToDevice, s));
        DN_CUDA(cudaMemcpyAsync(dtoff.p, las->toff, sizeof(int64_t) * n, cudaMemcpyHostToDevice, s));
        if (las->ntrace > 0) DN_CUDA(cudaMemcpyAsync(dtr.p, las->trace, sizeof(uint16_t) * las->ntrace, cudaMemcpyHostToDevice, s));
        DN_CUDA(cudaMemcpyAsync(dman.p, mask_anno, sizeof(int64_t) * (na + 1), cudaMemcpyHostToDevice, s));
        DN_CUDA(cudaMemcpyAsync(dmd.p, mask_data, mbytes, cudaMemcpyHostToDevice, s));
        DN_CUDA(cudaMemcpyAsync(dbl.p, blen, sizeof(int32_t) * nb, cudaMemcpyHostToDevice, s));
        DN_CUDA(cudaMemcpyAsync(dboff.p, boff.data(), sizeof(int64_t) * (nb + 1), cudaMemcpyHostToDevice, s));
        diff.zero(s); nsel.zero(s);
        DN_LAUNCH(k_propagate, (unsigned)((n + 127) / 128), 128, 0, s, (const dn_las_record *)drec.p, (const int64_t *)dtoff.p,
                  (const uint16_t *)dtr.p, n, las->tspace, (const int64_t *)dman.p, (const int32_t *)dmd.p, (const int32_t *)dbl.p,
                  (const int64_t *)dboff.p, diff.p, nsel.p);
        int32_t *hd = nullptr;
        dense_mask(diff, boff, dboff.p, nb, 0.0, 0.0, nsel.p, ha, &hd, s);      // masked = covered by >= 1 propagated interval
        *anno = ha; *data = hd;
        return DN_OK;
    });
}
