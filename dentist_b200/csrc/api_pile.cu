// api_pile.cu -- C ABI of the per-pile stages: LAS filters, intrinsic QVs, consensus.
#include "api_internal.hpp"
#include "pile.cuh"
#include <string.h>
#include <stdlib.h>
#include <algorithm>

using namespace dn;
using namespace dnapi;

namespace {

struct DevLasIn {               // a host dn_las_buf mirrored in the arena
    DBuf<dn_las_record> rec; DBuf<int64_t> toff; DBuf<uint16_t> trace;
    void upload(const dn_las_buf *l, bool with_trace, cudaStream_t s) {
        rec.alloc(l->nrec + 1); toff.alloc(l->nrec + 1);
        if (l->nrec) {
            DN_CUDA(cudaMemcpyAsync(rec.p, l->rec, sizeof(dn_las_record) * l->nrec, cudaMemcpyHostToDevice, s));
            DN_CUDA(cudaMemcpyAsync(toff.p, l->toff, sizeof(int64_t) * l->nrec, cudaMemcpyHostToDevice, s));
        }
        if (with_trace) {
            trace.alloc(l->ntrace + 2);
            if (l->ntrace) DN_CUDA(cudaMemcpyAsync(trace.p, l->trace, sizeof(uint16_t) * l->ntrace, cudaMemcpyHostToDevice, s));
        }
    }
};

template <typename T> T *to_device(DBuf<T> &d, const T *h, size_t n, cudaStream_t s) {
    d.alloc(n + 1);
    if (n) DN_CUDA(cudaMemcpyAsync(d.p, h, sizeof(T) * n, cudaMemcpyHostToDevice, s));
    return d.p;
}

// Records of a caller-supplied LAS are checked on the host before a kernel indexes reads, traces or vote columns by
// them: ids, coordinates against the read lengths, tile count against the span, trace extent against the buffer.
const char *validate_las(const dn_las_buf *l, const int32_t *alen, int64_t na, const int32_t *blen, int64_t nb, bool with_trace) {
    if (dnapi::g_trusted_las) return nullptr;
    const int ts = l->tspace;
    for (int64_t i = 0; i < l->nrec; i++) {
        const dn_las_record &r = l->rec[i];
        if (r.aread < 0 || r.aread >= na || r.bread < 0 || r.bread >= nb) return "contig id out of bounds";
        if (r.abpos < 0 || r.abpos > r.aepos || r.aepos > alen[r.aread]) return "A coordinates outside the read";
        if (r.bbpos < 0 || r.bbpos > r.bepos || r.bepos > blen[r.bread]) return "B coordinates outside the read";
        if (!with_trace) continue;
        const int nt = r.aepos > r.abpos ? (r.aepos + ts - 1) / ts - r.abpos / ts : 0;
        if (r.tlen != 2 * nt) return "tlen does not match the number of trace tiles";
        if (l->toff[i] < 0 || l->toff[i] + r.tlen > l->ntrace) return "trace outside the trace buffer";
        int64_t sb = 0; const uint16_t *t = l->trace + l->toff[i];
        for (int q = 0; q < nt; q++) sb += t[2 * q + 1];
        if (sb != r.bepos - r.bbpos) return "trace B bases do not add up to the B span";
    }
    return nullptr;
}

int filter_common(dn_las_buf *las, int mode, double max_err, const int32_t *alen, int32_t na, const int32_t *blen, int32_t nb, int allowance) {
    if (!las) return fail(DN_ERR_INVALID, "null argument");
    if (mode == 1) {
        if (!alen || !blen) return fail(DN_ERR_INVALID, "null read lengths");
        for (int64_t i = 0; i < las->nrec; i++)
            if (las->rec[i].aread < 0 || las->rec[i].aread >= na || las->rec[i].bread < 0 || las->rec[i].bread >= nb)
                return fail(DN_ERR_INVALID, "contig id out of bounds");            // dazzler.d:1765-1778
    }
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = ensure_device()) return rc;
    return guarded([&] {
        cudaSetDevice(g_device); arena().reset();
        DevLasIn d; d.upload(las, false, g_stream);
        DBuf<int32_t> da, db; 
        if (mode == 1) { to_device(da, alen, na, g_stream); to_device(db, blen, nb, g_stream); }
        DBuf<dn_las_record> orec(las->nrec + 1); DBuf<int64_t> otoff(las->nrec + 1);
        int64_t n_out = 0;
        las_filter_device(d.rec.p, d.toff.p, las->nrec, mode, max_err, da.p, db.p, allowance, orec.p, otoff.p, &n_out, g_stream);
        if (n_out) {
            DN_CUDA(cudaMemcpyAsync(las->rec, orec.p, sizeof(dn_las_record) * n_out, cudaMemcpyDeviceToHost, g_stream));
            DN_CUDA(cudaMemcpyAsync(las->toff, otoff.p, sizeof(int64_t) * n_out, cudaMemcpyDeviceToHost, g_stream));
        }
        DN_CUDA(cudaStreamSynchronize(g_stream));
        las->nrec = n_out;
        return DN_OK;
    });
}
}  // namespace

extern "C" {

void dn_free(void *p) { hcache_free(p); }

int dn_las_filter_error(dn_las_buf *las, double max_err) { return filter_common(las, 0, max_err, nullptr, 0, nullptr, 0, 0); }

int dn_las_filter_pileup(dn_las_buf *las, const int32_t *alen, int32_t na, const int32_t *blen, int32_t nb, int32_t allowance) {
    return filter_common(las, 1, 0.0, alen, na, blen, nb, allowance);
}

int dn_compute_qvs(const int32_t *rlen, int32_t nreads, const dn_las_buf *las, int32_t coverage, uint8_t **qv, int64_t **qoff) {
    return dn_compute_qvs_v(rlen, nreads, las, coverage, nullptr, qv, qoff);
}

int dn_compute_qvs_v(const int32_t *rlen, int32_t nreads, const dn_las_buf *las, int32_t coverage, const int32_t *cov_per_read,
                     uint8_t **qv, int64_t **qoff) {
    if (!rlen || !las || !qv || !qoff || nreads < 0) return fail(DN_ERR_INVALID, "null argument");
    if (las->tspace < 1) return fail(DN_ERR_INVALID, "bad trace spacing");
    if (const char *bad = validate_las(las, rlen, nreads, rlen, nreads, true)) return fail(DN_ERR_INVALID, bad);
    for (int64_t i = 1; i < las->nrec; i++)
        if (las->rec[i].aread < las->rec[i - 1].aread) return fail(DN_ERR_INVALID, "LAS not sorted by A read");
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = ensure_device()) return rc;
    return guarded([&] {
        cudaSetDevice(g_device); arena().reset();
        const int ts = las->tspace;
        int64_t *ho = (int64_t *)hcache_alloc(sizeof(int64_t) * (nreads + 1));
        ho[0] = 0; for (int r = 0; r < nreads; r++) ho[r + 1] = ho[r] + (rlen[r] + ts - 1) / ts;
        uint8_t *hq = (uint8_t *)hcache_alloc(ho[nreads] + 1);
        DevLasIn d; d.upload(las, true, g_stream);
        DBuf<int32_t> dl; to_device(dl, rlen, nreads, g_stream);
        DBuf<int64_t> dq; to_device(dq, (const int64_t *)ho, nreads + 1, g_stream);
        DBuf<uint8_t> q(ho[nreads] + 1);
        DBuf<int32_t> dcv; if (cov_per_read) to_device(dcv, cov_per_read, nreads, g_stream);
        qv_device(dl.p, nreads, d.rec.p, las->nrec, d.toff.p, d.trace.p, ts, coverage, cov_per_read ? dcv.p : nullptr, dq.p, q.p, g_stream);
        if (ho[nreads]) DN_CUDA(cudaMemcpyAsync(hq, q.p, ho[nreads], cudaMemcpyDeviceToHost, g_stream));
        DN_CUDA(cudaStreamSynchronize(g_stream));
        *qv = hq; *qoff = ho;
        return DN_OK;
    });
}

int dn_las_chain_mapper(dn_las_buf *las, int32_t nb_reads, int32_t max_indel, int32_t max_gap) {
    if (!las || nb_reads < 0) return fail(DN_ERR_INVALID, "null argument");
    for (int64_t i = 0; i < las->nrec; i++)
        if (las->rec[i].bread < 0 || las->rec[i].bread >= nb_reads) return fail(DN_ERR_INVALID, "contig id out of bounds");
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = ensure_device()) return rc;
    return guarded([&] {
        cudaSetDevice(g_device); arena().reset();
        DBuf<dn_las_record> d(las->nrec + 1);
        if (las->nrec) {
            DN_CUDA(cudaMemcpyAsync(d.p, las->rec, sizeof(dn_las_record) * las->nrec, cudaMemcpyHostToDevice, g_stream));
            mapper_chain_device(d.p, las->nrec, nb_reads, max_indel, max_gap, g_stream);
            DN_CUDA(cudaMemcpyAsync(las->rec, d.p, sizeof(dn_las_record) * las->nrec, cudaMemcpyDeviceToHost, g_stream));
            DN_CUDA(cudaStreamSynchronize(g_stream));
        }
        return DN_OK;
    });
}

// damapper reports, per mapped read, the best chain and -- with -n<f> -- every chain whose score reaches the fraction f of
// the best (dazzler.d:5920-5923); DENTIST passes no -n (commandline.d:2943-2955).  Host logic over the START/NEXT/BEST
// flags dn_las_chain_mapper wrote: chain score = sum of the A spans of its records (the score BEST was chosen by).
int dn_las_keep_best_chains(dn_las_buf *las, int32_t nb_reads, double n_frac) {
    if (!las || nb_reads < 0) return fail(DN_ERR_INVALID, "null argument");
    return guarded([&]() -> int {
        std::vector<long long> best(nb_reads, -1);
        std::vector<long long> score(las->nrec, 0);
        for (int64_t i = 0; i < las->nrec;) {
            if (las->rec[i].bread < 0 || las->rec[i].bread >= nb_reads) return fail(DN_ERR_INVALID, "contig id out of bounds");
            if (!(las->rec[i].flags & DN_LAS_START)) return fail(DN_ERR_INVALID, "chain flags missing: call dn_las_chain_mapper first");
            int64_t j = i; long long sc = 0;
            do { sc += las->rec[j].aepos - las->rec[j].abpos; j++; } while (j < las->nrec && (las->rec[j].flags & DN_LAS_NEXT));
            for (int64_t x = i; x < j; x++) score[x] = sc;
            if (las->rec[i].flags & DN_LAS_BEST) best[las->rec[i].bread] = sc;
            i = j;
        }
        int64_t w = 0; bool keep = false;
        for (int64_t i = 0; i < las->nrec; i++) {
            if (las->rec[i].flags & DN_LAS_START) {
                const bool is_best = (las->rec[i].flags & DN_LAS_BEST) != 0;
                keep = is_best || (n_frac > 0.0 && (double)score[i] >= n_frac * (double)best[las->rec[i].bread]);
            }
            if (keep) { las->rec[w] = las->rec[i]; las->toff[w] = las->toff[i]; w++; }
        }
        las->nrec = w;
        return DN_OK;
    });
}

int dn_las_force_flat(dn_las_buf *las) {
    if (!las) return fail(DN_ERR_INVALID, "null argument");
    {   // already in FlatLocalAlignment order (base.d:1787-1809), e.g. straight from the aligner and order-preserving
        // filters: only the chain flags go (dazzler.d:4084-4093), no sort
        bool sorted = true;
        for (int64_t i = 1; i < las->nrec && sorted; i++) {
            const dn_las_record &p = las->rec[i - 1], &q = las->rec[i];
            const int pc = p.flags & DN_LAS_COMP ? 1 : 0, qc = q.flags & DN_LAS_COMP ? 1 : 0;
            const int32_t a[8] = {p.aread, p.bread, pc, p.abpos, p.aepos, p.bbpos, p.bepos, p.diffs};
            const int32_t b[8] = {q.aread, q.bread, qc, q.abpos, q.aepos, q.bbpos, q.bepos, q.diffs};
            for (int f = 0; f < 8; f++) { if (a[f] < b[f]) break; if (a[f] > b[f]) { sorted = false; break; } }
        }
        if (sorted) {
            for (int64_t i = 0; i < las->nrec; i++) las->rec[i].flags &= (DN_LAS_COMP | DN_LAS_ELIM);
            return DN_OK;
        }
    }
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = ensure_device()) return rc;
    return guarded([&] { cudaSetDevice(g_device); force_flat_device(las->rec, las->toff, las->nrec, g_stream); return DN_OK; });
}

// findReferenceReadCandidates (processPileUps/package.d:518-568) for a batch of pile-ups: host logic on the
// QV bytes (stays on the host in DENTIST too).  rank[pile_off[p] .. pile_off[p+1]) = reads of pile p ordered by
// (numBadQVs, meanQV, readId).
int dn_reference_read_candidates(const uint8_t *qv, const int64_t *qoff, const int32_t *group, int32_t nreads, int32_t npiles,
                                 double bad_fraction, int32_t *rank, int64_t *pile_off) {
    if (!qv || !qoff || !group || !rank || !pile_off || nreads < 0 || npiles < 0) return fail(DN_ERR_INVALID, "null argument");
    return guarded([&]() -> int {
        const int MAXQV = 50;                                         // DbRecord.maxQV, dazzler.d:2873
        std::vector<std::vector<int32_t>> members(npiles);
        for (int r = 0; r < nreads; r++) {
            if (group[r] < 0) continue;                               // not in allowedReferenceReadIds (package.d:456-468, 520-523)
            if (group[r] >= npiles) return fail(DN_ERR_INVALID, "pile id out of bounds");
            members[group[r]].push_back(r);
        }
        int64_t o = 0;
        for (int p = 0; p < npiles; p++) {
            pile_off[p] = o;
            long long hist[50] = {0}, total = 0;
            for (int r : members[p]) for (int64_t q = qoff[r]; q < qoff[r + 1]; q++) if (qv[q] < MAXQV) { hist[qv[q]]++; total++; }
            const size_t bad_thres = (size_t)(bad_fraction * (double)total);
            long long cum = 0; int idx = -1;
            for (int v = MAXQV - 1; v >= 0; v--) { cum += hist[v]; if ((size_t)cum >= bad_thres) { idx = MAXQV - 1 - v; break; } }
            const int bad_qv = MAXQV - 1 - idx;                       // package.d:536-540
            struct S { long long nbad; double mean; int32_t id; };
            std::vector<S> sc;
            for (int r : members[p]) {
                long long nb = 0, sum = 0; const int64_t n = qoff[r + 1] - qoff[r];
                for (int64_t q = qoff[r]; q < qoff[r + 1]; q++) { nb += qv[q] >= bad_qv; sum += qv[q]; }
                sc.push_back(S{nb, n ? (double)sum / (double)n : 0.0, r});
            }
            std::sort(sc.begin(), sc.end(), [](const S &a, const S &b) { if (a.nbad != b.nbad) return a.nbad < b.nbad; if (a.mean != b.mean) return a.mean < b.mean; return a.id < b.id; });
            for (const S &x : sc) rank[o++] = x.id;
        }
        pile_off[npiles] = o;
        return DN_OK;
    });
}

// damapper -C (dazzler.d:5931-5936): the records of B.A.las from those of A.B.las -- roles swapped, coordinates mirrored for
// complemented alignments, trace points re-laid on the new A read by per-tile realignment (spec: oracle/pile_oracle.c,
// orc_transpose), result in LAsort order.  Chain flags are cleared (chain the result again: the criteria are symmetric).
// `daligner -B` (dazzler.d:5823-5824): neighbouring records of one (aread, bread, comp) separated by a short gap become
// one record; specification in oracle/pile_oracle.c (orc_bridge).  The bridges' DPs run on the device (one thread each),
// the records and traces are re-assembled on the host (bridges are few).
int dn_las_bridge(const dn_block *a, const dn_block *b, dn_las_buf *las, int32_t cdiff, int64_t *nbridged) {
    if (!a || !b || !las) return fail(DN_ERR_INVALID, "null argument");
    const DevBlock &A = a->b, &B = b->b;
    const int ts = las->tspace;
    if (ts < 1) return fail(DN_ERR_INVALID, "bad trace spacing");
    if (cdiff < 1) return fail(DN_ERR_INVALID, "bad cdiff");
    if (const char *bad = validate_las(las, A.h_len.data(), A.nreads, B.h_len.data(), B.nreads, true)) return fail(DN_ERR_INVALID, bad);
    if (nbridged) *nbridged = 0;
    const int64_t n = las->nrec;
    std::vector<uint8_t> br(n > 0 ? n : 1, 0);
    std::vector<BridgeTask> tasks; int64_t nrows = 0;
    for (int64_t r = 1; r < n; r++) {
        const dn_las_record &p = las->rec[r - 1], &q = las->rec[r];
        const int gA = q.abpos - p.aepos, gB = q.bbpos - p.bepos;
        const int dg = gA > gB ? gA - gB : gB - gA, mx = gA > gB ? gA : gB;
        const long long lim = std::max<long long>(16ll * cdiff, 6ll * mx);
        if (p.aread != q.aread || p.bread != q.bread || ((p.flags ^ q.flags) & DN_LAS_COMP) || gA < 0 || gA > 128 || gB < 0 || gB > 250 ||
            (long long)dg * cdiff > lim) continue;
        br[r] = 1;
        BridgeTask t; t.ga = A.h_off[q.aread] + p.aepos; t.gb = B.h_off[q.bread] + p.bepos; t.n = gA; t.m = gB; t.a0 = p.aepos;
        t.comp = (q.flags & DN_LAS_COMP) ? 1 : 0;
        t.nrow = (p.aepos + gA) / ts - p.aepos / ts;                       // multiples of ts in (aepos, aepos + gA]
        t.row_off = nrows; nrows += t.nrow;
        tasks.push_back(t);
    }
    if (tasks.empty()) return DN_OK;
    std::vector<int32_t> total(tasks.size()); std::vector<int2> rows((size_t)nrows + 1);
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (int rc = ensure_device()) return rc;
        int rc = guarded([&]() -> int {
            cudaSetDevice(g_device); arena().reset();
            cudaStream_t s = g_stream;
            DBuf<BridgeTask> dt; to_device(dt, tasks.data(), tasks.size(), s);
            DBuf<int32_t> dtot(tasks.size()); DBuf<int2> drows((size_t)nrows + 1);
            DBuf<u32> scratch((size_t)cons_vote_threads() * 2048);
            launch_bridge(dt.p, (int64_t)tasks.size(), A.fwd.p, B.fwd.p, B.rc.p, ts, scratch.p, dtot.p, drows.p, s);
            DN_CUDA(cudaMemcpyAsync(total.data(), dtot.p, sizeof(int32_t) * tasks.size(), cudaMemcpyDeviceToHost, s));
            if (nrows) DN_CUDA(cudaMemcpyAsync(rows.data(), drows.p, sizeof(int2) * (size_t)nrows, cudaMemcpyDeviceToHost, s));
            DN_CUDA(cudaStreamSynchronize(s));
            return DN_OK;
        });
        if (rc) return rc;
    }
    return guarded([&]() -> int {
        // host: records and traces re-assembled (orc_bridge's fold)
        dn_las_record *orec = (dn_las_record *)hcache_alloc(sizeof(dn_las_record) * (size_t)(n + 1));
        int64_t *otoff = (int64_t *)hcache_alloc(sizeof(int64_t) * (size_t)(n + 1));
        uint16_t *otr = (uint16_t *)hcache_alloc(sizeof(uint16_t) * (size_t)(las->ntrace + 2 * (nrows + (int64_t)tasks.size()) + 2));
        int64_t no = 0, to = 0; size_t ti = 0; bool open = false;
        for (int64_t r = 0; r < n; r++) {
            const dn_las_record &q = las->rec[r];
            const uint16_t *tq = las->trace + las->toff[r];
            const int ntq = q.tlen / 2;
            if (!br[r]) {
                orec[no] = q; otoff[no] = to;
                for (int t = 0; t < 2 * ntq; t++) otr[to++] = tq[t];
                open = q.aepos % ts != 0; no++;
                continue;
            }
            const BridgeTask &T = tasks[ti]; const int tot = total[ti]; ti++;
            int pj = 0, pd = 0;
            for (int k = 0; k <= T.nrow; k++) {
                const int ej = k < T.nrow ? rows[T.row_off + k].x : T.m, ed = k < T.nrow ? rows[T.row_off + k].y : tot;
                const int dd = ed - pd, bb = ej - pj;
                if (k < T.nrow || dd > 0 || bb > 0) {
                    if (open) { otr[to - 2] = (uint16_t)(otr[to - 2] + dd); otr[to - 1] = (uint16_t)(otr[to - 1] + bb); }
                    else { otr[to++] = (uint16_t)dd; otr[to++] = (uint16_t)bb; open = true; }
                    if (k < T.nrow) open = false;
                }
                pj = ej; pd = ed;
            }
            for (int t = 0; t < ntq; t++) {
                if (t == 0 && open) { otr[to - 2] = (uint16_t)(otr[to - 2] + tq[0]); otr[to - 1] = (uint16_t)(otr[to - 1] + tq[1]); }
                else { otr[to++] = tq[2 * t]; otr[to++] = tq[2 * t + 1]; }
            }
            open = q.aepos % ts != 0;
            dn_las_record &o = orec[no - 1];
            o.aepos = q.aepos; o.bepos = q.bepos; o.diffs += tot + q.diffs; o.tlen = (int32_t)(to - otoff[no - 1]);
        }
        hcache_free(las->rec); hcache_free(las->toff); hcache_free(las->trace);
        las->rec = orec; las->toff = otoff; las->trace = otr; las->nrec = no; las->ntrace = to;
        if (nbridged) *nbridged = (int64_t)tasks.size();
        return DN_OK;
    });
}

int dn_las_transpose(const dn_block *a, const dn_block *b, const dn_las_buf *las, dn_las_buf *out) {
    if (!a || !b || !las || !out) return fail(DN_ERR_INVALID, "null argument");
    const DevBlock &A = a->b, &B = b->b;
    const int ts = las->tspace;
    if (ts < 1) return fail(DN_ERR_INVALID, "bad trace spacing");
    if (const char *bad = validate_las(las, A.h_len.data(), A.nreads, B.h_len.data(), B.nreads, true)) return fail(DN_ERR_INVALID, bad);
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = ensure_device()) return rc;
    return guarded([&]() -> int {
        cudaSetDevice(g_device); arena().reset();
        cudaStream_t s = g_stream;
        const int64_t n = las->nrec;
        std::vector<int32_t> vla(n); std::vector<int64_t> task0(n + 1, 0), otoff(n + 1, 0);
        int maxm = 0;
        for (int64_t i = 0; i < n; i++) {
            const dn_las_record &r = las->rec[i];
            vla[i] = (int32_t)i; task0[i + 1] = task0[i] + r.tlen / 2;
            for (int q = 0; q < r.tlen / 2; q++) maxm = std::max<int>(maxm, las->trace[las->toff[i] + 2 * q + 1]);
            const bool comp = (r.flags & DN_LAS_COMP) != 0;
            const int lb = B.h_len[r.bread];
            const int ab2 = comp ? lb - r.bepos : r.bbpos, ae2 = comp ? lb - r.bbpos : r.bepos;
            otoff[i + 1] = otoff[i] + (ae2 > ab2 ? 2 * ((ae2 + ts - 1) / ts - ab2 / ts) : 0);
        }
        const int64_t ntasks = task0[n], ntr = otoff[n];
        const int kmax = maxm / ts + 2;
        HostLas h;
        if (n == 0) {
            h.rec = (dn_las_record *)hcache_alloc(64); h.toff = (int64_t *)hcache_alloc(64); h.trace = (uint16_t *)hcache_alloc(64);
        } else {
            DevLasIn d; d.upload(las, true, s);
            DBuf<int32_t> d_vla; DBuf<int64_t> d_task0, d_otoff;
            to_device(d_vla, vla.data(), (size_t)n, s); to_device(d_task0, task0.data(), (size_t)n + 1, s); to_device(d_otoff, otoff.data(), (size_t)n + 1, s);
            DBuf<ConsTask> tasks((size_t)ntasks + 1);
            launch_cons_tasks(d.rec.p, d.toff.p, d.trace.p, d_vla.p, (int)n, d_task0.p, ts, tasks.p, s);
            TrGeom G{A.fwd.p, B.fwd.p, B.rc.p, A.off.p, B.off.p, A.len.p, B.len.p};
            DBuf<u32> scratch((size_t)cons_vote_threads() * 2048);
            DBuf<int4> cross((size_t)ntasks * kmax + 1); DBuf<int32_t> ncross((size_t)ntasks + 1), tcost((size_t)ntasks + 1), status(1);
            status.zero(s);
            launch_tr_tiles(tasks.p, ntasks, d.rec.p, d.toff.p, d.trace.p, d_task0.p, G, ts, kmax, scratch.p, cross.p, ncross.p, tcost.p, s);
            DBuf<dn_las_record> orec((size_t)n); DBuf<uint16_t> otr((size_t)ntr + 2);
            launch_tr_assemble(d.rec.p, n, d_task0.p, G, kmax, cross.p, ncross.p, tcost.p, d_otoff.p, orec.p, otr.p, status.p, s);
            int32_t st = 0;
            DN_CUDA(cudaMemcpyAsync(&st, status.p, 4, cudaMemcpyDeviceToHost, s)); DN_CUDA(cudaStreamSynchronize(s));
            if (st) return fail(DN_ERR_INVALID, "transpose: more trace-point crossings in one tile than its B bases allow");
            // LAsort order of the transposed records (aread' = B read ...), traces gathered, one download
            merge_las_device(orec.p, n, otr.p, ntr, B.maxlen, A.maxlen, B.nreads, A.nreads, h, s, false);
            for (int64_t i = 0; i < h.nrec; i++) h.rec[i].flags &= (DN_LAS_COMP | DN_LAS_ELIM);
        }
        memset(out, 0, sizeof *out);
        out->nrec = h.nrec; out->ntrace = h.ntrace; out->tspace = ts;
        out->rec = h.rec; out->toff = h.toff; out->trace = h.trace;
        h.rec = nullptr; h.toff = nullptr; h.trace = nullptr;
        return DN_OK;
    });
}

void dn_seq_free(dn_seq_buf *b) { if (!b) return; hcache_free(b->off); hcache_free(b->bases); memset(b, 0, sizeof *b); }

int dn_consensus(const dn_block *db, const dn_las_buf *las, const int32_t *reads, int32_t nreads, dn_seq_buf *out) {
    if (!db || !las || !reads || !out || nreads < 0) return fail(DN_ERR_INVALID, "null argument");
    const DevBlock &B = db->b;
    if (las->tspace < 1 || las->tspace > 128) return fail(DN_ERR_INVALID, "consensus needs trace spacing <= 128");
    for (int i = 0; i < nreads; i++) if (reads[i] < 0 || reads[i] >= B.nreads) return fail(DN_ERR_INVALID, "read id out of bounds");
    if (const char *bad = validate_las(las, B.h_len.data(), B.nreads, B.h_len.data(), B.nreads, true)) return fail(DN_ERR_INVALID, bad);
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = ensure_device()) return rc;
    return guarded([&] {
        cudaSetDevice(g_device); arena().reset();
        cudaStream_t s = g_stream;
        const int ts = las->tspace;
        // host: which LAs vote, for which target; vote-column and task offsets
        std::vector<int32_t> target_of(B.nreads, -1);
        for (int i = 0; i < nreads; i++) target_of[reads[i]] = i;          // a read listed twice: last wins
        std::vector<int64_t> vote_off(nreads + 1, 0);
        for (int i = 0; i < nreads; i++) vote_off[i + 1] = vote_off[i] + B.h_len[reads[i]] + 1;
        const int64_t ncols = vote_off[nreads];
        std::vector<int32_t> vla, la_target(las->nrec + 1, -1); std::vector<int64_t> task_off;
        std::vector<int32_t> voters(nreads, 0);
        int64_t ntasks = 0;
        for (int64_t x = 0; x < las->nrec; x++) {
            int tg = target_of[las->rec[x].aread];
            if (tg < 0 || reads[tg] != las->rec[x].aread) continue;
            la_target[x] = tg; vla.push_back((int32_t)x); task_off.push_back(ntasks); ntasks += las->rec[x].tlen / 2;
            voters[tg]++;
        }
        DevLasIn d; d.upload(las, true, s);
        DBuf<int32_t> d_vla, d_lat, d_targets; DBuf<int64_t> d_toff, d_voff;
        to_device(d_vla, vla.data(), vla.size(), s); to_device(d_lat, la_target.data(), la_target.size(), s);
        to_device(d_targets, reads, nreads, s); to_device(d_toff, task_off.data(), task_off.size(), s);
        to_device(d_voff, vote_off.data(), vote_off.size(), s);
        ConsGeom G{B.fwd.p, B.rc.p, B.off.p, B.len.p, d_voff.p};
        DBuf<int32_t> cnt(ncols * 5 + 1), ins(ncols * 4 + 1), insn(ncols + 1), cov(ncols + 1);
        cnt.zero(s); ins.zero(s); insn.zero(s); cov.zero(s);
        if (ntasks > 0) {
            DBuf<ConsTask> tasks(ntasks);
            launch_cons_tasks(d.rec.p, d.toff.p, d.trace.p, d_vla.p, (int)vla.size(), d_toff.p, ts, tasks.p, s);
            DBuf<u32> scratch((size_t)cons_vote_threads() * 2048);
            launch_cons_vote(tasks.p, ntasks, d.rec.p, d_lat.p, G, scratch.p, cnt.p, ins.p, insn.p, cov.p, s);
        }
        DBuf<int32_t> nemit(ncols + 1), eoff(ncols + 1), tot(1); DBuf<uint8_t> sym(2 * ncols + 2);
        int32_t total = 0;
        if (ncols > 0) {
            launch_cons_count(G, d_targets.p, nreads, ncols, cnt.p, ins.p, insn.p, cov.p, nemit.p, sym.p, s);
            exclusive_scan_i32(nemit.p, eoff.p, ncols, tot.p, s);
            DN_CUDA(cudaMemcpyAsync(&total, tot.p, 4, cudaMemcpyDeviceToHost, s)); DN_CUDA(cudaStreamSynchronize(s));
        }
        DBuf<uint8_t> dout(total + 1);
        if (ncols > 0) launch_cons_write(ncols, nemit.p, eoff.p, sym.p, dout.p, s);
        // sequence offsets = eoff at each target's first column
        std::vector<int32_t> heoff(nreads + 1, total);
        int32_t *h_eoff = (int32_t *)hcache_alloc(sizeof(int32_t) * (size_t)(ncols + 1));      // pinned: one copy instead of one per target
        if (ncols > 0) DN_CUDA(cudaMemcpyAsync(h_eoff, eoff.p, sizeof(int32_t) * (size_t)ncols, cudaMemcpyDeviceToHost, s));
        memset(out, 0, sizeof *out);
        out->nseq = nreads;
        out->off = (int64_t *)hcache_alloc(sizeof(int64_t) * (nreads + 1));
        out->bases = (uint8_t *)hcache_alloc(total + 1);
        if (total) DN_CUDA(cudaMemcpyAsync(out->bases, dout.p, total, cudaMemcpyDeviceToHost, s));
        DN_CUDA(cudaStreamSynchronize(s));
        for (int i = 0; i < nreads; i++) heoff[i] = h_eoff[vote_off[i]];
        hcache_free(h_eoff);
        // a read nothing aligns to has no consensus (daccord prints nothing -> "consensus could not be computed" and the
        // next reference read candidate, package.d:600-619, 307-329): its columns are squeezed out of the result
        int64_t w = 0;
        for (int i = 0; i < nreads; i++) {
            const int64_t b = heoff[i], e = heoff[i + 1];
            out->off[i] = w;
            if (voters[i] > 0) { if (w != b) memmove(out->bases + w, out->bases + b, (size_t)(e - b)); w += e - b; }
        }
        out->off[nreads] = w;
        return DN_OK;
    });
}

}  // extern "C"
