// api_batch.cu -- dn_process_pileups: the device part of PileUpProcessor.processPileUp
// (commands/processPileUps/package.d:283-374) for a whole BATCH of pile-ups in one call.
//
// The reference runs, per pile-up, computeQVs (:474-516) -> findReferenceReadCandidates (:518-568) ->
// selectReferenceRead / computeConsensus with retry (:307-329, 600-619) -> alignConsensusToFlankingContigs
// (:621-667), forking a tool at every step.  Here every step runs ONCE for all pile-ups of the batch on blocks
// whose reads carry their pile-up id (`group`): only reads of one pile-up are ever compared.  The D host keeps
// crop() in front of this call and getInsertionAlignment() / makeInsertion() behind it.
#include "api_internal.hpp"
#include <string.h>
#include <stdlib.h>
#include <memory>
#include <vector>
#include <chrono>

using namespace dn;
using namespace dnapi;

namespace {

struct BlockGuard { dn_block *b = nullptr; ~BlockGuard() { if (b) dn_block_free(b); } };
struct LasGuard { dn_las_buf l; LasGuard() { memset(&l, 0, sizeof l); } ~LasGuard() { dn_las_free(&l); } };

void count_per_pile(const dn_las_buf &las, const std::vector<int32_t> &group, std::vector<int64_t> &cnt) {
    std::fill(cnt.begin(), cnt.end(), 0);
    for (int64_t i = 0; i < las.nrec; i++) cnt[group[las.rec[i].aread]]++;
}

// drop the records of pile-ups that were skipped (their reads must not vote or rank any more); order preserving
void drop_piles(dn_las_buf &las, const std::vector<int32_t> &group, const dn_insertion_out *out) {
    int64_t w = 0;
    for (int64_t i = 0; i < las.nrec; i++)
        if (out[group[las.rec[i].aread]].status == DN_PILE_OK) { las.rec[w] = las.rec[i]; las.toff[w] = las.toff[i]; w++; }
    las.nrec = w;
}

// DN_TRACE=1: wall clock per stage of the batch (every stage entry point returns synchronised)
struct StageClock {
    bool on = getenv("DN_TRACE") != nullptr; std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void mark(const char *what) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[dn batch] %-34s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

}  // namespace

extern "C" {

void dn_pileup_params_default(dn_pileup_params *p) {
    memset(p, 0, sizeof *p);
    p->max_alignment_error = 0.3;             // commandline.d:1808
    p->min_anchor_length = 500;               // commandline.d:2036
    p->tspace = 126;                          // forceLargeTracePointType, dazzler.d:154
    p->proper_alignment_allowance = 126;      // default: the alignment's trace point distance, commandline.d:2317-2332
    p->bad_fraction = 0.08;                   // commandline.d:1101
    p->min_qv_coverage = 4;                   // dazzler.d:3771
    p->dust = 1;                              // dbdust(croppedDb) + -mdust, package.d:476-481
    p->max_indel = 1000; p->max_chain_gap = 10000; p->max_rel_overlap = 0.3; p->min_rel_score = 1.0; p->min_score = 0;   // commandline.d:2820-2830 (0: tspace)
    p->k = 14; p->flank_k = 14;
    p->bridge = 1;                            // daligner -B in pileUpAlignmentOptions and postConsensusAlignmentOptions
}

void dn_insertion_free(dn_insertion_out *out, int32_t n) {
    if (!out) return;
    for (int32_t i = 0; i < n; i++) { hcache_free(out[i].consensus); dn_las_free(&out[i].flank_las); }
    memset(out, 0, sizeof(dn_insertion_out) * (size_t)(n > 0 ? n : 0));
}

const char *dn_pile_status_string(int32_t status) {
    switch (status) {
        case DN_PILE_OK: return "ok";
        case DN_PILE_EMPTY_ALIGNMENT: return "empty pileup alignment";                          // package.d:487-490
        case DN_PILE_EMPTY_AFTER_FILTER: return "empty pileup alignment after filtering";       // package.d:512-515
        case DN_PILE_NO_REFERENCE_READ: return "no valid reference read found";                 // package.d:331-339
        default: return "unknown";
    }
}

int dn_block_add_mask(dn_block *blk, const int64_t *mask_anno, const int32_t *mask_data) {
    if (!blk || !mask_anno || !mask_data) return fail(DN_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = ensure_device()) return rc;
    return guarded([&] { cudaSetDevice(g_device); block_add_mask(blk->b, mask_anno, mask_data, g_stream); return DN_OK; });
}

int dn_process_pileups(const dn_block *ref, const dn_pileup_desc *piles, int32_t n, const dn_pileup_params *pp, dn_insertion_out *out) {
    if (n < 0 || (n && (!piles || !out))) return fail(DN_ERR_INVALID, "null argument");
    dn_pileup_params P; dn_pileup_params_default(&P);
    if (pp) P = *pp;
    if (P.tspace < 1 || P.tspace > 128) return fail(DN_ERR_INVALID, "pile-up trace spacing must be in [1,128]");
    return guarded([&]() -> int {
        memset(out, 0, sizeof(dn_insertion_out) * (size_t)n);
        StageClock sc;
        // ---- all cropped reads of all pile-ups as ONE block, pile-up id per read -------------------------------
        std::vector<int64_t> first(n + 1, 0), ffirst(n + 1, 0);
        for (int p = 0; p < n; p++) {
            if (piles[p].nreads < 0 || piles[p].nflanks < 0 || (piles[p].nreads && (!piles[p].rlen || !piles[p].bases)) ||
                (piles[p].nflanks && !piles[p].flank_read))
                return fail(DN_ERR_INVALID, "dn_pileup_desc: null field");
            first[p + 1] = first[p] + piles[p].nreads; ffirst[p + 1] = ffirst[p] + piles[p].nflanks;
            out[p].reference_read = -1;
            if (piles[p].nreads == 0) out[p].status = DN_PILE_EMPTY_ALIGNMENT;
        }
        const int64_t nr = first[n];
        if (nr >= (1ll << 31)) return fail(DN_ERR_INVALID, "too many reads in one batch");
        std::vector<int32_t> rlen(nr), group(nr); std::vector<int64_t> boff(nr + 1, 0); std::vector<uint8_t> allowed(nr, 1);
        for (int p = 0; p < n; p++)
            for (int i = 0; i < piles[p].nreads; i++) {
                const int64_t r = first[p] + i;
                if (piles[p].rlen[i] < 0) return fail(DN_ERR_INVALID, "negative read length");
                rlen[r] = piles[p].rlen[i]; group[r] = p; boff[r + 1] = boff[r] + rlen[r];
                if (piles[p].allowed) allowed[r] = piles[p].allowed[i] ? 1 : 0;
            }
        // gathered into a page-locked block (recycled): the upload then runs at the link's speed instead of through a staging copy
        struct HostBlockGuard { void *p; ~HostBlockGuard() { hcache_free(p); } } bases_guard{hcache_alloc((size_t)boff[nr] + 1)};
        uint8_t *bases = (uint8_t *)bases_guard.p;
        for (int p = 0; p < n; p++)
            if (piles[p].nreads) memcpy(bases + boff[first[p]], piles[p].bases, (size_t)(boff[first[p + 1]] - boff[first[p]]));
        dn_block_desc d; memset(&d, 0, sizeof d);
        d.nreads = (int32_t)nr; d.format = DN_SEQ_BYTES; d.rlen = rlen.data(); d.boff = boff.data(); d.data = bases;
        d.data_bytes = boff[nr]; d.group = group.data();
        sc.mark("gather reads (host)");
        TrustedLas trusted;            // every LAS below is produced by this call itself
        BlockGuard g;
        if (int rc = dn_block_upload(&d, &g.b)) return rc;
        sc.mark("upload");
        if (P.dust) { if (int rc = dn_block_mask_dust(g.b, 64, 2.0, 10, nullptr)) return rc; }          // package.d:476
        sc.mark("dust");

        // ---- computeQVs (package.d:474-516) ---------------------------------------------------------------------
        // daligner -B -s126 -l<minAnchorLength> -e<1 - maxAlignmentError> -mdust X X   (pileUpAlignmentOptions, commandline.d:2886-2902)
        dn_align_params ap; dn_align_params_default(&ap);
        ap.k = P.k; ap.tspace = P.tspace; ap.minlen = P.min_anchor_length; ap.e = 1.0 - P.max_alignment_error; ap.self_block = 1;
        LasGuard las;
        if (int rc = dn_align_blocks(g.b, g.b, &ap, &las.l)) return rc;
        if (P.bridge) { if (int rc = dn_las_bridge(g.b, g.b, &las.l, (int32_t)(6.0 / (1.0 - ap.e) + 0.5), nullptr)) return rc; }
        sc.mark("pile alignment");
        if (int rc = dn_las_filter_error(&las.l, P.max_alignment_error)) return rc;                      // :483-485
        sc.mark("filter error");
        std::vector<int64_t> cnt(n > 0 ? n : 1);
        count_per_pile(las.l, group, cnt);
        for (int p = 0; p < n; p++) if (out[p].status == DN_PILE_OK && cnt[p] == 0) out[p].status = DN_PILE_EMPTY_ALIGNMENT;   // :487-490
        if (las.l.nrec > 0)
            if (int rc = dn_las_chain(&las.l, P.max_indel, P.max_chain_gap, P.max_rel_overlap, P.min_rel_score,
                                      P.min_score > 0 ? P.min_score : P.tspace)) return rc;              // :492-496
        sc.mark("chain");
        // coverage = |allowedReferenceReadIds|, raised to minQVCoverage for pile-ups of >= minQVCoverage reads (:498-501)
        std::vector<int32_t> cov(nr);
        for (int p = 0; p < n; p++) {
            int32_t na = 0;
            for (int64_t r = first[p]; r < first[p + 1]; r++) na += allowed[r];
            if (na < P.min_qv_coverage && piles[p].nreads >= P.min_qv_coverage) na = P.min_qv_coverage;
            for (int64_t r = first[p]; r < first[p + 1]; r++) cov[r] = na;
        }
        uint8_t *qv = nullptr; int64_t *qoff = nullptr;
        if (int rc = dn_compute_qvs_v(rlen.data(), (int32_t)nr, &las.l, 0, cov.data(), &qv, &qoff)) return rc;       // :503
        sc.mark("qvs");
        struct QvGuard { uint8_t *q; int64_t *o; ~QvGuard() { dn_free(q); dn_free(o); } } qg{qv, qoff};
        if (int rc = dn_las_filter_pileup(&las.l, rlen.data(), (int32_t)nr, rlen.data(), (int32_t)nr, P.proper_alignment_allowance)) return rc;   // :505-510
        sc.mark("filter pile-up");
        if (int rc = dn_las_force_flat(&las.l)) return rc;
        sc.mark("force flat");
        count_per_pile(las.l, group, cnt);
        for (int p = 0; p < n; p++) if (out[p].status == DN_PILE_OK && cnt[p] == 0) out[p].status = DN_PILE_EMPTY_AFTER_FILTER;   // :512-515
        drop_piles(las.l, group, out);

        // ---- findReferenceReadCandidates (:518-568), selectReferenceRead / computeConsensus with retry (:307-329) ----
        std::vector<int32_t> cgroup(nr), rank(nr > 0 ? nr : 1); std::vector<int64_t> poff(n + 1, 0);
        for (int64_t r = 0; r < nr; r++) cgroup[r] = (allowed[r] && out[group[r]].status == DN_PILE_OK) ? group[r] : -1;
        if (int rc = dn_reference_read_candidates(qv, qoff, cgroup.data(), (int32_t)nr, n, P.bad_fraction, rank.data(), poff.data())) return rc;
        sc.mark("reference read candidates");
        std::vector<std::vector<uint8_t>> cons(n);
        std::vector<int32_t> pending;
        for (int p = 0; p < n; p++) if (out[p].status == DN_PILE_OK) pending.push_back(p);
        for (int attempt = 0; !pending.empty(); attempt++) {
            std::vector<int32_t> targets, who;
            for (int p : pending) {
                if (poff[p] + attempt < poff[p + 1]) { targets.push_back(rank[poff[p] + attempt]); who.push_back(p); }
                else { out[p].status = DN_PILE_NO_REFERENCE_READ; out[p].ntries = attempt; }
            }
            pending.clear();
            if (targets.empty()) break;
            dn_seq_buf sq; memset(&sq, 0, sizeof sq);
            if (int rc = dn_consensus(g.b, &las.l, targets.data(), (int32_t)targets.size(), &sq)) return rc;       // :600-619
            for (size_t i = 0; i < who.size(); i++) {
                const int p = who[i]; const int64_t len = sq.off[i + 1] - sq.off[i];
                if (len > 0) {
                    cons[p].assign(sq.bases + sq.off[i], sq.bases + sq.off[i + 1]);
                    out[p].reference_read = (int32_t)(targets[i] - first[p]); out[p].ntries = attempt + 1;
                } else pending.push_back(p);                                  // "consensus could not be computed": next candidate
            }
            dn_seq_free(&sq);
        }
        for (int p = 0; p < n; p++) {
            if (out[p].status != DN_PILE_OK) continue;
            out[p].cons_len = (int64_t)cons[p].size();
            out[p].consensus = (uint8_t *)hcache_alloc(cons[p].size() + 1);
            if (!cons[p].empty()) memcpy(out[p].consensus, cons[p].data(), cons[p].size());
        }

        sc.mark("consensus");
        // ---- alignConsensusToFlankingContigs (:621-667) ---------------------------------------------------------------
        // daligner -A -B -s126 -l126 -e0.7 -mdust -mrep F C   (postConsensusAlignmentOptions, commandline.d:2918-2935)
        const int64_t nf = ffirst[n];
        if (ref && nf > 0) {
            std::vector<int32_t> fread(nf), fbeg(nf, 0), fend(nf), fgroup(nf);
            std::vector<int64_t> manno(nf + 1, 0); std::vector<int32_t> mdata;
            for (int p = 0; p < n; p++)
                for (int i = 0; i < piles[p].nflanks; i++) {
                    const int64_t f = ffirst[p] + i; const int32_t rr = piles[p].flank_read[i];
                    if (rr < 0 || rr >= ref->b.nreads) return fail(DN_ERR_INVALID, "flanking contig id out of bounds");
                    fread[f] = rr; fend[f] = ref->b.h_len[rr]; fgroup[f] = p;
                    manno[f] = (int64_t)mdata.size() * 4;
                    if (piles[p].mask_anno && piles[p].mask_data)
                        for (int64_t x = piles[p].mask_anno[i] / 4; x < piles[p].mask_anno[i + 1] / 4; x++) mdata.push_back(piles[p].mask_data[x]);
                }
            manno[nf] = (int64_t)mdata.size() * 4;
            BlockGuard fb, cb;
            if (int rc = dn_block_crop(ref, (int32_t)nf, fread.data(), fbeg.data(), fend.data(), fgroup.data(), &fb.b)) return rc;
            if (!mdata.empty()) { if (int rc = dn_block_add_mask(fb.b, manno.data(), mdata.data())) return rc; }     // writeMask(..., "rep")
            if (int rc = dn_block_mask_dust(fb.b, 64, 2.0, 10, nullptr)) return rc;                                  // dbdust(flankingContigsDb)
            std::vector<int32_t> clen(n), cgrp(n); std::vector<int64_t> coff(n + 1, 0); std::vector<uint8_t> cbases;
            for (int p = 0; p < n; p++) { clen[p] = (int32_t)cons[p].size(); cgrp[p] = p; coff[p + 1] = coff[p] + clen[p]; }
            cbases.resize((size_t)coff[n] + 1);
            for (int p = 0; p < n; p++) if (clen[p]) memcpy(cbases.data() + coff[p], cons[p].data(), clen[p]);
            dn_block_desc cd; memset(&cd, 0, sizeof cd);
            cd.nreads = n; cd.format = DN_SEQ_BYTES; cd.rlen = clen.data(); cd.boff = coff.data(); cd.data = cbases.data();
            cd.data_bytes = coff[n]; cd.group = cgrp.data();
            if (int rc = dn_block_upload(&cd, &cb.b)) return rc;
            sc.mark("flank crop + masks + upload");
            dn_align_params fp; dn_align_params_default(&fp);
            fp.k = P.flank_k; fp.tspace = P.tspace; fp.minlen = P.tspace; fp.e = 0.7;
            LasGuard fl;
            if (int rc = dn_align_blocks(fb.b, cb.b, &fp, &fl.l)) return rc;
            if (P.bridge) { if (int rc = dn_las_bridge(fb.b, cb.b, &fl.l, (int32_t)(6.0 / (1.0 - fp.e) + 0.5), nullptr)) return rc; }
            sc.mark("flank alignment");
            // split by pile-up (bread = pile-up id): records keep their LAsort order, aread becomes the index into the
            // pile-up's own flank list (1-based contig id of the reference's flankingContigsDb minus one), bread = 0
            std::vector<int64_t> pn(n, 0), pt(n, 0);
            for (int64_t i = 0; i < fl.l.nrec; i++) { pn[fl.l.rec[i].bread]++; pt[fl.l.rec[i].bread] += fl.l.rec[i].tlen; }
            for (int p = 0; p < n; p++) {
                if (out[p].status != DN_PILE_OK) continue;
                dn_las_buf &o = out[p].flank_las;
                o.tspace = P.tspace; o.nrec = 0; o.ntrace = 0;
                o.rec = (dn_las_record *)hcache_alloc(sizeof(dn_las_record) * (size_t)(pn[p] + 1));
                o.toff = (int64_t *)hcache_alloc(sizeof(int64_t) * (size_t)(pn[p] + 1));
                o.trace = (uint16_t *)hcache_alloc(sizeof(uint16_t) * (size_t)(pt[p] + 1));
            }
            for (int64_t i = 0; i < fl.l.nrec; i++) {
                const int p = fl.l.rec[i].bread;
                if (out[p].status != DN_PILE_OK) continue;
                dn_las_buf &o = out[p].flank_las;
                dn_las_record r = fl.l.rec[i];
                r.aread -= (int32_t)ffirst[p]; r.bread = 0;
                o.rec[o.nrec] = r; o.toff[o.nrec] = o.ntrace;
                memcpy(o.trace + o.ntrace, fl.l.trace + fl.l.toff[i], sizeof(uint16_t) * (size_t)r.tlen);
                o.ntrace += r.tlen; o.nrec++;
            }
            sc.mark("split by pile-up");
        }
        return DN_OK;
    });
}

}  // extern "C"
