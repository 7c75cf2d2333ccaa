/*
 * pile_oracle.c -- CPU ORACLE (test infrastructure, NOT product code) for the per-pile-up stages of
 * processPileUps (commands/processPileUps/package.d:474-619):
 *
 *   orc_filter_error     filterLocalAlignments!(la => la.averageErrorRate <= maxAlignmentError)
 *                        dazzler.d:3885-3899, call package.d:483-485, rate base.d:1764-1767
 *   orc_filter_pileup    filterPileUpAlignments / isValidPileUpAlignment  dazzler.d:4084-4141
 *                        (FlatLocus.beginsWithin / endsWithin base.d:1713-1724)
 *   orc_qv               computeQVs -> `DAScover`, `DASqv -c<cov>`  dazzler.d:3782-3792, 6142-6156
 *                        and `computeintrinsicqv -d<depth>` dazzler.d:6172-6183 (same rule, other track)
 *   orc_consensus        getConsensus -> `daccord -f -I<i>,<i>`  dazzler.d:4213-4255, 6187-6220
 *
 * PARITY STATUS.  The two filters restate D source that IS in /root/reference and are exact.
 * orc_qv and orc_consensus are "parity unpinned": DASCRUBBER @ a53dbe87 and daccord 0.0.18 are absent
 * (no source, no binaries), and the reference holds no QV vectors.  They follow north_star's definition
 * ("per-gap pile-up consensus as a ... majority/intrinsic-QV vote over the aligned read pile") and the
 * DASqv rule recorded in SURVEY Appendix A.2; the one reference vector that exists for this stage, the
 * consensus KAT dazzler.d:4257-4299 (3 reads, two single-base errors => consensus == read 3), is
 * reproduced in tests/test_pile_oracle.py.
 *
 * Specification
 *  QV: for read a and tile t (A bases [t*ts, min((t+1)*ts, len))) every LA with aread == a whose trace
 *      has a FULL tile t (tile lies inside [abpos, aepos]) contributes
 *      v = min(50, (200*diffs + (alen+bb)/2) / (alen+bb)), alen = tile length, bb = its B bases.
 *      m = #contributions; if m*4 < cov the tile is uncovered: QV = 50.  Otherwise sort ascending and
 *      QV = round-half-up mean of the first n = clamp(cov/2, 1, m) values.
 *  Consensus of read r: every LA with aread == r votes.  Per tile, the A tile and the B tile are aligned
 *      globally with unit costs; traceback from the end prefers diagonal, then "A base unmatched"
 *      (deletion in B), then "B base inserted".  Each aligned column adds a vote {a,c,g,t} or {deleted}
 *      to its A position; the first base of an insertion run between A positions p-1 and p votes for an
 *      insertion in front of p.  Tiles with more than 250 B bases do not vote.  Read r itself votes its own
 *      base once.  Position p emits the winning symbol (ties: r's own base first, then lowest code;
 *      a winning "deleted" emits nothing); an insertion in front of p is emitted when more than half of
 *      (covering LAs + 1) vote for one.  Uncovered positions keep r's base (daccord -f).  A read without a single
 *      voting LA has no consensus at all (empty output).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t tlen, diffs, abpos, bbpos, aepos, bepos;
    uint32_t flags;
    int32_t aread, bread;
    int32_t pad;
} las_rec;                                  /* the 40-byte LAS record */

void orc_filter_error(const las_rec *la, int64_t n, double max_err, uint8_t *keep) {
    for (int64_t i = 0; i < n; i++) {
        double rate = (double)la[i].diffs / (double)(la[i].aepos - la[i].abpos);
        keep[i] = rate <= max_err;
    }
}

void orc_filter_pileup(const las_rec *la, int64_t n, const int32_t *alen, const int32_t *blen, int32_t allowance, uint8_t *keep) {
    for (int64_t i = 0; i < n; i++) {
        const las_rec *x = &la[i];
        int la_ = alen[x->aread], lb_ = blen[x->bread];
        int ab = x->abpos <= allowance, bb = x->bbpos <= allowance;
        int ae = x->aepos + allowance >= la_, be = x->bepos + allowance >= lb_;
        int left_anch = ab && bb, left_prop = ab || bb, right_anch = ae && be, right_prop = ae || be;
        keep[i] = x->aread != x->bread && ((left_anch && right_prop) || (right_anch && left_prop));
    }
}

static int cmp_int(const void *a, const void *b) { return *(const int *)a - *(const int *)b; }

/* qv_off[r] = first tile byte of read r (prefix sum of ceil(len/ts)); la sorted by aread */
void orc_qv(const int32_t *rlen, int32_t nreads, const las_rec *la, int64_t nla, const int64_t *toff,
            const uint16_t *trace, int32_t ts, int32_t cov, const int64_t *qv_off, uint8_t *qv)
{
    int *vals = malloc(sizeof(int) * (nla + 1));
    int64_t lo = 0;
    for (int r = 0; r < nreads; r++) {
        while (lo < nla && la[lo].aread < r) lo++;
        int64_t hi = lo;
        while (hi < nla && la[hi].aread == r) hi++;
        int nt = (rlen[r] + ts - 1) / ts;
        for (int t = 0; t < nt; t++) {
            int t0 = t * ts, t1 = (t + 1) * ts < rlen[r] ? (t + 1) * ts : rlen[r];
            int m = 0;
            for (int64_t x = lo; x < hi; x++) {
                if (la[x].abpos > t0 || la[x].aepos < t1) continue;
                int idx = t - la[x].abpos / ts;
                const uint16_t *tp = trace + toff[x] + 2 * idx;
                int alen = t1 - t0, den = alen + tp[1];
                int v = (200 * tp[0] + den / 2) / den;
                vals[m++] = v > 50 ? 50 : v;
            }
            int q = 50;
            if (m * 4 >= cov && m > 0) {
                qsort(vals, m, sizeof(int), cmp_int);
                int n = cov / 2; if (n < 1) n = 1; if (n > m) n = m;
                int s = 0; for (int i = 0; i < n; i++) s += vals[i];
                q = (2 * s + n) / (2 * n);
                if (q > 50) q = 50;
            }
            qv[qv_off[r] + t] = (uint8_t)q;
        }
        lo = hi;
    }
    free(vals);
}

/* votes: cnt[p*5 + sym] (sym 0..3 base, 4 deleted), ins[p*4 + base], insn[p], cov[p] */
static void vote_tile(const uint8_t *a, int n, const uint8_t *b, int m, int p0,
                      int32_t *cnt, int32_t *ins, int32_t *insn, int32_t *cov)
{
    if (m > 250) return;
    uint8_t D[128 + 1][256];                    /* on the stack: the oracle is called from several host threads */
    for (int j = 0; j <= m; j++) D[0][j] = (uint8_t)j;
    for (int i = 1; i <= n; i++) {
        D[i][0] = (uint8_t)i;
        for (int j = 1; j <= m; j++) {
            int d = D[i - 1][j - 1] + (a[i - 1] != b[j - 1]);
            int u = D[i - 1][j] + 1, l = D[i][j - 1] + 1;
            int v = d; if (u < v) v = u; if (l < v) v = l;
            D[i][j] = (uint8_t)v;
        }
    }
    int i = n, j = m;
    int pend = -1;                              /* first base (lowest j) of the insertion run being walked */
    while (i > 0 || j > 0) {
        if (i > 0 && j > 0 && D[i][j] == D[i - 1][j - 1] + (a[i - 1] != b[j - 1])) {
            if (pend >= 0) { ins[(p0 + i) * 4 + pend]++; insn[p0 + i]++; pend = -1; }
            cnt[(p0 + i - 1) * 5 + b[j - 1]]++; i--; j--;
        } else if (i > 0 && D[i][j] == D[i - 1][j] + 1) {
            if (pend >= 0) { ins[(p0 + i) * 4 + pend]++; insn[p0 + i]++; pend = -1; }
            cnt[(p0 + i - 1) * 5 + 4]++; i--;
        } else {
            pend = b[j - 1]; j--;               /* b[j-1] is inserted in front of A position p0+i */
        }
    }
    if (pend >= 0) { ins[(p0 + i) * 4 + pend]++; insn[p0 + i]++; }
    for (int x = 0; x < n; x++) cov[p0 + x]++;
}

/* bases/off: the DB block (codes 0..3); returns consensus length in *outlen (out must hold 2*len+16) */
int orc_consensus(const int64_t *off, const uint8_t *bases, const las_rec *la, int64_t nla, const int64_t *toff,
                  const uint16_t *trace, int32_t ts, int32_t r, uint8_t *out, int32_t *outlen)
{
    const int L = (int)(off[r + 1] - off[r]);
    const uint8_t *A = bases + off[r];
    int32_t *cnt = calloc((size_t)(L + 1) * 5, 4), *ins = calloc((size_t)(L + 2) * 4, 4);
    int32_t *insn = calloc(L + 2, 4), *cov = calloc(L + 2, 4);
    uint8_t *brc = NULL; int brc_cap = 0;
    int64_t voters = 0;
    for (int64_t x = 0; x < nla; x++) {
        if (la[x].aread != r) continue;
        voters++;
        const int b = la[x].bread, LB = (int)(off[b + 1] - off[b]);
        const uint8_t *B = bases + off[b];
        if (la[x].flags & 1u) {
            if (LB > brc_cap) { brc = realloc(brc, LB + 1); brc_cap = LB; }
            for (int i = 0; i < LB; i++) brc[i] = 3 - B[LB - 1 - i];
            B = brc;
        }
        int ap = la[x].abpos, bp = la[x].bbpos, nt = la[x].tlen / 2;
        for (int t = 0; t < nt; t++) {
            int aend = (t == nt - 1) ? la[x].aepos : (ap / ts + 1) * ts;
            int bb = trace[toff[x] + 2 * t + 1];
            vote_tile(A + ap, aend - ap, B + bp, bb, ap, cnt, ins, insn, cov);
            ap = aend; bp += bb;
        }
    }
    int o = 0;
    /* a read nothing aligns to has no consensus (daccord prints nothing; DENTIST then reports "consensus could not be
     * computed" and tries the next reference read candidate, package.d:600-619, 307-329) */
    for (int p = 0; voters > 0 && p <= L; p++) {
        /* insertion in front of p (p == L: after the last base) */
        int total = (p < L ? cov[p] : (L > 0 ? cov[L - 1] : 0)) + 1;
        if (2 * insn[p] > total) {
            int best = 0;
            for (int s = 1; s < 4; s++) if (ins[p * 4 + s] > ins[p * 4 + best]) best = s;
            out[o++] = (uint8_t)best;
        }
        if (p == L) break;
        int own = A[p];
        int bestsym = own, bestc = cnt[p * 5 + own] + 1;
        for (int s = 0; s < 5; s++) {
            int c = cnt[p * 5 + s] + (s == own ? 1 : 0);
            if (c > bestc) { bestc = c; bestsym = s; }
        }
        if (bestsym < 4) out[o++] = (uint8_t)bestsym;
    }
    *outlen = o;
    free(cnt); free(ins); free(insn); free(cov); free(brc);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * orc_transpose -- what `damapper -C` adds (dazzler.d:5931-5936): the file Y.X.las "that contains all the same matches
 * as in X.Y.las" with the roles of the reads swapped.  PARITY UNPINNED (DAMAPPER absent); specification:
 *   record: aread' = bread, bread' = aread, COMP kept; not complemented: (ab', ae', bb', be') = (bb, be, ab, ae);
 *           complemented: ab' = lb - be, ae' = lb - bb, bb' = la - ae, be' = la - ab (each read in its own frame).
 *   path:   every A tile (its A bases against its B bases) is aligned globally with unit costs, traceback from the end
 *           preferring diagonal, then "A base unmatched", then "B base inserted" -- the rule of the consensus vote.
 *           Tiles with more than 250 B bases or more than 128 A bases take the diagonal-first path with the tile's
 *           recorded diffs spread evenly along it.
 *   trace': a trace point at every multiple of ts of the NEW A read (= B) strictly inside the alignment: the other read's
 *           coordinate and the diffs at the moment the path first reaches that column, walking in the new A read's
 *           direction (for complemented alignments that is the old path walked backwards).  diffs' = sum of the tiles'.
 * Records come back in input order (the caller sorts them into LAsort order); out_trace needs sum(tlen') elements. */
typedef struct { int x, a, c; } orc_cross;

static int tile_path(const uint8_t *a, int n, const uint8_t *b, int m, int dt, int comp, int bp, int ts, int lb, int lo, int hi,
                     orc_cross *cr, int *total)
{
    /* fills cr[] with the crossings of this tile in DECREASING x; returns their number */
    int nc = 0;
    if (m > 250 || n > 128) {
        *total = dt;
        const int q = n < m ? n : m;
        for (int j = m; j >= 0; j--) {
            const int x = bp + j;
            if (x <= lo || x >= hi) continue;
            int i;
            if (!comp) { if (j == 0 || x % ts) continue; i = j <= q ? j : n; if (j == m && n > m) i = m; }
            else { if (j == m || (lb - x) % ts) continue; i = j <= q ? j : n; if (j == m && n > m) i = n; }
            const int from_start = (n + m) ? (int)((long long)dt * (i + j) / (n + m)) : 0;
            cr[nc].x = x; cr[nc].a = i; cr[nc].c = comp ? dt - from_start : from_start; nc++;
        }
        return nc;
    }
    static __thread uint8_t D[128 + 1][256];
    for (int j = 0; j <= m; j++) D[0][j] = (uint8_t)j;
    for (int i = 1; i <= n; i++) {
        D[i][0] = (uint8_t)i;
        for (int j = 1; j <= m; j++) {
            int d = D[i - 1][j - 1] + (a[i - 1] != b[j - 1]);
            int u = D[i - 1][j] + 1, l = D[i][j - 1] + 1;
            int v = d; if (u < v) v = u; if (l < v) v = l;
            D[i][j] = (uint8_t)v;
        }
    }
    *total = D[n][m];
    int i = n, j = m, cend = 0;
    /* comp: the first cell visited in a column; not comp: the last one (recorded when the walk leaves the column) */
    for (;;) {
        const int x = bp + j;
        int dir;                                     /* 0 diagonal, 1 up (A base unmatched), 2 left (B base inserted), 3 done */
        if (i == 0 && j == 0) dir = 3;
        else if (i > 0 && j > 0 && D[i][j] == D[i - 1][j - 1] + (a[i - 1] != b[j - 1])) dir = 0;
        else if (i > 0 && D[i][j] == D[i - 1][j] + 1) dir = 1;
        else dir = 2;
        if (!comp) {
            if ((dir == 0 || dir == 2) && j > 0 && x % ts == 0 && x > lo && x < hi) { cr[nc].x = x; cr[nc].a = i; cr[nc].c = *total - cend; nc++; }
        }
        if (dir == 3) break;
        if (dir == 0) { cend += (a[i - 1] != b[j - 1]); i--; j--; }
        else if (dir == 1) { cend += 1; i--; }
        else { cend += 1; j--; }
        if (comp && (dir == 0 || dir == 2)) {        /* just entered column j (< m) */
            const int xn = bp + j;
            if ((lb - xn) % ts == 0 && xn > lo && xn < hi) { cr[nc].x = xn; cr[nc].a = i; cr[nc].c = cend; nc++; }
        }
    }
    return nc;
}

void orc_transpose(const int64_t *aoff, const uint8_t *abases, const int64_t *boff, const uint8_t *bbases,
                   const las_rec *la, int64_t nla, const int64_t *toff, const uint16_t *trace, int32_t ts,
                   las_rec *out, int64_t *out_toff, uint16_t *out_trace)
{
    int64_t to = 0;
    uint8_t *brc = NULL; int brc_cap = 0;
    for (int64_t r = 0; r < nla; r++) {
        const las_rec *x = &la[r];
        const int LA = (int)(aoff[x->aread + 1] - aoff[x->aread]), LB = (int)(boff[x->bread + 1] - boff[x->bread]);
        const uint8_t *A = abases + aoff[x->aread], *B = bbases + boff[x->bread];
        const int comp = (int)(x->flags & 1u);
        if (comp) {
            if (LB > brc_cap) { brc = realloc(brc, LB + 1); brc_cap = LB; }
            for (int i = 0; i < LB; i++) brc[i] = 3 - B[LB - 1 - i];
            B = brc;
        }
        const int nt = x->tlen / 2;
        /* all crossings of the record, tile after tile */
        int cap = (x->bepos - x->bbpos) / ts + 4 + nt, ncr = 0;
        orc_cross *cr = malloc(sizeof(orc_cross) * cap), *tmp = malloc(sizeof(orc_cross) * cap);
        int *tcost = malloc(sizeof(int) * (nt + 1)), *tfirst = malloc(sizeof(int) * (nt + 2));
        int ap = x->abpos, bp = x->bbpos;
        for (int t = 0; t < nt; t++) {
            const int aend = (t == nt - 1) ? x->aepos : (ap / ts + 1) * ts;
            const int m = trace[toff[r] + 2 * t + 1], dt = trace[toff[r] + 2 * t];
            tfirst[t] = ncr;
            const int k = tile_path(A + ap, aend - ap, B + bp, m, dt, comp, bp, ts, LB, x->bbpos, x->bepos, tmp, &tcost[t]);
            for (int q = 0; q < k; q++) { cr[ncr] = tmp[q]; cr[ncr].a += ap; ncr++; }     /* a: global A coordinate */
            ap = aend; bp += m;
        }
        tfirst[nt] = ncr;
        las_rec o = *x;
        o.aread = x->bread; o.bread = x->aread;
        out_toff[r] = to;
        int total = 0; for (int t = 0; t < nt; t++) total += tcost[t];
        int ntp = 0;
        if (!comp) {
            o.abpos = x->bbpos; o.aepos = x->bepos; o.bbpos = x->abpos; o.bepos = x->aepos;
            int pa = x->abpos, pd = 0, base = 0;
            for (int t = 0; t < nt; t++) {
                for (int q = tfirst[t + 1] - 1; q >= tfirst[t]; q--) {           /* increasing x inside the tile */
                    const int d = base + cr[q].c;
                    out_trace[to++] = (uint16_t)(d - pd); out_trace[to++] = (uint16_t)(cr[q].a - pa); ntp++;
                    pd = d; pa = cr[q].a;
                }
                base += tcost[t];
            }
            if (x->bepos > x->bbpos) { out_trace[to++] = (uint16_t)(total - pd); out_trace[to++] = (uint16_t)(x->aepos - pa); ntp++; }
        } else {
            o.abpos = LB - x->bepos; o.aepos = LB - x->bbpos; o.bbpos = LA - x->aepos; o.bepos = LA - x->abpos;
            int pa = x->aepos, pd = 0, base = 0;
            for (int t = nt - 1; t >= 0; t--) {
                for (int q = tfirst[t]; q < tfirst[t + 1]; q++) {                /* decreasing x inside the tile */
                    const int d = base + cr[q].c;
                    out_trace[to++] = (uint16_t)(d - pd); out_trace[to++] = (uint16_t)(pa - cr[q].a); ntp++;
                    pd = d; pa = cr[q].a;
                }
                base += tcost[t];
            }
            if (x->bepos > x->bbpos) { out_trace[to++] = (uint16_t)(total - pd); out_trace[to++] = (uint16_t)(pa - x->abpos); ntp++; }
        }
        o.diffs = total; o.tlen = 2 * ntp;
        out[r] = o;
        free(cr); free(tmp); free(tcost); free(tfirst);
    }
    free(brc);
}

/* ---- bridging: what `daligner -B` adds (dazzler.d:5823-5824; DENTIST passes it for pile and flank alignments,
 * commandline.d:2886-2902, 2918-2935).  PARITY UNPINNED: DALIGNER's bridging rule is neither in the reference nor
 * documented there; this is an own deterministic specification of "bridge consecutive aligned segments into one if
 * possible".
 *   Records are in LAsort order.  Inside a run of records of one (aread, bread, comp), neighbours P, Q (in file order)
 *   bridge iff  gA = Q.abpos - P.aepos and gB = Q.bbpos - P.bepos satisfy 0 <= gA <= 128, 0 <= gB <= 250 and
 *   |gA - gB| * cdiff <= max(16 * cdiff, 6 * max(gA, gB))      (cdiff = round(6 / (1 - e)): the diagonal may drift by
 *   the alignment's own error budget, at least 16).  The test uses the ORIGINAL neighbours, so a run of bridges folds
 *   into one record.  The bridge is the unit-cost global alignment of A[P.aepos, Q.abpos) with B[P.bepos, Q.bbpos),
 *   traceback from the end preferring diagonal, then "A base unmatched", then "B base inserted" (as in the consensus).
 *   Merged record: P's begin, Q's end, diffs = sum of both + the bridge's cost, P's flags.  Trace: the concatenated path
 *   cut at the multiples of ts of A where it FIRST reaches them -- P's tiles, the bridge's pieces, Q's tiles, pieces
 *   that share a tile added up.
 * out / out_toff need nla entries, out_trace the input's trace length + 2 * (number of bridges) * (128 / ts + 2).
 * Returns the number of output records; *nbridged = bridges made. */
static void bridge_path(const uint8_t *a, int n, const uint8_t *b, int m, int a0 /* A coordinate of row 0 */, int ts,
                        int *nrow, int *row_j, int *row_d, int *total)
{
    /* rows r in (0, n] with (a0 + r) % ts == 0, increasing: column and cost at the first arrival on the row */
    static __thread uint8_t D[128 + 1][256];
    for (int j = 0; j <= m; j++) D[0][j] = (uint8_t)j;
    for (int i = 1; i <= n; i++) {
        D[i][0] = (uint8_t)i;
        for (int j = 1; j <= m; j++) {
            int d = D[i - 1][j - 1] + (a[i - 1] != b[j - 1]);
            int u = D[i - 1][j] + 1, l = D[i][j - 1] + 1;
            int v = d; if (u < v) v = u; if (l < v) v = l;
            D[i][j] = (uint8_t)v;
        }
    }
    *total = D[n][m];
    int cnt = 0;
    for (int r = 1; r <= n; r++) if ((a0 + r) % ts == 0) cnt++;
    *nrow = cnt;
    int i = n, j = m, k = cnt;
    while (i > 0 || j > 0) {
        int dir;
        if (i > 0 && j > 0 && D[i][j] == D[i - 1][j - 1] + (a[i - 1] != b[j - 1])) dir = 0;
        else if (i > 0 && D[i][j] == D[i - 1][j] + 1) dir = 1;
        else dir = 2;
        if (dir != 2 && (a0 + i) % ts == 0) { k--; row_j[k] = j; row_d[k] = D[i][j]; }   /* leaving row i upwards: its first cell */
        if (dir == 0) { i--; j--; } else if (dir == 1) i--; else j--;
    }
}

int64_t orc_bridge(const int64_t *aoff, const uint8_t *abases, const int64_t *boff, const uint8_t *bbases,
                   const las_rec *la, int64_t nla, const int64_t *toff, const uint16_t *trace, int32_t ts, int32_t cdiff,
                   las_rec *out, int64_t *out_toff, uint16_t *out_trace, int64_t *nbridged)
{
    int64_t no = 0, to = 0, nb = 0;
    uint8_t *brc = NULL; int brc_cap = 0;
    int open = 0;                                   /* the last tile written still takes pieces (does not end on the grid) */
    for (int64_t r = 0; r < nla; r++) {
        const las_rec *q = &la[r];
        int bridged = 0;
        if (r > 0) {
            const las_rec *p = &la[r - 1];
            const int gA = q->abpos - p->aepos, gB = q->bbpos - p->bepos;
            const int dg = gA > gB ? gA - gB : gB - gA, mx = gA > gB ? gA : gB;
            const int lim = 16 * cdiff > 6 * mx ? 16 * cdiff : 6 * mx;
            if (p->aread == q->aread && p->bread == q->bread && ((p->flags ^ q->flags) & 1u) == 0 &&
                gA >= 0 && gA <= 128 && gB >= 0 && gB <= 250 && dg * cdiff <= lim) bridged = 1;
        }
        const uint16_t *tq = trace + toff[r];
        const int ntq = q->tlen / 2;
        if (!bridged) {
            out[no] = *q; out_toff[no] = to;
            for (int t = 0; t < 2 * ntq; t++) out_trace[to++] = tq[t];
            open = q->aepos % ts != 0;
            no++;
            continue;
        }
        nb++;
        las_rec *o = &out[no - 1];
        const las_rec *p = &la[r - 1];
        const int LB = (int)(boff[q->bread + 1] - boff[q->bread]);
        const uint8_t *A = abases + aoff[q->aread], *B = bbases + boff[q->bread];
        if (q->flags & 1u) {
            if (LB > brc_cap) { brc = realloc(brc, LB + 1); brc_cap = LB; }
            for (int i = 0; i < LB; i++) brc[i] = 3 - B[LB - 1 - i];
            B = brc;
        }
        const int gA = q->abpos - p->aepos, gB = q->bbpos - p->bepos;
        int nrow = 0, row_j[130], row_d[130], total = 0;
        bridge_path(A + p->aepos, gA, B + p->bepos, gB, p->aepos, ts, &nrow, row_j, row_d, &total);
        /* pieces of the bridge: up to each grid row, then the rest up to (gA, gB) */
        int pj = 0, pd = 0;
        for (int k = 0; k <= nrow; k++) {
            const int ej = k < nrow ? row_j[k] : gB, ed = k < nrow ? row_d[k] : total;
            const int dd = ed - pd, bb = ej - pj;
            if (k < nrow || dd > 0 || bb > 0) {     /* the last piece is empty when the path ends with its arrival on a grid row */
                if (open) { out_trace[to - 2] += (uint16_t)dd; out_trace[to - 1] += (uint16_t)bb; }
                else { out_trace[to++] = (uint16_t)dd; out_trace[to++] = (uint16_t)bb; open = 1; }
                if (k < nrow) open = 0;             /* the piece ends on the grid: the tile is complete */
            }
            pj = ej; pd = ed;
        }
        /* Q's tiles: the first joins an open tile */
        for (int t = 0; t < ntq; t++) {
            if (t == 0 && open) { out_trace[to - 2] += tq[0]; out_trace[to - 1] += tq[1]; }
            else { out_trace[to++] = tq[2 * t]; out_trace[to++] = tq[2 * t + 1]; }
        }
        open = q->aepos % ts != 0;
        o->aepos = q->aepos; o->bepos = q->bepos; o->diffs += total + q->diffs;
        o->tlen = (int32_t)(to - out_toff[no - 1]);
    }
    free(brc);
    if (nbridged) *nbridged = nb;
    return no;
}
