/*
 * align_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the local-alignment hot path that DENTIST reaches through
 *   getDalignment()  source/dentist/dazzler.d:3829-3844  (-> dalign()   :6131-6140)
 *   getDamapping()   source/dentist/dazzler.d:3855-3866  (-> damapper() :6163-6170)
 * i.e. what the external tools `daligner` / `damapper` compute.
 *
 * PARITY STATUS: "parity unpinned" for the third-party arithmetic.  The algorithm lives in
 * thegenemyers/DALIGNER @ c2b47da6b3c9 and thegenemyers/DAMAPPER @ b2c9d7fd64bb
 * (conda/recipes/{daligner,damapper}/meta.yaml); neither source nor binaries are under
 * /root/reference, and the reference repo holds no golden LAS for given sequences.  This file
 * therefore restates the PUBLISHED algorithm (Myers, WABI 2014: k-mer tuple sort -> merge ->
 * diagonal-band filter -> furthest-reaching O(ND) wave extension with trace points every
 * `tspace` A-bases) as one fully deterministic specification, and the CUDA path must match it
 * bit for bit.  What IS pinned against the reference's own vectors is the output contract:
 * the trace-point/tile semantics (base.d:185-242, KAT base.d:881-944) and the LAS record
 * layout (dazzler.d:1988-2032) -- see oracle/las.py and tests/test_oracle_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * use this file.  The product (dentist_b200/csrc) never links or calls it.
 *
 * Specification (all integer arithmetic):
 *   1. tuples: every k-mer start p of every read (p+k <= len, no masked base in [p,p+k));
 *      B contributes both strands (strand 1 = reverse complement, coordinates in the
 *      complemented read's frame, as daligner does by complementing the block).
 *   2. k-mers that occur more than `t` times in A are ignored.
 *   3. hit = (aread, bread, strand, apos, bpos) for every equal k-mer pair; self-pairs
 *      (aread == bread) are dropped when `self` is set.
 *   4. per (bread,strand,aread) hits are ordered by (diagonal, apos); each hit adds
 *      min(k, apos - prev.apos on same diagonal) covered bases to its band (diagonal >> w);
 *      a band is hot when it and a neighbour band together cover >= h bases; a cluster is a
 *      maximal run of consecutive hot bands; its seed is the median hit of the cluster.
 *   5. extension: forward O(ND) furthest-reaching waves from the seed in both directions
 *      (the backward one runs forward over the reverse-complemented pair). Score
 *      S = 3*(i+j) - C*d with C = round(6/(1-e)); the path ends at the maximum S (earliest
 *      wave, then lowest diagonal on ties); cells with S < best - xdrop die; the live window
 *      is capped at `wmax` diagonals around the wave's best cell.
 *   6. trace: every time a cell's A coordinate reaches a multiple of tspace (absolute A
 *      coordinate) it records (B offset, diffs so far); tiles are differences of records.
 *   7. an alignment is kept when (aepos-abpos)+(bepos-bbpos) >= 2*minlen; contained
 *      duplicates are dropped; hits covered by a kept alignment are retired and the whole
 *      select/extend step repeats on the remaining hits for up to `rounds` rounds; a round that
 *      keeps no alignment ends the loop (its clusters would only be retried one hit poorer).
 *   8. output order = LAsort order (aread, bread, comp, abpos, aepos, bbpos, bepos, diffs)
 *      = FlatLocalAlignment.opCmp, base.d:1787-1809.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

typedef struct {
    int32_t k;        /* k-mer length (<= 31)                       */
    int32_t w;        /* log2 band width                             */
    int32_t h;        /* min covered bases in a band pair            */
    int32_t t;        /* max k-mer multiplicity in A                 */
    int32_t tspace;   /* trace spacing                               */
    int32_t minlen;   /* -l                                          */
    int32_t cdiff;    /* C in S = 3*(i+j) - C*d                      */
    int32_t xdrop;    /* X                                           */
    int32_t wmax;     /* max live diagonals                          */
    int32_t rounds;   /* max select/extend rounds                    */
    int32_t self;     /* drop aread == bread                         */
    int32_t poolmul;  /* record pool = poolmul * (ntiles_bound + 2)  */
} orc_params;

typedef struct {
    int32_t nreads;
    const int64_t *off;     /* nreads+1, base offsets into bases[] */
    const uint8_t *bases;   /* 0..3 */
    const uint8_t *mask;    /* per base 0/1 or NULL */
    const int32_t *group;   /* per read pile id or NULL: hits between different groups are dropped */
} orc_block;

typedef struct {
    int32_t tlen, diffs, abpos, bbpos, aepos, bepos;
    uint32_t flags;
    int32_t aread, bread;
    int32_t toff;           /* offset (in uint16 units) into trace[] */
} orc_la;

typedef struct {
    int64_t nla;
    orc_la *la;
    int64_t ntrace;         /* uint16 elements */
    uint16_t *trace;
    int64_t nhits, nseeds, next;   /* statistics */
} orc_result;

/* ---------------------------------------------------------------- helpers */

typedef struct { uint64_t kmer; int32_t read; int32_t pos; int32_t strand; } tup_t;
typedef struct { int32_t bs; int32_t a; int32_t diag; int32_t apos; int32_t bpos; int32_t free_; } hit_t;

static int cmp_tup(const void *x, const void *y) {
    const tup_t *a = x, *b = y;
    if (a->kmer != b->kmer) return a->kmer < b->kmer ? -1 : 1;
    if (a->read != b->read) return a->read < b->read ? -1 : 1;
    if (a->strand != b->strand) return a->strand < b->strand ? -1 : 1;
    return (a->pos > b->pos) - (a->pos < b->pos);
}
static int cmp_hit(const void *x, const void *y) {
    const hit_t *a = x, *b = y;
    if (a->bs != b->bs) return a->bs < b->bs ? -1 : 1;
    if (a->a != b->a) return a->a < b->a ? -1 : 1;
    if (a->diag != b->diag) return a->diag < b->diag ? -1 : 1;
    return (a->apos > b->apos) - (a->apos < b->apos);
}
static int cmp_la(const void *x, const void *y) {
    const orc_la *a = x, *b = y;
#define C_(f) if (a->f != b->f) return a->f < b->f ? -1 : 1;
    C_(aread) C_(bread)
    { int ca = a->flags & 1, cb = b->flags & 1; if (ca != cb) return ca < cb ? -1 : 1; }
    C_(abpos) C_(aepos) C_(bbpos) C_(bepos) C_(diffs)
#undef C_
    return 0;
}

static uint8_t *revcomp_block(const orc_block *B) {
    int64_t n = B->off[B->nreads];
    uint8_t *rc = malloc(n > 0 ? n : 1);
    for (int r = 0; r < B->nreads; r++) {
        int64_t o = B->off[r]; int len = (int)(B->off[r + 1] - o);
        for (int i = 0; i < len; i++) rc[o + i] = 3 - B->bases[o + len - 1 - i];
    }
    return rc;
}

static int64_t emit_tuples(const orc_block *B, const uint8_t *seq, int strand, int k, tup_t *out) {
    int64_t n = 0;
    for (int r = 0; r < B->nreads; r++) {
        int64_t o = B->off[r]; int len = (int)(B->off[r + 1] - o);
        for (int p = 0; p + k <= len; p++) {
            uint64_t km = 0; int ok = 1;
            for (int i = 0; i < k; i++) {
                km = (km << 2) | seq[o + p + i];
                if (B->mask) {
                    /* mask is given in forward coordinates */
                    int64_t f = strand ? (o + len - 1 - (p + i)) : (o + p + i);
                    if (B->mask[f]) ok = 0;
                }
            }
            if (!ok) continue;
            if (out) { out[n].kmer = km; out[n].read = r; out[n].pos = p; out[n].strand = strand; }
            n++;
        }
    }
    return n;
}

/* ------------------------------------------------------------ extension */

typedef struct { int32_t prev, j, d; } rec_t;

typedef struct {
    int32_t i_end, j_end, d_end, ntiles;   /* ntiles includes the trailing partial tile */
} ext_out;

#define NEGV (-(1 << 29))

/* a, b point at the start position; la, lb = remaining lengths; firstT = relative A offset of
 * the first tile boundary (> 0).  tile_bb / tile_df receive tiles from the seed outward. */
static void extend(const uint8_t *a, int la, const uint8_t *b, int lb, int firstT,
                   const orc_params *P, int poolcap, rec_t *pool,
                   int *Vb, int *Tb, int span,
                   ext_out *out, int32_t *tile_bb, int32_t *tile_df)
{
    const int ts = P->tspace, C = P->cdiff, X = P->xdrop, WM = P->wmax;
    /* V arrays indexed by k + span (two buffers of size 2*span+3) */
    int *V0 = Vb, *V1 = Vb + (2 * span + 3), *T0 = Tb, *T1 = Tb + (2 * span + 3);
#define NB(i) ((i) >= firstT ? ((i) - firstT) / ts + 1 : 0)
    int npool = 0;
    int lo = 0, hi = 0, d = 0;
    int bestS, besti, bestk, bestd, bestT;
    /* wave 0 */
    {
        int i = 0, lim = la < lb ? la : lb;
        while (i < lim && a[i] == b[i]) i++;
        int T = -1, n = NB(i);
        if (n > poolcap) { out->i_end = out->j_end = out->d_end = out->ntiles = 0; return; }
        for (int q = 1; q <= n; q++) {
            pool[npool].prev = T; pool[npool].j = firstT + (q - 1) * ts; pool[npool].d = 0; T = npool++;
        }
        V0[0 + span] = i; T0[0 + span] = T;
        bestS = 3 * (2 * i); besti = i; bestk = 0; bestd = 0; bestT = T;
        if (i == la || i == lb) { lo = 1; hi = 0; }
    }
    int *Vo = V0, *Vn = V1, *To = T0, *Tn = T1;
    while (lo <= hi) {
        d++;
        int nlo = lo - 1, nhi = hi + 1;
        if (-nlo >= span || nhi >= span) break;            /* cannot happen: span >= dmax+2 */
        /* pass 1: compute cells, count records */
        int need = 0;
        for (int k = nlo; k <= nhi; k++) {
            int vs = (k >= lo && k <= hi) ? Vo[k + span] : NEGV;
            int vd = (k - 1 >= lo && k - 1 <= hi) ? Vo[k - 1 + span] : NEGV;
            int vi = (k + 1 >= lo && k + 1 <= hi) ? Vo[k + 1 + span] : NEGV;
            int i = NEGV, pi = NEGV, pT = -1;
            if (vs > NEGV) { i = vs + 1; pi = vs; pT = To[k + span]; }
            if (vd > NEGV && vd + 1 > i) { i = vd + 1; pi = vd; pT = To[k - 1 + span]; }
            if (vi > NEGV && vi > i) { i = vi; pi = vi; pT = To[k + 1 + span]; }
            if (i > NEGV) { int j = i - k; if (i > la || j > lb || j < 0) i = NEGV; }
            if (i > NEGV) {
                int j = i - k, lim = (la - i) < (lb - j) ? (la - i) : (lb - j), s = 0;
                while (s < lim && a[i + s] == b[j + s]) s++;
                i += s;
                need += NB(i) - NB(pi);
            }
            Vn[k + span] = i; Tn[k + span] = pT;            /* T fixed up in pass 2 */
            /* stash predecessor i in To? -- recompute in pass 2 instead */
        }
        if (npool + need > poolcap) break;                   /* pool exhausted: stop before this wave */
        /* pass 2: write records, score */
        int waveS = -(1 << 30), wavek = 0;
        for (int k = nlo; k <= nhi; k++) {
            int i = Vn[k + span];
            if (i == NEGV) continue;
            /* recover predecessor i (same selection as pass 1) */
            int vs = (k >= lo && k <= hi) ? Vo[k + span] : NEGV;
            int vd = (k - 1 >= lo && k - 1 <= hi) ? Vo[k - 1 + span] : NEGV;
            int vi = (k + 1 >= lo && k + 1 <= hi) ? Vo[k + 1 + span] : NEGV;
            int c = NEGV, pi = NEGV;
            if (vs > NEGV) { c = vs + 1; pi = vs; }
            if (vd > NEGV && vd + 1 > c) { c = vd + 1; pi = vd; }
            if (vi > NEGV && vi > c) { c = vi; pi = vi; }
            int T = Tn[k + span];
            for (int q = NB(pi) + 1; q <= NB(i); q++) {
                int bi = firstT + (q - 1) * ts;               /* boundary (relative A offset) */
                pool[npool].prev = T; pool[npool].j = bi - k; pool[npool].d = d; T = npool++;
            }
            Tn[k + span] = T;
            int S = 3 * (2 * i - k) - C * d;
            if (S > waveS) { waveS = S; wavek = k; }
            if (S > bestS) { bestS = S; besti = i; bestk = k; bestd = d; bestT = T; }
        }
        /* trim */
        int alo = 1 << 30, ahi = -(1 << 30);
        for (int k = nlo; k <= nhi; k++) {
            int i = Vn[k + span];
            if (i == NEGV) continue;
            int S = 3 * (2 * i - k) - C * d;
            int j = i - k;
            if (S < bestS - X || i == la || j == lb) { Vn[k + span] = NEGV; continue; }
            if (k < alo) alo = k;
            if (k > ahi) ahi = k;
        }
        if (alo > ahi) break;
        if (ahi - alo + 1 > WM) {
            int l2 = wavek - (WM / 2 - 1); if (l2 < alo) l2 = alo;
            int h2 = l2 + WM - 1; if (h2 > ahi) h2 = ahi;
            l2 = h2 - WM + 1; if (l2 < alo) l2 = alo;
            alo = l2; ahi = h2;
        }
        lo = alo; hi = ahi;
        { int *t_ = Vo; Vo = Vn; Vn = t_; t_ = To; To = Tn; Tn = t_; }
    }
    /* finalize: walk the record chain of the best cell */
    out->i_end = besti; out->j_end = besti - bestk; out->d_end = bestd;
    int n = NB(besti);
    int T = bestT;
    int lastj = 0, lastd = 0;
    if (T >= 0) { lastj = pool[T].j; lastd = pool[T].d; }
    for (int q = n; q >= 1; q--) {
        int pj = 0, pd = 0, pv = pool[T].prev;
        if (pv >= 0) { pj = pool[pv].j; pd = pool[pv].d; }
        tile_bb[q - 1] = pool[T].j - pj; tile_df[q - 1] = pool[T].d - pd;
        T = pv;
    }
    int lastB = n > 0 ? firstT + (n - 1) * ts : 0;
    if (besti > lastB) {
        tile_bb[n] = (besti - bestk) - lastj; tile_df[n] = bestd - lastd; n++;
    }
    out->ntiles = n;
#undef NB
}

/* upper bound used to size the record pool: an extension cannot span more A bases than
 * min(la, 1.5*lb + 64) before the x-drop rule stops it (capacity only; overflow is handled). */
static int ext_span(int la, int lb) { int64_t s = (int64_t)lb + lb / 2 + 64; return la < s ? la : (int)s; }

/* ------------------------------------------------------------ main entry */

typedef struct { int32_t a, bs, apos, bpos; } seed_t;

typedef struct {
    orc_la la; int32_t dmin, dmax; int32_t *bb; int32_t *df;
} cand_t;

int orc_align(const orc_block *A, const orc_block *B, const orc_params *P, orc_result *R)
{
    const int k = P->k, ts = P->tspace;
    memset(R, 0, sizeof *R);
    uint8_t *Arc = revcomp_block(A), *Brc = revcomp_block(B);

    /* 1. tuples */
    int64_t nta = emit_tuples(A, A->bases, 0, k, NULL);
    int64_t ntb = emit_tuples(B, B->bases, 0, k, NULL) + emit_tuples(B, Brc, 1, k, NULL);
    tup_t *TA = malloc(sizeof(tup_t) * (nta + 1)), *TB = malloc(sizeof(tup_t) * (ntb + 1));
    emit_tuples(A, A->bases, 0, k, TA);
    { int64_t n0 = emit_tuples(B, B->bases, 0, k, TB); emit_tuples(B, Brc, 1, k, TB + n0); }
    qsort(TA, nta, sizeof(tup_t), cmp_tup);
    qsort(TB, ntb, sizeof(tup_t), cmp_tup);

    /* 2+3. merge -> hits */
    int64_t nh = 0, caph = 1 << 16;
    hit_t *H = malloc(sizeof(hit_t) * caph);
    for (int64_t ia = 0, ib = 0; ia < nta && ib < ntb;) {
        if (TA[ia].kmer < TB[ib].kmer) { ia++; continue; }
        if (TA[ia].kmer > TB[ib].kmer) { ib++; continue; }
        int64_t ea = ia, eb = ib; uint64_t km = TA[ia].kmer;
        while (ea < nta && TA[ea].kmer == km) ea++;
        while (eb < ntb && TB[eb].kmer == km) eb++;
        if (ea - ia <= P->t) {
            for (int64_t x = ia; x < ea; x++) for (int64_t y = ib; y < eb; y++) {
                if (P->self && TA[x].read == TB[y].read) continue;
                if (A->group && B->group && A->group[TA[x].read] != B->group[TB[y].read]) continue;
                if (nh == caph) { caph *= 2; H = realloc(H, sizeof(hit_t) * caph); }
                hit_t *q = &H[nh++];
                q->a = TA[x].read; q->bs = TB[y].read * 2 + TB[y].strand;
                q->apos = TA[x].pos; q->bpos = TB[y].pos; q->diag = TA[x].pos - TB[y].pos; q->free_ = 1;
            }
        }
        ia = ea; ib = eb;
    }
    free(TA); free(TB);
    qsort(H, nh, sizeof(hit_t), cmp_hit);
    R->nhits = nh;

    /* work buffers for the extension */
    int maxlenA = 1, maxlenB = 1;
    for (int r = 0; r < A->nreads; r++) { int l = (int)(A->off[r + 1] - A->off[r]); if (l > maxlenA) maxlenA = l; }
    for (int r = 0; r < B->nreads; r++) { int l = (int)(B->off[r + 1] - B->off[r]); if (l > maxlenB) maxlenB = l; }
    int span = maxlenA + maxlenB + 4;
    int *Vb = malloc(sizeof(int) * 2 * (2 * span + 3)), *Tb = malloc(sizeof(int) * 2 * (2 * span + 3));
    int maxtiles = maxlenA / ts + 3;
    int32_t *fbb = malloc(sizeof(int32_t) * maxtiles * 4);
    int32_t *fdf = fbb + maxtiles, *rbb = fdf + maxtiles, *rdf = rbb + maxtiles;
    rec_t *pool = malloc(sizeof(rec_t) * (size_t)P->poolmul * (maxtiles + 6));

    int64_t ncand = 0, capc = 1024;
    cand_t *Cn = malloc(sizeof(cand_t) * capc);       /* kept alignments (all rounds) */

    for (int round = 0; round < P->rounds; round++) {
        /* 4. band filter over the free hits (H stays sorted; retired hits are removed) */
        int64_t nseeds = 0; seed_t *S = malloc(sizeof(seed_t) * (nh + 1));
        int64_t g0 = 0;
        while (g0 < nh) {
            int64_t g1 = g0;
            while (g1 < nh && H[g1].bs == H[g0].bs && H[g1].a == H[g0].a) g1++;
            /* bands inside the group */
            int64_t nb = 0; int64_t *bfirst = malloc(sizeof(int64_t) * (g1 - g0 + 1));
            int32_t *bkey = malloc(sizeof(int32_t) * (g1 - g0)), *bsc = malloc(sizeof(int32_t) * (g1 - g0));
            for (int64_t x = g0; x < g1; x++) {
                /* shift diagonal to be non-negative before banding */
                int32_t band = (H[x].diag + (1 << 30)) >> P->w;
                int c = k;
                if (x > g0 && H[x - 1].diag == H[x].diag && H[x].apos - H[x - 1].apos < k) c = H[x].apos - H[x - 1].apos;
                if (nb == 0 || bkey[nb - 1] != band) { bkey[nb] = band; bsc[nb] = 0; bfirst[nb] = x; nb++; }
                bsc[nb - 1] += c;
            }
            bfirst[nb] = g1;
            uint8_t *hot = calloc(nb + 1, 1);
            for (int64_t q = 0; q < nb; q++) {
                int p = bsc[q]; int adj = (q + 1 < nb && bkey[q + 1] == bkey[q] + 1);
                if (adj) p += bsc[q + 1];
                if (p >= P->h) { hot[q] = 1; if (adj) hot[q + 1] = 1; }
            }
            for (int64_t q = 0; q < nb;) {
                if (!hot[q]) { q++; continue; }
                int64_t e = q;
                while (e + 1 < nb && hot[e + 1] && bkey[e + 1] == bkey[e] + 1) e++;
                int64_t f = bfirst[q], l = bfirst[e + 1];       /* hits [f,l) */
                int64_t m = f + (l - f - 1) / 2;
                S[nseeds].a = H[m].a; S[nseeds].bs = H[m].bs; S[nseeds].apos = H[m].apos; S[nseeds].bpos = H[m].bpos;
                nseeds++;
                H[m].free_ = 0;                                  /* a seed is consumed */
                q = e + 1;
            }
            free(hot); free(bfirst); free(bkey); free(bsc);
            g0 = g1;
        }
        R->nseeds += nseeds;
        if (nseeds == 0) { free(S); break; }

        /* 5. extend every seed */
        int64_t nnew0 = ncand;
        for (int64_t s = 0; s < nseeds; s++) {
            int a = S[s].a, bs = S[s].bs, br = bs >> 1, st = bs & 1;
            int la = (int)(A->off[a + 1] - A->off[a]), lb = (int)(B->off[br + 1] - B->off[br]);
            const uint8_t *af = A->bases + A->off[a], *ar = Arc + A->off[a];
            const uint8_t *bf = (st ? Brc : B->bases) + B->off[br], *brv = (st ? B->bases : Brc) + B->off[br];
            int ap = S[s].apos, bp = S[s].bpos;
            ext_out fo, ro;
            int firstT = (ap / ts + 1) * ts - ap;
            int poolcap = P->poolmul * (ext_span(la - ap, lb - bp) / ts + 4);
            extend(af + ap, la - ap, bf + bp, lb - bp, firstT, P, poolcap, pool, Vb, Tb, span, &fo, fbb, fdf);
            R->next++;
            if (ap > 0 && bp > 0) {
                int firstTr = ap - ((ap - 1) / ts) * ts;          /* distance down to the boundary below ap */
                poolcap = P->poolmul * (ext_span(ap, bp) / ts + 4);
                extend(ar + (la - ap), ap, brv + (lb - bp), bp, firstTr, P, poolcap, pool, Vb, Tb, span, &ro, rbb, rdf);
                R->next++;
            } else { ro.i_end = ro.j_end = ro.d_end = ro.ntiles = 0; }
            int ab = ap - ro.i_end, bb = bp - ro.j_end, ae = ap + fo.i_end, be = bp + fo.j_end;
            if ((ae - ab) + (be - bb) < 2 * P->minlen) continue;
            /* join tiles: reverse part (outermost first) then forward part */
            int merge = (ap % ts != 0) && ro.ntiles > 0 && fo.ntiles > 0;
            int nt = ro.ntiles + fo.ntiles - (merge ? 1 : 0);
            if (ncand == capc) { capc *= 2; Cn = realloc(Cn, sizeof(cand_t) * capc); }
            cand_t *c = &Cn[ncand];
            c->bb = malloc(sizeof(int32_t) * (nt + 1) * 2); c->df = c->bb + nt + 1;
            int o = 0;
            for (int q = ro.ntiles - 1; q >= (merge ? 1 : 0); q--) { c->bb[o] = rbb[q]; c->df[o] = rdf[q]; o++; }
            if (merge) { c->bb[o] = rbb[0] + fbb[0]; c->df[o] = rdf[0] + fdf[0]; o++; }
            for (int q = merge ? 1 : 0; q < fo.ntiles; q++) { c->bb[o] = fbb[q]; c->df[o] = fdf[q]; o++; }
            c->la.tlen = 2 * nt; c->la.diffs = fo.d_end + ro.d_end;
            c->la.abpos = ab; c->la.bbpos = bb; c->la.aepos = ae; c->la.bepos = be;
            c->la.flags = st ? 1u : 0u; c->la.aread = a; c->la.bread = br; c->la.toff = 0;
            /* diagonal range over tile boundaries */
            { int dg = ab - bb, mn = dg, mx = dg, apos_ = ab, bpos_ = bb;
              for (int q = 0; q < nt; q++) {
                  int aend = (q == nt - 1) ? ae : (apos_ / ts + 1) * ts;
                  bpos_ += c->bb[q]; apos_ = aend; dg = apos_ - bpos_;
                  if (dg < mn) mn = dg;
                  if (dg > mx) mx = dg;
              }
              c->dmin = mn; c->dmax = mx; }
            ncand++;
        }
        free(S);

        /* 7b. retire hits covered by any alignment found this round (kept or dropped --
         * a dropped one is contained in a kept one anyway) */
        {
            int64_t o = 0;
            for (int64_t x = 0; x < nh; x++) {
                int keep = H[x].free_;
                for (int64_t j = nnew0; j < ncand && keep; j++) {
                    orc_la *y = &Cn[j].la;
                    if (y->aread != H[x].a || y->bread * 2 + (int)(y->flags & 1) != H[x].bs) continue;
                    if (H[x].apos >= y->abpos && H[x].apos <= y->aepos &&
                        H[x].diag >= Cn[j].dmin - (1 << P->w) && H[x].diag <= Cn[j].dmax + (1 << P->w)) keep = 0;
                }
                if (keep) H[o++] = H[x];
            }
            nh = o;
        }
        if (ncand == nnew0) break;                     /* nothing kept this round: stop */
    }

    /* 7a. drop contained duplicates among ALL candidates of the same (a, b, strand):
     * candidate j is dropped when any other candidate i (dropped or not -- containment is
     * transitive, so the result does not depend on evaluation order) contains it in both
     * coordinates; for identical intervals the lower candidate index wins. */
    {
        uint8_t *dropf = calloc(ncand + 1, 1);
        for (int64_t j = 0; j < ncand; j++) {
            orc_la *x = &Cn[j].la;
            for (int64_t i = 0; i < ncand && !dropf[j]; i++) {
                if (i == j) continue;
                orc_la *y = &Cn[i].la;
                if (y->aread != x->aread || y->bread != x->bread || (y->flags & 1) != (x->flags & 1)) continue;
                if (y->abpos <= x->abpos && x->aepos <= y->aepos && y->bbpos <= x->bbpos && x->bepos <= y->bepos) {
                    int same = (y->abpos == x->abpos && x->aepos == y->aepos && y->bbpos == x->bbpos && x->bepos == y->bepos);
                    if (!same || i < j) dropf[j] = 1;
                }
            }
        }
        for (int64_t j = 0; j < ncand; j++) if (dropf[j]) Cn[j].la.tlen = -1;
        free(dropf);
    }

    /* 8. output */
    int64_t nla = 0, ntr = 0;
    for (int64_t j = 0; j < ncand; j++) if (Cn[j].la.tlen >= 0) { nla++; ntr += Cn[j].la.tlen; }
    R->la = malloc(sizeof(orc_la) * (nla + 1)); R->trace = malloc(sizeof(uint16_t) * (ntr + 1));
    { int64_t o = 0;
      for (int64_t j = 0; j < ncand; j++) if (Cn[j].la.tlen >= 0) { R->la[o] = Cn[j].la; R->la[o].toff = (int32_t)j; o++; } }
    qsort(R->la, nla, sizeof(orc_la), cmp_la);
    { int64_t to = 0;
      for (int64_t o = 0; o < nla; o++) {
          cand_t *c = &Cn[R->la[o].toff]; int nt = c->la.tlen / 2;
          R->la[o].toff = (int32_t)to;
          for (int q = 0; q < nt; q++) { R->trace[to++] = (uint16_t)c->df[q]; R->trace[to++] = (uint16_t)c->bb[q]; }
      } }
    R->nla = nla; R->ntrace = ntr;
    for (int64_t j = 0; j < ncand; j++) free(Cn[j].bb);
    free(Cn); free(H); free(Vb); free(Tb); free(fbb); free(pool); free(Arc); free(Brc);
    return 0;
}

void orc_free(orc_result *R) { free(R->la); free(R->trace); memset(R, 0, sizeof *R); }
