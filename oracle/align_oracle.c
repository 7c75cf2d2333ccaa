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
 *      select/extend step repeats on the remaining hits for up to `rounds` rounds.  The rounds are
 *      counted per (bread, strand, aread) group: a group whose round selects no seed or keeps no
 *      alignment is finished (its clusters would only be retried one hit poorer), whatever the other
 *      groups of the block do -- so the output for a read never depends on the reads it shares a
 *      block with (shard invariance, tests/test_gpu_parity.py::test_shard_invariance).
 *   8. output order = LAsort order (aread, bread, comp, abpos, aepos, bbpos, bepos, diffs)
 *      = FlatLocalAlignment.opCmp, base.d:1787-1809.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>

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
    const uint64_t kmask = k < 32 ? (((uint64_t)1 << (2 * k)) - 1) : ~(uint64_t)0;
    for (int r = 0; r < B->nreads; r++) {
        int64_t o = B->off[r]; int len = (int)(B->off[r + 1] - o);
        uint64_t km = 0; int valid = 0;                  /* valid = consecutive unmasked bases ending at q */
        for (int q = 0; q < len; q++) {
            km = ((km << 2) | seq[o + q]) & kmask;
            int masked = 0;
            if (B->mask) { int64_t f = strand ? (o + len - 1 - q) : (o + q); masked = B->mask[f]; }   /* mask is given in forward coordinates */
            valid = masked ? 0 : valid + 1;
            if (valid < k) continue;                     /* also covers q + 1 < k */
            if (out) { out[n].kmer = km; out[n].read = r; out[n].pos = q + 1 - k; out[n].strand = strand; }
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

typedef struct { cand_t *c; int64_t n, cap; } candvec_t;

static void cv_push(candvec_t *v, const cand_t *c) {
    if (v->n == v->cap) { v->cap = v->cap ? 2 * v->cap : 256; v->c = realloc(v->c, sizeof(cand_t) * v->cap); }
    v->c[v->n++] = *c;
}

/* The sorted A index: tuples ordered by (kmer, read, pos) -- emitted in (read, pos) order and radix sorted
 * stably by k-mer -- plus a direct-address table over the top `tb` k-mer bits.  Built once per call and shared,
 * read-only, by the threads that each map their own B reads. */
typedef struct { tup_t *t; int64_t n; int64_t *tbl; int tb, sh; } aindex_t;

/* stable LSD byte radix sort of t[0..n) on k-mer bits [0, bits); tmp holds n tuples */
static void radix_sort_tuples(tup_t *t, tup_t *tmp, int64_t n, int bits) {
    tup_t *src = t, *dst = tmp;
    for (int lo = 0; lo < bits; lo += 8) {
        int64_t cnt[257]; memset(cnt, 0, sizeof cnt);
        for (int64_t i = 0; i < n; i++) cnt[((src[i].kmer >> lo) & 255) + 1]++;
        for (int d = 0; d < 256; d++) cnt[d + 1] += cnt[d];
        for (int64_t i = 0; i < n; i++) dst[cnt[(src[i].kmer >> lo) & 255]++] = src[i];
        tup_t *x = src; src = dst; dst = x;
    }
    if (src != t) memcpy(t, src, sizeof(tup_t) * n);
}

typedef struct { tup_t *t, *tmp; const int64_t *beg; int nb, bits; volatile int next; } sortjob_t;
static void *sort_worker(void *arg) {
    sortjob_t *J = arg;
    for (;;) {
        const int b = __sync_fetch_and_add(&J->next, 1);
        if (b >= J->nb) break;
        radix_sort_tuples(J->t + J->beg[b], J->tmp + J->beg[b], J->beg[b + 1] - J->beg[b], J->bits);
    }
    return NULL;
}

/* One stable MSD split on the top k-mer bits (sequential: keeps the (read, pos) emission order inside every
 * bucket), then the buckets are LSD-sorted independently by the host threads. */
static void build_index(const orc_block *A, int k, aindex_t *X, int nthreads) {
    X->n = emit_tuples(A, A->bases, 0, k, NULL);
    tup_t *raw = malloc(sizeof(tup_t) * (X->n + 1));
    X->t = malloc(sizeof(tup_t) * (X->n + 1));
    emit_tuples(A, A->bases, 0, k, raw);
    const int top = 2 * k < 8 ? 2 * k : 8, low = 2 * k - top, nb = 1 << top;
    int64_t beg[257]; memset(beg, 0, sizeof beg);
    for (int64_t i = 0; i < X->n; i++) beg[(raw[i].kmer >> low) + 1]++;
    for (int d = 0; d < nb; d++) beg[d + 1] += beg[d];
    { int64_t at[256]; memcpy(at, beg, sizeof(int64_t) * nb);
      for (int64_t i = 0; i < X->n; i++) X->t[at[raw[i].kmer >> low]++] = raw[i]; }
    sortjob_t J; J.t = X->t; J.tmp = raw; J.beg = beg; J.nb = nb; J.bits = low; J.next = 0;
    pthread_t th[256]; if (nthreads > 256) nthreads = 256;
    for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, sort_worker, &J);
    sort_worker(&J);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    free(raw);
    X->tb = 2 * k < 24 ? 2 * k : 24; X->sh = 2 * k - X->tb;
    int64_t nq = (int64_t)1 << X->tb;
    X->tbl = malloc(sizeof(int64_t) * (nq + 1));
    int64_t i = 0;
    for (int64_t q = 0; q <= nq; q++) {
        while (i < X->n && (int64_t)(X->t[i].kmer >> X->sh) < q) i++;
        X->tbl[q] = i;
    }
}

/* [s, e) = entries of the index with this k-mer */
static void index_range(const aindex_t *X, uint64_t km, int64_t *s, int64_t *e) {
    int64_t lo = X->tbl[km >> X->sh], hi = X->tbl[(km >> X->sh) + 1], a = lo, b = hi;
    while (a < b) { int64_t m = (a + b) >> 1; if (X->t[m].kmer < km) a = m + 1; else b = m; }
    *s = a;
    while (a < hi && X->t[a].kmer == km) a++;
    *e = a;
}

typedef struct { int32_t a, diag, apos, bpos; } ghit_t;      /* hit inside one (bread, strand) segment */
static int cmp_ghit(const void *x, const void *y) {
    const ghit_t *a = x, *b = y;
    if (a->a != b->a) return a->a < b->a ? -1 : 1;
    if (a->diag != b->diag) return a->diag < b->diag ? -1 : 1;
    return (a->apos > b->apos) - (a->apos < b->apos);
}

/* per-thread work buffers of the extension */
typedef struct {
    int *Vb, *Tb; int span; int32_t *fbb, *fdf, *rbb, *rdf; rec_t *pool;
    ghit_t *hits; int64_t caph; seed_t *S; int64_t capS;
    int64_t *bfirst; int32_t *bkey, *bsc; uint8_t *hot; int64_t capb;
    int64_t nhits, nseeds, next;
} work_t;

/* Steps 4-7 for ONE (bread, strand, aread) group: its hits H[0..nh) sorted by (diag, apos).  The rounds of a
 * group end when a round selects no seed or keeps no alignment -- the rule is evaluated per group, so a read's
 * output never depends on which other reads share its block (shard invariance). */
static void process_group(const orc_block *A, const orc_block *B, const uint8_t *Arc, const uint8_t *Brc, const orc_params *P,
                          work_t *W, ghit_t *H, int64_t nh, int a, int br, int st, candvec_t *out)
{
    const int k = P->k, ts = P->tspace;
    const int la = (int)(A->off[a + 1] - A->off[a]), lb = (int)(B->off[br + 1] - B->off[br]);
    const uint8_t *af = A->bases + A->off[a], *ar = Arc + A->off[a];
    const uint8_t *bf = (st ? Brc : B->bases) + B->off[br], *brv = (st ? B->bases : Brc) + B->off[br];
    const int64_t first_cand = out->n;
    if (nh + 1 > W->capb) {
        W->capb = 2 * (nh + 1);
        W->bfirst = realloc(W->bfirst, sizeof(int64_t) * (W->capb + 1)); W->bkey = realloc(W->bkey, sizeof(int32_t) * W->capb);
        W->bsc = realloc(W->bsc, sizeof(int32_t) * W->capb); W->hot = realloc(W->hot, W->capb + 1);
    }
    if (nh + 1 > W->capS) { W->capS = 2 * (nh + 1); W->S = realloc(W->S, sizeof(seed_t) * W->capS); }
    uint8_t *consumed = calloc(nh + 1, 1);
    for (int round = 0; round < P->rounds && nh > 0; round++) {
        /* 4. band filter */
        int64_t nb = 0, nseeds = 0;
        for (int64_t x = 0; x < nh; x++) {
            int32_t band = (H[x].diag + (1 << 30)) >> P->w;           /* shift the diagonal to be non-negative before banding */
            int c = k;
            if (x > 0 && H[x - 1].diag == H[x].diag && H[x].apos - H[x - 1].apos < k) c = H[x].apos - H[x - 1].apos;
            if (nb == 0 || W->bkey[nb - 1] != band) { W->bkey[nb] = band; W->bsc[nb] = 0; W->bfirst[nb] = x; nb++; }
            W->bsc[nb - 1] += c;
        }
        W->bfirst[nb] = nh;
        memset(W->hot, 0, nb + 1);
        for (int64_t q = 0; q < nb; q++) {
            int p = W->bsc[q]; int adj = (q + 1 < nb && W->bkey[q + 1] == W->bkey[q] + 1);
            if (adj) p += W->bsc[q + 1];
            if (p >= P->h) { W->hot[q] = 1; if (adj) W->hot[q + 1] = 1; }
        }
        memset(consumed, 0, nh);
        for (int64_t q = 0; q < nb;) {
            if (!W->hot[q]) { q++; continue; }
            int64_t e = q;
            while (e + 1 < nb && W->hot[e + 1] && W->bkey[e + 1] == W->bkey[e] + 1) e++;
            int64_t f = W->bfirst[q], l = W->bfirst[e + 1];            /* hits [f,l) */
            int64_t m = f + (l - f - 1) / 2;
            W->S[nseeds].a = a; W->S[nseeds].bs = 2 * br + st; W->S[nseeds].apos = H[m].apos; W->S[nseeds].bpos = H[m].bpos;
            nseeds++;
            consumed[m] = 1;                                           /* a seed is consumed */
            q = e + 1;
        }
        W->nseeds += nseeds;
        if (nseeds == 0) break;

        /* 5. extend every seed */
        const int64_t nnew0 = out->n;
        for (int64_t s = 0; s < nseeds; s++) {
            int ap = W->S[s].apos, bp = W->S[s].bpos;
            ext_out fo, ro;
            int firstT = (ap / ts + 1) * ts - ap;
            int poolcap = P->poolmul * (ext_span(la - ap, lb - bp) / ts + 4);
            extend(af + ap, la - ap, bf + bp, lb - bp, firstT, P, poolcap, W->pool, W->Vb, W->Tb, W->span, &fo, W->fbb, W->fdf);
            W->next++;
            if (ap > 0 && bp > 0) {
                int firstTr = ap - ((ap - 1) / ts) * ts;          /* distance down to the boundary below ap */
                poolcap = P->poolmul * (ext_span(ap, bp) / ts + 4);
                extend(ar + (la - ap), ap, brv + (lb - bp), bp, firstTr, P, poolcap, W->pool, W->Vb, W->Tb, W->span, &ro, W->rbb, W->rdf);
                W->next++;
            } else { ro.i_end = ro.j_end = ro.d_end = ro.ntiles = 0; }
            int ab = ap - ro.i_end, bb = bp - ro.j_end, ae = ap + fo.i_end, be = bp + fo.j_end;
            if ((ae - ab) + (be - bb) < 2 * P->minlen) continue;
            /* join tiles: reverse part (outermost first) then forward part */
            int merge = (ap % ts != 0) && ro.ntiles > 0 && fo.ntiles > 0;
            int nt = ro.ntiles + fo.ntiles - (merge ? 1 : 0);
            cand_t cn; cand_t *c = &cn;
            c->bb = malloc(sizeof(int32_t) * (nt + 1) * 2); c->df = c->bb + nt + 1;
            int o = 0;
            for (int q = ro.ntiles - 1; q >= (merge ? 1 : 0); q--) { c->bb[o] = W->rbb[q]; c->df[o] = W->rdf[q]; o++; }
            if (merge) { c->bb[o] = W->rbb[0] + W->fbb[0]; c->df[o] = W->rdf[0] + W->fdf[0]; o++; }
            for (int q = merge ? 1 : 0; q < fo.ntiles; q++) { c->bb[o] = W->fbb[q]; c->df[o] = W->fdf[q]; o++; }
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
            cv_push(out, c);
        }
        if (out->n == nnew0) break;                    /* 7. nothing kept this round: the group is finished */

        /* 7b. retire the seeds and the hits covered by an alignment kept this round */
        int64_t o = 0;
        for (int64_t x = 0; x < nh; x++) {
            int keep = !consumed[x];
            for (int64_t j = nnew0; j < out->n && keep; j++) {
                const cand_t *y = &out->c[j];
                if (H[x].apos >= y->la.abpos && H[x].apos <= y->la.aepos &&
                    H[x].diag >= y->dmin - (1 << P->w) && H[x].diag <= y->dmax + (1 << P->w)) keep = 0;
            }
            if (keep) H[o++] = H[x];
        }
        nh = o;
    }
    free(consumed);

    /* 7a. drop contained duplicates among the candidates of the group: candidate j is dropped when any other
     * candidate i (dropped or not -- containment is transitive, so the result does not depend on evaluation
     * order) contains it in both coordinates; for identical intervals the lower candidate index (earlier
     * round, then earlier cluster) wins. */
    for (int64_t j = first_cand; j < out->n; j++) {
        orc_la *x = &out->c[j].la;
        int drop = 0;
        for (int64_t i = first_cand; i < out->n && !drop; i++) {
            if (i == j) continue;
            const orc_la *y = &out->c[i].la;
            if (y->abpos <= x->abpos && x->aepos <= y->aepos && y->bbpos <= x->bbpos && x->bepos <= y->bepos) {
                int same = (y->abpos == x->abpos && x->aepos == y->aepos && y->bbpos == x->bbpos && x->bepos == y->bepos);
                if (!same || i < j) drop = 1;
            }
        }
        if (drop) x->toff = -1;                         /* tlen stays intact for the containment tests of the others */
    }
}

/* all groups of one B read (both strands) */
static void process_bread(const orc_block *A, const orc_block *B, const uint8_t *Arc, const uint8_t *Brc, const orc_params *P,
                          const aindex_t *X, work_t *W, int br, candvec_t *out)
{
    const int k = P->k;
    const int64_t o = B->off[br]; const int len = (int)(B->off[br + 1] - o);
    const uint64_t kmask = k < 32 ? (((uint64_t)1 << (2 * k)) - 1) : ~(uint64_t)0;
    for (int st = 0; st < 2; st++) {
        const uint8_t *seq = (st ? Brc : B->bases) + o;
        int64_t nh = 0;
        /* 1-3. this read's k-mers against the A index */
        uint64_t km = 0; int valid = 0;                  /* valid = consecutive unmasked bases ending at p+k-1 */
        for (int q = 0; q < len; q++) {
            km = ((km << 2) | seq[q]) & kmask;
            int masked = 0;
            if (B->mask) { int64_t f = st ? (o + len - 1 - q) : (o + q); masked = B->mask[f]; }   /* mask is in forward coordinates */
            valid = masked ? 0 : valid + 1;
            if (q + 1 < k || valid < k) continue;
            const int p = q + 1 - k;
            int64_t s, e; index_range(X, km, &s, &e);
            if (e - s == 0 || e - s > P->t) continue;
            for (int64_t x = s; x < e; x++) {
                const int a = X->t[x].read;
                if (P->self && a == br) continue;
                if (A->group && B->group && A->group[a] != B->group[br]) continue;
                if (nh == W->caph) { W->caph = W->caph ? 2 * W->caph : 4096; W->hits = realloc(W->hits, sizeof(ghit_t) * W->caph); }
                ghit_t *h = &W->hits[nh++];
                h->a = a; h->apos = X->t[x].pos; h->bpos = p; h->diag = h->apos - p;
            }
        }
        W->nhits += nh;
        /* 4. per (bread, strand, aread) group, hits ordered by (diagonal, apos) */
        qsort(W->hits, nh, sizeof(ghit_t), cmp_ghit);
        for (int64_t g0 = 0; g0 < nh;) {
            int64_t g1 = g0;
            while (g1 < nh && W->hits[g1].a == W->hits[g0].a) g1++;
            process_group(A, B, Arc, Brc, P, W, W->hits + g0, g1 - g0, W->hits[g0].a, br, st, out);
            g0 = g1;
        }
    }
}

typedef struct {
    const orc_block *A, *B; const uint8_t *Arc, *Brc; const orc_params *P; const aindex_t *X; int span, maxtiles;
    volatile int next;                  /* next chunk of B reads to hand out */
} job_t;
typedef struct { job_t *job; candvec_t *out; pthread_t th; int64_t nhits, nseeds, next; } worker_t;

static void *worker_main(void *arg) {
    worker_t *me = arg; job_t *J = me->job; const orc_params *P = J->P;
    work_t W; memset(&W, 0, sizeof W);
    W.span = J->span;
    W.Vb = malloc(sizeof(int) * 2 * (2 * J->span + 3)); W.Tb = malloc(sizeof(int) * 2 * (2 * J->span + 3));
    W.fbb = malloc(sizeof(int32_t) * J->maxtiles * 4);
    W.fdf = W.fbb + J->maxtiles; W.rbb = W.fdf + J->maxtiles; W.rdf = W.rbb + J->maxtiles;
    W.pool = malloc(sizeof(rec_t) * (size_t)P->poolmul * (J->maxtiles + 6));
    for (;;) {
        const int b0 = __sync_fetch_and_add(&J->next, 4);
        if (b0 >= J->B->nreads) break;
        for (int br = b0; br < b0 + 4 && br < J->B->nreads; br++) process_bread(J->A, J->B, J->Arc, J->Brc, P, J->X, &W, br, me->out);
    }
    me->nhits = W.nhits; me->nseeds = W.nseeds; me->next = W.next;
    free(W.Vb); free(W.Tb); free(W.fbb); free(W.pool); free(W.hits); free(W.S); free(W.bfirst); free(W.bkey); free(W.bsc); free(W.hot);
    return NULL;
}

/* `nthreads` host threads share the A index; B reads are dealt out dynamically; the result is independent of
 * the thread count (every group is processed by exactly one thread, the output is sorted at the end). */
int orc_align_mt(const orc_block *A, const orc_block *B, const orc_params *P, orc_result *R, int nthreads)
{
    const int ts = P->tspace;
    memset(R, 0, sizeof *R);
    if (nthreads < 1) nthreads = 1;
    uint8_t *Arc = revcomp_block(A), *Brc = revcomp_block(B);
    aindex_t X; build_index(A, P->k, &X, nthreads);

    int maxlenA = 1, maxlenB = 1;
    for (int r = 0; r < A->nreads; r++) { int l = (int)(A->off[r + 1] - A->off[r]); if (l > maxlenA) maxlenA = l; }
    for (int r = 0; r < B->nreads; r++) { int l = (int)(B->off[r + 1] - B->off[r]); if (l > maxlenB) maxlenB = l; }
    const int span = maxlenA + maxlenB + 4;
    const int maxtiles = maxlenA / ts + 3;

    candvec_t *CV = calloc(nthreads, sizeof(candvec_t));
    int64_t st_hits = 0, st_seeds = 0, st_ext = 0;
    job_t J; J.A = A; J.B = B; J.Arc = Arc; J.Brc = Brc; J.P = P; J.X = &X; J.span = span; J.maxtiles = maxtiles; J.next = 0;
    worker_t *wk = calloc(nthreads, sizeof(worker_t));
    for (int t = 0; t < nthreads; t++) { wk[t].job = &J; wk[t].out = &CV[t]; }
    for (int t = 1; t < nthreads; t++) pthread_create(&wk[t].th, NULL, worker_main, &wk[t]);
    worker_main(&wk[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(wk[t].th, NULL);
    for (int t = 0; t < nthreads; t++) { st_hits += wk[t].nhits; st_seeds += wk[t].nseeds; st_ext += wk[t].next; }
    free(wk);
    R->nhits = st_hits; R->nseeds = st_seeds; R->next = st_ext;

    /* 8. output in LAsort order */
    int64_t nla = 0, ntr = 0;
    for (int t = 0; t < nthreads; t++)
        for (int64_t j = 0; j < CV[t].n; j++) if (CV[t].c[j].la.toff >= 0) { nla++; ntr += CV[t].c[j].la.tlen; }
    R->la = malloc(sizeof(orc_la) * (nla + 1)); R->trace = malloc(sizeof(uint16_t) * (ntr + 1));
    cand_t **src = malloc(sizeof(cand_t *) * (nla + 1));
    orc_la *tmp = malloc(sizeof(orc_la) * (nla + 1));
    { int64_t o = 0;
      for (int t = 0; t < nthreads; t++)
          for (int64_t j = 0; j < CV[t].n; j++) if (CV[t].c[j].la.toff >= 0) { src[o] = &CV[t].c[j]; tmp[o] = CV[t].c[j].la; tmp[o].toff = (int32_t)o; o++; } }
    qsort(tmp, nla, sizeof(orc_la), cmp_la);
    { int64_t to = 0;
      for (int64_t o = 0; o < nla; o++) {
          const cand_t *c = src[tmp[o].toff]; const int nt = c->la.tlen / 2;
          R->la[o] = tmp[o]; R->la[o].toff = (int32_t)to;
          for (int q = 0; q < nt; q++) { R->trace[to++] = (uint16_t)c->df[q]; R->trace[to++] = (uint16_t)c->bb[q]; }
      } }
    R->nla = nla; R->ntrace = ntr;
    for (int t = 0; t < nthreads; t++) { for (int64_t j = 0; j < CV[t].n; j++) free(CV[t].c[j].bb); free(CV[t].c); }
    free(CV); free(src); free(tmp); free(X.t); free(X.tbl); free(Arc); free(Brc);
    return 0;
}

int orc_align(const orc_block *A, const orc_block *B, const orc_params *P, orc_result *R) { return orc_align_mt(A, B, P, R, 1); }

void orc_free(orc_result *R) { free(R->la); free(R->trace); memset(R, 0, sizeof *R); }
