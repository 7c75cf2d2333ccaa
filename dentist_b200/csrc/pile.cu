// pile.cu -- per-pile-up stages of processPileUps on the device:
//   k_las_filter   averageErrorRate filter (dazzler.d:3885-3899, base.d:1764-1767) and
//                  isValidPileUpAlignment (dazzler.d:4126-4141)
//   k_qv           per-tile intrinsic QVs (what DAScover/DASqv and computeintrinsicqv produce behind
//                  dazzler.d:3782-3792, 6142-6183)
//   k_cons_*       consensus of a reference read over its pile (what daccord produces behind
//                  dazzler.d:4213-4255): per-tile global alignment from the trace points, column votes,
//                  majority emit
// Specification = oracle/pile_oracle.c (header); everything is integer work except the one fp64 compare
// of the error-rate filter, which is evaluated with the same IEEE operations as the D source.
#include "engine.cuh"
#include "pile.cuh"
#include <stdlib.h>

namespace dn {
namespace {

__global__ void __launch_bounds__(256) k_las_filter(const dn_las_record *__restrict__ rec, int64_t n, int mode, double max_err,
                                                    const int32_t *__restrict__ alen, const int32_t *__restrict__ blen,
                                                    int allowance, int32_t *__restrict__ keep) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const dn_las_record x = rec[i];
    int k;
    if (mode == 0) {
        double rate = (double)x.diffs / (double)(x.aepos - x.abpos);
        k = rate <= max_err;
    } else {
        const int la = alen[x.aread], lb = blen[x.bread];
        const bool ab = x.abpos <= allowance, bb = x.bbpos <= allowance;
        const bool ae = x.aepos + allowance >= la, be = x.bepos + allowance >= lb;
        k = x.aread != x.bread && (((ab && bb) && (ae || be)) || ((ae && be) && (ab || bb)));
    }
    keep[i] = k;
}

__global__ void __launch_bounds__(256) k_las_compact(const dn_las_record *__restrict__ rec, const int64_t *__restrict__ toff, int64_t n,
                                                     const int32_t *__restrict__ keep, const int32_t *__restrict__ kidx,
                                                     dn_las_record *__restrict__ orec, int64_t *__restrict__ otoff) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    orec[kidx[i]] = rec[i]; otoff[kidx[i]] = toff[i];          // traces stay where they are
}

// one CTA per read; records sorted by aread
__global__ void __launch_bounds__(128) k_qv(const int32_t *__restrict__ rlen, const dn_las_record *__restrict__ rec, int64_t nla,
                                            const int64_t *__restrict__ toff, const uint16_t *__restrict__ trace, int ts, int cov_all,
                                            const int32_t *__restrict__ cov_per_read,
                                            const int64_t *__restrict__ qoff, uint8_t *__restrict__ qv) {
    const int r = blockIdx.x;
    const int cov = cov_per_read ? cov_per_read[r] : cov_all;
    __shared__ int64_t s_lo, s_hi;
    if (threadIdx.x == 0) {
        int64_t a = 0, b = nla;
        while (a < b) { int64_t m = (a + b) >> 1; if (rec[m].aread < r) a = m + 1; else b = m; }
        s_lo = a; b = nla;
        while (a < b) { int64_t m = (a + b) >> 1; if (rec[m].aread <= r) a = m + 1; else b = m; }
        s_hi = a;
    }
    __syncthreads();
    const int L = rlen[r], nt = (L + ts - 1) / ts;
    for (int t = threadIdx.x; t < nt; t += blockDim.x) {
        const int t0 = t * ts, t1 = min((t + 1) * ts, L);
        unsigned char hist[51];
#pragma unroll
        for (int i = 0; i < 51; i++) hist[i] = 0;
        int m = 0;
        for (int64_t x = s_lo; x < s_hi; x++) {
            const int ab = rec[x].abpos, ae = rec[x].aepos;
            if (ab > t0 || ae < t1) continue;
            const uint16_t *tp = trace + toff[x] + 2 * (t - ab / ts);
            const int den = (t1 - t0) + tp[1];
            int v = (200 * (int)tp[0] + den / 2) / den;
            if (v > 50) v = 50;
            if (hist[v] < 255) hist[v]++;
            m++;
        }
        int q = 50;
        if (m * 4 >= cov && m > 0) {
            int n = cov / 2; if (n < 1) n = 1; if (n > m) n = m;
            int s = 0, left = n;
            for (int v = 0; v <= 50 && left > 0; v++) { int c = min((int)hist[v], left); s += c * v; left -= c; }
            q = (2 * s + n) / (2 * n);
            if (q > 50) q = 50;
        }
        qv[qoff[r] + t] = (uint8_t)q;
    }
}

// ------------------------------------------------------------------------------- consensus

__device__ __forceinline__ int base_at(const u32 *__restrict__ w, int64_t g) { return (int)((w[g >> 4] >> ((g & 15) << 1)) & 3u); }

// thread per voting LA: expand its trace into tile tasks
__global__ void __launch_bounds__(256) k_cons_tasks(const dn_las_record *__restrict__ rec, const int64_t *__restrict__ toff,
                                                    const uint16_t *__restrict__ trace, const int32_t *__restrict__ vla, int nvla,
                                                    const int64_t *__restrict__ task_off, int ts, ConsTask *__restrict__ tasks) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvla) return;
    const int x = vla[v];
    const dn_las_record la = rec[x];
    int ap = la.abpos, bp = la.bbpos; const int nt = la.tlen / 2;
    ConsTask *o = tasks + task_off[v];
    for (int t = 0; t < nt; t++) {
        const int aend = (t == nt - 1) ? la.aepos : (ap / ts + 1) * ts;
        const int bb = trace[toff[x] + 2 * t + 1];
        o[t] = ConsTask{x, ap, aend - ap, bp, bb};
        ap = aend; bp += bb;
    }
}

// thread per tile task: unit-cost global alignment (full DP, directions kept in an interleaved scratch),
// traceback with priority diagonal > deletion > insertion, atomic column votes
__global__ void __launch_bounds__(128) k_cons_vote(const ConsTask *__restrict__ tasks, int64_t ntasks,
                                                   const dn_las_record *__restrict__ rec, const int32_t *__restrict__ la_target,
                                                   ConsGeom G, u32 *__restrict__ scratch, int32_t *__restrict__ cnt,
                                                   int32_t *__restrict__ ins, int32_t *__restrict__ insn, int32_t *__restrict__ cov) {
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    u32 *dirs = scratch + tid;                               // word w of this thread lives at dirs[w * nthreads]
    for (int64_t task = tid; task < ntasks; task += nthreads) {
        const ConsTask T = tasks[task];
        const int n = T.alen, m = T.bb;
        if (m > 250 || n > 128) continue;
        const dn_las_record la = rec[T.la];
        const int tg = la_target[T.la];                       // index of the target (consensus) read
        const int64_t vbase = G.vote_off[tg];                 // first column of this target in the vote arrays
        const u32 *Aw = G.fwd, *Bw = (la.flags & DN_LAS_COMP) ? G.rc : G.fwd;
        const int64_t ga = G.off[la.aread] + T.ap, gb = G.off[la.bread] + T.bp;
        unsigned char bq[256], row[256];
        for (int j = 0; j < m; j++) bq[j] = (unsigned char)base_at(Bw, gb + j);
        for (int j = 0; j <= m; j++) row[j] = (unsigned char)j;
        const int wpr = (m + 16) >> 4;                        // direction words per row (16 cells per word)
        for (int i = 1; i <= n; i++) {
            const int ai = base_at(Aw, ga + i - 1);
            int diag = row[0]; row[0] = (unsigned char)i;
            int left = i;
            u32 word = 0;
            for (int j = 1; j <= m; j++) {
                const int up = row[j];
                const int d = diag + (ai != bq[j - 1]), u = up + 1, l = left + 1;
                int v = d; if (u < v) v = u; if (l < v) v = l;
                const u32 dir = (v == d) ? 0u : ((v == u) ? 1u : 2u);
                word |= dir << ((j & 15) << 1);
                if ((j & 15) == 15 || j == m) { dirs[(int64_t)((i - 1) * wpr + (j >> 4)) * nthreads] = word; word = 0; }
                row[j] = (unsigned char)v; diag = up; left = v;
            }
        }
        int i = n, j = m, pend = -1;
        while (i > 0 || j > 0) {
            u32 dir;
            if (i == 0) dir = 2u; else if (j == 0) dir = 1u;
            else dir = (dirs[(int64_t)((i - 1) * wpr + (j >> 4)) * nthreads] >> ((j & 15) << 1)) & 3u;
            if (dir != 2u && pend >= 0) { atomicAdd(&ins[(vbase + T.ap + i) * 4 + pend], 1); atomicAdd(&insn[vbase + T.ap + i], 1); pend = -1; }
            if (dir == 0u) { atomicAdd(&cnt[(vbase + T.ap + i - 1) * 5 + bq[j - 1]], 1); i--; j--; }
            else if (dir == 1u) { atomicAdd(&cnt[(vbase + T.ap + i - 1) * 5 + 4], 1); i--; }
            else { pend = bq[j - 1]; j--; }
        }
        if (pend >= 0) { atomicAdd(&ins[(vbase + T.ap + i) * 4 + pend], 1); atomicAdd(&insn[vbase + T.ap + i], 1); }
        for (int x = 0; x < n; x++) atomicAdd(&cov[vbase + T.ap + x], 1);
    }
}

// The same votes from a bit-parallel DP (Myers 1999 in Hyyro's edit-distance form): the A tile (<= 128 bases) is the
// pattern, one 128-bit pair (VP, VN) of vertical +1 / -1 deltas per B column replaces the column of cells, so a column
// costs ~60 instructions instead of 126 cells x ~15, and nothing lives in local memory.  The traceback needs exact cell
// values to repeat the cell DP's choice (diagonal > deletion > insertion among the minima): D[i][j] = j + popc(VP_j & low i
// bits) - popc(VN_j & low i bits), so the directions -- hence every vote -- are identical to k_cons_vote's.
struct U128 { unsigned long long lo, hi; };
__device__ __forceinline__ U128 u_and(U128 a, U128 b) { return U128{a.lo & b.lo, a.hi & b.hi}; }
__device__ __forceinline__ U128 u_or(U128 a, U128 b) { return U128{a.lo | b.lo, a.hi | b.hi}; }
__device__ __forceinline__ U128 u_xor(U128 a, U128 b) { return U128{a.lo ^ b.lo, a.hi ^ b.hi}; }
__device__ __forceinline__ U128 u_not(U128 a) { return U128{~a.lo, ~a.hi}; }
__device__ __forceinline__ U128 u_add(U128 a, U128 b) { U128 r; r.lo = a.lo + b.lo; r.hi = a.hi + b.hi + (r.lo < a.lo ? 1ull : 0ull); return r; }
__device__ __forceinline__ U128 u_shl1(U128 a) { return U128{a.lo << 1, (a.hi << 1) | (a.lo >> 63)}; }
__device__ __forceinline__ int u_bit(U128 a, int i) { return (int)(((i < 64 ? a.lo >> i : a.hi >> (i - 64))) & 1ull); }
// number of set bits among the lowest i bits (0 <= i <= 128)
__device__ __forceinline__ int u_popc_low(U128 a, int i) {
    if (i >= 64) return __popcll(a.lo) + (i >= 128 ? __popcll(a.hi) : __popcll(a.hi & ((1ull << (i - 64)) - 1ull)));
    return __popcll(a.lo & ((1ull << i) - 1ull));
}

// The DP and its traceback, shared by the consensus vote, the `-B` bridges and the transposition.  Rows = A[ga, ga + n),
// n <= 128; columns = B[gb, gb + m), m <= 250; `cols` = this thread's slice of the interleaved column store (word w of
// column j at cols[((j - 1) * 4 + w) * nthreads]).  Returns D[n][m].  The visitor sees every cell of the path from (n, m)
// down to (0, 0), the last one with dir = 3:  step(dir, i, j, d_here, d_next)  with dir 0 = diagonal, 1 = A base unmatched,
// 2 = B base inserted -- the cell DP's choice among the minima -- and the exact cell values of this cell and the next.
template <class Visitor>
__device__ __forceinline__ int bv_align(const u32 *__restrict__ Aw, int64_t ga, int n, const u32 *__restrict__ Bw, int64_t gb, int m,
                                        unsigned long long *__restrict__ cols, int64_t nthreads, Visitor &&visit) {
    U128 P0{0, 0}, P1{0, 0}, P2{0, 0}, P3{0, 0};                      // positions of each base in the A tile
    for (int i = 0; i < n; i++) {
        const int a = base_at(Aw, ga + i);
        const unsigned long long bl = i < 64 ? 1ull << i : 0ull, bh = i < 64 ? 0ull : 1ull << (i - 64);
        if (a == 0) { P0.lo |= bl; P0.hi |= bh; } else if (a == 1) { P1.lo |= bl; P1.hi |= bh; }
        else if (a == 2) { P2.lo |= bl; P2.hi |= bh; } else { P3.lo |= bl; P3.hi |= bh; }
    }
    U128 VP{~0ull, ~0ull}, VN{0, 0};
    for (int j = 1; j <= m; j++) {
        const int b = base_at(Bw, gb + j - 1);
        const U128 Eq = b == 0 ? P0 : (b == 1 ? P1 : (b == 2 ? P2 : P3));
        const U128 D0 = u_or(u_or(u_xor(u_add(u_and(Eq, VP), VP), VP), Eq), VN);
        const U128 HP = u_or(VN, u_not(u_or(D0, VP))), HN = u_and(D0, VP);
        U128 X = u_shl1(HP); X.lo |= 1ull;                            // row 0 grows by one per column (global alignment)
        VP = u_or(u_shl1(HN), u_not(u_or(D0, X))); VN = u_and(D0, X);
        unsigned long long *c = cols + (int64_t)((j - 1) * 4) * nthreads;
        c[0] = VP.lo; c[nthreads] = VP.hi; c[2 * nthreads] = VN.lo; c[3 * nthreads] = VN.hi;
    }
    auto column = [&](int j, U128 &vp, U128 &vn) {
        if (j == 0) { vp = U128{~0ull, ~0ull}; vn = U128{0, 0}; return; }
        const unsigned long long *c = cols + (int64_t)((j - 1) * 4) * nthreads;
        vp = U128{c[0], c[nthreads]}; vn = U128{c[2 * nthreads], c[3 * nthreads]};
    };
    int i = n, j = m;
    U128 vpj, vnj; column(m, vpj, vnj);
    int c0 = m + u_popc_low(vpj, n) - u_popc_low(vnj, n);             // D[n][m]
    const int total = c0;
    for (;;) {
        u32 dir; int nc0 = 0;
        U128 vpl{0, 0}, vnl{0, 0};
        if (i == 0 && j == 0) dir = 3u;
        else if (i == 0) { dir = 2u; nc0 = j - 1; column(j - 1, vpl, vnl); }
        else if (j == 0) { dir = 1u; nc0 = i - 1; }
        else {
            column(j - 1, vpl, vnl);
            const int up = c0 - (u_bit(vpj, i - 1) - u_bit(vnj, i - 1));
            const int left = (j - 1) + u_popc_low(vpl, i) - u_popc_low(vnl, i);
            const int dg = left - (u_bit(vpl, i - 1) - u_bit(vnl, i - 1));
            const int d = dg + (base_at(Aw, ga + i - 1) != base_at(Bw, gb + j - 1)), u = up + 1, l = left + 1;
            int v = d; if (u < v) v = u; if (l < v) v = l;
            dir = (v == d) ? 0u : ((v == u) ? 1u : 2u);
            nc0 = dir == 0u ? dg : (dir == 1u ? up : left);
        }
        visit(dir, i, j, c0, nc0);
        if (dir == 3u) break;
        if (dir == 0u) { i--; j--; vpj = vpl; vnj = vnl; }
        else if (dir == 1u) i--;
        else { j--; vpj = vpl; vnj = vnl; }
        c0 = nc0;
    }
    return total;
}

__global__ void __launch_bounds__(128) k_cons_vote_bv(const ConsTask *__restrict__ tasks, int64_t ntasks,
                                                      const dn_las_record *__restrict__ rec, const int32_t *__restrict__ la_target,
                                                      ConsGeom G, u32 *__restrict__ scratch, int32_t *__restrict__ cnt,
                                                      int32_t *__restrict__ ins, int32_t *__restrict__ insn, int32_t *__restrict__ cov) {
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long *cols = (unsigned long long *)scratch + tid;
    for (int64_t task = tid; task < ntasks; task += nthreads) {
        const ConsTask T = tasks[task];
        const int n = T.alen, m = T.bb;
        if (m > 250 || n > 128) continue;
        const dn_las_record la = rec[T.la];
        const int tg = la_target[T.la];
        const int64_t vbase = G.vote_off[tg];
        const u32 *Aw = G.fwd, *Bw = (la.flags & DN_LAS_COMP) ? G.rc : G.fwd;
        const int64_t ga = G.off[la.aread] + T.ap, gb = G.off[la.bread] + T.bp;
        int pend = -1;                                               // first base (lowest j) of the insertion run being walked
        bv_align(Aw, ga, n, Bw, gb, m, cols, nthreads, [&](u32 dir, int i, int j, int, int) {
            if (dir != 2u && pend >= 0) { atomicAdd(&ins[(vbase + T.ap + i) * 4 + pend], 1); atomicAdd(&insn[vbase + T.ap + i], 1); pend = -1; }
            if (dir == 0u) atomicAdd(&cnt[(vbase + T.ap + i - 1) * 5 + base_at(Bw, gb + j - 1)], 1);
            else if (dir == 1u) atomicAdd(&cnt[(vbase + T.ap + i - 1) * 5 + 4], 1);
            else if (dir == 2u) pend = base_at(Bw, gb + j - 1);
        });
        for (int x = 0; x < n; x++) atomicAdd(&cov[vbase + T.ap + x], 1);
    }
}

// `daligner -B` bridges (specification: oracle/pile_oracle.c, orc_bridge): thread per bridge, the same DP over
// A[P.aepos, Q.abpos) x B[P.bepos, Q.bbpos); the traceback records, for every multiple of ts of A inside the bridge, the
// B column and the cost at the path's FIRST cell on that row (the cell it leaves the row from, walking backwards).
__global__ void __launch_bounds__(128) k_bridge(const BridgeTask *__restrict__ tasks, int64_t ntasks, const u32 *__restrict__ a_fwd,
                                                const u32 *__restrict__ b_fwd, const u32 *__restrict__ b_rc, int ts,
                                                u32 *__restrict__ scratch, int32_t *__restrict__ total, int2 *__restrict__ rows) {
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long *cols = (unsigned long long *)scratch + tid;
    for (int64_t task = tid; task < ntasks; task += nthreads) {
        const BridgeTask T = tasks[task];
        int k = T.nrow;
        total[task] = bv_align(a_fwd, T.ga, T.n, T.comp ? b_rc : b_fwd, T.gb, T.m, cols, nthreads, [&](u32 dir, int i, int j, int d_here, int) {
            if (dir < 2u && (T.a0 + i) % ts == 0) { k--; rows[T.row_off + k] = make_int2(j, d_here); }
        });
    }
}

// thread per vote column (targets concatenated, L+1 columns each): how many symbols does it emit?
__global__ void __launch_bounds__(256) k_cons_count(ConsGeom G, const int32_t *__restrict__ targets, int ntargets, int64_t ncols,
                                                    const int32_t *__restrict__ cnt, const int32_t *__restrict__ ins,
                                                    const int32_t *__restrict__ insn, const int32_t *__restrict__ cov,
                                                    int32_t *__restrict__ nemit, uint8_t *__restrict__ sym /* 2 per column */) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    int lo = 0, hi = ntargets;                                // target of this column
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (G.vote_off[mid] <= c) lo = mid; else hi = mid; }
    const int r = targets[lo];
    const int p = (int)(c - G.vote_off[lo]);
    const int L = G.len[r];
    int ne = 0;
    const int total = (p < L ? cov[c] : (L > 0 ? cov[c - 1] : 0)) + 1;
    if (2 * insn[c] > total) {
        int best = 0;
        for (int s = 1; s < 4; s++) if (ins[c * 4 + s] > ins[c * 4 + best]) best = s;
        sym[2 * c + ne++] = (uint8_t)best;
    }
    if (p < L) {
        const int own = base_at(G.fwd, G.off[r] + p);
        int bestsym = own, bestc = cnt[c * 5 + own] + 1;
        for (int s = 0; s < 5; s++) { int v = cnt[c * 5 + s] + (s == own ? 1 : 0); if (v > bestc) { bestc = v; bestsym = s; } }
        if (bestsym < 4) sym[2 * c + ne++] = (uint8_t)bestsym;
    }
    nemit[c] = ne;
}

__global__ void __launch_bounds__(256) k_cons_write(int64_t ncols, const int32_t *__restrict__ nemit, const int32_t *__restrict__ eoff,
                                                    const uint8_t *__restrict__ sym, uint8_t *__restrict__ out) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    for (int e = 0; e < nemit[c]; e++) out[eoff[c] + e] = sym[2 * c + e];
}

// ---- transposition (damapper -C, dazzler.d:5931-5936): specification in oracle/pile_oracle.c (orc_transpose) -----------
// thread per A tile: the same unit-cost DP and traceback as the consensus vote; instead of votes it records where the
// path crosses the trace-point columns of the NEW A read (= B): {x, a, c} = column, the other read's coordinate there,
// and the diffs from the tile's start (not complemented) or from its end (complemented: the new read runs backwards).
__global__ void __launch_bounds__(128) k_tr_tiles(const ConsTask *__restrict__ tasks, int64_t ntasks, const dn_las_record *__restrict__ rec,
                                                  const int64_t *__restrict__ toff, const uint16_t *__restrict__ trace,
                                                  const int64_t *__restrict__ la_task0, TrGeom G, int ts, int kmax,
                                                  u32 *__restrict__ scratch, int4 *__restrict__ cross, int32_t *__restrict__ ncross,
                                                  int32_t *__restrict__ tcost) {
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long *cols = (unsigned long long *)scratch + tid;
    for (int64_t task = tid; task < ntasks; task += nthreads) {
        const ConsTask T = tasks[task];
        const int n = T.alen, m = T.bb;
        const dn_las_record la = rec[T.la];
        const bool comp = (la.flags & DN_LAS_COMP) != 0;
        const int lb = G.b_len[la.bread], lo = la.bbpos, hi = la.bepos, bp = T.bp;
        int4 *cr = cross + task * kmax; int nc = 0;
        if (m > 250 || n > 128) {                             // diagonal-first path, the tile's recorded diffs spread along it
            const int dt = trace[toff[T.la] + 2 * (task - la_task0[T.la])];
            const int q = n < m ? n : m;
            for (int j = m; j >= 0; j--) {
                const int x = bp + j;
                if (x <= lo || x >= hi) continue;
                int i;
                if (!comp) { if (j == 0 || x % ts) continue; i = j <= q ? j : n; if (j == m && n > m) i = m; }
                else { if (j == m || (lb - x) % ts) continue; i = j <= q ? j : n; }
                const int from_start = (n + m) ? (int)((long long)dt * (i + j) / (n + m)) : 0;
                if (nc < kmax) cr[nc] = make_int4(x, T.ap + i, comp ? dt - from_start : from_start, 0);
                nc++;
            }
            ncross[task] = nc; tcost[task] = dt;
            continue;
        }
        const u32 *Aw = G.a_fwd, *Bw = comp ? G.b_rc : G.b_fwd;
        const int64_t ga = G.a_off[la.aread] + T.ap, gb = G.b_off[la.bread] + T.bp;
        // the crossings of a non-complemented tile carry costs from the tile's start (d_here); those of a complemented one
        // costs from its end (D[n][m] - d_next): D[n][m] is added once the walk has returned it
        const int total = bv_align(Aw, ga, n, Bw, gb, m, cols, nthreads, [&](u32 dir, int i, int j, int d_here, int d_next) {
            const int x = bp + j;
            if (!comp && (dir == 0u || dir == 2u) && j > 0 && x % ts == 0 && x > lo && x < hi) {      // leaving column j: its last (lowest) cell
                if (nc < kmax) cr[nc] = make_int4(x, T.ap + i, d_here, 0);
                nc++;
            }
            if (comp && (dir == 0u || dir == 2u)) {                                                  // entering column j - 1: its first (highest) cell
                const int xn = bp + j - 1, in = dir == 0u ? i - 1 : i;
                if ((lb - xn) % ts == 0 && xn > lo && xn < hi) { if (nc < kmax) cr[nc] = make_int4(xn, T.ap + in, -d_next, 0); nc++; }
            }
        });
        if (comp) for (int q = 0; q < nc && q < kmax; q++) cr[q].z += total;
        ncross[task] = nc; tcost[task] = total;
    }
}

// thread per record: strings the tiles' crossings together into the trace of the transposed record
__global__ void __launch_bounds__(128) k_tr_assemble(const dn_las_record *__restrict__ rec, int64_t nla, const int64_t *__restrict__ la_task0,
                                                     TrGeom G, int kmax, const int4 *__restrict__ cross, const int32_t *__restrict__ ncross,
                                                     const int32_t *__restrict__ tcost, const int64_t *__restrict__ out_toff,
                                                     dn_las_record *__restrict__ out, uint16_t *__restrict__ out_trace, int32_t *__restrict__ status) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nla) return;
    const dn_las_record x = rec[r];
    const int nt = x.tlen / 2;
    const int64_t t0 = la_task0[r];
    const bool comp = (x.flags & DN_LAS_COMP) != 0;
    const int LA = G.a_len[x.aread], LB = G.b_len[x.bread];
    dn_las_record o = x;
    o.aread = x.bread; o.bread = x.aread;
    uint16_t *tr = out_trace + out_toff[r];
    int total = 0, ntp = 0, w = 0;
    for (int t = 0; t < nt; t++) { total += tcost[t0 + t]; if (ncross[t0 + t] > kmax) atomicExch(status, 1); }
    if (!comp) {
        o.abpos = x.bbpos; o.aepos = x.bepos; o.bbpos = x.abpos; o.bepos = x.aepos;
        int pa = x.abpos, pd = 0, base = 0;
        for (int t = 0; t < nt; t++) {
            const int4 *cr = cross + (t0 + t) * kmax;
            for (int q = min(ncross[t0 + t], kmax) - 1; q >= 0; q--) {
                const int d = base + cr[q].z;
                tr[w++] = (uint16_t)(d - pd); tr[w++] = (uint16_t)(cr[q].y - pa); ntp++;
                pd = d; pa = cr[q].y;
            }
            base += tcost[t0 + t];
        }
        if (x.bepos > x.bbpos) { tr[w++] = (uint16_t)(total - pd); tr[w++] = (uint16_t)(x.aepos - pa); ntp++; }
    } else {
        o.abpos = LB - x.bepos; o.aepos = LB - x.bbpos; o.bbpos = LA - x.aepos; o.bepos = LA - x.abpos;
        int pa = x.aepos, pd = 0, base = 0;
        for (int t = nt - 1; t >= 0; t--) {
            const int4 *cr = cross + (t0 + t) * kmax;
            const int k = min(ncross[t0 + t], kmax);
            for (int q = 0; q < k; q++) {
                const int d = base + cr[q].z;
                tr[w++] = (uint16_t)(d - pd); tr[w++] = (uint16_t)(pa - cr[q].y); ntp++;
                pd = d; pa = cr[q].y;
            }
            base += tcost[t0 + t];
        }
        if (x.bepos > x.bbpos) { tr[w++] = (uint16_t)(total - pd); tr[w++] = (uint16_t)(pa - x.abpos); ntp++; }
    }
    o.diffs = total; o.tlen = 2 * ntp;
    out[r] = o;
}

}  // namespace

void launch_tr_tiles(const ConsTask *tasks, int64_t ntasks, const dn_las_record *rec, const int64_t *toff, const uint16_t *trace,
                     const int64_t *la_task0, TrGeom G, int ts, int kmax, u32 *scratch, int4 *cross, int32_t *ncross, int32_t *tcost, cudaStream_t s) {
    DN_LAUNCH(k_tr_tiles, sm_count() * 4, 128, 0, s, tasks, ntasks, rec, toff, trace, la_task0, G, ts, kmax, scratch, cross, ncross, tcost);
}
void launch_tr_assemble(const dn_las_record *rec, int64_t nla, const int64_t *la_task0, TrGeom G, int kmax, const int4 *cross,
                        const int32_t *ncross, const int32_t *tcost, const int64_t *out_toff, dn_las_record *out, uint16_t *out_trace,
                        int32_t *status, cudaStream_t s) {
    DN_LAUNCH(k_tr_assemble, (unsigned)((nla + 127) / 128), 128, 0, s, rec, nla, la_task0, G, kmax, cross, ncross, tcost, out_toff, out, out_trace, status);
}

void las_filter_device(const dn_las_record *rec, const int64_t *toff, int64_t n, int mode, double max_err, const int32_t *alen,
                       const int32_t *blen, int allowance, dn_las_record *orec, int64_t *otoff, int64_t *n_out, cudaStream_t s) {
    *n_out = 0;
    if (n == 0) return;
    DBuf<int32_t> keep(n), kidx(n), tot(1);
    DN_LAUNCH(k_las_filter, (unsigned)((n + 255) / 256), 256, 0, s, rec, n, mode, max_err, alen, blen, allowance, keep.p);
    exclusive_scan_i32(keep.p, kidx.p, n, tot.p, s);
    DN_LAUNCH(k_las_compact, (unsigned)((n + 255) / 256), 256, 0, s, rec, toff, n, (const int32_t *)keep.p, (const int32_t *)kidx.p, orec, otoff);
    int32_t h; DN_CUDA(cudaMemcpyAsync(&h, tot.p, 4, cudaMemcpyDeviceToHost, s)); DN_CUDA(cudaStreamSynchronize(s));
    *n_out = h;
}

void qv_device(const int32_t *rlen, int nreads, const dn_las_record *rec, int64_t nla, const int64_t *toff, const uint16_t *trace,
               int ts, int cov, const int32_t *cov_per_read, const int64_t *qoff, uint8_t *qv, cudaStream_t s) {
    if (nreads == 0) return;
    DN_LAUNCH(k_qv, nreads, 128, 0, s, rlen, rec, nla, toff, trace, ts, cov, cov_per_read, qoff, qv);
}

void launch_cons_tasks(const dn_las_record *rec, const int64_t *toff, const uint16_t *trace, const int32_t *vla, int nvla,
                       const int64_t *task_off, int ts, ConsTask *tasks, cudaStream_t s) {
    DN_LAUNCH(k_cons_tasks, (nvla + 255) / 256, 256, 0, s, rec, toff, trace, vla, nvla, task_off, ts, tasks);
}
int cons_vote_threads() { return sm_count() * 4 * 128; }
void launch_bridge(const BridgeTask *tasks, int64_t ntasks, const u32 *a_fwd, const u32 *b_fwd, const u32 *b_rc, int ts, u32 *scratch,
                   int32_t *total, int2 *rows, cudaStream_t s) {
    DN_LAUNCH(k_bridge, sm_count() * 4, 128, 0, s, tasks, ntasks, a_fwd, b_fwd, b_rc, ts, scratch, total, rows);
}
void launch_cons_vote(const ConsTask *tasks, int64_t ntasks, const dn_las_record *rec, const int32_t *la_target, ConsGeom G,
                      u32 *scratch, int32_t *cnt, int32_t *ins, int32_t *insn, int32_t *cov, cudaStream_t s) {
    static const bool cell_dp = getenv("DN_CONS_CELL_DP") != nullptr;         // the thread-per-cell-row DP the bit-parallel kernel replaced
    if (cell_dp) DN_LAUNCH(k_cons_vote, sm_count() * 4, 128, 0, s, tasks, ntasks, rec, la_target, G, scratch, cnt, ins, insn, cov);
    else DN_LAUNCH(k_cons_vote_bv, sm_count() * 4, 128, 0, s, tasks, ntasks, rec, la_target, G, scratch, cnt, ins, insn, cov);
}
void launch_cons_count(ConsGeom G, const int32_t *targets, int ntargets, int64_t ncols, const int32_t *cnt, const int32_t *ins,
                       const int32_t *insn, const int32_t *cov, int32_t *nemit, uint8_t *sym, cudaStream_t s) {
    DN_LAUNCH(k_cons_count, (unsigned)((ncols + 255) / 256), 256, 0, s, G, targets, ntargets, ncols, cnt, ins, insn, cov, nemit, sym);
}
void launch_cons_write(int64_t ncols, const int32_t *nemit, const int32_t *eoff, const uint8_t *sym, uint8_t *out, cudaStream_t s) {
    DN_LAUNCH(k_cons_write, (unsigned)((ncols + 255) / 256), 256, 0, s, ncols, nemit, eoff, sym, out);
}

}  // namespace dn

// ------------------------------------------------------------------------------- mapper chains
// damapper reports, per read, chains of local alignments (START / NEXT flags, BEST on the top chain;
// decoded by DENTIST at dazzler.d:1738-1755 and packed into AlignmentChains at :708-743).
// Specification (oracle/chain_oracle.py; DAMAPPER is absent -> parity unpinned): records are in LAsort
// order; record i+1 CONTINUES the chain of record i when both have the same (aread, bread, comp), both
// coordinates advance (abpos, aepos, bbpos, bepos all strictly larger), the gap difference
// |(ab'-ae) - (bb'-be)| <= max_indel and max(|ab'-ae|, |bb'-be|) <= max_gap.  A chain's score is the sum
// of its (aepos-abpos); per B read the chain with the highest score (ties: first in file order) is BEST.
namespace dn {
namespace {

__device__ __forceinline__ bool continues(const dn_las_record &x, const dn_las_record &y, int max_indel, int max_gap) {
    if (x.aread != y.aread || x.bread != y.bread || ((x.flags ^ y.flags) & DN_LAS_COMP)) return false;
    if (!(y.abpos > x.abpos && y.aepos > x.aepos && y.bbpos > x.bbpos && y.bepos > x.bepos)) return false;
    const int ga = y.abpos - x.aepos, gb = y.bbpos - x.bepos;
    const int indel = ga > gb ? ga - gb : gb - ga;
    const int mg = max(ga < 0 ? -ga : ga, gb < 0 ? -gb : gb);
    return indel <= max_indel && mg <= max_gap;
}

// pass 1: chain starts; one thread per chain start walks its chain, scores it, competes for its read
__global__ void __launch_bounds__(256) k_chain_score(const dn_las_record *__restrict__ rec, int64_t n, int max_indel, int max_gap,
                                                     unsigned long long *__restrict__ best /* per B read */) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i > 0 && continues(rec[i - 1], rec[i], max_indel, max_gap)) return;         // not a chain start
    long long score = 0; int64_t j = i;
    for (;;) { score += rec[j].aepos - rec[j].abpos; if (j + 1 < n && continues(rec[j], rec[j + 1], max_indel, max_gap)) j++; else break; }
    // highest score wins, ties: lowest start index  ->  pack (score, ~index)
    const unsigned long long packed = ((unsigned long long)score << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
    atomicMax(&best[rec[i].bread], packed);
}

__global__ void __launch_bounds__(256) k_chain_flags(dn_las_record *__restrict__ rec, int64_t n, int max_indel, int max_gap,
                                                     const unsigned long long *__restrict__ best) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool cont = i > 0 && continues(rec[i - 1], rec[i], max_indel, max_gap);
    unsigned f = rec[i].flags & ~(DN_LAS_START | DN_LAS_NEXT | DN_LAS_BEST);
    if (cont) f |= DN_LAS_NEXT;
    else {
        f |= DN_LAS_START;
        const unsigned long long b = best[rec[i].bread];
        if ((unsigned)(0xffffffffu - (unsigned)(b & 0xffffffffu)) == (unsigned)i) f |= DN_LAS_BEST;
    }
    rec[i].flags = f;
}

}  // namespace

void mapper_chain_device(dn_las_record *rec, int64_t n, int nb_reads, int max_indel, int max_gap, cudaStream_t s) {
    if (n == 0) return;
    DBuf<unsigned long long> best(nb_reads + 1); best.zero(s);
    DN_LAUNCH(k_chain_score, (unsigned)((n + 255) / 256), 256, 0, s, (const dn_las_record *)rec, n, max_indel, max_gap, best.p);
    DN_LAUNCH(k_chain_flags, (unsigned)((n + 255) / 256), 256, 0, s, rec, n, max_indel, max_gap, (const unsigned long long *)best.p);
}

}  // namespace dn
