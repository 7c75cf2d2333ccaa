// extend.cu -- O(ND) furthest-reaching wave extension with trace points, one warp per
// (seed, direction) task; candidate assembly, hit retirement and duplicate removal.
// Stage K5/K6 of DESIGN.md: the device replacement for daligner's Local_Alignment wave
// (what `daligner`/`damapper` run behind dazzler.d:6131-6170), emitting daligner-compatible
// (diffs, bbases) pairs per tspace-tile of A (tile semantics base.d:185-242).
//
// Warp layout: the live wave window [lo,hi] (<= 62 diagonals, <= 64 after the +-1 growth) lives in
// shared memory, two buffers of V (furthest A offset) and T (trace record index) indexed by
// diagonal mod 64; lane l owns diagonals nlo+l and nlo+32+l of the new wave.  Sequence is read
// from the 2-bit packed block with 32-bit funnel-shifted windows (16 bases per compare); the whole
// packed block is L2-resident on B200 (126 MB L2), so slides are L1/L2 hits.
#include "seed.cuh"
#include <stdlib.h>

namespace dn {

int ext_ctas_per_sm();

namespace {

constexpr int EXT_WARPS = 9;                 // warps per CTA: 4 resident CTAs x 9 warps = 36 warps per SM measured best (32: +1.6 %, 40: +6.5 %, 27: +7 %)
constexpr int NEGV = -(1 << 29);
constexpr int NEGS = -(1 << 30);

// base offsets fit 32 bits: a block holds < 2^31 padded bases (block_upload checks 2 * total < 2^32)
__device__ __forceinline__ u32 fetch16(const u32 *__restrict__ w, u32 g) {
    const u32 wi = g >> 4; const int sh = (int)(g & 15u) << 1;
    u32 lo = __ldg(w + wi), hi = __ldg(w + wi + 1);
    return __funnelshift_r(lo, hi, sh);
}

__device__ __forceinline__ int slide(const u32 *__restrict__ A, u32 ga, const u32 *__restrict__ B, u32 gb, int lim) {
    int s = 0;
    while (s < lim) {
        u32 x = fetch16(A, ga + s) ^ fetch16(B, gb + s);
        if (x) { s += (__ffs(x) - 1) >> 1; break; }
        s += 16;
    }
    return s < lim ? s : lim;
}

// ---- experiment (DN_EXT_TMA=1): the first STAGE_BASES bases of a task's A and B windows staged in shared memory by one
// bulk copy each (cp.async.bulk -> SASS UBLKCP, completion on an mbarrier); slides read the staged words with LDS and fall
// back to the global path beyond them.  Same results; measured against the __ldg path in DESIGN.md.
constexpr int STAGE_WORDS = 512;                     // 2 KB = 8192 bases per stream and warp
constexpr int STAGE_BASES = STAGE_WORDS * 16;

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, u32 bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, u32 parity) {
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

struct Staged { const u32 *sa, *sb; u32 a_base, b_base; int a_n, b_n; };     // staged bases [x_base, x_base + x_n)

__device__ __forceinline__ u32 fetch16s(const u32 *__restrict__ w, u32 g, const u32 *__restrict__ sw, u32 base, int n) {
    const u32 rel = g - base;
    if ((int)rel + 32 <= n && g >= base) {
        const u32 wi = rel >> 4; const int sh = (int)(rel & 15u) << 1;
        return __funnelshift_r(sw[wi], sw[wi + 1], sh);
    }
    return fetch16(w, g);
}
__device__ __forceinline__ int slide_s(const u32 *__restrict__ A, u32 ga, const u32 *__restrict__ B, u32 gb, int lim, const Staged &S) {
    int s = 0;
    while (s < lim) {
        u32 x = fetch16s(A, ga + s, S.sa, S.a_base, S.a_n) ^ fetch16s(B, gb + s, S.sb, S.b_base, S.b_n);
        if (x) { s += (__ffs(x) - 1) >> 1; break; }
        s += 16;
    }
    return s < lim ? s : lim;
}

__host__ __device__ inline int ext_span(int la, int lb) { long long s = (long long)lb + lb / 2 + 64; return la < s ? la : (int)s; }

struct Task {
    const u32 *A, *B; u32 ga, gb; int la, lb, firstT;
};

__device__ __forceinline__ Task make_task(const Seed &sd, int dir, const ExtGeom &G) {
    Task t;
    const int st = sd.bs >= G.nb_reads ? 1 : 0, br = sd.bs - st * G.nb_reads;
    const int LA = G.a_len[sd.a], LB = G.b_len[br];
    const int64_t oa = G.a_off[sd.a], ob = G.b_off[br];
    if (dir == 0) {
        t.A = G.a_fwd; t.B = st ? G.b_rc : G.b_fwd;
        t.ga = (u32)(oa + sd.apos); t.gb = (u32)(ob + sd.bpos); t.la = LA - sd.apos; t.lb = LB - sd.bpos;
        t.firstT = (sd.apos / G.ts + 1) * G.ts - sd.apos;
    } else {
        t.A = G.a_rc; t.B = st ? G.b_fwd : G.b_rc;
        t.ga = (u32)(oa + (LA - sd.apos)); t.gb = (u32)(ob + (LB - sd.bpos)); t.la = sd.apos; t.lb = sd.bpos;
        t.firstT = sd.apos > 0 ? sd.apos - ((sd.apos - 1) / G.ts) * G.ts : G.ts;
        if (sd.apos == 0 || sd.bpos == 0) { t.la = 0; t.lb = 0; }
    }
    return t;
}

__global__ void __launch_bounds__(256) k_task_caps(const Seed *__restrict__ seeds, int nseeds, ExtGeom G, u32 *__restrict__ caps) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseeds) return;
    Seed sd = seeds[s];
#pragma unroll
    for (int dir = 0; dir < 2; dir++) {
        Task t = make_task(sd, dir, G);
        caps[2 * s + dir] = (u32)(ext_span(t.la, t.lb) / G.ts + 3);
    }
}

__global__ void __launch_bounds__(EXT_WARPS * 32) k_extend(const Seed *__restrict__ seeds, int ntasks, ExtGeom G,
                                                           const int64_t *__restrict__ tile_off, int2 *__restrict__ tiles,
                                                           ExtOut *__restrict__ outs, int4 *__restrict__ pool_all,
                                                           int64_t pool_stride, int *__restrict__ counter, const int *__restrict__ order) {
    __shared__ int sV[EXT_WARPS][2][64];      // furthest A offset per diagonal (two wave buffers)
    __shared__ int sT[EXT_WARPS][2][64];      // trace record index per diagonal
    __shared__ int sR[EXT_WARPS][2][64];      // next tile boundary (relative A offset) above the cell
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 FULL = 0xffffffffu;
    int4 *pool = pool_all + (int64_t)(blockIdx.x * EXT_WARPS + warp) * pool_stride;
    const int ts = G.ts, C = G.cdiff, X = G.xdrop, WM = G.wmax;

    for (;;) {
        int task = 0;
        if (lane == 0) { task = atomicAdd(counter, 1); if (task < ntasks && order) task = order[task]; }   // longest expected tasks first
        task = __shfl_sync(FULL, task, 0);
        if (task >= ntasks) break;
        const Seed sd = seeds[task >> 1];
        const Task tk = make_task(sd, task & 1, G);
        const int la = tk.la, lb = tk.lb, firstT = tk.firstT;
        const int poolcap = G.poolmul * (ext_span(la, lb) / ts + 4);
        // number of tile boundaries at or below relative A offset i
        auto NB = [&](int i) -> int { return i >= firstT ? (int)__umulhi((u32)(i - firstT), G.ts_magic) + 1 : 0; };

        int npool = 0;
        int lo = 0, hi = 0, d = 0;
        int bS = NEGS, bi = 0, bk = 0, bd = 0, bT = -1;     // lane-local best
        int gbest;                                            // warp-uniform best S so far
        int cur = 0;
        {   // wave 0 (uniform across lanes)
            int lim = la < lb ? la : lb;
            int i = slide(tk.A, tk.ga, tk.B, tk.gb, lim);
            int n = NB(i), T = -1;
            if (n > poolcap) {
                if (lane == 0) outs[task] = ExtOut{0, 0, 0, 0};
                continue;
            }
            if (lane == 0) {
                for (int q = 1; q <= n; q++) { pool[npool] = make_int4(T, firstT + (q - 1) * ts, 0, q); T = npool++; }
                sV[warp][0][0] = i; sT[warp][0][0] = T; sR[warp][0][0] = firstT + n * ts;
                bS = 3 * (2 * i); bi = i; bk = 0; bd = 0; bT = T;
            }
            npool = n;
            gbest = 3 * (2 * i);
            if (i == la || i == lb) { lo = 1; hi = 0; }
            __syncwarp();
        }
        while (lo <= hi) {
            d++;
            const int nlo = lo - 1, nhi = hi + 1;
            const bool two = nhi - nlo >= 32;                   // warp-uniform: second slot in use
            const int *Vo = sV[warp][cur], *To = sT[warp][cur], *Ro = sR[warp][cur];
            int *Vn = sV[warp][cur ^ 1], *Tn = sT[warp][cur ^ 1], *Rn = sR[warp][cur ^ 1];
            int ci[2] = {NEGV, NEGV}, cT[2] = {-1, -1}, cR[2] = {0, 0}, cn[2] = {0, 0};
            // one cell of the new wave: best of the three predecessors, slide, count crossed boundaries
            auto cell = [&](const int sl) {
                const int k = nlo + lane + 32 * sl;
                if (k > nhi) return;
                int i = NEGV, ps = k;
                const int vs = (k >= lo && k <= hi) ? Vo[k & 63] : NEGV;
                const int vd = (k - 1 >= lo) ? Vo[(k - 1) & 63] : NEGV;          // k-1 <= hi always holds
                const int vi = (k + 1 <= hi) ? Vo[(k + 1) & 63] : NEGV;          // k+1 >= lo always holds
                if (vs > NEGV) { i = vs + 1; }
                if (vd > NEGV && vd + 1 > i) { i = vd + 1; ps = k - 1; }
                if (vi > NEGV && vi > i) { i = vi; ps = k + 1; }
                if (i == NEGV) return;
                int j = i - k;
                if (i > la || j > lb || j < 0) return;
                i += slide(tk.A, tk.ga + i, tk.B, tk.gb + j, min(la - i, lb - j));
                int R = Ro[ps & 63], n = 0;
                while (i >= R) { n++; R += ts; }
                ci[sl] = i; cT[sl] = To[ps & 63]; cR[sl] = R; cn[sl] = n;
            };
            cell(0);
            if (two) cell(1);
            const int need_l = cn[0] + cn[1];
            if (__any_sync(FULL, need_l != 0)) {
                const int need = __reduce_add_sync(FULL, need_l);
                if (npool + need > poolcap) break;              // pool exhausted: stop before this wave
                int pre = need_l;                               // inclusive scan across lanes
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, pre, o); if (lane >= o) pre += t; }
                int at = npool + pre - need_l;
#pragma unroll
                for (int sl = 0; sl < 2; sl++) {
                    const int k = nlo + lane + 32 * sl;
                    int T = cT[sl];
                    for (int q = cn[sl]; q >= 1; q--) {
                        const int bnd = cR[sl] - q * ts;        // boundaries crossed, in increasing order
                        pool[at] = make_int4(T, bnd - k, d, 0); T = at++;
                    }
                    cT[sl] = T;
                }
                npool += need;
            }
            int S[2];
#pragma unroll
            for (int sl = 0; sl < 2; sl++) {
                const int k = nlo + lane + 32 * sl;
                S[sl] = ci[sl] > NEGV ? 3 * (2 * ci[sl] - k) - C * d : NEGS;
                if (S[sl] > bS) { bS = S[sl]; bi = ci[sl]; bk = k; bd = d; bT = cT[sl]; }
            }
            const int waveS = __reduce_max_sync(FULL, max(S[0], S[1]));
            gbest = max(gbest, waveS);
            bool alive[2];
#pragma unroll
            for (int sl = 0; sl < 2; sl++) {
                const int k = nlo + lane + 32 * sl;
                alive[sl] = ci[sl] > NEGV && !(S[sl] < gbest - X || ci[sl] == la || ci[sl] - k == lb);
                if (k <= nhi) { Vn[k & 63] = alive[sl] ? ci[sl] : NEGV; Tn[k & 63] = cT[sl]; Rn[k & 63] = cR[sl]; }
            }
            const u32 m0 = __ballot_sync(FULL, alive[0]), m1 = two ? __ballot_sync(FULL, alive[1]) : 0u;
            if ((m0 | m1) == 0u) break;
            int alo = nlo + (m0 ? __ffs(m0) - 1 : 32 + __ffs(m1) - 1);
            int ahi = nlo + (m1 ? 63 - __clz(m1) : 31 - __clz(m0));
            if (ahi - alo + 1 > WM) {
                int wk = 1 << 30;
                if (S[0] == waveS) wk = nlo + lane;
                else if (S[1] == waveS) wk = nlo + lane + 32;
                const int wavek = __reduce_min_sync(FULL, wk);
                int l2 = wavek - (WM / 2 - 1); if (l2 < alo) l2 = alo;
                int h2 = l2 + WM - 1; if (h2 > ahi) h2 = ahi;
                l2 = h2 - WM + 1; if (l2 < alo) l2 = alo;
                alo = l2; ahi = h2;
            }
            lo = alo; hi = ahi; cur ^= 1;
            __syncwarp();
        }
        // winner: max S, then min d, then min k
        const int mS = __reduce_max_sync(FULL, bS);
        const int md = __reduce_min_sync(FULL, bS == mS ? bd : (1 << 30));
        const int mk = __reduce_min_sync(FULL, (bS == mS && bd == md) ? bk : (1 << 30));
        const u32 win = __ballot_sync(FULL, bS == mS && bd == md && bk == mk);
        const int wl = __ffs(win) - 1;
        const int besti = __shfl_sync(FULL, bi, wl), bestk = __shfl_sync(FULL, bk, wl);
        const int bestd = __shfl_sync(FULL, bd, wl), bestT = __shfl_sync(FULL, bT, wl);
        __syncwarp();
        if (lane == 0) {
            int2 *tl = tiles + (tile_off ? tile_off[task] : (int64_t)task * G.tile_stride);
            int n = NB(besti);
            int T = bestT, lastj = 0, lastd = 0;
            if (T >= 0) { int4 r = pool[T]; lastj = r.y; lastd = r.z; }
            int4 curr = T >= 0 ? pool[T] : make_int4(-1, 0, 0, 0);
            for (int q = n; q >= 1; q--) {
                int pj = 0, pd = 0; int4 pr = make_int4(-1, 0, 0, 0);
                if (curr.x >= 0) { pr = pool[curr.x]; pj = pr.y; pd = pr.z; }
                tl[q - 1] = make_int2(curr.y - pj, curr.z - pd);
                curr = pr;
            }
            int lastB = n > 0 ? firstT + (n - 1) * ts : 0;
            if (besti > lastB) { tl[n] = make_int2((besti - bestk) - lastj, bestd - lastd); n++; }
            outs[task] = ExtOut{besti, besti - bestk, bestd, n};
        }
        __syncwarp();
    }
}

// Register-resident variant for wmax <= 30 (the default): the window never exceeds 32 diagonals, lane l
// owns the diagonal congruent to l mod 32, and V / T / R live in registers; neighbours are one shuffle
// away.  No shared memory, no modular addressing, one slot per lane.  Same specification, same results.
template <int MINB, bool STAGE>
__global__ void __launch_bounds__(EXT_WARPS * 32, MINB) k_extend32(const Seed *__restrict__ seeds, int ntasks, ExtGeom G,
                                                             const int64_t *__restrict__ tile_off, int2 *__restrict__ tiles,
                                                             ExtOut *__restrict__ outs, int4 *__restrict__ pool_all,
                                                             int64_t pool_stride, int *__restrict__ counter, const int *__restrict__ order) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 FULL = 0xffffffffu;
    int4 *pool = pool_all + (int64_t)(blockIdx.x * EXT_WARPS + warp) * pool_stride;
    const int ts = G.ts, C = G.cdiff, X = G.xdrop, WM = G.wmax;
    // lane constants held in registers: left to itself the compiler re-derives them from SR_TID in every wave (S2R + 4 ALU)
    int lane_l = (lane + 31) & 31, lane_r = (lane + 1) & 31, lane_k = lane;
    asm volatile("" : "+r"(lane_l), "+r"(lane_r), "+r"(lane_k));
    const u32 lt_mask = (1u << lane) - 1u;
    __shared__ __align__(16) u32 s_seq[STAGE ? EXT_WARPS : 1][2][STAGE ? STAGE_WORDS : 4];
    __shared__ uint64_t s_bar[STAGE ? EXT_WARPS : 1];
    u32 parity = 0;
    if (STAGE) { if (lane == 0) mbar_init(&s_bar[warp], 1); __syncwarp(); }

    for (;;) {
        int task = 0;
        if (lane == 0) { task = atomicAdd(counter, 1); if (task < ntasks && order) task = order[task]; }   // longest expected tasks first
        task = __shfl_sync(FULL, task, 0);
        if (task >= ntasks) break;
        const Seed sd = seeds[task >> 1];
        const Task tk = make_task(sd, task & 1, G);
        const int la = tk.la, lb = tk.lb, firstT = tk.firstT;
        Staged SG; SG.sa = s_seq[STAGE ? warp : 0][0]; SG.sb = s_seq[STAGE ? warp : 0][1]; SG.a_base = SG.b_base = 0; SG.a_n = SG.b_n = 0;
        if (STAGE) {
            // 16-byte aligned windows starting at the 64-base boundary below the task's first base, clipped to the arrays
            SG.a_base = tk.ga & ~63u; SG.b_base = tk.gb & ~63u;
            const u32 wa = SG.a_base >> 4, wb = SG.b_base >> 4;
            const u32 na = min((u32)STAGE_WORDS, (tk.A == G.a_fwd || tk.A == G.a_rc ? G.a_words : G.b_words) - wa) & ~3u;
            const u32 nb = min((u32)STAGE_WORDS, (tk.B == G.b_fwd || tk.B == G.b_rc ? G.b_words : G.a_words) - wb) & ~3u;
            SG.a_n = (int)na * 16; SG.b_n = (int)nb * 16;
            __syncwarp();                                       // the previous task's LDS reads are done
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&s_bar[warp], (na + nb) * 4u);
                if (na) bulk_g2s(s_seq[warp][0], tk.A + wa, na * 4u, &s_bar[warp]);
                if (nb) bulk_g2s(s_seq[warp][1], tk.B + wb, nb * 4u, &s_bar[warp]);
            }
            mbar_wait(&s_bar[warp], parity); parity ^= 1u;
        }
        const int poolcap = G.poolmul * (ext_span(la, lb) / ts + 4);
        auto NB = [&](int i) -> int { return i >= firstT ? (int)__umulhi((u32)(i - firstT), G.ts_magic) + 1 : 0; };

        int npool = 0, lo = 0, hi = 0, d = 0;
        int V = NEGV, T = -1, R = 0;                          // this lane's diagonal
        int bS = NEGS, bi = 0, bk = 0, bd = 0, bT = -1;     // lane-local best
        int gbest;
        {   // wave 0 (uniform): diagonal 0 lives in lane 0
            const int i = STAGE ? slide_s(tk.A, tk.ga, tk.B, tk.gb, la < lb ? la : lb, SG) : slide(tk.A, tk.ga, tk.B, tk.gb, la < lb ? la : lb);
            const int n = NB(i);
            if (n > poolcap) {
                if (lane == 0) outs[task] = ExtOut{0, 0, 0, 0};
                continue;
            }
            if (lane == 0) {
                int t = -1;
                for (int q = 1; q <= n; q++) { pool[npool] = make_int4(t, firstT + (q - 1) * ts, 0, q); t = npool++; }
                V = i; T = t; R = firstT + n * ts;
                bS = 3 * (2 * i); bi = i; bk = 0; bd = 0; bT = t;
            }
            npool = n;
            gbest = 3 * (2 * i);
            if (i == la || i == lb) { lo = 1; hi = 0; }
            __syncwarp();
        }
        int Cd = 0;                                           // C * d, carried instead of multiplied
        while (lo <= hi) {
            d++; Cd += C;
            const int nlo = lo - 1;
            const int k = nlo + ((lane_k - nlo) & 31);          // the diagonal = lane (mod 32) inside the new window
            // invariant: V == NEGV in every lane whose diagonal is outside [lo, hi] (kept at the end of each wave), so the
            // three predecessors need no range checks and lanes beyond hi + 1 fall out by themselves
            const int Vl = __shfl_sync(FULL, V, lane_l), Vr = __shfl_sync(FULL, V, lane_r);
            int i = V + 1, src = lane_k;
            if (Vl + 1 > i) { i = Vl + 1; src = lane_l; }
            if (Vr > i) { i = Vr; src = lane_r; }
            int cT = __shfl_sync(FULL, T, src), cR = __shfl_sync(FULL, R, src);
            int cn = 0;
            {
                const int j = i - k;
                if (i <= NEGV + 1 || i > la || j > lb || j < 0) i = NEGV;
                else {
                    i += STAGE ? slide_s(tk.A, tk.ga + i, tk.B, tk.gb + j, min(la - i, lb - j), SG)
                               : slide(tk.A, tk.ga + i, tk.B, tk.gb + j, min(la - i, lb - j));
                    // tile boundaries crossed: almost always 0 or 1 (a second one needs a slide of >= ts bases); written
                    // this way the common case costs one compare instead of the division the counted loop compiles to
                    // (ts_magic: exact floor(x / ts) for every x the length check in align_blocks admits)
                    if (i >= cR) {
                        cn = 1; cR += ts;
                        if (i >= cR) { const int more = (int)__umulhi((u32)(i - cR), G.ts_magic) + 1; cn += more; cR += more * ts; }
                    }
                }
            }
            // the wave's collectives sit together in one convergent stretch (one divergence check instead of three); nothing is
            // committed before the pool check below, so an exhausted pool still stops BEFORE this wave
            const int S = i > NEGV ? 3 * (2 * i - k) - Cd : NEGS;
            const u32 cm = __ballot_sync(FULL, cn != 0);
            const int waveS = __reduce_max_sync(FULL, S);
            const int nbest = max(gbest, waveS);
            const bool alive = S >= nbest - X && i != la && i - k != lb;      // S = NEGS (dead cell) fails the first test
            const u32 m = __ballot_sync(FULL, alive);
            if (cm) {
                if (!__any_sync(FULL, cn > 1)) {                // the usual case: at most one tile boundary per cell
                    const int need = __popc(cm);
                    if (npool + need > poolcap) break;          // pool exhausted: stop before this wave
                    if (cn) {
                        const int at = npool + __popc(cm & lt_mask);
                        pool[at] = make_int4(cT, cR - ts - k, d, 0); cT = at;
                    }
                    npool += need;
                } else {
                    const int need = __reduce_add_sync(FULL, cn);
                    int pre = cn;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, pre, o); if (lane >= o) pre += t; }
                    int at = npool + pre - cn;
                    if (npool + need > poolcap) break;
#pragma unroll 1
                    for (int q = cn; q >= 1; q--) { pool[at] = make_int4(cT, cR - q * ts - k, d, 0); cT = at++; }
                    npool += need;
                }
            }
            if (S > bS) { bS = S; bi = i; bk = k; bd = d; bT = cT; }
            gbest = nbest;
            V = alive ? i : NEGV; T = cT; R = cR;
            if (m == 0u) break;
            const u32 rot = __funnelshift_r(m, m, nlo & 31);    // bit j <-> diagonal nlo + j
            int alo = nlo + __ffs(rot) - 1, ahi = nlo + 31 - __clz(rot);
            if (ahi - alo + 1 > WM) {
                const u32 mb = __ballot_sync(FULL, S == waveS);
                const u32 rb = __funnelshift_r(mb, mb, nlo & 31);
                const int wavek = nlo + __ffs(rb) - 1;
                int l2 = wavek - (WM / 2 - 1); if (l2 < alo) l2 = alo;
                int h2 = l2 + WM - 1; if (h2 > ahi) h2 = ahi;
                l2 = h2 - WM + 1; if (l2 < alo) l2 = alo;
                alo = l2; ahi = h2;
                if (k < alo || k > ahi) V = NEGV;              // trimmed off the window
            }
            lo = alo; hi = ahi;
        }
        // winner: max S, then min d, then min k
        const int mS = __reduce_max_sync(FULL, bS);
        const int md = __reduce_min_sync(FULL, bS == mS ? bd : (1 << 30));
        const int mk = __reduce_min_sync(FULL, (bS == mS && bd == md) ? bk : (1 << 30));
        const u32 win = __ballot_sync(FULL, bS == mS && bd == md && bk == mk);
        const int wl = __ffs(win) - 1;
        const int besti = __shfl_sync(FULL, bi, wl), bestk = __shfl_sync(FULL, bk, wl);
        const int bestd = __shfl_sync(FULL, bd, wl), bestT = __shfl_sync(FULL, bT, wl);
        __syncwarp();
        if (lane == 0) {
            int2 *tl = tiles + (tile_off ? tile_off[task] : (int64_t)task * G.tile_stride);
            int n = NB(besti);
            int lastj = 0, lastd = 0;
            int4 curr = bestT >= 0 ? pool[bestT] : make_int4(-1, 0, 0, 0);
            if (bestT >= 0) { lastj = curr.y; lastd = curr.z; }
            for (int q = n; q >= 1; q--) {
                int pj = 0, pd = 0; int4 pr = make_int4(-1, 0, 0, 0);
                if (curr.x >= 0) { pr = pool[curr.x]; pj = pr.y; pd = pr.z; }
                tl[q - 1] = make_int2(curr.y - pj, curr.z - pd);
                curr = pr;
            }
            const int lastB = n > 0 ? firstT + (n - 1) * ts : 0;
            if (besti > lastB) { tl[n] = make_int2((besti - bestk) - lastj, bestd - lastd); n++; }
            outs[task] = ExtOut{besti, besti - bestk, bestd, n};
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------- task order
// The warps fetch tasks from one counter; a task is a chain of dependent waves that no second warp can help with, so the
// kernel ends when the LAST-started long task ends.  Handing the tasks out longest-expected-first (counting sort into
// half-octave classes of min(la, lb), largest class first) leaves only short tasks for the tail.  The order of the tasks
// has no influence on any result: every task writes its own output slots.
constexpr int ORD_CLASSES = 64;
__device__ __forceinline__ int order_class(const Seed &sd, int dir, const ExtGeom &G) {
    const Task t = make_task(sd, dir, G);
    const u32 span = (u32)min(t.la, t.lb) | 1u;
    const int msb = 31 - __clz(span);
    const int c = 2 * msb + (msb > 0 ? (int)((span >> (msb - 1)) & 1u) : 0);
    return ORD_CLASSES - 1 - c;                      // class 0 = longest
}
__global__ void __launch_bounds__(256) k_task_classes(const Seed *__restrict__ seeds, int ntasks, ExtGeom G, int *__restrict__ counts) {
    __shared__ int h[ORD_CLASSES];
    if (threadIdx.x < ORD_CLASSES) h[threadIdx.x] = 0;
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ntasks) atomicAdd(&h[order_class(seeds[t >> 1], t & 1, G)], 1);
    __syncthreads();
    if (threadIdx.x < ORD_CLASSES && h[threadIdx.x]) atomicAdd(&counts[threadIdx.x], h[threadIdx.x]);
}
__global__ void __launch_bounds__(256) k_task_order(const Seed *__restrict__ seeds, int ntasks, ExtGeom G, const int *__restrict__ counts,
                                                    int *__restrict__ cursors, int *__restrict__ order) {
    // a few classes hold almost every task: ranks are taken inside the CTA (shared-memory atomics), one global atomic per
    // (CTA, class) reserves the CTA's slots
    __shared__ int h[ORD_CLASSES], gbase[ORD_CLASSES];
    if (threadIdx.x < ORD_CLASSES) h[threadIdx.x] = 0;
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    int c = 0, r = 0;
    if (t < ntasks) { c = order_class(seeds[t >> 1], t & 1, G); r = atomicAdd(&h[c], 1); }
    __syncthreads();
    if (threadIdx.x < ORD_CLASSES && h[threadIdx.x]) {
        int a = 0;
        for (int q = 0; q < (int)threadIdx.x; q++) a += counts[q];
        gbase[threadIdx.x] = a + atomicAdd(&cursors[threadIdx.x], h[threadIdx.x]);
    }
    __syncthreads();
    if (t < ntasks) order[gbase[c] + r] = t;
}

// ---------------------------------------------------------------- candidate assembly

__global__ void __launch_bounds__(256) k_combine(const Seed *__restrict__ seeds, int nseeds, ExtGeom G, int minlen,
                                                 const int64_t *__restrict__ tile_off, const int2 *__restrict__ tiles,
                                                 const ExtOut *__restrict__ outs, Cand *__restrict__ cand,
                                                 int32_t *__restrict__ valid, u32 *__restrict__ ntl) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseeds) return;
    const Seed sd = seeds[s];
    const ExtOut fo = outs[2 * s], ro = outs[2 * s + 1];
    const int ts = G.ts;
    const int ab = sd.apos - ro.i_end, bb = sd.bpos - ro.j_end, ae = sd.apos + fo.i_end, be = sd.bpos + fo.j_end;
    const bool ok = (ae - ab) + (be - bb) >= 2 * minlen;
    const bool merge = (sd.apos % ts != 0) && ro.ntiles > 0 && fo.ntiles > 0;
    const int nt = ro.ntiles + fo.ntiles - (merge ? 1 : 0);
    Cand c;
    c.a = sd.a; c.bs = sd.bs; c.ab = ab; c.ae = ae; c.bb = bb; c.be = be; c.diffs = fo.d_end + ro.d_end; c.nt = nt; c.toff = 0;
    int dmin = ab - bb, dmax = dmin;
    if (ok) {
        const int2 *ft = tiles + (tile_off ? tile_off[2 * s] : (int64_t)(2 * s) * G.tile_stride), *rt = tiles + (tile_off ? tile_off[2 * s + 1] : (int64_t)(2 * s + 1) * G.tile_stride);
        int apos = ab, bpos = bb, q = 0;
        auto step = [&](int bbases) {
            int aend = (q == nt - 1) ? ae : (apos / ts + 1) * ts;
            bpos += bbases; apos = aend; q++;
            int dg = apos - bpos; dmin = min(dmin, dg); dmax = max(dmax, dg);
        };
        for (int x = ro.ntiles - 1; x >= (merge ? 1 : 0); x--) step(rt[x].x);
        if (merge) step(rt[0].x + ft[0].x);
        for (int x = merge ? 1 : 0; x < fo.ntiles; x++) step(ft[x].x);
    }
    c.dmin = dmin; c.dmax = dmax;
    cand[s] = c;
    valid[s] = ok ? 1 : 0;
    ntl[s] = ok ? (u32)nt : 0u;
}

__global__ void __launch_bounds__(256) k_write_traces(const Seed *__restrict__ seeds, int nseeds, ExtGeom G,
                                                      const int64_t *__restrict__ tile_off, const int2 *__restrict__ tiles,
                                                      const ExtOut *__restrict__ outs, const Cand *__restrict__ cand,
                                                      const int32_t *__restrict__ valid, const int32_t *__restrict__ vidx,
                                                      const int64_t *__restrict__ toff, Cand *__restrict__ cand_out,
                                                      uint16_t *__restrict__ trace) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseeds || !valid[s]) return;
    const Seed sd = seeds[s];
    const ExtOut fo = outs[2 * s], ro = outs[2 * s + 1];
    const bool merge = (sd.apos % G.ts != 0) && ro.ntiles > 0 && fo.ntiles > 0;
    const int2 *ft = tiles + (tile_off ? tile_off[2 * s] : (int64_t)(2 * s) * G.tile_stride), *rt = tiles + (tile_off ? tile_off[2 * s + 1] : (int64_t)(2 * s + 1) * G.tile_stride);
    Cand c = cand[s];
    c.toff = 2 * toff[s];
    uint16_t *o = trace + c.toff;
    for (int x = ro.ntiles - 1; x >= (merge ? 1 : 0); x--) { *o++ = (uint16_t)rt[x].y; *o++ = (uint16_t)rt[x].x; }
    if (merge) { *o++ = (uint16_t)(rt[0].y + ft[0].y); *o++ = (uint16_t)(rt[0].x + ft[0].x); }
    for (int x = merge ? 1 : 0; x < fo.ntiles; x++) { *o++ = (uint16_t)ft[x].y; *o++ = (uint16_t)ft[x].x; }
    cand_out[vidx[s]] = c;
}

// ---------------------------------------------------------------- hit retirement

__device__ __forceinline__ u64 gkey(int bs, int a) { return ((u64)(u32)bs << 32) | (u32)a; }

__device__ __forceinline__ int group_lower(const Cand *__restrict__ c, int lo, int hi, u64 key) {
    while (lo < hi) { int m = (lo + hi) >> 1; if (gkey(c[m].bs, c[m].a) < key) lo = m + 1; else hi = m; }
    return lo;
}

// The kept alignments of a hit's (bread, strand, aread) group, looked up ONCE per hot band (all hits of a band share the group)
// instead of once per hit: crange[band] = [first, last) candidate of the group in rc (empty: the group kept nothing this round).
__global__ void __launch_bounds__(256) k_band_candidates(const ulonglong2 *__restrict__ hits, const int32_t *__restrict__ bfirst,
                                                         const uint8_t *__restrict__ hot, int32_t nbands, const Cand *__restrict__ rc,
                                                         const int32_t *__restrict__ d_nrc, SeedGeom G, int2 *__restrict__ crange) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nbands) return;
    if (!hot[q]) { crange[q] = make_int2(0, 0); return; }
    const int nrc = *d_nrc;
    const ulonglong2 h = hits[bfirst[q]];
    const u64 gd = h.x & ((1ull << G.gdbits) - 1ull);
    const int bs = (int)(h.x >> G.gdbits);
    int lo = 0, hi = G.na;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if ((u64)G.a_dbase[mid] <= gd) lo = mid; else hi = mid; }
    const u64 key = gkey(bs, lo);
    int x = group_lower(rc, 0, nrc, key), e = x;
    while (e < nrc && gkey(rc[e].bs, rc[e].a) == key) e++;
    crange[q] = make_int2(x, e);
}

__global__ void __launch_bounds__(256) k_retire(const ulonglong2 *__restrict__ hits, int64_t n, const uint8_t *__restrict__ consumed,
                                                const Cand *__restrict__ rc, const int2 *__restrict__ crange, int w,
                                                const int32_t *__restrict__ bflag, const int32_t *__restrict__ bidx,
                                                const uint8_t *__restrict__ hot, int32_t *__restrict__ keep) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int kp = consumed[i] ? 0 : 1;
    // a hit in a cold band can never become part of a hot band later (scores only shrink): drop it for good.
    // Pure optimisation -- the later rounds see exactly the clusters they would have seen anyway.
    const int band = bflag[i] ? bidx[i] : bidx[i] - 1;
    if (kp && !hot[band]) kp = 0;
    if (kp) {
        const int2 cr = crange[band];
        // the rounds are counted per (bread, strand, aread) group (spec item 7): a group that kept no alignment in this
        // round is finished, whatever the other groups of the block do -- all its hits go
        if (cr.x >= cr.y) kp = 0;
        else {
            const ulonglong2 h = hits[i];
            const int apos = (int)(u32)h.y, bpos = (int)(h.y >> 32), diag = apos - bpos;
            for (int x = cr.x; x < cr.y; x++) {
                const Cand c = rc[x];
                if (apos >= c.ab && apos <= c.ae && diag >= c.dmin - (1 << w) && diag <= c.dmax + (1 << w)) { kp = 0; break; }
            }
        }
    }
    keep[i] = kp;
}

__global__ void __launch_bounds__(256) k_compact_hits(const ulonglong2 *__restrict__ hits, int64_t n, const int32_t *__restrict__ keep,
                                                      const int32_t *__restrict__ kidx, ulonglong2 *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (keep[i]) out[kidx[i]] = hits[i];
}

// drop[j] = some other candidate i of the same (a, bs) contains j in both coordinates (identical: lower index wins)
__global__ void __launch_bounds__(256) k_dedupe(const Cand *__restrict__ c, int n, const int32_t *__restrict__ rbeg, int nrounds,
                                                uint8_t *__restrict__ drop) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const Cand x = c[j];
    const u64 key = gkey(x.bs, x.a);
    int dr = 0;
    for (int r = 0; r < nrounds && !dr; r++) {
        const int e = rbeg[r + 1];
        for (int i = group_lower(c, rbeg[r], e, key); i < e && gkey(c[i].bs, c[i].a) == key; i++) {
            if (i == j) continue;
            const Cand y = c[i];
            if (y.ab <= x.ab && x.ae <= y.ae && y.bb <= x.bb && x.be <= y.be) {
                bool same = y.ab == x.ab && x.ae == y.ae && y.bb == x.bb && x.be == y.be;
                if (!same || i < j) { dr = 1; break; }
            }
        }
    }
    drop[j] = (uint8_t)dr;
}

// ---------------------------------------------------------------- final ordering on the device

// LAsort order (base.d:1787-1809) by four stable LSD sorts over (key, candidate index) items
__global__ void __launch_bounds__(256) k_final_setkey(const Cand *__restrict__ c, const uint8_t *__restrict__ drop,
                                                      ulonglong2 *__restrict__ items, int n, int field, FinalBits fb,
                                                      unsigned long long *__restrict__ ndrop) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 idx = (field == 0 || field == 4) ? (u64)i : items[i].y;      // fields 0 and 4 start a sort: items are not set yet
    const Cand x = c[idx];
    u64 key;
    if (field == 0) key = (u64)(u32)x.diffs;
    else if (field == 1) key = ((u64)(u32)x.bb << fb.nb) | (u32)x.be;
    else if (field == 2) key = ((u64)(x.bs >= fb.nb_reads ? 1 : 0) << (2 * fb.na)) | ((u64)(u32)x.ab << fb.na) | (u32)x.ae;
    else if (field == 3) {
        const u64 d = drop[idx] ? 1ull : 0ull;
        key = (d << (fb.nra + fb.nrb)) | ((u64)(u32)x.a << fb.nrb) | (u32)(x.bs >= fb.nb_reads ? x.bs - fb.nb_reads : x.bs);
        if (d) atomicAdd(ndrop, 1ull);
    } else {
        // field 4: the leading fields of the LAsort order in ONE key -- (dropped, aread, bread, comp, abpos); the few records
        // that tie on it are put in order by k_final_fixup
        const u64 d = drop[idx] ? 1ull : 0ull;
        const u64 comp = x.bs >= fb.nb_reads ? 1ull : 0ull;
        key = (d << (fb.nra + fb.nrb + 1 + fb.na)) | ((u64)(u32)x.a << (fb.nrb + 1 + fb.na)) |
              ((u64)(u32)(comp ? x.bs - fb.nb_reads : x.bs) << (1 + fb.na)) | (comp << fb.na) | (u64)(u32)x.ab;
        if (d) atomicAdd(ndrop, 1ull);
    }
    items[i] = make_ulonglong2(key, idx);
}

// runs of items with equal primary key (same aread, bread, comp, abpos: rare and tiny) ordered by the remaining LAsort
// fields (aepos, bbpos, bepos, diffs) and the candidate index: one thread per run start, insertion sort in place
__global__ void __launch_bounds__(256) k_final_fixup(const Cand *__restrict__ c, ulonglong2 *__restrict__ items, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (i > 0 && items[i - 1].x == items[i].x)) return;
    int j = i + 1;
    while (j < n && items[j].x == items[i].x) j++;
    if (j - i < 2) return;
    auto less = [&](u64 p, u64 q) {
        const Cand &x = c[p], &y = c[q];
        if (x.ae != y.ae) return x.ae < y.ae;
        if (x.bb != y.bb) return x.bb < y.bb;
        if (x.be != y.be) return x.be < y.be;
        if (x.diffs != y.diffs) return x.diffs < y.diffs;
        return p < q;
    };
    for (int a = i + 1; a < j; a++) {
        const ulonglong2 t = items[a]; int b = a;
        while (b > i && less(t.y, items[b - 1].y)) { items[b] = items[b - 1]; b--; }
        items[b] = t;
    }
}

__global__ void __launch_bounds__(256) k_final_records(const Cand *__restrict__ c, const ulonglong2 *__restrict__ items, int ncand,
                                                       int nb_reads, dn_las_record *__restrict__ rec, u32 *__restrict__ tl,
                                                       unsigned long long *__restrict__ ctr /* [0] dropped, [1] aligned bases, [2] ext bytes */) {
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= ncand) return;
    if (o >= ncand - (int)ctr[0]) { tl[o] = 0u; return; }          // dropped duplicates sort to the end: no record, no trace
    unsigned long long *acc = ctr + 1;
    const Cand x = c[items[o].y];
    dn_las_record q;
    q.tlen = 2 * x.nt; q.diffs = x.diffs; q.abpos = x.ab; q.bbpos = x.bb; q.aepos = x.ae; q.bepos = x.be;
    const int comp = x.bs >= nb_reads ? 1 : 0;
    q.flags = comp ? DN_LAS_COMP : 0u; q.aread = x.a; q.bread = x.bs - comp * nb_reads; q.pad_ = 0;
    rec[o] = q; tl[o] = (u32)(2 * x.nt);
    atomicAdd(&acc[0], (unsigned long long)(x.ae - x.ab));
    atomicAdd(&acc[1], (unsigned long long)((x.ae - x.ab) / 4 + (x.be - x.bb) / 4 + 40 + 4 * x.nt));
}

// one warp per record: copy its (diffs, bbases) pairs from the round buffer to the output position
__global__ void __launch_bounds__(256) k_final_traces(const Cand *__restrict__ c, const ulonglong2 *__restrict__ items, int ncand,
                                                      const unsigned long long *__restrict__ ctr,
                                                      const int64_t *__restrict__ toff, FinalGeom G, uint16_t *__restrict__ out) {
    const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (o >= ncand - (int)ctr[0]) return;
    const int j = (int)items[o].y;
    const Cand x = c[j];
    int r = 0;
    while (r + 1 < G.nrounds && j >= G.round_beg[r + 1]) r++;
    const uint16_t *src = G.round_trace[r] + x.toff;
    uint16_t *dst = out + toff[o];
    for (int q = lane; q < 2 * x.nt; q += 32) dst[q] = src[q];
}

}  // namespace

void launch_final_setkey(const Cand *c, const uint8_t *drop, ulonglong2 *items, int n, int field, FinalBits fb,
                         unsigned long long *ndrop, cudaStream_t s) {
    DN_LAUNCH(k_final_setkey, (n + 255) / 256, 256, 0, s, c, drop, items, n, field, fb, ndrop);
}
void launch_final_fixup(const Cand *c, ulonglong2 *items, int n, cudaStream_t s) {
    DN_LAUNCH(k_final_fixup, (n + 255) / 256, 256, 0, s, c, items, n);
}
void launch_final_records(const Cand *c, const ulonglong2 *items, int ncand, int nb_reads, dn_las_record *rec, u32 *tl, unsigned long long *ctr, cudaStream_t s) {
    DN_LAUNCH(k_final_records, (ncand + 255) / 256, 256, 0, s, c, items, ncand, nb_reads, rec, tl, ctr);
}
void launch_final_traces(const Cand *c, const ulonglong2 *items, int ncand, const unsigned long long *ctr, const int64_t *toff, FinalGeom G, uint16_t *out, cudaStream_t s) {
    DN_LAUNCH(k_final_traces, (unsigned)(((int64_t)ncand * 32 + 255) / 256), 256, 0, s, c, items, ncand, ctr, toff, G, out);
}
void launch_task_caps(const Seed *seeds, int nseeds, ExtGeom G, u32 *caps, cudaStream_t s) {
    DN_LAUNCH(k_task_caps, (nseeds + 255) / 256, 256, 0, s, seeds, nseeds, G, caps);
}
// order: 2 * nseeds ints; scratch: 2 * 64 ints, zeroed here
void launch_task_order(const Seed *seeds, int nseeds, ExtGeom G, int *scratch, int *order, cudaStream_t s) {
    DN_CUDA(cudaMemsetAsync(scratch, 0, sizeof(int) * 2 * ORD_CLASSES, s));
    DN_LAUNCH(k_task_classes, (2 * nseeds + 255) / 256, 256, 0, s, seeds, 2 * nseeds, G, scratch);
    DN_LAUNCH(k_task_order, (2 * nseeds + 255) / 256, 256, 0, s, seeds, 2 * nseeds, G, (const int *)scratch, scratch + ORD_CLASSES, order);
}
void launch_extend(const Seed *seeds, int nseeds, ExtGeom G, const int64_t *tile_off, int2 *tiles, ExtOut *outs,
                   int4 *pool, int64_t pool_stride, int nwarps_total, int *counter, const int *order, cudaStream_t s) {
    int ctas = nwarps_total / EXT_WARPS;
    // resident CTAs per SM: 4 by default (DN_EXT_CTAS = 2..6 for the sweep recorded in DESIGN.md; 6 = the 40-register build, a few spilled words)
    static const bool stage = getenv("DN_EXT_TMA") != nullptr;
    if (G.wmax <= 30 && stage) DN_LAUNCH((k_extend32<5, true>), ctas, EXT_WARPS * 32, 0, s, seeds, 2 * nseeds, G, tile_off, tiles, outs, pool, pool_stride, counter, order);
    else if (G.wmax <= 30 && ext_ctas_per_sm() == 2) DN_LAUNCH((k_extend32<2, false>), ctas, EXT_WARPS * 32, 0, s, seeds, 2 * nseeds, G, tile_off, tiles, outs, pool, pool_stride, counter, order);
    else if (G.wmax <= 30 && ext_ctas_per_sm() == 3) DN_LAUNCH((k_extend32<3, false>), ctas, EXT_WARPS * 32, 0, s, seeds, 2 * nseeds, G, tile_off, tiles, outs, pool, pool_stride, counter, order);
    else if (G.wmax <= 30 && ext_ctas_per_sm() == 4) DN_LAUNCH((k_extend32<4, false>), ctas, EXT_WARPS * 32, 0, s, seeds, 2 * nseeds, G, tile_off, tiles, outs, pool, pool_stride, counter, order);
    else if (G.wmax <= 30 && ext_ctas_per_sm() >= 6) DN_LAUNCH((k_extend32<6, false>), ctas, EXT_WARPS * 32, 0, s, seeds, 2 * nseeds, G, tile_off, tiles, outs, pool, pool_stride, counter, order);
    else if (G.wmax <= 30) DN_LAUNCH((k_extend32<5, false>), ctas, EXT_WARPS * 32, 0, s, seeds, 2 * nseeds, G, tile_off, tiles, outs, pool, pool_stride, counter, order);
    else DN_LAUNCH(k_extend, ctas, EXT_WARPS * 32, 0, s, seeds, 2 * nseeds, G, tile_off, tiles, outs, pool, pool_stride, counter, order);
}
void launch_combine(const Seed *seeds, int nseeds, ExtGeom G, int minlen, const int64_t *tile_off, const int2 *tiles,
                    const ExtOut *outs, Cand *cand_all, int32_t *valid, u32 *ntl, cudaStream_t s) {
    DN_LAUNCH(k_combine, (nseeds + 255) / 256, 256, 0, s, seeds, nseeds, G, minlen, tile_off, tiles, outs, cand_all, valid, ntl);
}
void launch_write_traces(const Seed *seeds, int nseeds, ExtGeom G, const int64_t *tile_off, const int2 *tiles,
                         const ExtOut *outs, const Cand *cand_all, const int32_t *valid, const int32_t *vidx,
                         const int64_t *toff, Cand *cand_out, uint16_t *trace, cudaStream_t s) {
    DN_LAUNCH(k_write_traces, (nseeds + 255) / 256, 256, 0, s, seeds, nseeds, G, tile_off, tiles, outs, cand_all, valid, vidx,
              toff, cand_out, trace);
}
void launch_retire(const ulonglong2 *hits, int64_t n, const uint8_t *consumed, const Cand *rc, const int32_t *d_nrc, SeedGeom G, int w,
                   const int32_t *bflag, const int32_t *bidx, const uint8_t *hot, const int32_t *bfirst, int32_t nbands, int2 *crange,
                   int32_t *keep, cudaStream_t s) {
    DN_LAUNCH(k_band_candidates, (unsigned)((nbands + 255) / 256), 256, 0, s, hits, bfirst, hot, nbands, rc, d_nrc, G, crange);
    DN_LAUNCH(k_retire, (unsigned)((n + 255) / 256), 256, 0, s, hits, n, consumed, rc, (const int2 *)crange, w, bflag, bidx, hot, keep);
}
void launch_compact_hits(const ulonglong2 *hits, int64_t n, const int32_t *keep, const int32_t *kidx, ulonglong2 *out, cudaStream_t s) {
    DN_LAUNCH(k_compact_hits, (unsigned)((n + 255) / 256), 256, 0, s, hits, n, keep, kidx, out);
}
void launch_dedupe(const Cand *cands, int ncand, const int32_t *round_beg, int nrounds, uint8_t *drop, cudaStream_t s) {
    DN_LAUNCH(k_dedupe, (ncand + 255) / 256, 256, 0, s, cands, ncand, round_beg, nrounds, drop);
}

int ext_warps_per_cta() { return EXT_WARPS; }
int ext_ctas_per_sm() { static const int v = getenv("DN_EXT_CTAS") ? atoi(getenv("DN_EXT_CTAS")) : 4; return v < 1 ? 1 : v; }

}  // namespace dn

// ---------------------------------------------------------------- LAS merge (what LAmerge does, Snakefile:1173-1200)
// Device-resident concatenation of LAS segments (e.g. the result of the NCCL all-gatherv) -> one LAS in
// LAsort order (base.d:1787-1809) with its traces gathered: four stable LSD sorts of (key, index) items.
namespace dn {
namespace {

__global__ void __launch_bounds__(256) k_merge_tlen(const dn_las_record *__restrict__ rec, int64_t n, u32 *__restrict__ tl) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) tl[i] = (u32)rec[i].tlen;
}
__global__ void __launch_bounds__(256) k_merge_setkey(const dn_las_record *__restrict__ rec, ulonglong2 *__restrict__ items, int64_t n,
                                                      int field, FinalBits fb) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 idx = field == 0 ? (u64)i : items[i].y;
    const dn_las_record x = rec[idx];
    u64 key;
    if (field == 0) key = (u64)(u32)x.diffs;
    else if (field == 1) key = ((u64)(u32)x.bbpos << fb.nb) | (u32)x.bepos;
    else if (field == 2) key = ((u64)(x.flags & DN_LAS_COMP) << (2 * fb.na)) | ((u64)(u32)x.abpos << fb.na) | (u32)x.aepos;
    else key = ((u64)(u32)x.aread << fb.nrb) | (u32)x.bread;
    items[i] = make_ulonglong2(key, idx);
}
__global__ void __launch_bounds__(256) k_merge_records(const dn_las_record *__restrict__ rec, const ulonglong2 *__restrict__ items, int64_t n,
                                                       dn_las_record *__restrict__ out, u32 *__restrict__ tl) {
    int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n) return;
    const dn_las_record x = rec[items[o].y];
    out[o] = x; tl[o] = (u32)x.tlen;
}
__global__ void __launch_bounds__(256) k_merge_traces(const ulonglong2 *__restrict__ items, int64_t n, const int64_t *__restrict__ src_toff,
                                                      const int64_t *__restrict__ dst_toff, const dn_las_record *__restrict__ out_rec,
                                                      const uint16_t *__restrict__ src, uint16_t *__restrict__ dst) {
    const int64_t o = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; const int lane = threadIdx.x & 31;
    if (o >= n) return;
    const uint16_t *s = src + src_toff[items[o].y];
    uint16_t *d = dst + dst_toff[o];
    for (int q = lane; q < out_rec[o].tlen; q += 32) d[q] = s[q];
}
int bits_of(uint64_t v) { int b = 0; while (v) { b++; v >>= 1; } return b; }

}  // namespace


// forceFlat (dazzler.d:4084-4093): clear the chain flags and put the records into FlatLocalAlignment order;
// the traces stay where they are (toff is permuted with the records).
namespace { __global__ void __launch_bounds__(256) k_flat_gather(const dn_las_record *__restrict__ rec, const int64_t *__restrict__ toff,
                                                                  const ulonglong2 *__restrict__ items, int64_t n,
                                                                  dn_las_record *__restrict__ orec, int64_t *__restrict__ otoff) {
    int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n) return;
    const u64 i = items[o].y;
    dn_las_record x = rec[i];
    x.flags &= (DN_LAS_COMP | DN_LAS_ELIM);
    orec[o] = x; otoff[o] = toff[i];
} }

void force_flat_device(dn_las_record *h_rec, int64_t *h_toff, int64_t n, cudaStream_t s) {
    arena().reset();
    if (n == 0) return;
    int64_t mxa = 1, mxb = 1, na = 1, nb = 1;
    for (int64_t i = 0; i < n; i++) {
        mxa = std::max<int64_t>(mxa, h_rec[i].aepos); mxb = std::max<int64_t>(mxb, h_rec[i].bepos);
        na = std::max<int64_t>(na, (int64_t)h_rec[i].aread + 1); nb = std::max<int64_t>(nb, (int64_t)h_rec[i].bread + 1);
    }
    DBuf<dn_las_record> drec(n), orec(n); DBuf<int64_t> dtoff(n), otoff(n); DBuf<ulonglong2> it1(n), it2(n);
    DN_CUDA(cudaMemcpyAsync(drec.p, h_rec, sizeof(dn_las_record) * n, cudaMemcpyHostToDevice, s));
    DN_CUDA(cudaMemcpyAsync(dtoff.p, h_toff, sizeof(int64_t) * n, cudaMemcpyHostToDevice, s));
    FinalBits fb{bits_of((uint64_t)mxa), bits_of((uint64_t)mxb), bits_of((uint64_t)na), bits_of((uint64_t)nb), 0};
    const int fbits[4] = {bits_of((uint64_t)mxa + mxb), 2 * fb.nb, 2 * fb.na + 1, fb.nra + fb.nrb};
    ulonglong2 *cur = it1.p, *oth = it2.p;
    for (int f = 0; f < 4; f++) {
        DN_LAUNCH(k_merge_setkey, (unsigned)((n + 255) / 256), 256, 0, s, (const dn_las_record *)drec.p, cur, n, f, fb);
        ulonglong2 *res = radix_sort_rec16(cur, oth, n, 0, 0, fbits[f], s);
        if (res != cur) { oth = cur; cur = res; }
    }
    DN_LAUNCH(k_flat_gather, (unsigned)((n + 255) / 256), 256, 0, s, (const dn_las_record *)drec.p, (const int64_t *)dtoff.p,
              (const ulonglong2 *)cur, n, orec.p, otoff.p);
    DN_CUDA(cudaMemcpyAsync(h_rec, orec.p, sizeof(dn_las_record) * n, cudaMemcpyDeviceToHost, s));
    DN_CUDA(cudaMemcpyAsync(h_toff, otoff.p, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, s));
    DN_CUDA(cudaStreamSynchronize(s));
}

void merge_las_device(const dn_las_record *d_rec, int64_t n, const uint16_t *d_trace, int64_t ntrace, int64_t max_alen, int64_t max_blen,
                      int64_t na_reads, int64_t nb_reads, HostLas &out, cudaStream_t s, bool reset_arena) {
    out = HostLas();
    if (reset_arena) arena().reset();                  // false: the inputs themselves live in the arena (dn_las_transpose)
    out.rec = (dn_las_record *)hcache_alloc(sizeof(dn_las_record) * (size_t)(n + 1));
    out.toff = (int64_t *)hcache_alloc(sizeof(int64_t) * (size_t)(n + 1));
    out.trace = (uint16_t *)hcache_alloc(sizeof(uint16_t) * (size_t)(ntrace + 1));
    out.nrec = n; out.ntrace = ntrace;
    if (n == 0) return;
    if (n >= (1ll << 31)) throw Error("too many records to merge");
    DBuf<u32> tl(n); DBuf<int64_t> stoff(n), dtoff(n), tot(1);
    DN_LAUNCH(k_merge_tlen, (unsigned)((n + 255) / 256), 256, 0, s, d_rec, n, tl.p);
    exclusive_scan_u32_to_i64(tl.p, stoff.p, n, tot.p, s);
    FinalBits fb{bits_of((uint64_t)max_alen), bits_of((uint64_t)max_blen), bits_of((uint64_t)na_reads), bits_of((uint64_t)nb_reads), 0};
    const int fbits[4] = {bits_of((uint64_t)max_alen + max_blen), 2 * fb.nb, 2 * fb.na + 1, fb.nra + fb.nrb};
    DBuf<ulonglong2> it1(n), it2(n);
    ulonglong2 *cur = it1.p, *oth = it2.p;
    for (int f = 0; f < 4; f++) {
        DN_LAUNCH(k_merge_setkey, (unsigned)((n + 255) / 256), 256, 0, s, d_rec, cur, n, f, fb);
        ulonglong2 *res = radix_sort_rec16(cur, oth, n, 0, 0, fbits[f], s);
        if (res != cur) { oth = cur; cur = res; }
    }
    DBuf<dn_las_record> orec(n); DBuf<uint16_t> otr((size_t)ntrace + 1);
    DN_LAUNCH(k_merge_records, (unsigned)((n + 255) / 256), 256, 0, s, d_rec, (const ulonglong2 *)cur, n, orec.p, tl.p);
    exclusive_scan_u32_to_i64(tl.p, dtoff.p, n, tot.p, s);
    DN_LAUNCH(k_merge_traces, (unsigned)((n * 32 + 255) / 256), 256, 0, s, (const ulonglong2 *)cur, n, (const int64_t *)stoff.p,
              (const int64_t *)dtoff.p, (const dn_las_record *)orec.p, d_trace, otr.p);
    DN_CUDA(cudaMemcpyAsync(out.rec, orec.p, sizeof(dn_las_record) * n, cudaMemcpyDeviceToHost, s));
    DN_CUDA(cudaMemcpyAsync(out.toff, dtoff.p, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, s));
    if (ntrace) DN_CUDA(cudaMemcpyAsync(out.trace, otr.p, sizeof(uint16_t) * ntrace, cudaMemcpyDeviceToHost, s));
    DN_CUDA(cudaStreamSynchronize(s));
}

}  // namespace dn
