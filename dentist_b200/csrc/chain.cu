// chain.cu -- chainLocalAlignments on the device: one thread per (contigA, contigB) group of local
// alignments restates source/dentist/common/alignments/chaining.d:152-334 (components, DAG single-source
// shortest paths over a DFS topological order with the "live" set iteration of util/math.d:1992-2047,
// chain selection, alternate chains, minimum score) and the flag writing of dazzler.d:2037-2083.
// Groups are tiny (1-20 records; at most 63 supported): integer work, sets are 64-bit masks.
// Specification = oracle/chaining.py (exact w.r.t. the D source except for unspecified tie orders).
#include "api_internal.hpp"
#include "pile.cuh"
#include <string.h>
#include <algorithm>
#include <vector>

namespace dn {
namespace {

struct ChainOpts { int max_indel, max_chain_gap, min_score; double max_rel_overlap, min_rel_score; };

__device__ __forceinline__ int effective_min_score(const ChainOpts &o, int best) {
    const double a = (double)o.min_score, b = o.min_rel_score * (double)best;
    return (int)(a > b ? a : b);
}
__device__ __forceinline__ bool chainable(const dn_las_record &x, const dn_las_record &y, const ChainOpts &o) {
    if ((x.flags ^ y.flags) & DN_LAS_COMP) return false;
    const int ga = y.abpos - x.aepos, gb = y.bbpos - x.bepos;
    const int indel = ga > gb ? ga - gb : gb - ga;
    const int mg = max(ga < 0 ? -ga : ga, gb < 0 ? -gb : gb);
    const int la = min(x.aepos - x.abpos, y.aepos - y.abpos), lb = min(x.bepos - x.bbpos, y.bepos - y.bbpos);
    return x.abpos < y.abpos && x.bbpos < y.bbpos && indel <= o.max_indel && mg <= o.max_chain_gap &&
           (double)max(0, -ga) <= o.max_rel_overlap * (double)la && (double)max(0, -gb) <= o.max_rel_overlap * (double)lb;
}
__device__ __forceinline__ int ascore(const dn_las_record &x) { return ((x.aepos - x.abpos) + (x.bepos - x.bbpos)) / 2; }
__device__ __forceinline__ int cscore(const dn_las_record &x, const dn_las_record &y) {
    const int ga = y.abpos - x.aepos, gb = y.bbpos - x.bepos;
    const int indel = ga > gb ? ga - gb : gb - ga;
    const int mg = max(ga < 0 ? -ga : ga, gb < 0 ? -gb : gb);
    return indel + mg / 10 - ascore(y);
}
__device__ __forceinline__ int next_member(u64 set, int after) {       // smallest member > after, or -1
    const u64 m = after >= 63 ? 0ull : (after < 0 ? set : set & ~((2ull << after) - 1ull));
    return m ? __ffsll((long long)m) - 1 : -1;
}

constexpr int CH_MAX = 63;

// gstart[g] .. gstart[g+1]: records of group g (already without ELIM records).  Output: for group g up to
// cap_per = 4 * n records at out_src/out_flags[4 * gstart[g] ...], count in out_cnt[g]; status[g] != 0 on error.
__global__ void __launch_bounds__(64) k_chain_groups(const dn_las_record *__restrict__ rec, const int32_t *__restrict__ gstart, int ngroups,
                                                     ChainOpts o, int32_t *__restrict__ out_src, u32 *__restrict__ out_flags,
                                                     int32_t *__restrict__ out_cnt, int32_t *__restrict__ status) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    const int s = gstart[g], n = gstart[g + 1] - s;
    out_cnt[g] = 0; status[g] = 0;
    if (n > CH_MAX) { status[g] = 1; return; }
    const dn_las_record *la = rec + s;
    u64 adj[CH_MAX], und[CH_MAX];
    for (int x = 0; x < n; x++) adj[x] = 0;
    for (int x = 0; x < n; x++) for (int y = 0; y < n; y++) if (x != y && chainable(la[x], la[y], o)) adj[x] |= 1ull << y;
    for (int x = 0; x < n; x++) { u64 u = adj[x]; for (int y = 0; y < n; y++) if (adj[y] >> x & 1ull) u |= 1ull << y; und[x] = u; }

    // selected chains over all components: end node, component mask, score, alternate flag; paths re-derived on emit
    int nsel = 0;
    short sel_path[4 * CH_MAX]; int sel_off[CH_MAX + 1], sel_score[CH_MAX]; unsigned char sel_alt[CH_MAX];
    int path_used = 0;
    sel_off[0] = 0;

    u64 unvisited = n >= 64 ? ~0ull : ((1ull << n) - 1ull);
    while (unvisited) {
        // connected component of the smallest unvisited node (graphalgo.d:43-160)
        u64 comp = 1ull << (__ffsll((long long)unvisited) - 1), frontier = comp;
        while (frontier) {
            u64 nxt = 0;
            for (u64 f = frontier; f; f &= f - 1) nxt |= und[__ffsll((long long)f) - 1];
            nxt &= unvisited & ~comp; comp |= nxt; frontier = nxt;
        }
        unvisited &= ~comp;
        int c[CH_MAX], m = 0;
        for (u64 f = comp; f; f &= f - 1) c[m++] = __ffsll((long long)f) - 1;
        const int M = m + 1;                                    // local nodes 0..m, 0 = source
        u64 ladj[CH_MAX + 1];
        ladj[0] = (M >= 64 ? ~0ull : ((1ull << M) - 1ull)) & ~1ull;
        for (int x = 1; x < M; x++) { u64 a = 0; for (int y = 1; y < M; y++) if (adj[c[x - 1]] >> c[y - 1] & 1ull) a |= 1ull << y; ladj[x] = a; }
        // DFS topological sort with live ascending iteration (graphalgo.d:1011-1052)
        int order[CH_MAX + 1], head = M;
        {
            u64 U = M >= 64 ? ~0ull : ((1ull << M) - 1ull);
            signed char st_node[CH_MAX + 1], st_cur[CH_MAX + 1]; int sp = 0;
            int oc = next_member(U, -1);
            while (oc >= 0) {
                st_node[0] = (signed char)oc; st_cur[0] = -1; sp = 1;
                while (sp > 0) {
                    const int v = st_node[sp - 1];
                    const int nx = next_member(U, st_cur[sp - 1]);
                    if (nx >= 0) {
                        st_cur[sp - 1] = (signed char)nx;
                        bool onstack = false;
                        for (int q = 0; q < sp; q++) onstack |= st_node[q] == nx;
                        if ((ladj[v] >> nx & 1ull) && !onstack) { st_node[sp] = (signed char)nx; st_cur[sp] = -1; sp++; }
                    } else { sp--; U &= ~(1ull << v); order[--head] = v; }
                }
                oc = next_member(U, oc);
            }
        }
        // shortest paths from the source in topological order (graphalgo.d:926-960)
        int dist[CH_MAX + 1], pred[CH_MAX + 1];
        for (int x = 0; x < M; x++) { dist[x] = 0x7fffffff; pred[x] = -1; }
        dist[0] = 0;
        int u0 = 0; while (order[u0] != 0) u0++;
        for (int u = u0; u < M; u++) for (int v = u + 1; v < M; v++) {
            const int nu = order[u], nv = order[v];
            if (!(ladj[nu] >> nv & 1ull) || dist[nu] == 0x7fffffff) continue;
            const int w = nu == 0 ? -ascore(la[c[nv - 1]]) : cscore(la[c[nu - 1]], la[c[nv - 1]]);
            if (dist[nv] > dist[nu] + w) { dist[nv] = dist[nu] + w; pred[nv] = nu; }
        }
        // end nodes from best to worst (ties: node index)
        int srt[CH_MAX + 1];
        for (int x = 0; x < M; x++) srt[x] = x;
        for (int x = 1; x < M; x++) { int t = srt[x], y = x; while (y > 0 && (dist[srt[y - 1]] > dist[t])) { srt[y] = srt[y - 1]; y--; } srt[y] = t; }
        const int max_distance = -effective_min_score(o, -dist[srt[0]]);
        u64 forbidden = 1ull;
        for (int q = 0; q < M; q++) {
            const int end = srt[q];
            if ((forbidden >> end & 1ull) || dist[end] > max_distance) continue;
            bool alt = false; int len = 0;
            for (int p = end; p >= 0; p = pred[p]) if (p > 0) { if (forbidden >> p & 1ull) alt = true; forbidden |= 1ull << p; len++; }
            if (nsel >= CH_MAX || path_used + len > 4 * CH_MAX) { status[g] = 2; return; }
            int w = path_used + len;                             // path stored front to back
            for (int p = end; p > 0; p = pred[p]) sel_path[--w] = (short)c[p - 1];
            path_used += len; sel_score[nsel] = -dist[end]; sel_alt[nsel] = alt; nsel++; sel_off[nsel] = path_used;
        }
    }
    // minimum score relative to the best chain (first maximum), then AlignmentChain.opCmp order
    int best = sel_score[0];
    for (int q = 1; q < nsel; q++) if (sel_score[q] > best) best = sel_score[q];
    const int min_score = effective_min_score(o, best);
    int acc[CH_MAX], nacc = 0;
    for (int q = 0; q < nsel; q++) if (min_score <= sel_score[q]) acc[nacc++] = q;
    auto less = [&](int p, int q) {
        const dn_las_record &pf = la[sel_path[sel_off[p]]], &pl = la[sel_path[sel_off[p + 1] - 1]];
        const dn_las_record &qf = la[sel_path[sel_off[q]]], &ql = la[sel_path[sel_off[q + 1] - 1]];
        if (pf.abpos != qf.abpos) return pf.abpos < qf.abpos;
        if (pf.bbpos != qf.bbpos) return pf.bbpos < qf.bbpos;
        if (pl.aepos != ql.aepos) return pl.aepos < ql.aepos;
        return pl.bepos < ql.bepos;
    };
    for (int x = 1; x < nacc; x++) { int t = acc[x], y = x; while (y > 0 && less(t, acc[y - 1])) { acc[y] = acc[y - 1]; y--; } acc[y] = t; }
    int cnt = 0;
    for (int x = 0; x < nacc; x++) {
        const int q = acc[x];
        // alternate chains repeat shared prefixes (chaining.d:252-275): the group's 4 * n output slots can run out
        if (cnt + (sel_off[q + 1] - sel_off[q]) > 4 * n) { out_cnt[g] = 0; status[g] = 2; return; }
        const u32 base = la[sel_path[sel_off[q]]].flags & (DN_LAS_COMP | DN_LAS_ELIM);
        for (int t = sel_off[q]; t < sel_off[q + 1]; t++) {
            out_src[4 * s + cnt] = s + sel_path[t];
            out_flags[4 * s + cnt] = base | (t == sel_off[q] ? (DN_LAS_START | (sel_alt[q] ? 0u : DN_LAS_BEST)) : DN_LAS_NEXT);
            cnt++;
        }
    }
    out_cnt[g] = cnt;
}

// The same algorithm on the host, without size limits, for the rare groups the kernel declines (more than 63 local
// alignments between one pair of reads, or more output than the group's 4 * n slots): chaining.d:152-334 restated with
// vectors instead of 64-bit set masks.  DENTIST runs this logic on the host too; the reference has no limit.
struct HostChainOpts { int max_indel, max_chain_gap, min_score; double max_rel_overlap, min_rel_score; };

bool h_chainable(const dn_las_record &x, const dn_las_record &y, const HostChainOpts &o) {
    if ((x.flags ^ y.flags) & DN_LAS_COMP) return false;
    const int ga = y.abpos - x.aepos, gb = y.bbpos - x.bepos;
    const int indel = ga > gb ? ga - gb : gb - ga;
    const int mg = std::max(ga < 0 ? -ga : ga, gb < 0 ? -gb : gb);
    const int la = std::min(x.aepos - x.abpos, y.aepos - y.abpos), lb = std::min(x.bepos - x.bbpos, y.bepos - y.bbpos);
    return x.abpos < y.abpos && x.bbpos < y.bbpos && indel <= o.max_indel && mg <= o.max_chain_gap &&
           (double)std::max(0, -ga) <= o.max_rel_overlap * (double)la && (double)std::max(0, -gb) <= o.max_rel_overlap * (double)lb;
}
int h_ascore(const dn_las_record &x) { return ((x.aepos - x.abpos) + (x.bepos - x.bbpos)) / 2; }
int h_cscore(const dn_las_record &x, const dn_las_record &y) {
    const int ga = y.abpos - x.aepos, gb = y.bbpos - x.bepos;
    const int indel = ga > gb ? ga - gb : gb - ga;
    const int mg = std::max(ga < 0 ? -ga : ga, gb < 0 ? -gb : gb);
    return indel + mg / 10 - h_ascore(y);
}
int h_min_score(const HostChainOpts &o, int best) {
    const double a = (double)o.min_score, b = o.min_rel_score * (double)best;
    return (int)(a > b ? a : b);
}

void chain_group_host(const dn_las_record *la, int n, const HostChainOpts &o, std::vector<int32_t> &out_src, std::vector<u32> &out_flags) {
    std::vector<std::vector<char>> ch(n, std::vector<char>(n, 0));
    for (int x = 0; x < n; x++) for (int y = 0; y < n; y++) ch[x][y] = x != y && h_chainable(la[x], la[y], o);
    struct Sel { std::vector<int> path; bool alt; int score; };
    std::vector<Sel> selected;
    std::vector<char> visited(n, 0);
    for (int root = 0; root < n; root++) {
        if (visited[root]) continue;
        std::vector<int> comp{root}; visited[root] = 1;                 // connected component of the smallest unvisited node
        for (size_t q = 0; q < comp.size(); q++)
            for (int y = 0; y < n; y++) if (!visited[y] && (ch[comp[q]][y] || ch[y][comp[q]])) { visited[y] = 1; comp.push_back(y); }
        std::sort(comp.begin(), comp.end());
        const int M = (int)comp.size() + 1;                               // local nodes 0..m, 0 = source
        auto edge = [&](int x, int y) { return y > 0 && (x == 0 || ch[comp[x - 1]][comp[y - 1]]); };
        // DFS topological sort with the live ascending iteration over the unvisited set (graphalgo.d:1011-1052)
        std::vector<int> order(M); int head = M;
        {
            std::vector<char> U(M, 1), onstack(M, 0);
            auto next_member = [&](int after) { for (int e = after + 1; e < M; e++) if (U[e]) return e; return -1; };
            std::vector<int> st_node, st_cur;
            for (int oc = next_member(-1); oc >= 0; oc = next_member(oc)) {
                st_node.assign(1, oc); st_cur.assign(1, -1); onstack[oc] = 1;
                while (!st_node.empty()) {
                    const int v = st_node.back();
                    const int nx = next_member(st_cur.back());
                    if (nx >= 0) {
                        st_cur.back() = nx;
                        if (edge(v, nx) && !onstack[nx]) { st_node.push_back(nx); st_cur.push_back(-1); onstack[nx] = 1; }
                    } else { st_node.pop_back(); st_cur.pop_back(); onstack[v] = 0; U[v] = 0; order[--head] = v; }
                }
            }
        }
        std::vector<int> dist(M, 0x7fffffff), pred(M, -1);
        dist[0] = 0;
        int u0 = 0; while (order[u0] != 0) u0++;
        for (int u = u0; u < M; u++) for (int v = u + 1; v < M; v++) {
            const int nu = order[u], nv = order[v];
            if (!edge(nu, nv) || dist[nu] == 0x7fffffff) continue;
            const int w = nu == 0 ? -h_ascore(la[comp[nv - 1]]) : h_cscore(la[comp[nu - 1]], la[comp[nv - 1]]);
            if (dist[nv] > dist[nu] + w) { dist[nv] = dist[nu] + w; pred[nv] = nu; }
        }
        std::vector<int> srt(M);
        for (int x = 0; x < M; x++) srt[x] = x;
        std::stable_sort(srt.begin(), srt.end(), [&](int a, int b) { return dist[a] < dist[b]; });
        const int max_distance = -h_min_score(o, -dist[srt[0]]);
        std::vector<char> forbidden(M, 0); forbidden[0] = 1;
        for (int q = 0; q < M; q++) {
            const int end = srt[q];
            if (forbidden[end] || dist[end] > max_distance) continue;
            Sel sl; sl.alt = false; sl.score = -dist[end];
            for (int p = end; p >= 0; p = pred[p]) if (p > 0) { if (forbidden[p]) sl.alt = true; forbidden[p] = 1; sl.path.push_back(comp[p - 1]); }
            std::reverse(sl.path.begin(), sl.path.end());
            selected.push_back(std::move(sl));
        }
    }
    if (selected.empty()) return;
    int best = selected[0].score;
    for (const Sel &c : selected) if (c.score > best) best = c.score;
    const int min_score = h_min_score(o, best);
    std::vector<const Sel *> acc;
    for (const Sel &c : selected) if (min_score <= c.score) acc.push_back(&c);
    std::stable_sort(acc.begin(), acc.end(), [&](const Sel *p, const Sel *q) {
        const dn_las_record &pf = la[p->path.front()], &pl = la[p->path.back()], &qf = la[q->path.front()], &ql = la[q->path.back()];
        if (pf.abpos != qf.abpos) return pf.abpos < qf.abpos;
        if (pf.bbpos != qf.bbpos) return pf.bbpos < qf.bbpos;
        if (pl.aepos != ql.aepos) return pl.aepos < ql.aepos;
        return pl.bepos < ql.bepos;
    });
    for (const Sel *c : acc) {
        const u32 base = la[c->path.front()].flags & (DN_LAS_COMP | DN_LAS_ELIM);
        for (size_t t = 0; t < c->path.size(); t++) {
            out_src.push_back(c->path[t]);
            out_flags.push_back(base | (t == 0 ? (DN_LAS_START | (c->alt ? 0u : DN_LAS_BEST)) : DN_LAS_NEXT));
        }
    }
}

}  // namespace
}  // namespace dn

using namespace dn;
using namespace dnapi;

extern "C" int dn_las_chain(dn_las_buf *las, int32_t max_indel, int32_t max_chain_gap, double max_rel_overlap, double min_rel_score,
                            int32_t min_score) {
    if (!las) return fail(DN_ERR_INVALID, "null argument");
    // host glue: drop disabled records, check the order (chaining.d:131-140), find the (A,B) groups
    std::vector<int64_t> idx; std::vector<int32_t> gstart;
    for (int64_t i = 0; i < las->nrec; i++) {
        if (las->rec[i].flags & DN_LAS_ELIM) continue;
        if (!idx.empty()) {
            const dn_las_record &p = las->rec[idx.back()], &q = las->rec[i];
            if (p.aread > q.aread || (p.aread == q.aread && p.bread > q.bread)) return fail(DN_ERR_INVALID, "local alignments are not ordered properly");
            if (p.aread != q.aread || p.bread != q.bread) gstart.push_back((int32_t)idx.size());
        } else gstart.push_back(0);
        idx.push_back(i);
    }
    const int64_t n = (int64_t)idx.size();
    gstart.push_back((int32_t)n);
    const int ngroups = (int)gstart.size() - 1;
    if (n == 0) { las->nrec = 0; return DN_OK; }
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = ensure_device()) return rc;
    return guarded([&]() -> int {
        cudaSetDevice(g_device); arena().reset();
        cudaStream_t s = g_stream;
        // page-locked staging for both directions (hcache blocks >= 256 KB are pinned and recycled)
        dn_las_record *in = (dn_las_record *)hcache_alloc(sizeof(dn_las_record) * (size_t)(n + 1));
        for (int64_t i = 0; i < n; i++) in[i] = las->rec[idx[i]];
        DBuf<dn_las_record> drec(n); DBuf<int32_t> dg(ngroups + 1), osrc(4 * n), ocnt(ngroups), ost(ngroups); DBuf<u32> ofl(4 * n);
        DN_CUDA(cudaMemcpyAsync(drec.p, in, sizeof(dn_las_record) * n, cudaMemcpyHostToDevice, s));
        DN_CUDA(cudaMemcpyAsync(dg.p, gstart.data(), sizeof(int32_t) * (ngroups + 1), cudaMemcpyHostToDevice, s));
        ChainOpts o{max_indel, max_chain_gap, min_score, max_rel_overlap, min_rel_score};
        DN_LAUNCH(k_chain_groups, (ngroups + 63) / 64, 64, 0, s, (const dn_las_record *)drec.p, (const int32_t *)dg.p, ngroups, o,
                  osrc.p, ofl.p, ocnt.p, ost.p);
        int32_t *hsrc = (int32_t *)hcache_alloc(sizeof(int32_t) * (size_t)(4 * n + 1));
        u32 *hfl = (u32 *)hcache_alloc(sizeof(u32) * (size_t)(4 * n + 1));
        int32_t *hcnt = (int32_t *)hcache_alloc(sizeof(int32_t) * (size_t)(2 * ngroups + 1)), *hst = hcnt + ngroups;
        struct Stage { void *a, *b, *c, *d; ~Stage() { hcache_free(a); hcache_free(b); hcache_free(c); hcache_free(d); } } stage{in, hsrc, hfl, hcnt};
        DN_CUDA(cudaMemcpyAsync(hsrc, osrc.p, sizeof(int32_t) * 4 * n, cudaMemcpyDeviceToHost, s));
        DN_CUDA(cudaMemcpyAsync(hfl, ofl.p, sizeof(u32) * 4 * n, cudaMemcpyDeviceToHost, s));
        DN_CUDA(cudaMemcpyAsync(hcnt, ocnt.p, sizeof(int32_t) * ngroups, cudaMemcpyDeviceToHost, s));
        DN_CUDA(cudaMemcpyAsync(hst, ost.p, sizeof(int32_t) * ngroups, cudaMemcpyDeviceToHost, s));
        DN_CUDA(cudaStreamSynchronize(s));
        // groups the kernel declined (> 63 records, or more output than 4 * n slots) go through the host restatement
        std::vector<int32_t> xgroup;                                   // declined groups, ascending
        std::vector<std::vector<int32_t>> xsrc; std::vector<std::vector<u32>> xfl;
        const HostChainOpts ho{max_indel, max_chain_gap, min_score, max_rel_overlap, min_rel_score};
        int64_t total = 0;
        for (int g = 0; g < ngroups; g++) {
            if (hst[g]) {
                xgroup.push_back(g); xsrc.emplace_back(); xfl.emplace_back();
                chain_group_host(in + gstart[g], gstart[g + 1] - gstart[g], ho, xsrc.back(), xfl.back());
                for (auto &v : xsrc.back()) v += gstart[g];
                hcnt[g] = (int32_t)xsrc.back().size();
            }
            total += hcnt[g];
        }
        // gather: records keep their trace (toff), flags are rewritten; a record may appear in two chains
        dn_las_record *nrec = (dn_las_record *)hcache_alloc(sizeof(dn_las_record) * (total + 1));
        int64_t *ntoff = (int64_t *)hcache_alloc(sizeof(int64_t) * (total + 1));
        int64_t w = 0; size_t xi = 0;
        for (int g = 0; g < ngroups; g++) {
            const bool host = hst[g] != 0;
            for (int t = 0; t < hcnt[g]; t++) {
                const int64_t src = idx[host ? xsrc[xi][t] : hsrc[4 * (int64_t)gstart[g] + t]];
                nrec[w] = las->rec[src]; nrec[w].flags = host ? xfl[xi][t] : hfl[4 * (int64_t)gstart[g] + t]; ntoff[w] = las->toff[src]; w++;
            }
            if (host) xi++;
        }
        hcache_free(las->rec); hcache_free(las->toff);
        las->rec = nrec; las->toff = ntoff; las->nrec = total;
        return DN_OK;
    });
}
