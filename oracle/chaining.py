"""chainLocalAlignments -- ORACLE (test infrastructure only): a restatement of
source/dentist/common/alignments/chaining.d:122-475 with util/graphalgo.d (connectedComponents :43-160,
topologicalSort :1011-1052 incl. the "live" NaturalNumberSet iteration util/math.d:1992-2047,
dagSingleSourceShortestPaths :926-960) and the LAS flag writing of dazzler.d:2037-2083.

The source IS in /root/reference, so this is exact -- with one documented exception: the reference sorts
candidate end nodes and accepted chains with Phobos' UNSTABLE `sort`; on exact ties (equal distance / equal
chain coordinates) its order is unspecified by the source.  Here ties keep index order.  The reference has no
unit test for chaining (0 in chaining.d), so there is no golden vector to pin beyond the D source itself.
"""
import numpy as np

COMP, START, NEXT, BEST, ELIM = 0x1, 0x4, 0x8, 0x10, 0x20
INF = 2 ** 31 - 1


class ChainingOptions:
    """commandline.d:2820-2830; defaults :1982 (maxIndel), :1819 (maxChainGap), :2014, :2153, :2158-2174."""

    def __init__(self, max_indel=1000, max_chain_gap=10000, max_rel_overlap=0.3, min_rel_score=1.0, min_score=126):
        self.max_indel, self.max_chain_gap = max_indel, max_chain_gap
        self.max_rel_overlap, self.min_rel_score, self.min_score = max_rel_overlap, min_rel_score, min_score

    def effective_min_score(self, best):                       # chaining.d:111-117
        return int(max(float(self.min_score), self.min_rel_score * best))


def _gap(x, y, s):                                             # chaining.d:367-370
    return int(y[s + "bpos"]) - int(x[s + "epos"])


def _length(x, s):
    return int(x[s + "epos"]) - int(x[s + "bpos"])


def are_chainable(x, y, o):                                    # chaining.d:434-451
    if (int(x["flags"]) ^ int(y["flags"])) & COMP:
        return False
    ga, gb = _gap(x, y, "a"), _gap(x, y, "b")
    return (x["abpos"] < y["abpos"] and x["bbpos"] < y["bbpos"] and
            abs(ga - gb) <= o.max_indel and
            max(abs(ga), abs(gb)) <= o.max_chain_gap and
            max(0, -ga) <= o.max_rel_overlap * min(_length(x, "a"), _length(y, "a")) and
            max(0, -gb) <= o.max_rel_overlap * min(_length(x, "b"), _length(y, "b")))


def alignment_score(x):                                        # chaining.d:455-461
    return (_length(x, "a") + _length(x, "b")) // 2


def chain_score(x, y):                                         # chaining.d:467-475
    ga, gb = _gap(x, y, "a"), _gap(x, y, "b")
    return abs(ga - gb) + max(abs(ga), abs(gb)) // 10 - alignment_score(y)


def _next_member(s, after):
    """live ElementsRange: the smallest member greater than `after` at the time of the call."""
    c = [e for e in s if e > after]
    return min(c) if c else None


def connected_components(n, und):                              # graphalgo.d:43-160
    unvisited = set(range(n))
    comps = []
    while unvisited:
        comp = set()

        def discover(cur):
            comp.add(cur); unvisited.discard(cur)
            e = _next_member(unvisited, -1)
            while e is not None:
                if und(cur, e):
                    discover(e)
                e = _next_member(unvisited, e)
        discover(min(unvisited))
        comps.append(sorted(comp))
    return comps


def topological_sort(n, has_edge):                             # graphalgo.d:1011-1052
    order = [None] * n
    head = [n]
    unvisited = set(range(n))
    temp = set()

    def visit(node):
        if node not in unvisited:
            return
        assert node not in temp, "cycle"
        temp.add(node)
        e = _next_member(unvisited, -1)
        while e is not None:
            if has_edge(node, e):
                visit(e)
            e = _next_member(unvisited, e)
        temp.discard(node); unvisited.discard(node)
        head[0] -= 1; order[head[0]] = node
    e = _next_member(unvisited, -1)
    while e is not None:
        visit(e)
        e = _next_member(unvisited, e)
    return order


def dag_sssp(n, has_edge, weight, start=0):                    # graphalgo.d:926-960
    order = topological_sort(n, has_edge)
    dist = [INF] * n; pred = [-1] * n
    dist[start] = 0
    for u in range(order.index(start), n):
        for v in range(u + 1, n):
            nu, nv = order[u], order[v]
            if has_edge(nu, nv) and dist[nu] < INF:
                d = dist[nu] + weight(nu, nv)
                if dist[nv] > d:
                    dist[nv] = d; pred[nv] = nu
    return dist, pred


def _reverse_path(pred, end):
    p = [end]
    while pred[p[-1]] >= 0:
        p.append(pred[p[-1]])
    return p


def build_alignment_chains(las, o):                            # chaining.d:152-334; las = records of one (A,B) group
    n = len(las)
    ch = [[are_chainable(las[x], las[y], o) for y in range(n)] for x in range(n)]
    comps = connected_components(n, lambda x, y: ch[x][y] or ch[y][x])
    selected = []                                               # (path as group indices, alternate, score)
    for comp in comps:
        m = len(comp)
        he = lambda x, y: (x == 0 and y > 0) or (x > 0 and y > 0 and ch[comp[x - 1]][comp[y - 1]])
        wt = lambda x, y: -alignment_score(las[comp[y - 1]]) if x == 0 else chain_score(las[comp[x - 1]], las[comp[y - 1]])
        dist, pred = dag_sssp(m + 1, he, wt)
        srt = sorted(range(m + 1), key=lambda e: (dist[e], e))
        max_distance = -o.effective_min_score(-dist[srt[0]])
        forbidden = {0}
        for end in srt:
            if end not in forbidden and dist[end] <= max_distance:
                alt = False
                for pn in _reverse_path(pred, end):
                    if pn > 0:
                        if pn in forbidden:
                            alt = True
                        forbidden.add(pn)
                path = _reverse_path(pred, end)[::-1][1:]
                selected.append(([comp[p - 1] for p in path], alt, -dist[end]))
    best = max(s for _, _, s in selected)                       # maxIndex: first maximum
    min_score = o.effective_min_score(best)
    accepted = [c for c in selected if min_score <= c[2]]
    key = lambda c: (int(las[c[0][0]]["abpos"]), int(las[c[0][0]]["bbpos"]), int(las[c[0][-1]]["aepos"]), int(las[c[0][-1]]["bepos"]))
    accepted.sort(key=key)                                      # AlignmentChain.opCmp, base.d:766-777 (ids equal inside a group)
    return accepted


def chain_local_alignments(rec, o=None):
    """rec: LAS records (numpy structured, LAsort order).  Returns (source index per output record, flags)."""
    o = o or ChainingOptions()
    idx = [i for i in range(len(rec)) if not (int(rec[i]["flags"]) & ELIM)]
    out_src, out_flags = [], []
    g0 = 0
    while g0 < len(idx):
        g1 = g0
        while g1 < len(idx) and rec[idx[g1]]["aread"] == rec[idx[g0]]["aread"] and rec[idx[g1]]["bread"] == rec[idx[g0]]["bread"]:
            g1 += 1
        grp = [rec[i] for i in idx[g0:g1]]
        for path, alt, _ in build_alignment_chains(grp, o):
            first = grp[path[0]]
            base = int(first["flags"]) & (COMP | ELIM)          # writeAlignmentChain, dazzler.d:2037-2083
            for t, p in enumerate(path):
                f = base | ((START | (0 if alt else BEST)) if t == 0 else NEXT)
                out_src.append(idx[g0 + p]); out_flags.append(f)
        g0 = g1
    return np.array(out_src, np.int64), np.array(out_flags, np.uint32)
