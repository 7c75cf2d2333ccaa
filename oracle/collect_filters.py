"""collectPileUps alignment filters -- ORACLE (test infrastructure only): a restatement of
source/dentist/commands/collectPileUps/filter.d:122-356 applied in the order of
collectPileUps/package.d:129-141 (LQ, Improper, WeaklyAnchored, Contained, Ambiguous, Redundant) on the
AlignmentChains that AlignmentChainPacker (dazzler.d:708-743) builds from a chained LAS.
Chain predicates: base.d:527-560 (isProper/beginsWith/endsWith), :563-603 (isFullyContained),
:661-700 (coveredBases/totalDiffs/averageErrorRate); regions common/package.d:228-289, util/region.d:474-499.
Exact w.r.t. the D source (filter.d has no unit tests; base.d's predicate KATs are replayed in
tests/test_collect_filters.py).  The reference runs ContainedAlignmentChainsFilter with `parallel`, which makes
its own result order-dependent in principle; the sequential order is taken here (containment is transitive).

Status per chain: 0 kept, 1 LQ, 2 improper, 3 weakly anchored, 4 contained, 5 ambiguous read, 6 redundant read.
"""
import numpy as np

COMP, START, NEXT, BEST, ELIM = 0x1, 0x4, 0x8, 0x10, 0x20


def chain_summaries(rec, alen, blen, mask):
    """mask: dict contig -> sorted disjoint [(b, e)] (the repeat mask as a ReferenceRegion)."""
    out = []
    i, n = 0, len(rec)
    while i < n:
        j = i + 1
        if int(rec[i]["flags"]) & (START | BEST):
            while j < n and (int(rec[j]["flags"]) & NEXT):
                j += 1
        las = [rec[x] for x in range(i, j)]
        a, b = int(las[0]["aread"]), int(las[0]["bread"])
        iv = sorted((int(l["abpos"]), int(l["aepos"])) for l in las)        # Region normalises: union of the LA intervals
        merged = []
        for s, e in iv:
            if merged and s <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], e)
            else:
                merged.append([s, e])
        uniq = 0
        for s, e in merged:
            cov = 0
            for ms, me in mask.get(a, []):
                cov += max(0, min(e, me) - max(s, ms))
            uniq += (e - s) - cov
        out.append(dict(first=i, a=a, b=b, comp=int(las[0]["flags"]) & COMP, fab=int(las[0]["abpos"]), fbb=int(las[0]["bbpos"]),
                        lae=int(las[-1]["aepos"]), lbe=int(las[-1]["bepos"]), alen=int(alen[a]), blen=int(blen[b]),
                        cov_a=sum(int(l["aepos"]) - int(l["abpos"]) for l in las), diffs=sum(int(l["diffs"]) for l in las),
                        uniq=uniq, disabled=bool(int(las[0]["flags"]) & ELIM)))
        i = j
    return out


def _b_interval(c):                                             # toInterval!(.., "contigB"), common/package.d:259-289
    return (c["blen"] - c["lbe"], c["blen"] - c["fbb"]) if c["comp"] else (c["fbb"], c["lbe"])


def collect_filter(rec, alen, blen, mask, max_err, allowance, min_anchor):
    ch = chain_summaries(rec, alen, blen, mask)
    st = [7 if c["disabled"] else 0 for c in ch]               # 7 = disabled on input
    for i, c in enumerate(ch):
        if st[i]:
            continue
        if c["diffs"] / c["cov_a"] > max_err:                                                    # filter.d:122-137
            st[i] = 1
        elif not ((c["fab"] <= allowance or c["fbb"] <= allowance) and
                  (c["lae"] + allowance >= c["alen"] or c["lbe"] + allowance >= c["blen"])):      # filter.d:141-160, base.d:527-531
            st[i] = 2
        elif c["uniq"] <= min_anchor:                                                            # filter.d:327-356
            st[i] = 3
    # ContainedAlignmentChainsFilter  filter.d:181-209 (stable sort by AlignmentChain.opCmp, base.d:766-777)
    order = sorted(range(len(ch)), key=lambda i: (ch[i]["a"], ch[i]["b"], ch[i]["fab"], ch[i]["fbb"], ch[i]["lae"], ch[i]["lbe"]))
    for p, i in enumerate(order):
        if st[i]:
            continue
        c1 = ch[i]; b1 = _b_interval(c1)
        for j in order[p + 1:]:
            c2 = ch[j]
            if not (c2["a"] == c1["a"] and c1["fab"] <= c2["fab"] and c2["lae"] <= c1["lae"]):
                break
            b2 = _b_interval(c2)
            if c2["comp"] == c1["comp"] and c2["b"] == c1["b"] and b1[0] <= b2[0] and b2[1] <= b1[1] and not st[j]:
                st[j] = 4
    used = set()
    by_read = {}
    for i, c in enumerate(ch):
        by_read.setdefault(c["b"], []).append(i)
    for b, idxs in by_read.items():                            # AmbiguousAlignmentChainsFilter  filter.d:232-318
        live = [i for i in idxs if not st[i]]
        amb = any(max(_b_interval(ch[x])[0], _b_interval(ch[y])[0]) < min(_b_interval(ch[x])[1], _b_interval(ch[y])[1])
                  for k, x in enumerate(live) for y in live[k + 1:])
        if amb:
            used.add(b)
            for i in live:
                st[i] = 5
    for b, idxs in by_read.items():                            # RedundantAlignmentChainsFilter  filter.d:166-177, base.d:563-603
        live = [i for i in idxs if not st[i]]
        if any(ch[i]["fbb"] <= ch[i]["fab"] and ch[i]["lae"] + ch[i]["blen"] - ch[i]["lbe"] < ch[i]["alen"] for i in live):
            used.add(b)
            for i in live:
                st[i] = 6
    return (np.array([c["first"] for c in ch], np.int64), np.array(st, np.uint8), sorted(used))
