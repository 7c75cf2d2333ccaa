"""Reader for `LAdump -cdtl` text (the format DENTIST's own LAS fixtures are written in, dazzler.d:965-1026) and a
record-level comparison of two alignment sets -- the `--compare-ladump` mode of SURVEY §8c: wherever the real
daligner / damapper binaries exist, dump their LAS with `LAdump -cdtl` and measure how far this engine's
(parity-unpinned) alignments are from theirs.  Product-side code: no oracle imports."""
import numpy as np

from ._lib import REC_DTYPE

COMP, START, NEXT, BEST = 0x1, 0x4, 0x8, 0x10


def parse(lines):
    """-> (tspace, records (REC_DTYPE, 0-based ids), [trace (n, 2) per record]).
    Line types: `X tspace`, `P a b n|c chain` (chain: > best start, + alternate start, - continuation, . none),
    `L alen blen`, `C ab ae bb be`, `D diffs`, `T n` + n lines `diffs bases`; +/%/@ header lines are sizes only."""
    tspace, rows, traces, cur = 100, [], [], None
    it = iter(lines)
    for ln in it:
        p = ln.split()
        if not p or p[0] in "+%@":
            continue
        if p[0] == "X":
            tspace = int(p[1])
        elif p[0] == "P":
            if cur is not None:
                rows.append(cur); traces.append(np.zeros((0, 2), np.uint16))
            fl = (COMP if p[3] == "c" else 0) | ({">": START | BEST, "+": START, "-": NEXT, ".": 0}[p[4]] if len(p) > 4 else 0)
            cur = dict(aread=int(p[1]) - 1, bread=int(p[2]) - 1, flags=fl, diffs=-1)
        elif p[0] == "C":
            cur.update(abpos=int(p[1]), aepos=int(p[2]), bbpos=int(p[3]), bepos=int(p[4]))
        elif p[0] == "D":
            cur["diffs"] = int(p[1])
        elif p[0] == "T":
            t = np.array([[int(x) for x in next(it).split()] for _ in range(int(p[1]))], np.uint16).reshape(-1, 2)
            if cur["diffs"] < 0:
                cur["diffs"] = int(t[:, 0].sum())
            rows.append(cur); traces.append(t); cur = None
    if cur is not None:
        rows.append(cur); traces.append(np.zeros((0, 2), np.uint16))
    rec = np.zeros(len(rows), REC_DTYPE)
    for i, r in enumerate(rows):
        for k in ("aread", "bread", "flags", "abpos", "aepos", "bbpos", "bepos"):
            rec[i][k] = r[k]
        rec[i]["diffs"] = max(r["diffs"], 0); rec[i]["tlen"] = 2 * len(traces[i])
    return tspace, rec, traces


def compare(rec_x, traces_x, rec_y, traces_y, min_overlap=0.5):
    """How much of alignment set X (e.g. the real tool's) is found in Y (ours)?  An X record is `matched` when a Y
    record of the same (aread, bread, strand) overlaps it on A by >= min_overlap of the longer of the two.
    Returns counts: x, y, matched, same_coords (all four coordinates equal), same_trace, a_bases_x / a_bases_y /
    a_bases_shared (A bases covered per read pair and strand, union), y_only (Y records no X record matches)."""
    def index(rec):
        d = {}
        for i in range(len(rec)):
            d.setdefault((int(rec[i]["aread"]), int(rec[i]["bread"]), int(rec[i]["flags"]) & COMP), []).append(i)
        return d

    def covered(rec, idx):
        iv = sorted((int(rec[i]["abpos"]), int(rec[i]["aepos"])) for i in idx)
        out = []
        for b, e in iv:
            if out and b <= out[-1][1]:
                out[-1][1] = max(out[-1][1], e)
            else:
                out.append([b, e])
        return out

    ix, iy = index(rec_x), index(rec_y)
    res = dict(x=len(rec_x), y=len(rec_y), matched=0, same_coords=0, same_trace=0, a_bases_x=0, a_bases_y=0, a_bases_shared=0, y_only=0)
    hit_y = set()
    for key, xs in ix.items():
        ys = iy.get(key, [])
        for i in xs:
            best, bj = 0.0, -1
            for j in ys:
                ov = min(int(rec_x[i]["aepos"]), int(rec_y[j]["aepos"])) - max(int(rec_x[i]["abpos"]), int(rec_y[j]["abpos"]))
                span = max(int(rec_x[i]["aepos"]) - int(rec_x[i]["abpos"]), int(rec_y[j]["aepos"]) - int(rec_y[j]["abpos"]), 1)
                if ov / span > best:
                    best, bj = ov / span, j
            if best >= min_overlap:
                res["matched"] += 1; hit_y.add(bj)
                if all(int(rec_x[i][f]) == int(rec_y[bj][f]) for f in ("abpos", "aepos", "bbpos", "bepos")):
                    res["same_coords"] += 1
                    if np.array_equal(np.asarray(traces_x[i]).reshape(-1, 2), np.asarray(traces_y[bj]).reshape(-1, 2)):
                        res["same_trace"] += 1
        cx, cy = covered(rec_x, xs), covered(rec_y, ys)
        res["a_bases_x"] += sum(e - b for b, e in cx)
        for b, e in cx:
            for b2, e2 in cy:
                res["a_bases_shared"] += max(0, min(e, e2) - max(b, b2))
    for key, ys in iy.items():
        res["a_bases_y"] += sum(e - b for b, e in covered(rec_y, ys))
    res["y_only"] = len(rec_y) - len(hit_y)
    return res
