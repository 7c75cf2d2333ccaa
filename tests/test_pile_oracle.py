"""CPU checks of the per-pile oracle (filters exact vs D source; consensus pinned by the reference's KAT)."""
import json
import os

import numpy as np

from dentist_b200 import synth
from oracle import oracle

CODE = {"a": 0, "c": 1, "g": 2, "t": 3}


def kat_block():
    k = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "consensus_kat.json")))
    seqs = [np.array([CODE[c] for c in r.lower()], np.uint8) for r in k["reads"]]
    off = np.zeros(len(seqs) + 1, np.int64); off[1:] = np.cumsum([len(s) for s in seqs])
    return k, synth.Block(off, np.concatenate(seqs))


def pile_las(blk, tspace, minlen, allowance):
    la, tr, _ = oracle.align(blk.off, blk.bases, blk.off, blk.bases, tspace=tspace, minlen=minlen, self=1)
    toff = la["toff"].astype(np.int64)
    lens = np.diff(blk.off)
    keep = oracle.filter_pileup(la, lens, lens, allowance)
    return la, toff, tr, keep


def test_reference_consensus_kat():
    # dazzler.d:4257-4299
    k, blk = kat_block()
    assert [len(blk.read(i)) for i in range(3)] == [1050] * 3
    assert (blk.read(0) != blk.read(2)).sum() == 1 and (blk.read(1) != blk.read(2)).sum() == 1
    la, toff, tr, keep = pile_las(blk, 100, k["minlen"], k["allowance"])
    assert keep.all() and len(la) == 6          # every pair, both directions, forward strand
    sel = np.flatnonzero(keep)
    for r in (0, 1, 2):
        cons = oracle.consensus(blk.off, blk.bases, la[sel], toff[sel], tr, 100, r)
        assert np.array_equal(cons, blk.read(k["expected_read"])), r


def test_consensus_corrects_noisy_pile():
    sc = synth.make_scaffolds(1, 12000, 5, n_repeats=0)
    truth_seq = sc[0][1000:9000]
    # 14 reads (either strand) over the same 8 kb window, 10 % error
    seqs = []
    for i in range(14):
        r, _ = synth.simulate_reads([truth_seq.copy()], 1.0, 8000, 1, 0.10, 100 + i, min_len=8000, lognormal=False)
        seqs.append(r)
    reads = []
    for i, r in enumerate(seqs):
        reads.append(r.read(0))
    off = np.zeros(len(reads) + 1, np.int64); off[1:] = np.cumsum([len(x) for x in reads])
    blk = synth.Block(off, np.concatenate(reads))
    la, toff, tr, keep = pile_las(blk, 126, 500, 126)
    sel = np.flatnonzero(keep & oracle.filter_error(la, 0.3))
    assert (la[sel]["aread"] == 0).sum() >= 8
    cons = oracle.consensus(blk.off, blk.bases, la[sel], toff[sel], tr, 126, 0)
    t = truth_seq
    rc = lambda s: (3 - s)[::-1]
    raw = blk.read(0)
    # compare identities with a cheap proxy instead: k-mer containment against truth
    def kmers(s, k=12):
        v = np.zeros(len(s) - k + 1, np.int64)
        for i in range(k):
            v = v * 4 + s[i:len(s) - k + 1 + i]
        return set(v.tolist())
    kt = kmers(t) | kmers(rc(t))
    f_raw = np.mean([x in kt for x in kmers(raw)])
    f_cons = np.mean([x in kt for x in kmers(cons)])
    assert f_raw < 0.45 and f_cons > 0.80, (f_raw, f_cons)


def test_filters_match_definitions():
    rng = np.random.default_rng(3)
    n = 500
    rec = np.zeros(n, oracle.LAS40)
    rec["aread"] = rng.integers(0, 6, n); rec["bread"] = rng.integers(0, 6, n)
    lens = rng.integers(2000, 9000, 6).astype(np.int32)
    rec["abpos"] = rng.integers(0, 400, n); rec["bbpos"] = rng.integers(0, 400, n)
    rec["aepos"] = lens[rec["aread"]] - rng.integers(0, 400, n); rec["bepos"] = lens[rec["bread"]] - rng.integers(0, 400, n)
    rec["diffs"] = rng.integers(0, 3000, n)
    keep = oracle.filter_error(rec, 0.3)
    assert np.array_equal(keep, rec["diffs"] / (rec["aepos"] - rec["abpos"]) <= 0.3)
    kp = oracle.filter_pileup(rec, lens, lens, 126)
    ab, bb = rec["abpos"] <= 126, rec["bbpos"] <= 126
    ae, be = rec["aepos"] + 126 >= lens[rec["aread"]], rec["bepos"] + 126 >= lens[rec["bread"]]
    exp = (rec["aread"] != rec["bread"]) & (((ab & bb) & (ae | be)) | ((ae & be) & (ab | bb)))
    assert np.array_equal(kp, exp) and 0 < kp.sum() < n


def test_qv_rule_on_hand_made_pile():
    # read 0 (len 300, ts 100) covered by 4 LAs with known tile diffs
    rlen = np.array([300, 300, 300, 300, 300], np.int32)
    rec = np.zeros(4, oracle.LAS40); traces = []
    for i, d in enumerate([(10, 4, 2), (20, 6, 2), (30, 8, 50), (40, 10, 2)]):
        rec[i]["aread"] = 0; rec[i]["bread"] = i + 1; rec[i]["abpos"] = 0; rec[i]["aepos"] = 300
        rec[i]["bbpos"] = 0; rec[i]["bepos"] = 300; rec[i]["tlen"] = 6
        traces += [d[0], 100, d[1], 100, d[2], 100]
    toff = np.arange(4, dtype=np.int64) * 6
    q, qoff = oracle.qv(rlen, rec, toff, np.array(traces, np.uint16), 100, 4)
    # tile values: 200*d/200 = d ; best cov/2 = 2 lowest averaged, rounded half up
    assert q[qoff[0]:qoff[1]].tolist() == [15, 5, 2]
    assert q[qoff[1]:].tolist() == [50] * 12          # uncovered reads
    q2, _ = oracle.qv(rlen, rec, toff, np.array(traces, np.uint16), 100, 20)
    assert q2[:3].tolist() == [50, 50, 50]            # 4 LAs * 4 < cov 20 -> uncovered


def test_transposed_las_keeps_the_data_model_invariants():
    """orc_transpose (what damapper -C writes as Y.X.las, dazzler.d:5931-5936): every transposed record satisfies the LAS /
    trace contract pinned from the reference (tile count from the span, sum of tile B bases == B span, sum of tile diffs ==
    diffs: base.d:434-458, 196-219), coordinates mirror exactly, and transposing twice returns the original coordinates."""
    from dentist_b200 import synth
    sc = synth.make_scaffolds(1, 120000, 71, n_repeats=0)
    ref, _ = synth.contigs_from(sc, synth.make_gaps(sc, 1, 72))
    reads, _ = synth.simulate_reads(sc, 3, 7000, 2000, 0.13, 73)
    la, tr, _ = oracle.align(ref.off, ref.bases, reads.off, reads.bases, tspace=100, minlen=500)
    assert len(la) > 30 and (la["flags"] & 1).any() and not (la["flags"] & 1).all()
    t, ttoff, ttr = oracle.transpose(ref.off, ref.bases, reads.off, reads.bases, la, la["toff"].astype(np.int64), tr, 100)
    alen, blen = np.diff(ref.off), np.diff(reads.off)
    for i in range(len(la)):
        x, y = la[i], t[i]
        assert (y["aread"], y["bread"], y["flags"] & 1) == (x["bread"], x["aread"], x["flags"] & 1)
        if x["flags"] & 1:
            assert (y["abpos"], y["aepos"]) == (blen[x["bread"]] - x["bepos"], blen[x["bread"]] - x["bbpos"])
            assert (y["bbpos"], y["bepos"]) == (alen[x["aread"]] - x["aepos"], alen[x["aread"]] - x["abpos"])
        else:
            assert (y["abpos"], y["aepos"], y["bbpos"], y["bepos"]) == (x["bbpos"], x["bepos"], x["abpos"], x["aepos"])
        tiles = ttr[ttoff[i]:ttoff[i] + y["tlen"]].reshape(-1, 2).astype(np.int64)
        assert len(tiles) == -(-int(y["aepos"]) // 100) - int(y["abpos"]) // 100
        assert tiles[:, 1].sum() == y["bepos"] - y["bbpos"] and tiles[:, 0].sum() == y["diffs"]
        assert y["diffs"] <= x["diffs"]                              # per-tile optimal paths never cost more than the extension's path
        assert (tiles[:, 1] >= 0).all() and (tiles[:, 0] >= 0).all() and tiles[:, 1].max() < 250
    back, btoff, btr = oracle.transpose(reads.off, reads.bases, ref.off, ref.bases, t, ttoff, ttr, 100)
    for f in ("aread", "bread", "abpos", "aepos", "bbpos", "bepos"):
        assert np.array_equal(back[f], la[f]), f


def _edit_path(a, b):
    """unit-cost global alignment, traceback diagonal > A base unmatched > B base inserted: first column and cost on every row"""
    n, m = len(a), len(b)
    D = np.zeros((n + 1, m + 1), np.int32)
    D[0, :] = np.arange(m + 1); D[:, 0] = np.arange(n + 1)
    for i in range(1, n + 1):
        for j in range(1, m + 1):
            D[i, j] = min(D[i - 1, j - 1] + (a[i - 1] != b[j - 1]), D[i - 1, j] + 1, D[i, j - 1] + 1)
    first = {}
    i, j = n, m
    while i > 0 or j > 0:
        if i > 0 and j > 0 and D[i, j] == D[i - 1, j - 1] + (a[i - 1] != b[j - 1]):
            first[i] = (j, int(D[i, j])); i -= 1; j -= 1
        elif i > 0 and D[i, j] == D[i - 1, j] + 1:
            first[i] = (j, int(D[i, j])); i -= 1
        else:
            j -= 1
    return int(D[n, m]), first


def test_bridging_by_hand():
    """orc_bridge (`daligner -B`): two exact alignments either side of a junk stretch become one record whose tiles are the
    pieces of the concatenated path; checked against an independent DP for three positions of the gap on the tile grid."""
    rng = np.random.default_rng(11)
    ts = 100
    for p_end, gap_a, gap_b, q_len in ((300, 50, 45, 350), (270, 60, 71, 330), (250, 128, 120, 222), (300, 0, 7, 200), (233, 0, 0, 167)):
        A = rng.integers(0, 4, p_end + gap_a + q_len).astype(np.uint8)
        junk = rng.integers(0, 4, gap_b).astype(np.uint8)
        B = np.concatenate([A[:p_end], junk, A[p_end + gap_a:]])
        aoff = np.array([0, len(A)], np.int64); boff = np.array([0, len(B)], np.int64)

        def exact(ab, ae, bb):
            tiles, p = [], ab
            while p < ae:
                e = min((p // ts + 1) * ts, ae); tiles.append((0, e - p)); p = e
            return dict(abpos=ab, aepos=ae, bbpos=bb, bepos=bb + (ae - ab), diffs=0, tlen=2 * len(tiles)), tiles
        P, tp = exact(0, p_end, 0)
        Q, tq = exact(p_end + gap_a, len(A), p_end + gap_b)
        rec = np.zeros(2, oracle.LAS40)
        for i, r in enumerate((P, Q)):
            for f, v in r.items():
                rec[i][f] = v
        trace = np.array([x for t in tp + tq for x in t], np.uint16)
        toff = np.array([0, 2 * len(tp)], np.int64)
        out, otoff, otr, nb = oracle.bridge(aoff, A, boff, B, rec, toff, trace, ts)
        assert nb == 1 and len(out) == 1
        ed, first = _edit_path(A[p_end:p_end + gap_a], B[p_end:p_end + gap_b])
        o = out[0]
        assert (o["abpos"], o["aepos"], o["bbpos"], o["bepos"], o["diffs"]) == (0, len(A), 0, len(B), ed)
        # expected tiles: cumulative (B offset, cost) of the whole path at every multiple of ts of A, first arrival
        cum = []
        for g in range(ts, len(A), ts):
            if g <= p_end:
                cum.append((g, 0))
            elif g <= p_end + gap_a:
                j, d = first[g - p_end]; cum.append((p_end + j, d))
            else:
                cum.append((g - gap_a + gap_b, ed))
        cum.append((len(B), ed))
        exp, pb, pd = [], 0, 0
        for b, d in cum:
            exp += [d - pd, b - pb]; pb, pd = b, d
        assert otr.tolist() == exp and o["tlen"] == len(exp)
    # no bridge: other strand, other read, overlapping or far-apart neighbours, too much diagonal drift
    rec = np.zeros(2, oracle.LAS40)
    base = dict(abpos=0, aepos=100, bbpos=0, bepos=100, diffs=0, tlen=2)
    for i in range(2):
        for f, v in base.items():
            rec[i][f] = v
    A = rng.integers(0, 4, 1000).astype(np.uint8); aoff = np.array([0, 1000], np.int64)
    for ab, bb, flags in ((90, 90, 0), (100 + 129, 100 + 129, 0), (110, 100 + 251, 0), (150, 100, 0), (110, 110, 1)):
        rec[1]["abpos"], rec[1]["aepos"], rec[1]["bbpos"], rec[1]["bepos"], rec[1]["flags"] = ab, ab + 100, bb, bb + 100, flags
        tr = np.array([0, 100, 0, (ab // ts + 1) * ts - ab, 0, 100 - ((ab // ts + 1) * ts - ab)], np.uint16)
        n2 = 2 if ab % ts else 1
        rec[1]["tlen"] = 2 * n2
        out, _, _, nb = oracle.bridge(aoff, A, aoff, A, rec, np.array([0, 2], np.int64), tr[:2 + 2 * n2], ts)
        assert nb == 0 and len(out) == 2


def test_bridging_random_gaps_against_independent_dp():
    """orc_bridge over seeded random gap geometries and three trace spacings: merged coordinates, diffs and every tile against
    the path of an independent DP (first arrival on each trace-point row of A)."""
    rng = np.random.default_rng(2024)
    done = 0
    for case in range(60):
        ts = int(rng.choice([40, 100, 126]))
        p_end = int(rng.integers(150, 400)); q_len = int(rng.integers(150, 400))
        gap_a = int(rng.integers(0, 129)); gap_b = int(rng.integers(0, 251))
        if abs(gap_a - gap_b) * 20 > max(320, 6 * max(gap_a, gap_b)):
            gap_b = max(0, min(250, gap_a + int(rng.integers(-10, 11))))            # keep the pair bridgeable
        A = rng.integers(0, 4, p_end + gap_a + q_len).astype(np.uint8)
        # the bridged stretch of B: a noisy copy of A's (some bridges cheap, some junk)
        core = A[p_end:p_end + gap_a]
        junk = rng.integers(0, 4, gap_b).astype(np.uint8)
        if gap_a and gap_b and case % 2:
            junk[:min(gap_a, gap_b)] = core[:min(gap_a, gap_b)]
            flip = rng.random(gap_b) < 0.15; junk[flip] = (junk[flip] + 1) % 4
        B = np.concatenate([A[:p_end], junk, A[p_end + gap_a:]])
        aoff = np.array([0, len(A)], np.int64); boff = np.array([0, len(B)], np.int64)

        def exact(ab, ae, bb):
            tiles, p = [], ab
            while p < ae:
                e = min((p // ts + 1) * ts, ae); tiles.append((0, e - p)); p = e
            return dict(abpos=ab, aepos=ae, bbpos=bb, bepos=bb + (ae - ab), diffs=0, tlen=2 * len(tiles)), tiles
        P, tp = exact(0, p_end, 0)
        Q, tq = exact(p_end + gap_a, len(A), p_end + gap_b)
        rec = np.zeros(2, oracle.LAS40)
        for i, r in enumerate((P, Q)):
            for f, v in r.items():
                rec[i][f] = v
        trace = np.array([x for t in tp + tq for x in t], np.uint16)
        out, otoff, otr, nb = oracle.bridge(aoff, A, boff, B, rec, np.array([0, 2 * len(tp)], np.int64), trace, ts)
        assert nb == 1 and len(out) == 1, (case, gap_a, gap_b)
        ed, first = _edit_path(A[p_end:p_end + gap_a], B[p_end:p_end + gap_b])
        o = out[0]
        assert (o["abpos"], o["aepos"], o["bbpos"], o["bepos"], o["diffs"]) == (0, len(A), 0, len(B), ed)
        cum = []
        for g in range(ts, len(A), ts):
            if g <= p_end:
                cum.append((g, 0))
            elif g <= p_end + gap_a:
                j, d = first[g - p_end]; cum.append((p_end + j, d))
            else:
                cum.append((g - gap_a + gap_b, ed))
        cum.append((len(B), ed))
        exp, pb, pd = [], 0, 0
        for b, d in cum:
            exp += [d - pd, b - pb]; pb, pd = b, d
        assert otr.tolist() == exp, (case, ts, p_end, gap_a, gap_b)
        done += 1
    assert done == 60


def test_bit_parallel_dp_reproduces_the_cell_dp(tmp_path):
    """The identity the CUDA kernels' bit-parallel DP rests on (tests/cpp/bitparallel_dp_check.c): all cell values and
    every traceback direction of the plain unit-cost DP, from 128-bit vertical-delta columns."""
    import subprocess
    src = os.path.join(os.path.dirname(__file__), "cpp", "bitparallel_dp_check.c")
    exe = str(tmp_path / "bvcheck")
    subprocess.check_call(["gcc", "-O2", "-o", exe, src])
    r = subprocess.run([exe, "20000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "all ok" in r.stdout, r.stdout
