"""Pins the oracle's LAS codec / flag mapping / chain packing / trace-point arithmetic against the
reference's OWN golden vectors (extracted by tests/golden/make_golden.py from dazzler.d / base.d)."""
import json
import os

import numpy as np
import pytest

from oracle import las

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "las_golden.json")))


def _parsed():
    return las.parse_ladump(G["ladump"])


def test_ladump_to_flat_matches_reference_expectation():
    # dazzler.d:1029-1169: dumpLA(text) -> LocalAlignmentReader == expected FlatLocalAlignment[]
    tspace, recs, traces = _parsed()
    assert tspace == G["tspace"] == 100
    buf = las.encode(recs, traces, tspace)
    ts2, r2, t2 = las.decode(buf)
    assert ts2 == 100 and len(r2) == len(G["flat"]) == 11
    for i, exp in enumerate(G["flat"]):
        assert exp["id"] == i
        assert r2[i]["aread"] + 1 == exp["contigA"] and r2[i]["bread"] + 1 == exp["contigB"]
        assert (r2[i]["abpos"], r2[i]["aepos"], r2[i]["bbpos"], r2[i]["bepos"]) == \
            (exp["abpos"], exp["aepos"], exp["bbpos"], exp["bepos"])
        assert las.dentist_flags(int(r2[i]["flags"])) == set(exp["flags"])
        assert t2[i].tolist() == exp["trace"]
        assert r2[i]["diffs"] == sum(t[0] for t in exp["trace"])   # dazzler.d:2143
        # inverse flag mapping (dazzler.d:2100-2114)
        assert las.las_flags(set(exp["flags"])) == int(r2[i]["flags"])


def test_chain_packing_matches_reference_expectation():
    # dazzler.d:482-657
    tspace, recs, traces = _parsed()
    _, r2, t2 = las.decode(las.encode(recs, traces, tspace))
    groups = las.chains(r2)
    assert len(groups) == len(G["chains"]) == 7
    for g, exp in zip(groups, G["chains"]):
        first = r2[g[0]]
        assert first["aread"] + 1 == exp["contigA"] and first["bread"] + 1 == exp["contigB"]
        fl = las.dentist_flags(int(first["flags"])) - {"chainContinuation"}
        assert fl == set(exp["flags"])
        assert len(g) == len(exp["las"])
        for idx, e in zip(g, exp["las"]):
            assert [r2[idx]["abpos"], r2[idx]["aepos"], r2[idx]["bbpos"], r2[idx]["bepos"],
                    int(t2[idx][:, 0].sum())] == e


def test_record_layout_is_40_bytes_and_header_12():
    # dazzler.d:1717-1725 (DazzlerOverlap[8..48)), :1672-1688
    tspace, recs, traces = _parsed()
    buf = las.encode(recs, traces, tspace)
    ntp = sum(len(t) for t in traces)
    assert len(buf) == 12 + 40 * len(recs) + 2 * ntp * 1          # tspace 100 -> uint8 traces
    buf2 = las.encode(recs, traces, 126)
    assert len(buf2) == 12 + 40 * len(recs) + 2 * ntp * 2         # tspace > 125 -> uint16 traces
    assert int.from_bytes(buf[:8], "little") == 11 and int.from_bytes(buf[8:12], "little") == 100


@pytest.mark.parametrize("tspace", [100, 200, 1337])
def test_roundtrip_both_trace_widths(tspace):
    # dazzler.d:1864-1905 (tspace 100 and 200), header test :6049-6112 (1337)
    recs = [dict(aread=0, bread=1, flags=las.START | las.BEST, abpos=3, aepos=4, bbpos=5, bepos=6),
            dict(aread=0, bread=1, flags=las.NEXT, abpos=12, aepos=13, bbpos=14, bepos=15),
            dict(aread=18, bread=19, flags=las.COMP | las.START, abpos=21, aepos=22, bbpos=23, bepos=24),
            dict(aread=18, bread=19, flags=las.COMP | las.NEXT, abpos=30, aepos=31, bbpos=32, bepos=33)]
    traces = [[(7, 1)], [(16, 1)], [(25, 1)], [(0, 1)]]
    ts, r, t = las.decode(las.encode(recs, traces, tspace))
    assert ts == tspace
    assert [int(x) for x in r["diffs"]] == [7, 16, 25, 0]
    assert [x.tolist() for x in t] == [[list(p) for p in tr] for tr in traces]
    assert las.dentist_flags(int(r[2]["flags"])) == {"complement", "alternateChain"}


def test_truncated_las_raises():
    tspace, recs, traces = _parsed()
    buf = las.encode(recs, traces, tspace)
    with pytest.raises(ValueError):
        las.decode(buf[:-1])


def test_trace_point_translation_kat():
    # base.d:881-944 -- a real daligner trace (21 tiles) with 12 assertions
    k = G["trace_kat"]
    tr = k["trace"]
    assert len(tr) == las.num_tiles(k["abpos"], k["aepos"], k["tspace"]) == 21
    assert sum(t[0] for t in tr) == k["diffs"]
    assert sum(t[1] for t in tr) == k["bepos"] - k["bbpos"]        # invariant base.d:434-458
    f = lambda pos, mode: las.translate_trace_point(k["abpos"], k["aepos"], k["bbpos"], k["tspace"], tr, pos, mode)
    for a in k["asserts"]:
        assert f(a["pos"], a["mode"]) == (a["a"], a["b"]), a
    for p1, m1, p2, m2 in k["equal_pairs"]:
        assert f(p1, m1) == f(p2, m2)
    for p in k["throws"]:
        with pytest.raises(ValueError):
            f(p, "floor")


def test_oracle_alignment_satisfies_trace_invariants():
    """The aligner oracle's output honours the same tile semantics as the reference's real trace."""
    from dentist_b200 import synth
    from oracle import oracle
    sc = synth.make_scaffolds(1, 60000, 5, n_repeats=0)
    ref, _ = synth.contigs_from(sc, synth.make_gaps(sc, 1, 6))
    reads, _ = synth.simulate_reads(sc, 3, 5000, 1500, 0.13, 7)
    for ts in (100, 126):
        la, tr, _ = oracle.align(ref.off, ref.bases, reads.off, reads.bases, tspace=ts, minlen=500)
        assert len(la) > 10
        for r in la:
            t = tr[r["toff"]:r["toff"] + r["tlen"]].reshape(-1, 2)
            assert len(t) == las.num_tiles(int(r["abpos"]), int(r["aepos"]), ts)
            assert int(t[:, 0].sum()) == r["diffs"]
            assert int(t[:, 1].sum()) == r["bepos"] - r["bbpos"]
        key = np.stack([la["aread"], la["bread"], la["flags"] & 1, la["abpos"]], 1).tolist()
        assert key == sorted(key)                                   # LAsort order, base.d:1787-1809


def test_oracle_is_shard_and_thread_invariant():
    """Spec item 7: the rounds are counted per (bread, strand, aread) group, so the oracle's output for a read does
    not depend on the other reads of its block (one-read shards merge to the whole-block result), nor on the number
    of host threads sharing the A index."""
    from dentist_b200 import sharding, synth
    from oracle import oracle
    from dentist_b200._lib import REC_DTYPE
    sc = synth.make_scaffolds(2, 100000, 81, n_repeats=2, repeat_copies=6)
    ref, _ = synth.contigs_from(sc, synth.make_gaps(sc, 2, 82))
    reads, _ = synth.simulate_reads(sc, 1.5, 6000, 2000, 0.13, 83)
    la, tr, st = oracle.align(ref.off, ref.bases, reads.off, reads.bases, tspace=100, minlen=500)
    la4, tr4, st4 = oracle.align(ref.off, ref.bases, reads.off, reads.bases, tspace=100, minlen=500, threads=4)
    assert len(la) > 100 and la.tobytes() == la4.tobytes() and np.array_equal(tr, tr4) and st == st4
    recs, trs = [], []
    for r in range(reads.nreads):
        o = reads.off[r:r + 2] - reads.off[r]
        l1, t1, _ = oracle.align(ref.off, ref.bases, o, reads.bases[reads.off[r]:reads.off[r + 1]], tspace=100, minlen=500)
        rec = np.zeros(len(l1), REC_DTYPE)
        for f in ("tlen", "diffs", "abpos", "bbpos", "aepos", "bepos", "flags", "aread", "bread"):
            rec[f] = l1[f]
        rec["bread"] += r
        recs.append(rec); trs.append(t1)
    mrec, mtoff, mtr = sharding.merge_las(recs, trs)
    for f in ("tlen", "diffs", "abpos", "bbpos", "aepos", "bepos", "flags", "aread", "bread"):
        assert np.array_equal(mrec[f], la[f]), f
    assert np.array_equal(mtr, tr) and np.array_equal(mtoff, la["toff"])
