"""`dentist mask-repetitive-regions`: the oracle against the reference's own vectors
(commands/maskRepetitiveRegions.d:395-411, 582-617) on CPU, and the device coverage mask against the oracle."""
import json
import os

import numpy as np
import pytest

from dentist_b200 import synth
from oracle import mask_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "maskcov_kat.json")))


def _kat():
    return [tuple(x) for x in KAT["alignments"]], [tuple(x) for x in KAT["contigs"]]


def test_oracle_reproduces_the_reference_vectors():
    iv, ct = _kat()
    assert [list(x) for x in mask_oracle.coverage_changes(iv, ct)] == KAT["changes"]
    assert [list(x) for x in mask_oracle.bad_coverage_mask(iv, ct, *KAT["bounds"])] == KAT["mask"]
    assert mask_oracle.bad_coverage_mask([], ct, 3, 5) == []
    # upper bound only (the defaults are [0, maxCoverage], commandline.d:1889): only the > 5 stretches remain
    assert mask_oracle.bad_coverage_mask(iv, ct, 0, 5) == [(1, 10, 18), (1, 20, 30), (2, 0, 3), (2, 5, 15)]


def _las_of(intervals, flags=None):
    from dentist_b200 import _lib, dazzler
    rec = np.zeros(len(intervals), _lib.REC_DTYPE)
    for i, (c, b, e) in enumerate(intervals):
        rec[i]["aread"], rec[i]["abpos"], rec[i]["aepos"], rec[i]["bread"], rec[i]["bepos"] = c, b, e, 0, e - b
        rec[i]["flags"] = 0 if flags is None else flags[i]
    buf = _lib.LasBuf(); buf.nrec = len(rec); buf.tspace = 100
    import ctypes as C
    buf.rec = C.cast(rec.ctypes.data, C.POINTER(_lib.LasRecord))
    class _L:                                   # caller-owned LAS buffer (no traces needed by this entry point)
        pass
    las = _L(); las._buf = buf; las._keep = rec
    return las


@pytest.mark.gpu
def test_device_mask_equals_the_reference_vector():
    from dentist_b200 import dazzler
    iv, ct = _kat()
    las = _las_of([(c - 1, b, e) for c, b, e in iv])
    alen = [e for _, _, e in ct]
    got = dazzler.maskRepetitiveRegions(las, alen, [40], KAT["bounds"])
    assert [[c + 1, b, e] for c in range(3) for b, e in got[c]] == KAT["mask"]
    assert dazzler.maskRepetitiveRegions(_las_of([]), alen, [40], (3, 5)) == [[], [], []]
    # improper pass that selects no chain gives an empty mask even with a positive lower bound
    full = _las_of([(0, 0, 30)])
    assert dazzler.maskRepetitiveRegions(full, [30], [30], (0, 9), improper_coverage_bounds=(2, 9), proper_alignment_allowance=0) == [[]]
    assert dazzler.maskRepetitiveRegions(full, [30], [30], (2, 9)) == [[(0, 30)]]


@pytest.mark.gpu
def test_device_mask_on_a_mapping_with_repeats():
    from dentist_b200 import dazzler
    sc = synth.make_scaffolds(2, 150000, 77, n_repeats=3)
    ref, _ = synth.contigs_from(sc, synth.make_gaps(sc, 2, 78))
    reads, _ = synth.simulate_reads(sc, 12, 6000, 2000, 0.12, 79)
    alen, blen = np.diff(ref.off), np.diff(reads.off)
    las = dazzler.align(dazzler.Block(ref.off, ref.bases), dazzler.Block(reads.off, reads.bases), tspace=100, minlen=500)
    las.chainMapper(reads.nreads)
    assert len(las) > 500
    for bounds, imp in (((0, 18), None), ((4, 18), (0, 3)), ((0, 30), (0, 1))):
        got = dazzler.maskRepetitiveRegions(las, alen, blen, bounds, imp, 100)
        want = mask_oracle.mask_repetitive_regions(las.rec, alen, blen, bounds, imp, 100)
        assert [(c, b, e) for c in range(len(alen)) for b, e in got[c]] == want
    assert any(got)                                                   # the planted repeats are over-covered somewhere


def test_propagate_oracle_on_the_trace_kat():
    """One alignment = the reference's 21-tile KAT (base.d:886-941): floor / ceil translation of mask ends, clipping of the
    first and last interval to the alignment, mirroring for a complement alignment."""
    from dentist_b200 import _lib
    g = json.load(open(os.path.join(HERE, "golden", "las_golden.json")))["trace_kat"]
    tr = np.array(g["trace"], np.uint16)
    rec = np.zeros(2, _lib.REC_DTYPE)
    for i in range(2):
        rec[i]["abpos"], rec[i]["aepos"], rec[i]["bbpos"], rec[i]["bepos"] = g["abpos"], g["aepos"], g["bbpos"], g["bepos"]
        rec[i]["bread"], rec[i]["flags"], rec[i]["tlen"] = i, i, 2 * len(tr)
    blen = [g["bepos"] + 50, g["bepos"] + 50]
    want = {a["pos"]: a["b"] for a in g["asserts"] if a["mode"] == "floor"}
    # the reference asserts translate(699, ceil) == translate(701, floor) (base.d:936-941), and floor(701) == floor(700)
    assert [699, "ceil", 701, "floor"] in g["equal_pairs"]
    p0, p1 = 600, 699
    out = mask_oracle.propagate_mask(rec, [tr, tr], g["tspace"], [[(p0, p1)]], blen)
    assert out[0] == [(want[600], want[700])] == [(23, 132)]
    assert out[1] == [(blen[1] - want[700], blen[1] - want[600])]
    # a mask that sticks out on both sides is clipped to the alignment: the whole B interval
    out = mask_oracle.propagate_mask(rec[:1], [tr], g["tspace"], [[(0, g["aepos"] + 1000)]], blen)
    assert out[0] == [(g["bbpos"], g["bepos"])]


@pytest.mark.gpu
def test_device_propagation_equals_the_oracle():
    from dentist_b200 import dazzler
    sc = synth.make_scaffolds(2, 150000, 87, n_repeats=3)
    ref, _ = synth.contigs_from(sc, synth.make_gaps(sc, 2, 88))
    reads, _ = synth.simulate_reads(sc, 10, 6000, 2000, 0.12, 89)
    alen, blen = np.diff(ref.off), np.diff(reads.off)
    las = dazzler.align(dazzler.Block(ref.off, ref.bases), dazzler.Block(reads.off, reads.bases), tspace=100, minlen=500)
    rng = np.random.default_rng(5)
    mask = []
    for L in alen:                                                     # random disjoint intervals, some tiny, some spanning tiles
        cuts = np.sort(rng.choice(int(L), size=40, replace=False))
        mask.append([(int(cuts[i]), int(cuts[i + 1])) for i in range(0, 40, 2)])
    got = dazzler.propagateMask(las, mask, len(alen), blen)
    want = mask_oracle.propagate_mask(las.rec, las.traces(), 100, mask, blen)
    assert got == want and sum(len(x) for x in got) > 100
    # and back: the propagated read mask through the transposed roles gives something inside the contigs
    assert dazzler.propagateMask(las, [[] for _ in alen], len(alen), blen) == [[] for _ in blen]
