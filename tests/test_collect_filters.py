"""A15: collectPileUps alignment filters.  CPU: the oracle replays the reference's own predicate KATs
(base.d:605-657 isFullyContained, :687-700 averageErrorRate).  GPU: device filters == oracle."""
import numpy as np
import pytest

from dentist_b200 import synth
from oracle import collect_filters as cf

START, BEST, NEXT = 0x4, 0x10, 0x8
DT = np.dtype([("tlen", "<i4"), ("diffs", "<i4"), ("abpos", "<i4"), ("bbpos", "<i4"), ("aepos", "<i4"), ("bepos", "<i4"),
               ("flags", "<u4"), ("aread", "<i4"), ("bread", "<i4"), ("pad", "<i4")])


def _chain(las, a=0, b=0, flags=0):
    rec = np.zeros(len(las), DT)
    for i, (ab, ae, bb, be, d) in enumerate(las):
        rec[i] = (0, d, ab, bb, ae, be, flags | ((START | BEST) if i == 0 else NEXT), a, b, 0)
    return rec


def test_is_fully_contained_reference_kats():
    # base.d:605-657: Contig A length 50, read length 15
    cases = [([(30, 35, 5, 10, 1)], True), ([(10, 20, 5, 10, 1), (30, 40, 5, 10, 1)], True), ([(0, 10, 5, 10, 1)], False),
             ([(40, 50, 5, 10, 1)], False), ([(0, 20, 5, 10, 1), (30, 50, 5, 10, 1)], False)]
    for las, expected in cases:
        _, st, used = cf.collect_filter(_chain(las), [50], [15], {}, max_err=1.0, allowance=10 ** 6, min_anchor=0)
        assert (st[0] == 6) == expected and (used == [0]) == expected


def test_error_rate_and_proper_kats():
    # base.d:687-700: two LAs (1..3, 5..10) with 1 + 2 diffs -> averageErrorRate == 3/7
    rec = _chain([(1, 3, 1, 3, 1), (5, 10, 5, 10, 2)])
    assert cf.collect_filter(rec, [10], [10], {}, max_err=3.0 / 7.0, allowance=10, min_anchor=0)[1][0] == 0
    assert cf.collect_filter(rec, [10], [10], {}, max_err=0.42, allowance=10, min_anchor=0)[1][0] == 1
    # improper: neither begins within the allowance nor ends within it
    rec = _chain([(200, 700, 300, 800, 10)])
    assert cf.collect_filter(rec, [1000], [1100], {}, 0.3, 100, 0)[1][0] == 2
    assert cf.collect_filter(rec, [1000], [820], {}, 0.3, 100, 0)[1][0] == 2          # ends with B, does not begin with anything
    rec = _chain([(200, 700, 50, 550, 10)])
    assert cf.collect_filter(rec, [720], [600], {}, 0.3, 100, 0)[1][0] == 0           # begins with B, ends with A; the read sticks out of the contig
    assert cf.collect_filter(rec, [1000], [600], {}, 0.3, 100, 0)[1][0] == 6          # same read inside a longer contig: redundant
    # weakly anchored: 500 aligned bases, 450 of them masked
    assert cf.collect_filter(rec, [720], [600], {0: [(150, 650)]}, 0.3, 100, 100)[1][0] == 3


def _scenario(seed):
    sc = synth.make_scaffolds(1, 150000, seed, n_repeats=2, repeat_len=1500, repeat_copies=4)
    ref, _ = synth.contigs_from(sc, synth.make_gaps(sc, 2, seed + 1))
    reads, _ = synth.simulate_reads(sc, 4, 9000, 3000, 0.12, seed + 2)
    return ref, reads


@pytest.mark.gpu
def test_device_collect_filters_match_oracle():
    from dentist_b200 import dazzler
    ref, reads = _scenario(501)
    ga, gb = dazzler.Block(ref.off, ref.bases), dazzler.Block(reads.off, reads.bases)
    las = dazzler.align(ga, gb, tspace=100, minlen=500)
    las.chainMapper(reads.nreads)
    alen, blen = np.diff(ref.off), np.diff(reads.off)
    mask = [[(1000, 4000), (20000, 20500)] if r % 2 == 0 else [] for r in range(ref.nreads)]
    for kw in (dict(), dict(max_alignment_error=0.125, min_anchor_length=3000, proper_alignment_allowance=30)):
        first, st, used = dazzler.collectFilter(las, alen, blen, repeat_mask=mask, **kw)
        of, ost, oused = cf.collect_filter(las.rec, alen, blen, {r: m for r, m in enumerate(mask)},
                                           kw.get("max_alignment_error", 0.3), kw.get("proper_alignment_allowance", 100),
                                           kw.get("min_anchor_length", 500))
        assert np.array_equal(first, of) and np.array_equal(st, ost) and used == oused
        assert len(first) > 100 and (st == 0).sum() > 10      # most reads lie inside a contig: redundant
    assert len(set(st.tolist())) >= 4                      # several filters actually fire in the strict setting
