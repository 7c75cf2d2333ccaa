"""The whole hot path on a synthetic assembly, driven like `dentist collect | dentist process`:
map reads to contigs (dn_damap path) -> collect filters -> pile-ups (PileUpDb file) -> crop at common trace points ->
batched processPileUp (dust, align, filter, chain, QVs, reference read, consensus, flank alignment) -> InsertionDb.
The insertions close the gaps: each consensus contains the true gap sequence with >= 97 % identity although every
read carries 13 % errors, and its overlaps anchor it properly in both flanking contigs."""
import numpy as np
import pytest

from dentist_b200 import binio, process, synth

pytestmark = pytest.mark.gpu


def _identity(a, b):
    """1 - edit distance / len via numpy row DP (sequences of a few kb)."""
    prev = np.arange(len(b) + 1)
    for i in range(1, len(a) + 1):
        cur = np.minimum(prev[:-1] + (a[i - 1] != b), prev[1:] + 1)
        cur = np.concatenate([[i], cur])
        cur = np.minimum.accumulate(cur - np.arange(len(cur))) + np.arange(len(cur))     # insertion chain
        prev = cur
    return 1.0 - prev[-1] / max(len(a), len(b))


def _build_pileups(las, alen, blen, kept, n_gaps):
    """`dentist collect` in miniature: the kept chains of every read become read alignments (collectReadAlignments,
    collectPileUps/pileups.d:821-888, with SeededAlignment.from seeds); read alignments with the same join -- gap
    (contig g end, contig g+1 begin) or an extension into that gap from either side -- form one pile-up (what bundling
    the scaffold graph's edges does, pileups.d:650-666)."""
    rec, toff, trace = las.rec, las.toff, las.trace
    by_read = {}
    for i in kept:
        j = i + 1
        while j < len(rec) and (int(rec[j]["flags"]) & 0x8):
            j += 1
        sub = slice(int(i), j)
        chain = binio.seeded_alignments_from_las(rec[sub], toff[sub], trace, alen, blen, 100, lambda f, l: "front")[0]
        chain.pop("seed"); chain["id"] = int(i)
        by_read.setdefault(int(rec[i]["bread"]), []).append(chain)
    piles = [[] for _ in range(n_gaps)]
    for b in sorted(by_read):
        ras, reason = process.collect_read_alignments(by_read[b])
        for ra in ras:
            if process._is_gap(ra) and not process.is_in_order(ra):
                ra = [ra[1], ra[0]]                                   # ReadAlignment.getInOrder, base.d:2177-2183
            start, end = process.make_join(ra)
            if process._is_gap(ra):
                if start[1] == "end" and end[1] == "begin" and end[0] == start[0] + 1 and process._is_parallel(ra):
                    piles[start[0] - 1].append(ra)
            elif ra[0]["seed"] == "back" and ra[0]["contigA"][0] <= n_gaps:
                piles[ra[0]["contigA"][0] - 1].append(ra)             # extends contig g beyond its end, into gap g
            elif ra[0]["seed"] == "front" and ra[0]["contigA"][0] >= 2:
                piles[ra[0]["contigA"][0] - 2].append(ra)             # extends contig g+1 beyond its begin, into gap g
    return piles


def test_collect_process_output_round_trip(tmp_path):
    from dentist_b200 import dazzler
    sc = synth.make_scaffolds(1, 300000, 901, n_repeats=0)
    gaps = synth.make_gaps(sc, 3, 902, min_len=300, max_len=900)
    ref, meta = synth.contigs_from(sc, gaps)
    reads, truth = synth.simulate_reads(sc, 25, 9000, 2500, 0.13, 903)
    alen, blen = np.diff(ref.off), np.diff(reads.off)
    ga, gb = dazzler.Block(ref.off, ref.bases), dazzler.Block(reads.off, reads.bases)
    las = dazzler.align(ga, gb, tspace=100, minlen=500)                      # ref vs reads (damapper role)
    las.chainMapper(reads.nreads)
    first, st, used = dazzler.collectFilter(las, alen, blen, max_alignment_error=0.3, proper_alignment_allowance=100, min_anchor_length=500)
    piles = _build_pileups(las, alen, blen, first[st == 0], len(gaps[0]))
    assert sum(len(ra) == 2 for p in piles for ra in p) >= 6 * len(piles)
    assert all(len(p) >= 6 for p in piles) and any(len(ra) == 1 for p in piles for ra in p)
    db = str(tmp_path / "pileups.db")
    binio.write_pileup_db(db, piles)
    insertions, skipped = process.process_pileup_db(binio.read_pileup_db(db), reads, ref)
    assert not skipped and len(insertions) == len(gaps[0])
    out = str(tmp_path / "insertions.db")
    binio.write_insertion_db(out, insertions)
    back = binio.read_insertion_db(out)
    for g, ins in enumerate(back):
        assert ins["start"] == (g + 1, "end") and ins["end"] == (g + 2, "begin")            # makeJoin of a parallel gap
        assert len(ins["overlaps"]) == 2 and [o["seed"] for o in ins["overlaps"]] == ["back", "front"]
        left, right = ins["overlaps"]
        assert left["contigA"] == (g + 1, int(alen[g])) and right["contigA"] == (g + 2, int(alen[g + 1]))
        assert left["contigB"] == (1, len(ins["sequence"])) == right["contigB"]
        assert list(ins["read_ids"]) == sorted(ra[0]["contigB"][0] for ra in piles[g])
        # the consensus, cut between its two anchors' inner ends, is the gap (+ a few anchor bases)
        comp = bool(left["flags"] & 1)
        cons = ins["sequence"] if not comp else (3 - ins["sequence"])[::-1]
        _, lbeg = meta[g]; _, rbeg = meta[g + 1]
        t = sc[0][lbeg + left["las"][0]["ab"]:rbeg + right["las"][-1]["ae"]]
        ident = _identity(cons, t)
        assert ident >= 0.97, (g, ident, len(cons), len(t))
