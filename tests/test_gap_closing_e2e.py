"""The whole hot path on a synthetic assembly, driven like `dentist collect | dentist process`:
map reads to contigs (dn_damap path) -> collect filters -> pile-ups (PileUpDb file) -> crop at common trace points ->
batched processPileUp (dust, align, filter, chain, QVs, reference read, consensus, flank alignment) -> InsertionDb.
The insertions close the gaps: each consensus contains the true gap sequence with >= 97 % identity although every
read carries 13 % errors, and its overlaps anchor it properly in both flanking contigs."""
import numpy as np
import pytest

from dentist_b200 import binio, process, synth

pytestmark = pytest.mark.gpu


from tests.pipeline_util import build_pileups as _build_pileups, identity as _identity  # noqa: E402


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


def test_example_dataset_excerpt_gaps_are_closed(tmp_path):
    """BASELINE configs[0] in miniature: 1.5 Mbp of the reference's own example assembly with the four gaps its gaps.bed
    places there (tests/golden/example_excerpt.npz, cut by tests/golden/make_example_excerpt.py), reads from our generator
    with the example's simulator settings (-m25000 -s12500 -e.13 -c20, example/Makefile:13; the DAZZ_DB simulator itself is
    absent).  map -> collect filters -> PileUpDb -> dn_process_pileups -> InsertionDb closes all four gaps."""
    import os
    from tests import pipeline_util
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "example_excerpt.npz"))
    n = int(z["length"])
    p = z["packed"]
    scaffold = np.stack([(p >> 6) & 3, (p >> 4) & 3, (p >> 2) & 3, p & 3], 1).reshape(-1)[:n].astype(np.uint8)
    gaps = [tuple(int(x) for x in g) for g in z["gaps"]]
    assert n == 1500000 and len(gaps) == 4
    ref, meta = synth.contigs_from([scaffold], [gaps])
    reads, _ = synth.simulate_reads([scaffold], 20, 25000, 12500, 0.13, 19339)
    ins, skipped, piles, st = pipeline_util.close_gaps(ref, reads, len(gaps), tmp_path, k=20, minlen=1000)
    assert all(len(pl) >= 6 for pl in piles), [len(pl) for pl in piles]
    assert not skipped and len(ins) == 4
    for g, i in enumerate(ins):
        assert i["start"] == (g + 1, "end") and i["end"] == (g + 2, "begin") and len(i["overlaps"]) == 2
        ident, lc, lt = pipeline_util.gap_identity(i, scaffold, meta, g)
        assert ident >= 0.97, (g, ident, lc, lt)
