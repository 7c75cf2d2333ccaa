"""GPU parity of the per-pile stages (filters, intrinsic QVs, consensus) against oracle/pile_oracle.c,
including the reference's own consensus KAT (dazzler.d:4257-4299)."""
import numpy as np
import pytest

from dentist_b200 import synth
from tests.test_pile_oracle import kat_block

pytestmark = pytest.mark.gpu


def _pile(seed, cov=10, glen=20000, rl=6000, err=0.13):
    sc = synth.make_scaffolds(1, glen, seed, n_repeats=0)
    reads, _ = synth.simulate_reads(sc, cov, rl, rl // 4, err, seed + 1)
    return reads


def _gpu_las(blk, tspace, minlen):
    from dentist_b200 import dazzler
    g = dazzler.Block(blk.off, blk.bases)
    return g, dazzler.align(g, g, tspace=tspace, minlen=minlen, self_block=1)


def test_filters_match_oracle_and_reference_definition():
    from oracle import oracle
    blk = _pile(81)
    g, las = _gpu_las(blk, 126, 500)
    lens = np.diff(blk.off)
    rec0 = las.rec.copy(); toff0 = las.toff.copy()
    ke = oracle.filter_error(rec0, 0.22)
    assert 0 < ke.sum() < len(rec0)
    las.filterLocalAlignments(0.22)
    assert np.array_equal(las.rec, rec0[ke]) and np.array_equal(las.toff, toff0[ke])
    rec1 = las.rec.copy()
    kp = oracle.filter_pileup(rec1, lens, lens, 126)
    assert 0 < kp.sum() < len(rec1)
    las.filterPileUpAlignments(lens, lens, 126)
    assert np.array_equal(las.rec, rec1[kp])
    # traces still reachable through toff
    for r, t in zip(las.rec[:20], las.traces()[:20]):
        assert int(t[:, 0].sum()) == r["diffs"]


def test_qvs_match_oracle():
    from dentist_b200 import dazzler
    from oracle import oracle
    blk = _pile(91, cov=14)
    g, las = _gpu_las(blk, 126, 500)
    lens = np.diff(blk.off)
    las.filterLocalAlignments(0.3)
    for cov in (4, blk.nreads):
        q, qoff = dazzler.computeQVs(lens, las, cov)
        oq, ooff = oracle.qv(lens, las.rec, las.toff, las.trace, 126, cov)
        assert np.array_equal(qoff, ooff) and np.array_equal(q, oq)
        assert q.min() < 30 and q.max() == 50


def test_reference_consensus_kat_on_gpu():
    from dentist_b200 import dazzler
    k, blk = kat_block()
    g, las = _gpu_las(blk, 100, k["minlen"])
    lens = np.diff(blk.off)
    las.filterPileUpAlignments(lens, lens, k["allowance"])
    assert len(las) == 6
    cons = dazzler.getConsensus(g, las, [0, 1, 2])
    for c in cons:
        assert np.array_equal(c, blk.read(k["expected_read"]))


def test_consensus_matches_oracle_on_noisy_piles():
    from dentist_b200 import dazzler
    from oracle import oracle
    blk = _pile(95, cov=12, glen=15000, rl=7000)
    g, las = _gpu_las(blk, 126, 500)
    lens = np.diff(blk.off)
    las.filterLocalAlignments(0.3).filterPileUpAlignments(lens, lens, 126)
    targets = [0, 3, blk.nreads - 1, 5]
    cons = dazzler.getConsensus(g, las, targets)
    for r, c in zip(targets, cons):
        oc = oracle.consensus(blk.off, blk.bases, las.rec, las.toff, las.trace, 126, r)
        assert np.array_equal(c, oc), r
        if (las.rec["aread"] == r).any():
            assert abs(len(c) - lens[r]) < 0.1 * lens[r]
        else:
            assert len(c) == 0                                      # nothing aligns to it any more: no consensus
    # a read nothing aligns to has no consensus ("empty consensus", dazzler.d:4232-4235 -> the next candidate, package.d:307-329)
    none = dazzler.align(g, g, tspace=126, minlen=10 ** 6, self_block=1)
    empty = dazzler.getConsensus(g, none, [2, 0])
    assert len(empty[0]) == 0 and len(empty[1]) == 0
    assert len(oracle.consensus(blk.off, blk.bases, none.rec, none.toff, none.trace, 126, 2)) == 0
    # several targets in one call give what single calls give
    multi = dazzler.getConsensus(g, las, [0, 3, 5])
    assert np.array_equal(multi[0], cons[0]) and np.array_equal(multi[1], cons[1]) and np.array_equal(multi[2], cons[3])


def test_batched_pileups_equal_per_pile_processing():
    """All pile-ups of a batch in ONE block (pile id per read) give exactly what the reference's
    one-pile-at-a-time loop (package.d:153) gives: per-pile oracle runs, concatenated."""
    from dentist_b200 import pileups
    from oracle import chaining, oracle
    sc = synth.make_scaffolds(1, 120000, 301, n_repeats=0)
    gaps = synth.make_gaps(sc, 4, 302, min_len=200, max_len=1500)
    reads, group, regions = synth.make_pile_batch(sc, gaps, 303, depth=9, anchor=1200)
    ref, _ = synth.contigs_from(sc, gaps)
    res = pileups.process_pileups(reads, group, flanks=ref)
    lens = np.diff(reads.off)
    assert len(res["consensus"]) == 4
    for p in range(4):
        members = np.flatnonzero(group == p)
        lo, hi = members[0], members[-1] + 1
        off = reads.off[lo:hi + 1] - reads.off[lo]
        bases = reads.bases[reads.off[lo]:reads.off[hi]]
        la, tr, _ = oracle.align(off, bases, off, bases, tspace=126, minlen=500, self=1)
        toff = la["toff"].astype(np.int64)
        keep = oracle.filter_error(la, 0.3)
        la, toff = la[keep], toff[keep]
        src, fl = chaining.chain_local_alignments(la, chaining.ChainingOptions(min_score=126))
        la, toff = la[src].copy(), toff[src]
        la["flags"] = fl
        q, qoff = oracle.qv(lens[lo:hi], la, toff, tr, 126, max(len(members), 4) if len(members) >= 4 else len(members))
        assert np.array_equal(q, res["qv"][res["qoff"][lo]:res["qoff"][hi]])
        kp = oracle.filter_pileup(la, lens[lo:hi], lens[lo:hi], 126)
        la, toff = la[kp], toff[kp]
        la["flags"] &= 0x21                                              # Yes.forceFlat: unchained + FlatLocalAlignment order
        o = np.lexsort((la["diffs"], la["bepos"], la["bbpos"], la["aepos"], la["abpos"], la["flags"] & 1, la["bread"], la["aread"]))
        la, toff = la[o], toff[o]
        sel = (res["las"].rec["aread"] >= lo) & (res["las"].rec["aread"] < hi)
        grec = res["las"].rec[sel]
        assert len(grec) == len(la)
        assert np.array_equal(grec["aread"] - lo, la["aread"]) and np.array_equal(grec["bread"] - lo, la["bread"])
        for f in ("abpos", "aepos", "bbpos", "bepos", "diffs", "flags"):
            assert np.array_equal(grec[f], la[f]), f
        cand = pileups.find_reference_read_candidates(q, qoff, np.arange(len(members)))
        assert cand[0] + lo == res["reference_read"][p]
        oc = oracle.consensus(off, bases, la, toff, tr, 126, cand[0])
        assert np.array_equal(oc, res["consensus"][p])
        # the consensus is closer to the truth than the raw reference read: compare 12-mers with the region
        si, rb, re_ = regions[p]
        t = sc[si][rb:re_]
        def kmers(s, k=12):
            v = np.zeros(len(s) - k + 1, np.int64)
            for i in range(k):
                v = v * 4 + s[i:len(s) - k + 1 + i]
            return set(v.tolist())
        kt = kmers(t) | kmers((3 - t)[::-1])
        raw = reads.read(res["reference_read"][p])
        assert np.mean([x in kt for x in kmers(oc)]) > np.mean([x in kt for x in kmers(raw)]) + 0.2
    # every consensus aligns to its two flanking contigs (package.d:699-769 needs one proper overlap per flank)
    fl = res["flank_las"].rec
    for p in range(4):
        assert len(set(fl["aread"][fl["bread"] == p].tolist())) >= 2


def test_dust_mask_matches_oracle():
    from dentist_b200 import dazzler
    from oracle import dust
    rng = np.random.default_rng(17)
    seqs = []
    for i in range(12):
        s = rng.integers(0, 4, int(rng.integers(300, 3000)), dtype=np.uint8)
        if i % 3 == 0:                       # plant a homopolymer run, a dinucleotide repeat and a short tandem
            s[100:190] = 0
            s[250:290] = np.tile([1, 2], 20)
        if i % 4 == 1 and len(s) > 900:
            s[700:820] = np.tile([0, 3, 3], 40)
        seqs.append(s)
    seqs.append(np.zeros(20, np.uint8)); seqs.append(np.zeros(0, np.uint8)); seqs.append(np.full(700, 2, np.uint8))
    off = np.zeros(len(seqs) + 1, np.int64); off[1:] = np.cumsum([len(s) for s in seqs])
    blk = synth.Block(off, np.concatenate(seqs))
    g = dazzler.Block(blk.off, blk.bases)
    for kw in (dict(), dict(window=32, threshold=1.5, minlen=40)):
        got = dazzler.dbdust(g, **kw)
        okw = dict(w=kw.get("window", 64), threshold=kw.get("threshold", 2.0), minlen=kw.get("minlen", 10))
        exp = dust.dust_block(blk.off, blk.bases, **okw)
        assert got == exp
    d = dazzler.dbdust(g)
    assert any(b <= 100 and e >= 190 for b, e in d[0]) and d[-1] == [(0, 700)] and d[-2] == [] and d[2] == []
    # the mask plugs into the aligner exactly like `-mdust`
    g2 = dazzler.Block(blk.off, blk.bases, mask=d)
    assert g2.nreads == g.nreads


def test_mapper_chain_flags_match_oracle_and_pack_into_chains():
    from dentist_b200 import dazzler
    from oracle import chain_oracle, las as olas
    # reads with a large insertion in the middle align as two colinear local alignments -> one chain
    sc = synth.make_scaffolds(1, 150000, 401, n_repeats=1, repeat_copies=3)
    ref, _ = synth.contigs_from(sc, synth.make_gaps(sc, 1, 402))
    reads, _ = synth.simulate_reads(sc, 3, 9000, 2500, 0.12, 403)
    rng = np.random.default_rng(5)
    seqs = []
    for r in range(reads.nreads):
        s = reads.read(r)
        if r % 3 == 0 and len(s) > 5000:           # splice 300 random bases into the middle
            m = len(s) // 2
            s = np.concatenate([s[:m], rng.integers(0, 4, 300, dtype=np.uint8), s[m:]])
        seqs.append(s)
    off = np.zeros(len(seqs) + 1, np.int64); off[1:] = np.cumsum([len(s) for s in seqs])
    rd = synth.Block(off, np.concatenate(seqs))
    ga, gb = dazzler.Block(ref.off, ref.bases), dazzler.Block(rd.off, rd.bases)
    las = dazzler.align(ga, gb, tspace=100, minlen=500)
    before = las.rec.copy()
    las.chainMapper(rd.nreads)
    exp = chain_oracle.mapper_chain_flags(before)
    assert np.array_equal(las.rec["flags"], exp)
    groups = olas.chains(las.rec)                      # AlignmentChainPacker accepts it (dazzler.d:708-743)
    assert sum(len(g) for g in groups) == len(las) and any(len(g) > 1 for g in groups)
    # exactly one BEST chain per mapped read
    starts = las.rec[(las.rec["flags"] & olas.BEST) != 0]
    assert sorted(starts["bread"].tolist()) == sorted(set(las.rec["bread"].tolist()))


def test_chain_local_alignments_matches_oracle():
    """chaining.d restated on the device == oracle/chaining.py (exact restatement of the D source)."""
    from dentist_b200 import dazzler
    from oracle import chaining, las as olas
    sc = synth.make_scaffolds(1, 20000, 21, n_repeats=1, repeat_len=800, repeat_copies=3)
    reads, _ = synth.simulate_reads(sc, 10, 7000, 2000, 0.13, 22)
    rng = np.random.default_rng(1)
    seqs = []
    for r in range(reads.nreads):
        s = reads.read(r)
        if r % 2 == 0 and len(s) > 4000:                  # junk in the middle breaks alignments into chainable pieces
            m = len(s) // 2
            s = np.concatenate([s[:m], rng.integers(0, 4, 200, dtype=np.uint8), s[m:]])
        seqs.append(s)
    off = np.zeros(len(seqs) + 1, np.int64); off[1:] = np.cumsum([len(s) for s in seqs])
    g = dazzler.Block(off, np.concatenate(seqs))
    for opts in (dict(), dict(min_rel_score=0.5, max_rel_overlap=0.1), dict(max_indel=50, max_chain_gap=300, min_rel_score=0.2)):
        las = dazzler.align(g, g, tspace=126, minlen=500, self_block=1)
        before, toff0 = las.rec.copy(), las.toff.copy()
        assert len(before) > 500
        las.chainLocalAlignments(**opts)
        o = chaining.ChainingOptions(max_indel=opts.get("max_indel", 1000), max_chain_gap=opts.get("max_chain_gap", 10000),
                                     max_rel_overlap=opts.get("max_rel_overlap", 0.3), min_rel_score=opts.get("min_rel_score", 1.0), min_score=126)
        src, fl = chaining.chain_local_alignments(before, o)
        assert len(las) == len(src)
        assert np.array_equal(las.rec["flags"], fl)
        for f in ("aread", "bread", "abpos", "aepos", "bbpos", "bepos", "diffs", "tlen"):
            assert np.array_equal(las.rec[f], before[src][f]), f
        assert np.array_equal(las.toff, toff0[src])
        chains = olas.chains(las.rec)                       # what AlignmentChainPacker (dazzler.d:708-743) builds from it
        assert sum(len(c) for c in chains) == len(las) and any(len(c) > 1 for c in chains)
    # disabled records are ignored, unordered input is rejected like chaining.d:131-140
    las = dazzler.align(g, g, tspace=126, minlen=500, self_block=1)
    las.rec["flags"][::3] |= 0x20
    b2 = las.rec.copy()
    las.chainLocalAlignments()
    src, fl = chaining.chain_local_alignments(b2, chaining.ChainingOptions(min_score=126))
    assert np.array_equal(las.rec["flags"], fl) and len(las) == len(src)
    las = dazzler.align(g, g, tspace=126, minlen=500, self_block=1)
    las.rec[:] = las.rec[::-1].copy()
    with pytest.raises(dazzler.DnError, match="not ordered properly"):
        las.chainLocalAlignments()


def _batch_case():
    """4 gap pile-ups + one pile-up of unrelated reads (no alignments) + one whose reads are all outside
    allowedReferenceReadIds; flanking contigs with a repeat mask on one of them."""
    sc = synth.make_scaffolds(1, 120000, 501, n_repeats=0)
    gaps = synth.make_gaps(sc, 4, 502, min_len=200, max_len=1500)
    reads, group, regions = synth.make_pile_batch(sc, gaps, 503, depth=9, anchor=1200)
    ref, _ = synth.contigs_from(sc, gaps)
    rng = np.random.default_rng(9)
    piles = []
    for p in range(4):
        members = np.flatnonzero(group == p)
        piles.append(dict(reads=[reads.read(int(r)) for r in members], flanks=[p, p + 1],
                          mask=[[(100, 400)], []] if p == 1 else None,
                          allowed=[i % 5 != 0 for i in range(len(members))] if p == 2 else None))
    piles.append(dict(reads=[rng.integers(0, 4, 2500, dtype=np.uint8) for _ in range(4)], flanks=[0, 1], mask=None, allowed=None))
    members = np.flatnonzero(group == 0)
    piles.append(dict(reads=[reads.read(int(r)) for r in members], flanks=[0, 1], mask=None, allowed=[False] * len(members)))
    return ref, piles


def test_batch_entry_point_equals_stepwise_path():
    """dn_process_pileups (one C call for the batch) == the same stages called one by one through the per-stage entry
    points (pileups.process_pileups), pile-up by pile-up: status, reference read, consensus bytes, flank alignments and
    their traces.  Failing pile-ups are reported with the reference's reasons and do not disturb the others."""
    from dentist_b200 import dazzler, pileups
    ref, piles = _batch_case()
    ref_block = dazzler.Block(ref.off, ref.bases)
    got = dazzler.processPileUps(ref_block, piles)
    assert [o["status"] for o in got] == [0, 0, 0, 0, 1, 3]
    assert got[4]["reason"] == "empty pileup alignment" and got[5]["reason"] == "no valid reference read found"
    # the stepwise mirror over the same batch
    seqs = [r for p in piles for r in p["reads"]]
    group = np.concatenate([[i] * len(p["reads"]) for i, p in enumerate(piles)]).astype(np.int32)
    allowed = np.concatenate([p["allowed"] if p["allowed"] is not None else [True] * len(p["reads"]) for p in piles]).astype(bool)
    off = np.zeros(len(seqs) + 1, np.int64); off[1:] = np.cumsum([len(x) for x in seqs])
    first = np.concatenate([[0], np.cumsum([len(p["reads"]) for p in piles])])
    fl_seq, fl_group, fl_mask = [], [], []
    for i, p in enumerate(piles):
        for j, c in enumerate(p["flanks"]):
            fl_seq.append(ref.read(c)); fl_group.append(i); fl_mask.append(p["mask"][j] if p["mask"] is not None else [])
    foff = np.zeros(len(fl_seq) + 1, np.int64); foff[1:] = np.cumsum([len(x) for x in fl_seq])
    ffirst = np.concatenate([[0], np.cumsum([len(p["flanks"]) for p in piles])])
    fb = dazzler.Block(foff, np.concatenate(fl_seq), mask=fl_mask, group=np.array(fl_group, np.int32))
    fb.maskDust()
    step = pileups.process_pileups(synth.Block(off, np.concatenate(seqs)), group, flanks=fb, allowed=allowed, dust=True)
    assert step["status"].tolist() == [o["status"] for o in got]
    frec, ftr = step["flank_las"].rec, step["flank_las"].traces()
    for i, o in enumerate(got):
        if o["status"]:
            assert o["reference_read"] == -1 and len(o["consensus"]) == 0 and len(o["flank_las"]) == 0
            continue
        assert o["reference_read"] + first[i] == step["reference_read"][i]
        assert np.array_equal(o["consensus"], step["consensus"][i]) and len(o["consensus"]) > 2000
        sel = np.flatnonzero(frec["bread"] == i)
        r = o["flank_las"].rec
        assert len(r) == len(sel) >= 2 and (r["bread"] == 0).all()
        assert np.array_equal(r["aread"] + ffirst[i], frec["aread"][sel])
        for f in ("abpos", "aepos", "bbpos", "bepos", "diffs", "tlen", "flags"):
            assert np.array_equal(r[f], frec[f][sel]), f
        for a, b in zip(o["flank_las"].traces(), [ftr[j] for j in sel]):
            assert np.array_equal(a, b)
    assert got[2]["reference_read"] % 5 != 0                       # the reference read honours allowedReferenceReadIds
    # non-default options reach the aligner and the filters: a stricter error bound changes the result
    strict = dazzler.processPileUps(ref_block, piles[:2], max_alignment_error=0.22, min_anchor_length=800, proper_alignment_allowance=50)
    assert [o["status"] for o in strict] == [0, 0]
    assert any(not np.array_equal(a["consensus"], b["consensus"]) for a, b in zip(strict, got[:2]))


def _synthetic_las(tmp_path, rec, tspace=126):
    """A Las object over hand-made records (written and read back through the LAS codec; traces are dummies)."""
    import ctypes as C
    from dentist_b200 import _lib, dazzler
    nt = -(-rec["aepos"] // tspace) - rec["abpos"] // tspace
    rec["tlen"] = 2 * nt
    toff = np.cumsum(rec["tlen"]) - rec["tlen"]
    path = str(tmp_path / "synthetic.las")
    dazzler.write_las(path, tspace, rec, toff, np.zeros(int(rec["tlen"].sum()), np.uint16))
    buf = _lib.LasBuf()
    _lib.check(_lib.lib().dn_las_read(path.encode(), C.byref(buf)))
    return dazzler.Las(buf)


def test_chaining_of_oversized_groups_goes_through_the_host_restatement(tmp_path):
    """one (A,B) pair with more than 63 local alignments, and a group whose alternate chains need more than
    4 * n output records, used to fail the whole call / overrun the group's slots; both now equal the oracle."""
    from dentist_b200._lib import REC_DTYPE
    from oracle import chaining
    rng = np.random.default_rng(77)

    def colinear(n, a, b, step, ln, jitter):
        r = np.zeros(n, REC_DTYPE)
        r["aread"], r["bread"] = a, b
        r["abpos"] = 1000 + step * np.arange(n) + rng.integers(0, jitter, n)
        r["bbpos"] = 500 + step * np.arange(n) + rng.integers(0, jitter, n)
        r["aepos"] = r["abpos"] + ln; r["bepos"] = r["bbpos"] + ln + rng.integers(-20, 20, n)
        r["diffs"] = ln // 10
        return r
    big = colinear(90, 0, 1, 700, 600, 60)                       # 90 > 63 records between one pair of reads
    # many overlapping alternatives: every record chains with many later ones -> alternate chains repeat prefixes
    alt = colinear(16, 0, 2, 300, 900, 5)
    alt["abpos"][8:] = alt["abpos"][:8] + 40; alt["bbpos"][8:] = alt["bbpos"][:8] + 55
    alt["aepos"][8:] = alt["aepos"][:8] + 40; alt["bepos"][8:] = alt["bepos"][:8] + 55
    small = colinear(5, 1, 2, 2000, 1500, 30)
    rec = np.concatenate([big, alt, small])
    order = np.lexsort((rec["abpos"], rec["bread"], rec["aread"]))
    rec = rec[order]
    for opts in (dict(), dict(min_rel_score=0.05, max_rel_overlap=0.9), dict(min_rel_score=0.3, max_indel=200)):
        las = _synthetic_las(tmp_path, rec.copy())
        before = las.rec.copy()
        las.chainLocalAlignments(**opts)
        o = chaining.ChainingOptions(max_indel=opts.get("max_indel", 1000), max_rel_overlap=opts.get("max_rel_overlap", 0.3),
                                     min_rel_score=opts.get("min_rel_score", 1.0), min_score=126)
        src, fl = chaining.chain_local_alignments(before, o)
        assert len(las) == len(src) and np.array_equal(las.rec["flags"], fl)
        for f in ("aread", "bread", "abpos", "aepos", "bbpos", "bepos"):
            assert np.array_equal(las.rec[f], before[src][f]), f
    assert len(src) > 0


def test_transposed_las_matches_oracle():
    """dn_las_transpose (damapper -C's second file, dazzler.d:5931-5936) == oracle/pile_oracle.c orc_transpose: records,
    mirrored coordinates, re-laid trace points; LAsort order of the new (read, contig) numbering; and the double
    transposition returns to the original coordinates."""
    from dentist_b200 import dazzler
    from oracle import oracle
    sc = synth.make_scaffolds(1, 150000, 171, n_repeats=1, repeat_copies=3)
    ref, _ = synth.contigs_from(sc, synth.make_gaps(sc, 2, 172))
    reads, _ = synth.simulate_reads(sc, 4, 8000, 2500, 0.13, 173)
    ga, gb = dazzler.Block(ref.off, ref.bases), dazzler.Block(reads.off, reads.bases)
    for ts in (100, 126):
        las = dazzler.align(ga, gb, tspace=ts, minlen=500)
        assert len(las) > 60 and (las.rec["flags"] & 1).any()
        t = las.transpose(ga, gb)
        orec, otoff, otr = oracle.transpose(ref.off, ref.bases, reads.off, reads.bases, las.rec, las.toff, las.trace, ts)
        order = np.lexsort((orec["diffs"], orec["bepos"], orec["bbpos"], orec["aepos"], orec["abpos"], orec["flags"] & 1, orec["bread"], orec["aread"]))
        assert len(t) == len(orec)
        for f in ("aread", "bread", "abpos", "aepos", "bbpos", "bepos", "diffs", "tlen"):
            assert np.array_equal(t.rec[f], orec[f][order]), f
        assert np.array_equal(t.rec["flags"], orec["flags"][order] & 1)
        for i, tr in zip(order, t.traces()):
            assert np.array_equal(tr.reshape(-1), otr[otoff[i]:otoff[i] + orec["tlen"][i]])
        back = t.transpose(gb, ga)
        for f in ("aread", "bread", "abpos", "aepos", "bbpos", "bepos"):
            assert np.array_equal(back.rec[f], las.rec[f]), f
    empty = dazzler.align(ga, gb, tspace=100, minlen=10 ** 6).transpose(ga, gb)
    assert len(empty) == 0


def test_bridging_matches_oracle():
    """`daligner -B` (dazzler.d:5823-5824) as dn_las_bridge: reads with a stretch of junk break their alignments in two; the
    bridged LAS equals the oracle's (records, diffs, every trace point) and keeps the trace invariants."""
    from dentist_b200 import dazzler
    from oracle import las as olas, oracle
    for ts, seed in ((126, 301), (100, 302), (40, 303)):
        blk = _pile(seed, cov=8)
        rng = np.random.default_rng(seed)
        bases = blk.bases.copy()
        for r in range(0, blk.nreads, 2):                              # junk in the middle of every second read
            L = int(blk.off[r + 1] - blk.off[r])
            if L > 3000:
                s = int(blk.off[r]) + L // 2
                n = int(rng.integers(20, 110))
                bases[s:s + n] = rng.integers(0, 4, n)
        blk = synth.Block(blk.off, bases)
        g, las = _gpu_las(blk, ts, 500)
        rec0, toff0, tr0 = las.rec.copy(), las.toff.copy(), las.trace.copy()
        exp, etoff, etr, enb = oracle.bridge(blk.off, blk.bases, blk.off, blk.bases, rec0, toff0, tr0, ts)
        nb = las.bridge(g, g)
        assert nb == enb > 10 and len(las) == len(exp) == len(rec0) - nb
        for f in ("aread", "bread", "abpos", "aepos", "bbpos", "bepos", "diffs", "tlen", "flags"):
            assert np.array_equal(las.rec[f], exp[f]), f
        got_tr = np.concatenate([las.trace[int(o):int(o) + int(t)] for o, t in zip(las.toff, las.rec["tlen"])])
        assert np.array_equal(got_tr, etr)
        for x, o in zip(las.rec, las.toff):
            t = las.trace[int(o):int(o) + int(x["tlen"])].reshape(-1, 2)
            assert len(t) == olas.num_tiles(int(x["abpos"]), int(x["aepos"]), ts)
            assert int(t[:, 1].sum()) == x["bepos"] - x["bbpos"] and int(t[:, 0].sum()) == x["diffs"]
        # bridging an LAS without bridgeable neighbours changes nothing
        again = las.bridge(g, g)
        rest = oracle.bridge(blk.off, blk.bases, blk.off, blk.bases, las.rec.copy(), las.toff.copy(), las.trace.copy(), ts)[3]
        assert again == rest
        g.free()
