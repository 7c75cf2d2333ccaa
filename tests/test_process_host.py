"""Host logic of the batched `dentist process` driver (dentist_b200/process.py) on the reference's own pile-up and
insertion test data (common/binio/_testdata/*.d) -- no GPU."""
import json
import os

import numpy as np
import pytest

from dentist_b200 import pileups, process

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "binio_kat.json")))


def _seeded(d):
    return dict(id=d["id"], contigA=tuple(d["contigA"]), contigB=tuple(d["contigB"]), flags=d["flags"], tpd=d["tpd"] or 100, seed=d["seed"],
                las=[dict(ab=l["ab"], ae=l["ae"], bb=l["bb"], be=l["be"], diffs=l["diffs"], trace=np.array(l["trace"], np.uint16).reshape(-1, 2))
                     for l in d["las"]])


class _Seqs:
    """random sequences of given lengths, 1-based ids -> read(i) 0-based"""

    def __init__(self, lengths, seed):
        rng = np.random.default_rng(seed)
        self.seq = {i: rng.integers(0, 4, n).astype(np.uint8) for i, n in lengths.items()}

    def read(self, i):
        return self.seq[i + 1]


def test_common_trace_point_rules():
    ctp = pileups.common_trace_point
    assert ctp([(250, 1000), (130, 1000)], "back", 100, 1000) == 300            # first grid point inside every alignment
    assert ctp([(0, 730), (0, 655)], "front", 100, 5000) == 600                 # last grid point for a front seed
    assert ctp([(0, 730), (0, 655)], "front", 100, 5000, repeat_mask=[(550, 700)]) == 500     # avoids the repeat mask ...
    assert ctp([(0, 730), (0, 655)], "front", 100, 5000, repeat_mask=[(0, 5000)]) == 600      # ... unless nothing is left
    assert ctp([[(0, 210), (390, 900)], (0, 900)], "back", 100, 900) == 0        # chain with a hole: region = its local alignments
    assert ctp([[(10, 210), (390, 900)], (380, 900)], "back", 100, 900) == 400
    assert ctp([(0, 300), (400, 900)], "back", 100, 900) == -1                   # no common region
    assert ctp([(850, 955)], "back", 100, 955) == 900
    assert ctp([(910, 955)], "back", 100, 955) == 955                            # only the contig end is left as a candidate


def test_crop_reference_pileups():
    piles = [[[_seeded(s) for s in ra] for ra in p] for p in KAT["pileupdb"]["data"]]
    contigs = {sa["contigA"][0]: sa["contigA"][1] for p in piles for ra in p for sa in ra}
    rlens = {sa["contigB"][0]: sa["contigB"][1] for p in piles for ra in p for sa in ra}
    ref, reads = _Seqs(contigs, 1), _Seqs(rlens, 2)
    for pile in piles:
        crop = process.crop_pileup(pile, ref, reads, {}, 500)
        assert [c for c, _ in crop["ref_positions"]] == sorted({sa["contigA"][0] for ra in pile for sa in ra})
        for (cid, pos), seed in zip(crop["ref_positions"], crop["seeds"]):
            assert pos % 100 == 0 or pos == contigs[cid]
            for ra in pile:
                for sa in ra:
                    if sa["contigA"][0] == cid:
                        assert sa["seed"] == seed and any(la["ab"] <= pos <= la["ae"] for la in sa["las"])
        for ra, seq, ok in zip(pile, crop["sequences"], crop["allowed"]):
            assert 0 < len(seq) <= ra[0]["contigB"][1] + 2 * 500
            assert ok == (len(ra) == len(crop["ref_positions"]))
    # second pile-up: a gap between contigs 1 and 2; the spanning read keeps exactly what lies between both crop points
    pile = piles[1]
    crop = process.crop_pileup(pile, ref, reads, {}, 500)
    span = [i for i, ra in enumerate(pile) if len(ra) == 2][0]
    sa_l, sa_r = pile[span]
    b0, e0 = process.chain_cropping_slice(sa_l, dict(crop["ref_positions"])[1])
    b1, e1 = process.chain_cropping_slice(sa_r, dict(crop["ref_positions"])[2])
    assert len(crop["sequences"][span]) == min(e0, e1) - max(b0, b1)
    with pytest.raises(process.PileUpSkipped):
        process.crop_pileup([[dict(pile[0][0], las=[dict(pile[0][0]["las"][0], ab=8290, ae=8300)])], pile[1]], ref, reads, {}, 500)


def test_support_patches_extend_short_anchors():
    ref, reads = _Seqs({1: 3000}, 3), _Seqs({1: 4000, 2: 4000}, 4)
    tr = np.array([[0, 100]] * 3, np.uint16)
    mk = lambda rid, comp: dict(id=rid, contigA=(1, 3000), contigB=(rid, 4000), flags=1 if comp else 0, tpd=100, seed="front",
                                las=[dict(ab=0, ae=300, bb=3700, be=4000, diffs=0, trace=tr)])
    crop = process.crop_pileup([[mk(1, False)], [mk(2, True)]], ref, reads, {}, 500)
    assert crop["ref_positions"] == [(1, 200)] and crop["seeds"] == ["front"]
    patch = ref.read(0)[200:500]                                                     # contig[pos .. minAnchorLength)
    fwd, rev = crop["sequences"]
    assert np.array_equal(fwd, np.concatenate([reads.read(0)[:3900], patch]))         # read kept up to the crop point + patch behind it
    assert np.array_equal(rev, np.concatenate([(3 - patch)[::-1], reads.read(1)[100:]]))   # complement: mirrored slice, patch in front


def test_adjust_repeat_mask():
    mask = {1: [(0, 900)], 2: [(0, 100)]}
    out = process.adjust_repeat_mask(mask, {1: 5000, 2: 5000}, [(1, 1200), (2, 4000)], ["front", "back"], 500)
    assert 1 not in out and out[2] == [(0, 100)]                                     # 300 unmasked anchor bases < 500 -> mask dropped


def test_insertion_alignment_on_reference_insertions():
    n = 0
    for ins in KAT["insertiondb"]["data"]:
        ov = [_seeded(s) for s in ins["overlaps"]]
        if not ov:
            continue
        ref_read = [dict(contigA=o["contigA"], contigB=(7, 1), flags=o["flags"], seed=o["seed"]) for o in ov]
        positions = [(o["contigA"][0], 0) for o in ov]
        chains = [dict(o, seed="front", flags=o["flags"]) for o in ov]
        got = process.insertion_alignment(chains, ref_read, positions, 100)
        assert [(g["contigA"], g["seed"], g["las"][0]["ab"]) for g in got] == [(o["contigA"], o["seed"], o["las"][0]["ab"]) for o in ov]
        start, end = process.make_join(ref_read)
        assert (list(start), list(end)) == (ins["start"], ins["end"])
        # a consensus that maps to the other strand than the reference read did is rejected
        with pytest.raises(process.PileUpSkipped):
            shifted = dict(chains[0], flags=chains[0]["flags"] ^ 1)
            process.insertion_alignment([shifted] + [dict(c) for c in chains[1:]], ref_read, positions, 100)
        n += 1
    assert n >= 10


def test_chain_order_follows_the_reference_kat():
    """AlignmentChain.opCmp (base.d:766-777) on the reference's two ordered lists (base.d:780-857): the key used for
    splitAlignmentsByContigA / filterContainedAlignmentChains here and for the accepted chains in oracle/chaining.py."""
    import itertools
    from oracle import chaining
    lists = json.load(open(os.path.join(HERE, "golden", "chain_order_kat.json")))
    assert [len(x) for x in lists] == [4, 7]
    for chains in lists:
        sas = [dict(contigA=(c["contigA"], 10), contigB=(c["contigB"], 10),
                    las=[dict(ab=l[0], ae=l[1], bb=l[2], be=l[3]) for l in c["las"]]) for c in chains]
        keys = [process._chain_key(sa) for sa in sas]
        assert keys == sorted(keys) and len(set(keys)) == len(keys)                     # strictly ascending as listed
        for perm in itertools.islice(itertools.permutations(range(len(sas))), 50):
            assert [i for i in sorted(perm, key=lambda i: keys[i])] == list(range(len(sas)))
        # the chaining oracle orders accepted chains of one (A, B) group by the same coordinates
        if len({(c["contigA"], c["contigB"]) for c in chains}) == 1:
            ck = [(s["las"][0]["ab"], s["las"][0]["bb"], s["las"][-1]["ae"], s["las"][-1]["be"]) for s in sas]
            assert ck == sorted(ck)


def test_read_alignment_typing_follows_the_reference_kat():
    """SeededAlignment.from + ReadAlignment.{isInOrder, isValid, type, isExtension, isFrontExtension, isBackExtension, isGap,
    isParallel, isAntiParallel} on the reference's eight cases (base.d:2324-2676)."""
    kat = json.load(open(os.path.join(HERE, "golden", "read_alignment_kat.json")))
    assert len(kat["cases"]) == 8
    for name, chains in kat["cases"].items():
        ra = []
        for c in chains:
            chain = dict(id=c["id"], contigA=tuple(c["contigA"]), contigB=tuple(c["contigB"]), flags=c["flags"], tpd=100, las=c["las"])
            ra.append(process.seeds_from(chain)[0])                                   # `.front` of the filtered range
        e = kat["expect"][name]
        got = [process.is_in_order(ra), process.is_valid(ra), None, process.is_extension(ra),
               len(ra) == 1 and ra[0]["seed"] == "front", len(ra) == 1 and ra[0]["seed"] == "back",
               process._is_gap(ra), process._is_parallel(ra), process.is_anti_parallel(ra)]
        for i, ch in enumerate(e):
            if i == 2:
                assert {"F": "front", "B": "back", "G": "gap"}[ch] == process._type(ra), name
            else:
                assert got[i] == (ch == "+"), (name, i)
        start, end = process.make_join(ra)                                            # makeJoin, base.d:2680-2721
        if process._is_gap(ra):
            part = lambda s: "begin" if s == "front" else "end"
            assert (start, end) == ((ra[0]["contigA"][0], part(ra[0]["seed"])), (ra[1]["contigA"][0], part(ra[1]["seed"])))
        else:
            assert start[0] == end[0] == ra[0]["contigA"][0]


def test_collect_read_alignments_follows_the_reference_kat():
    """collectReadAlignments (collectPileUps/pileups.d:821-888) on the reference's five cases (:890-1098)."""
    cases = json.load(open(os.path.join(HERE, "golden", "collect_read_alignments_kat.json")))
    assert len(cases) == 5
    for c in cases:
        chains = [dict(id=x["id"], contigA=tuple(x["contigA"]), contigB=tuple(x["contigB"]), flags=x["flags"], tpd=100, las=x["las"])
                  for x in c["chains"]]
        ras, reason = process.collect_read_alignments(chains)
        assert reason is None
        got = [[[chains.index({k: v for k, v in sa.items() if k != "seed"}), sa["seed"]] for sa in ra] for ra in ras]
        assert got == c["expect"]
    # two chains that use the same stretch of the read: the read is not touched at all
    a = dict(id=0, contigA=(1, 20), contigB=(1, 60), flags=0, tpd=100, las=[dict(ab=10, ae=20, bb=0, be=10, diffs=0)])
    b = dict(id=1, contigA=(2, 20), contigB=(1, 60), flags=0, tpd=100, las=[dict(ab=0, ae=20, bb=5, be=25, diffs=0)])
    assert process.collect_read_alignments([a, b]) == ([], "alignments overlap on read")
    assert process.collect_read_alignments([]) == ([], "empty input")


def test_singular_pile_ups_use_the_single_read():
    """--allow-single-reads (shouldProcessSingularPileUp, package.d:292-305, 376-379): the read itself is the insertion."""
    kat = json.load(open(os.path.join(HERE, "golden", "read_alignment_kat.json")))
    piles, names = [], []
    for name, chains in kat["cases"].items():
        ra = [process.seeds_from(dict(id=c["id"], contigA=tuple(c["contigA"]), contigB=tuple(c["contigB"]), flags=c["flags"], tpd=100, las=c["las"]))[0]
              for c in chains]
        piles.append([ra]); names.append(name)
    # an invalid read alignment (two alignments to ONE contig: neither extension nor gap, base.d:2190-2193)
    bad = [dict(sa) for sa in next(ra for (ra,) in piles if len(ra) == 2)]
    bad[1]["contigA"] = bad[0]["contigA"]
    assert not process.is_valid(bad)
    piles.append([bad])
    reads = _Seqs({ra[0]["contigB"][0]: ra[0]["contigB"][1] for (ra,) in piles}, 5)
    # without the flag every one-read pile-up falls to minReadsPerPileUp (shouldSkipSmallPileUp, :381-397)
    ins, skipped = process.process_pileup_db(piles, reads, None, min_reads_per_pileup=3)
    assert ins == [] and all(skipped[p] == "minReadsPerPileUp" for p in range(len(piles)))
    ins, skipped = process.process_pileup_db(piles, reads, None, min_reads_per_pileup=3, allow_single_reads=True)
    valid = [p for p, (ra,) in enumerate(piles) if process.is_valid(ra)]
    assert 0 < len(valid) < len(piles)
    assert sorted(i["pile_up"] for i in ins) == valid
    assert all(skipped[p] == "consensus alignment is invalid" for p in range(len(piles)) if p not in valid)
    for i in ins:
        (ra,) = piles[i["pile_up"]]
        assert (i["start"], i["end"]) == process.make_join(ra)
        assert i["read_ids"] == [ra[0]["contigB"][0]] and len(i["overlaps"]) == len(ra)
        assert np.array_equal(i["sequence"], reads.read(ra[0]["contigB"][0] - 1)) and len(i["sequence"]) == i["overlaps"][0]["contigB"][1]
    part = {"pre": 0, "begin": 1, "end": 2, "post": 3}
    keys = [(i["start"][0], part[i["start"][1]], i["end"][0], part[i["end"][1]]) for i in ins]
    assert keys == sorted(keys)                                                       # insertions.sort(), package.d:156
