"""`dentist process` for a whole pile-up DB in one batch (commands/processPileUps/package.d:100-160, 283-374):
pile-ups in (binio.read_pileup_db), insertions out (binio.write_insertion_db).  The per-pile-up host logic of the
reference -- cropping at a common trace point (cropper.d), allowedReferenceReadIds, repeat-mask adjustment,
post-consensus checks, getInsertionAlignment, makeInsertion -- is restated here; every alignment / QV / consensus
step runs once for ALL pile-ups on the device through pileups.process_pileups.

Sequences are engine base codes (a0 c1 g2 t3); contig and read ids are 1-based as in DENTIST."""
import numpy as np

from . import dazzler, pileups

FLAG_COMPLEMENT, FLAG_DISABLED = 1, 2


class PileUpSkipped(Exception):
    """What processPileUp logs as `pileUpSkipped` (package.d:351-373)."""


def _first(sa):
    return sa["las"][0]


def _last(sa):
    return sa["las"][-1]


def _chain_key(sa):                                   # AlignmentChain.opCmp, base.d:766-777
    return (sa["contigA"][0], sa["contigB"][0], _first(sa)["ab"], _first(sa)["bb"], _last(sa)["ae"], _last(sa)["be"])


def _revcomp(x):
    return (3 - np.asarray(x, np.uint8))[::-1]


def chain_cropping_slice(sa, pos):
    """getCroppingSlice (cropper.d:503-550) for a chain: the first local alignment covering `pos` translates it
    (AlignmentChain.translateTracePoint, base.d:866-879; coveringLocalAlignmentIndex :1136-1154)."""
    for la in sa["las"]:
        if la["ab"] <= pos <= la["ae"]:
            return pileups.get_cropping_slice(la["ab"], la["ae"], la["bb"], sa["tpd"], la["trace"], bool(sa["flags"] & FLAG_COMPLEMENT),
                                              sa["seed"], sa["contigB"][1], pos)
    raise PileUpSkipped("cannot translate coordinate due to lack of alignment coverage")


def crop_pileup(pile, ref, reads, repeat_mask, min_anchor_length):
    """cropPileUp (cropper.d:97-380).  Returns dict(ref_positions=[(contig id, pos)], seeds, sequences (one per read
    alignment: support patch + read slice + support patch), allowed (bool per read alignment))."""
    by_contig = {}
    for sa in sorted((sa for ra in pile for sa in ra), key=_chain_key):           # splitAlignmentsByContigA :426-443
        by_contig.setdefault(sa["contigA"][0], []).append(sa)
    ref_positions, seeds = [], []
    for cid in sorted(by_contig):
        grp = by_contig[cid]
        pos = pileups.common_trace_point([[(la["ab"], la["ae"]) for la in sa["las"]] for sa in grp], grp[0]["seed"], grp[0]["tpd"],
                                         grp[0]["contigA"][1], repeat_mask.get(cid, ()))
        if pos < 0:
            raise PileUpSkipped("could not find a common trace point")               # cropper.d:176-206
        ref_positions.append((cid, pos)); seeds.append(grp[0]["seed"])
    patches = {}                                                                     # fetchSupportPatches :208-243
    for (cid, pos), seed in zip(ref_positions, seeds):
        contig = ref.read(cid - 1)
        b = e = 0
        if seed == "front" and pos < min_anchor_length:
            b, e = pos, min_anchor_length
        elif seed == "back" and len(contig) - pos < min_anchor_length:
            b, e = len(contig) - min_anchor_length, pos
        patches[cid] = contig[b:e] if b < e else contig[:0]      # a reversed slice is empty in spirit; D would throw a RangeError
    pos_of = dict(ref_positions)
    seqs, allowed = [], []
    for ra in pile:
        b, e = 0, ra[0]["contigB"][1]
        sides = []
        for sa in ra:
            sb, se = chain_cropping_slice(sa, pos_of[sa["contigA"][0]])
            b, e = max(b, sb), min(e, se)                                            # fold!"a & b" :309-316
            comp = bool(sa["flags"] & FLAG_COMPLEMENT)
            read_seed = 1 if (sa["seed"] == "front") ^ comp else 0                   # getSingleReadPatch :334-349 (front 0 < back 1)
            patch = patches[sa["contigA"][0]]
            sides.append((read_seed, tuple(int(v) for v in (_revcomp(patch) if comp else patch))))
        if b >= e:
            raise PileUpSkipped("invalid/empty read cropping slice")
        sides.sort()
        if len(sides) == 2:
            pre, post = sides[0][1], sides[1][1]
        elif sides[0][0] == 0:
            pre, post = sides[0][1], ()
        else:
            pre, post = (), sides[0][1]
        read = reads.read(ra[0]["contigB"][0] - 1)
        seqs.append(np.concatenate([np.array(pre, np.uint8), read[b:e], np.array(post, np.uint8)]).astype(np.uint8))
        allowed.append(len(ra) == len(ref_positions) and [sa["contigA"][0] for sa in ra] == [c for c, _ in ref_positions])   # package.d:456-468
    return dict(ref_positions=ref_positions, seeds=seeds, sequences=seqs, allowed=allowed)


def adjust_repeat_mask(mask, contigs, ref_positions, seeds, min_anchor_length):
    """adjustRepeatMaskToMakeMappingPossible (package.d:428-454): drop a contig's mask when fewer than
    minAnchorLength unmasked bases remain in the part of the contig that the cropped reads keep."""
    out = dict(mask)
    for (cid, pos), seed in zip(ref_positions, seeds):
        clen = contigs[cid]
        p = max(pos, min_anchor_length) if seed == "front" else min(pos, clen - min_anchor_length)
        iv = (0, p) if seed == "front" else (p, clen)
        free = pileups._subtract([iv] if iv[0] < iv[1] else [], pileups._normalise(out.get(cid, ())))
        if sum(e - b for b, e in free) < min_anchor_length:
            out.pop(cid, None)
    return out


def seeds_from(chain):
    """SeededAlignment.from (base.d:2002-2014): a chain seeds the contig's FRONT when the read sticks out beyond the
    contig begin (first.contigB.begin > first.contigA.begin, isFrontExtension :2030-2036), its BACK when the read sticks
    out beyond the contig end (:2042-2048); both are possible.  Returns the seeded alignment dicts in (front, back) order."""
    first, last = chain["las"][0], chain["las"][-1]
    out = []
    if first["bb"] > first["ab"]:
        out.append(dict(chain, seed="front"))
    if chain["contigB"][1] - last["be"] > chain["contigA"][1] - last["ae"]:
        out.append(dict(chain, seed="back"))
    return out


def collect_read_alignments(chains):
    """collectReadAlignments (commands/collectPileUps/pileups.d:821-888): all alignment chains of ONE read -> its read
    alignments.  Every chain is seeded (seeds_from), the seeded alignments are ordered along the read (forward read
    coordinates, then seed), no region of the read may be used by two different chains, and consecutive seeded
    alignments pair up into gaps -- starting with a lone extension when the first one does not begin at the read's begin.
    Returns (read alignments, reason): an empty list comes with the reference's reason string."""
    def begin_b(sa):
        return sa["contigB"][1] - sa["las"][-1]["be"] if sa["flags"] & FLAG_COMPLEMENT else sa["las"][0]["bb"]

    def end_b(sa):
        return sa["contigB"][1] - sa["las"][0]["bb"] if sa["flags"] & FLAG_COMPLEMENT else sa["las"][-1]["be"]

    def seed_b(sa):
        v = 0 if sa["seed"] == "front" else 1
        return -v if sa["flags"] & FLAG_COMPLEMENT else v

    seeded = [sa for c in chains for sa in seeds_from(c)]
    seeded.sort(key=lambda sa: (begin_b(sa), end_b(sa), seed_b(sa)))
    if not seeded:
        return [], "empty input"
    for x, y in zip(seeded, seeded[1:]):
        same_chain = x["id"] == y["id"] and x["contigA"] == y["contigA"] and x["seed"] != y["seed"]
        if end_b(x) > begin_b(y) and not same_chain:
            return [], "alignments overlap on read"
    start = 1 if begin_b(seeded[0]) > 0 else 0
    ras = ([seeded[0:1]] if start else []) + [seeded[i:i + 2] for i in range(start, len(seeded), 2)]
    if any(not is_valid(ra) for ra in ras):
        return [], "invalid read alignment"
    return ras, None


def is_extension(ra):                                                                # base.d:2226-2230
    return len(ra) == 1


def is_in_order(ra):                                                                 # base.d:2170-2173
    return not _is_gap(ra) or ra[0]["contigA"][0] < ra[1]["contigA"][0]


def is_valid(ra):                                                                    # base.d:2190-2193
    return is_extension(ra) != _is_gap(ra)


def is_anti_parallel(ra):                                                            # base.d:2314-2322
    return _is_gap(ra) and ra[0]["seed"] == ra[1]["seed"] and (ra[0]["flags"] & 1) != (ra[1]["flags"] & 1)


def _is_gap(ra):
    return len(ra) == 2 and ra[0]["contigA"][0] != ra[1]["contigA"][0] and ra[0]["contigB"][0] == ra[1]["contigB"][0]


def _is_parallel(ra):                                                                # base.d:2300-2306
    return _is_gap(ra) and ra[0]["seed"] != ra[1]["seed"] and (ra[0]["flags"] & 1) == (ra[1]["flags"] & 1)


def _type(ra):                                                                       # base.d:2201-2217
    return "gap" if _is_gap(ra) else ra[0]["seed"]


def make_join(ra):
    """makeJoin!Insertion(referenceRead) (base.d:2680-2721) -> (start node, end node)."""
    part = lambda seed: "begin" if seed == "front" else "end"
    if _is_gap(ra):
        return (ra[0]["contigA"][0], part(ra[0]["seed"])), (ra[1]["contigA"][0], part(ra[1]["seed"]))
    c = ra[0]["contigA"][0]
    return ((c, "pre"), (c, "begin")) if ra[0]["seed"] == "front" else ((c, "end"), (c, "post"))


def insertion_alignment(chains, ref_read, ref_positions, allowance):
    """alignConsensusToFlankingContigs' post-processing + getInsertionAlignment (package.d:669-760).
    chains: the consensus-vs-flanks alignments of one pile-up as seeded-alignment dicts (global contig ids, seed unset)."""
    chains = sorted(chains, key=_chain_key)
    iv_a = lambda c: (_first(c)["ab"], _last(c)["ae"])
    iv_b = lambda c: (_first(c)["bb"], _last(c)["be"])
    for i, c1 in enumerate(chains):                                                  # filterContainedAlignmentChains, filter.d:178-211
        if c1["flags"] & FLAG_DISABLED:
            continue
        for c2 in chains[i + 1:]:
            if not (c1["contigA"][0] == c2["contigA"][0] and iv_a(c1)[0] <= iv_a(c2)[0] and iv_a(c2)[1] <= iv_a(c1)[1]):
                break
            if (c1["flags"] & 1) == (c2["flags"] & 1) and c1["contigB"][0] == c2["contigB"][0] and iv_b(c1)[0] <= iv_b(c2)[0] and iv_b(c2)[1] <= iv_b(c1)[1]:
                c2["flags"] |= FLAG_DISABLED
    by_contig = {sa["contigA"][0]: sa for sa in ref_read}
    for c in chains:
        proper = ((_first(c)["ab"] <= allowance or _first(c)["bb"] <= allowance) and
                  (_last(c)["ae"] + allowance >= c["contigA"][1] or _last(c)["be"] + allowance >= c["contigB"][1]))   # isProper, base.d:537-557
        rr = by_contig.get(c["contigA"][0])
        if not proper or (rr is not None and (c["flags"] & 1) != (rr["flags"] & 1)):
            c["flags"] |= FLAG_DISABLED
    if all(c["flags"] & FLAG_DISABLED for c in chains):
        raise PileUpSkipped("consensus does not align to flanking contigs")
    out = [None] * len(ref_read)
    for cid, _ in ref_positions:
        idx = [i for i, sa in enumerate(ref_read) if sa["contigA"][0] == cid][0]
        seed = ref_read[idx]["seed"]
        if seed == "front":
            good = lambda c: _first(c)["ab"] <= allowance and abs(_last(c)["be"] - c["contigB"][1]) <= allowance
        else:
            good = lambda c: abs(_last(c)["ae"] - c["contigA"][1]) <= allowance and _first(c)["bb"] <= allowance
        cand = [c for c in chains if not (c["flags"] & FLAG_DISABLED) and c["contigA"][0] == cid and good(c)]
        if not cand:
            raise PileUpSkipped("consensus does not align to flanking contig %d" % cid)
        if len(cand) > 1:
            raise PileUpSkipped("consensus ambiguously aligns to flanking contig %d" % cid)
        out[idx] = dict(cand[0], seed=seed)
    if _type(out) != _type(ref_read) or _is_parallel(out) != _is_parallel(ref_read):
        raise PileUpSkipped("consensus alignment has an unexpected type")
    return out


class _HostBlock:
    def __init__(self, seqs):
        self.off = np.zeros(len(seqs) + 1, np.int64)
        self.off[1:] = np.cumsum([len(x) for x in seqs])
        self.bases = np.concatenate(seqs).astype(np.uint8) if seqs else np.zeros(0, np.uint8)


def process_pileup_db(pile_ups, reads, ref, repeat_mask=None, min_reads_per_pileup=3, min_anchor_length=500,
                      proper_alignment_allowance=100, max_alignment_error=0.3, allow_single_reads=False):
    """PileUpsProcessor.run (package.d:100-160) for all pile-ups at once.
    pile_ups: binio.read_pileup_db() nesting; reads / ref: host blocks with .read(i) (0-based) -> base codes;
    repeat_mask: {contig id: [(begin, end), ...]}.  Returns (insertions sorted like `insertions.sort()`, skipped =
    {pile-up index: reason}).
    Singular pile-ups (`--allow-single-reads`, shouldProcessSingularPileUp package.d:376-379): the single read stands in for
    the pile-up -- postConsensusAlignment = its own alignments (:294), insertion sequence = the read as stored
    (getInsertionSequence :764-772), referenceRead = pileUp[0] (referenceReadIdx keeps its initial 0, :231, :587-590).  The
    reference then builds the insertion alignment from croppingPositions, which crop() never filled on this path
    (:704-747), i.e. from default-initialised SeededAlignments; what is returned here is the evidently intended value, the
    read's own seeded alignments, checked with the same validity / type rule (:749-760).  No device work is involved."""
    repeat_mask = dict(repeat_mask or {})
    skipped, crops, live, singular = {}, {}, [], []
    for p, pile in enumerate(pile_ups):
        if allow_single_reads and len(pile) == 1:                                   # shouldProcessSingularPileUp, package.d:292-305
            singular.append(p); continue
        if len(pile) < min_reads_per_pileup:                                        # shouldSkipSmallPileUp, package.d:381-397
            skipped[p] = "minReadsPerPileUp"; continue
        contigs = {sa["contigA"][0]: sa["contigA"][1] for ra in pile for sa in ra}
        local_mask = {c: repeat_mask[c] for c in contigs if c in repeat_mask}       # reduceRepeatMaskToFlankingContigs :399-409
        try:
            crops[p] = crop_pileup(pile, ref, reads, local_mask, min_anchor_length)
        except PileUpSkipped as e:
            skipped[p] = str(e); continue
        crops[p]["mask"] = adjust_repeat_mask(local_mask, contigs, crops[p]["ref_positions"], crops[p]["seeds"], min_anchor_length)
        crops[p]["contigs"] = contigs
        live.append(p)
    insertions = []
    for p in singular:
        ra = pile_ups[p][0]
        if not is_valid(ra):                                                         # insertionAlignment.isValid, package.d:749-753
            skipped[p] = "consensus alignment is invalid"; continue
        start, end = make_join(ra)
        rid = ra[0]["contigB"][0]
        insertions.append(dict(start=start, end=end, sequence=np.asarray(reads.read(rid - 1), np.uint8), contig_length=0,
                               overlaps=[dict(sa) for sa in ra], read_ids=[rid], pile_up=p))
    if not live:
        return _sorted_insertions(insertions), skipped
    # ONE call for all pile-ups: pile alignment, filters, chaining, QVs, reference read, consensus (with retry) and the
    # consensus-vs-flanks alignment (-mdust -mrep) run grouped on the device (dn_process_pileups, package.d:303-341)
    piles_in = []
    for p in live:
        cids = [cid for cid, _ in crops[p]["ref_positions"]]
        piles_in.append(dict(reads=crops[p]["sequences"], allowed=crops[p]["allowed"], flanks=[cid - 1 for cid in cids],
                             mask=[pileups._normalise(crops[p]["mask"].get(cid, ())) for cid in cids]))
    ref_block = ref if isinstance(ref, dazzler.Block) else dazzler.Block(ref.off, ref.bases)
    res = dazzler.processPileUps(ref_block, piles_in, max_alignment_error=max_alignment_error, min_anchor_length=min_anchor_length,
                                 proper_alignment_allowance=proper_alignment_allowance)
    if ref_block is not ref:
        ref_block.free()
    for g, p in enumerate(live):
        pile, crop, out = pile_ups[p], crops[p], res[g]
        try:
            if out["status"] != 0:
                raise PileUpSkipped(out["reason"])
            ref_read = pile[out["reference_read"]]
            cons = out["consensus"]
            fl = out["flank_las"]
            rec, traces = fl.rec, fl.traces()
            chains = []
            for i in range(len(rec)):
                cid = crop["ref_positions"][int(rec[i]["aread"])][0]
                chains.append(dict(id=len(chains), contigA=(cid, crop["contigs"][cid]), contigB=(1, len(cons)), flags=int(rec[i]["flags"]) & 1,
                                   tpd=pileups.TSPACE, seed="front",
                                   las=[dict(ab=int(rec[i]["abpos"]), ae=int(rec[i]["aepos"]), bb=int(rec[i]["bbpos"]), be=int(rec[i]["bepos"]),
                                             diffs=int(rec[i]["diffs"]), trace=np.asarray(traces[i], np.uint16).reshape(-1, 2))]))
            if not chains:
                raise PileUpSkipped("consensus does not align to flanking contigs")
            overlaps = insertion_alignment(chains, ref_read, crop["ref_positions"], proper_alignment_allowance)
            start, end = make_join(ref_read)
            insertions.append(dict(start=start, end=end, sequence=cons, contig_length=0, overlaps=overlaps,
                                   read_ids=sorted(ra[0]["contigB"][0] for ra in pile), pile_up=p))      # makeInsertion :787-803
        except PileUpSkipped as e:
            skipped[p] = str(e)
    return _sorted_insertions(insertions), skipped


def _sorted_insertions(insertions):
    """`insertions.sort()` (package.d:156): by start node, then end node."""
    part = {"pre": 0, "begin": 1, "end": 2, "post": 3}
    insertions.sort(key=lambda i: (i["start"][0], part[i["start"][1]], i["end"][0], part[i["end"][1]]))
    return insertions
