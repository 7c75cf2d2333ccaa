"""Host-side mirror of `PileUpProcessor.processPileUp` (commands/processPileUps/package.d:283-374) for a
whole BATCH of pile-ups at once: the reference runs the steps below per pile-up, forking a tool at
each one (>= 12 fork/execs per pile-up); here every step is one call on a block that
holds the cropped reads of all pile-ups (pile id per read = `group`).

  computeQVs                   package.d:474-516   A1 + A5 + A9 + A13
  findReferenceReadCandidates  package.d:518-568   A10 (host logic, mirrored here in numpy)
  computeConsensus             package.d:600-619   A11
  alignConsensusToFlankingContigs  package.d:621-697   A12
  chainLocalAlignments         package.d:492       A6 (chaining.d restated on the device)
"""
import numpy as np

from . import dazzler

MAX_QV = 50            # DbRecord.maxQV, dazzler.d:2873
MIN_QV_COVERAGE = 4    # dazzler.d:3771
BAD_FRACTION = 0.08    # commandline.d:1101
TSPACE = 126           # forceLargeTracePointType, dazzler.d:154
MIN_ANCHOR = 500       # commandline.d:2036, 2894


def translate_trace_point(abpos, aepos, bbpos, tspace, trace, pos):
    """Trace.translateTracePoint!\"contigA\"(pos, RoundingMode.floor) (base.d:185-229): the read coordinate that the
    trace points map the last tile boundary at or below `pos` to.  Returns (contigA pos, contigB pos)."""
    if not (abpos <= pos <= aepos):
        raise ValueError("position outside the alignment")
    second = (abpos // tspace) * tspace + tspace
    if pos < second:
        idx = 0
    elif pos < aepos:
        idx = 1 + (pos - second) // tspace
    else:
        idx = len(trace)
    b = bbpos + int(sum(int(t[1]) for t in trace[:idx]))
    a = abpos if idx == 0 else ((abpos // tspace) * tspace + idx * tspace if idx < len(trace) else aepos)
    return a, b


def get_cropping_slice(abpos, aepos, bbpos, tspace, trace, complement, seed, read_len, crop_ref_pos):
    """getCroppingSlice (cropper.d:503-550): the part of the read that is kept when the pile-up is cropped at
    reference position `crop_ref_pos`; seed = 'front' | 'back' (AlignmentLocationSeed).  Forward read coordinates."""
    _, cpos = translate_trace_point(abpos, aepos, bbpos, tspace, trace, crop_ref_pos)
    if seed == "front":
        b, e = 0, cpos
    else:
        b, e = cpos, read_len
    if complement:
        b, e = read_len - e, read_len - b
    return b, e


def find_reference_read_candidates(qv, qoff, reads):
    """package.d:518-568 for one pile: reads ranked by (numBadQVs, meanQV, readId); returns read ids."""
    hist = np.zeros(MAX_QV, np.int64)
    per = [qv[qoff[r]:qoff[r + 1]] for r in reads]
    for q in per:
        hist += np.bincount(q[q < MAX_QV], minlength=MAX_QV)[:MAX_QV]
    bad_thres = int(BAD_FRACTION * int(hist.sum()))
    cum = np.cumsum(hist[::-1])
    idx = int(np.argmax(cum >= bad_thres)) if (cum >= bad_thres).any() else -1
    bad_qv = MAX_QV - 1 - idx
    scored = sorted((int((q >= bad_qv).sum()), float(q.mean()) if len(q) else 0.0, int(r)) for q, r in zip(per, reads))
    return [r for _, _, r in scored]


def process_pileups_batch(piles, ref=None, **params):
    """The batch path of `dentist process` through the C ABI (dn_process_pileups): see dazzler.processPileUps."""
    return dazzler.processPileUps(ref, piles, **params)


def process_pileups(reads, group, max_alignment_error=0.3, flanks=None, allowed=None, dust=False, candidates=False,
                    min_anchor_length=MIN_ANCHOR, proper_alignment_allowance=TSPACE, bridge=True):
    """The same path STEP BY STEP through the per-stage entry points (what a D host that keeps processPileUp's
    structure would call; tests assert it equals dn_process_pileups byte for byte).
    reads: synth.Block-like (off, bases) of all cropped reads; group: pile id per read.
    allowed: bool per read = member of allowedReferenceReadIds (package.d:456-468; default all); dust: DUST-mask the
    cropped reads first (package.d:476); candidates: also return the ranked reference read candidates of every pile.
    Returns dict(consensus=[codes per pile], reference_read=[read id per pile], las=Las, flank_las=Las|None)."""
    group = np.ascontiguousarray(group, np.int32)
    lens = np.diff(reads.off).astype(np.int32)
    npiles = int(group.max()) + 1 if len(group) else 0
    g = dazzler.Block(reads.off, reads.bases, group=group)
    if dust:
        g.maskDust()                                                     # dbdust(croppedDb) + -mdust, package.d:476-481
    # daligner -T<n> -B -s126 -l500 -e0.7 -mdust X X   (pileUpAlignmentOptions, commandline.d:2886-2902)
    las = dazzler.align(g, g, tspace=TSPACE, minlen=min_anchor_length, e=1.0 - max_alignment_error, self_block=1)
    if bridge:
        las.bridge(g, g, e=1.0 - max_alignment_error)                    # -B
    las.filterLocalAlignments(max_alignment_error)                      # package.d:483-485
    status = np.zeros(npiles, np.int32)
    status[np.bincount(group[las.rec["aread"]], minlength=npiles) == 0] = 1     # "empty pileup alignment", package.d:487-490 (per pile-up)
    if len(las):
        las.chainLocalAlignments(min_score=TSPACE)                       # package.d:492-496, chainingOptions commandline.d:2820-2830
    # coverage = |allowedReferenceReadIds|, raised to minQVCoverage for piles of >= 4 reads (package.d:498-501)
    psize = np.bincount(group, minlength=npiles)
    ok = np.ones(len(group), bool) if allowed is None else np.ascontiguousarray(allowed, bool)
    nallowed = np.bincount(group[ok], minlength=npiles)
    cov_pile = np.where((nallowed < MIN_QV_COVERAGE) & (psize >= MIN_QV_COVERAGE), MIN_QV_COVERAGE, nallowed)
    qv, qoff = dazzler.computeQVs(lens, las, cov_pile[group])            # package.d:498-503
    las.filterPileUpAlignments(lens, lens, proper_alignment_allowance)   # package.d:505-510 (Yes.forceFlat)
    las.forceFlat()
    status[(status == 0) & (np.bincount(group[las.rec["aread"]], minlength=npiles) == 0)] = 2   # "... after filtering", :512-515
    ranked = dazzler.findReferenceReadCandidates(qv, qoff, np.where(ok & (status[group] == 0), group, -1), npiles, BAD_FRACTION)   # package.d:518-568
    # selectReferenceRead / computeConsensus with retry on the next candidate (package.d:307-329)
    ref_reads = [-1] * npiles
    cons = [np.zeros(0, np.uint8)] * npiles
    pending = [p for p in range(npiles) if status[p] == 0]
    attempt = 0
    while pending:
        todo = [p for p in pending if attempt < len(ranked[p])]
        for p in pending:
            if attempt >= len(ranked[p]):
                status[p] = 3                                                # "no valid reference read found"
        got = dazzler.getConsensus(g, las, [int(ranked[p][attempt]) for p in todo]) if todo else []   # package.d:600-619
        nxt = []
        for p, c in zip(todo, got):
            if len(c):
                cons[p] = c; ref_reads[p] = int(ranked[p][attempt])
            else:
                nxt.append(p)                                                # "consensus could not be computed": next candidate
        pending = nxt; attempt += 1
    out = dict(consensus=cons, reference_read=ref_reads, las=las, qv=qv, qoff=qoff, flank_las=None, status=status)
    if candidates:
        out["candidates"] = ranked
    if flanks is not None:
        # daligner -A -B -s126 -l126 -e0.7 F C   (postConsensusAlignmentOptions, commandline.d:2918-2935)
        coff = np.zeros(len(cons) + 1, np.int64); coff[1:] = np.cumsum([len(c) for c in cons])
        cb = dazzler.Block(coff, np.concatenate(cons) if cons else np.zeros(0, np.uint8),
                           group=None if not getattr(flanks, "has_group", False) else np.arange(npiles, dtype=np.int32))
        fb = flanks if isinstance(flanks, dazzler.Block) else dazzler.Block(flanks.off, flanks.bases)
        out["flank_las"] = dazzler.align(fb, cb, tspace=TSPACE, minlen=TSPACE, e=0.7)
        if bridge:
            out["flank_las"].bridge(fb, cb, e=0.7)                       # -B
    g.free()
    return out


def _subtract(intervals, mask):
    out = []
    for b, e in intervals:
        cur = b
        for mb, me in mask:
            if me <= cur or mb >= e:
                continue
            if mb > cur:
                out.append((cur, mb))
            cur = max(cur, me)
        if cur < e:
            out.append((cur, e))
    return out


def _normalise(intervals):
    """Region semantics (util/region.d): sorted, overlapping or touching intervals merged, empty ones dropped."""
    out = []
    for b, e in sorted((int(b), int(e)) for b, e in intervals if e > b):
        if out and b <= out[-1][1]:
            out[-1] = (out[-1][0], max(out[-1][1], e))
        else:
            out.append((b, e))
    return out


def _intersect(x, y):
    out, i, j = [], 0, 0
    while i < len(x) and j < len(y):
        b, e = max(x[i][0], y[j][0]), min(x[i][1], y[j][1])
        if b < e:
            out.append((b, e))
        if x[i][1] < y[j][1]:
            i += 1
        else:
            j += 1
    return out


def common_trace_point(regions, seed, tspace, contig_len, repeat_mask=()):
    """getCommonTracePoint (cropper.d:446-500): a trace point (multiple of `tspace`, or the contig end) that lies in
    the contig-A region of every alignment chain and -- if possible -- outside `repeat_mask`; the last one for seed
    'front', the first one for seed 'back'; -1 if there is none.
    regions: per chain either one (abpos, aepos) or the list of its local alignments' (abpos, aepos)."""
    common = None
    for r in regions:
        r = _normalise([r] if isinstance(r[0], (int, np.integer)) else r)
        common = r if common is None else _intersect(common, r)
    common = common or []
    for region in (_subtract(common, _normalise(repeat_mask)), common):
        if not region:
            continue
        lo = -(-region[0][0] // tspace) * tspace
        sup = -(-region[-1][1] // tspace) * tspace
        cands = list(range(lo, sup, tspace)) + ([contig_len] if sup > contig_len else [])
        ok = [c for c in cands if any(rb <= c < re_ for rb, re_ in region) or c == region[-1][1]]
        if ok:
            return ok[-1] if seed == "front" else ok[0]
    return -1
