"""BadAlignmentCoverageAssessor -- ORACLE (test infrastructure only): the event-based restatement of
source/dentist/commands/maskRepetitiveRegions.d:116-232 (alignment intervals of chains, improper-only pass),
:258-420 (assessor state machine, coverage zones) and :436-560 (CoverageChangeRange).  The D source is in
/root/reference and carries two golden tests (:395-411 mask for bounds (3, 5); :582-617 coverage changes), both
extracted to tests/golden/maskcov_kat.json -> parity PINNED."""
COMP, START, NEXT, BEST = 0x1, 0x4, 0x8, 0x10


def coverage_changes(intervals, contigs):
    """intervals: [(contig, begin, end)], contigs: [(contig, 0, length)] -> [(contig, pos, current, new)]  (:436-560)"""
    if not intervals:
        return []
    ev = []
    for c, b, e in intervals:
        ev += [(c, b, 1), (c, e, -1)]
    for c, b, e in contigs:
        ev += [(c, 0, 0), (c, e - b, 0)]
    ev.sort()
    out, cur, i = [], 0, 0
    while i < len(ev):
        c, p, d = ev[i][0], ev[i][1], 0
        while i < len(ev) and ev[i][0] == c and ev[i][1] == p:
            d += ev[i][2]; i += 1
        out.append((c, p, cur, cur + d)); cur += d
    return out


def bad_coverage_mask(intervals, contigs, lower, upper):
    """opCall (:343-378): maximal stretches whose coverage is < lower or > upper, as a normalised region."""
    zone = lambda cov: 0 if cov < lower else (2 if cov > upper else 1)
    ch = coverage_changes(intervals, contigs)
    if not ch:
        return []
    acc, masking, mc, ms, last = [], False, 0, 0, ch[0]
    for e in ch:
        cz, nz = zone(e[2]), zone(e[3])
        if masking and e[0] != last[0]:
            acc.append((mc, ms, last[1])); masking = False
        if not masking and (nz != 1 or (cz == 1 and cz != nz)):
            masking, mc, ms = True, e[0], e[1]
        elif masking and cz != 1 and nz == 1:
            acc.append((mc, ms, e[1])); masking = False
        last = e
    if masking:
        acc.append((mc, ms, last[1]))
    out = []                                                           # ReferenceRegion(...) normalises: sort, merge, drop empties
    for c, b, e in sorted(x for x in acc if x[2] > x[1]):
        if out and out[-1][0] == c and b <= out[-1][2]:
            out[-1] = (c, out[-1][1], max(out[-1][2], e))
        else:
            out.append((c, b, e))
    return out


def chain_intervals(rec, alen, blen, improper_only=False, allowance=0):
    """alignmentIntervals (:186-205): one (contigA, first.begin, last.end) per chain; improperOnly keeps chains that
    are not isProper(allowance) (base.d:537-557).  Contig ids stay 0-based here."""
    out, i, n = [], 0, len(rec)
    while i < n:
        j = i + 1
        if int(rec[i]["flags"]) & (START | BEST):
            while j < n and (int(rec[j]["flags"]) & NEXT):
                j += 1
        f, l = rec[i], rec[j - 1]
        a, b = int(f["aread"]), int(f["bread"])
        proper = ((int(f["abpos"]) <= allowance or int(f["bbpos"]) <= allowance) and
                  (int(l["aepos"]) + allowance >= int(alen[a]) or int(l["bepos"]) + allowance >= int(blen[b])))
        if not improper_only or not proper:
            out.append((a, int(f["abpos"]), int(l["aepos"])))
        i = j
    return out


def mask_repetitive_regions(rec, alen, blen, bounds, improper_bounds=None, allowance=0):
    """assessRepeatStructure + writeRepeatMask (:135-232): union of the all-chains mask and (reads only) the
    improper-chains mask."""
    contigs = [(c, 0, int(alen[c])) for c in range(len(alen))]
    m = bad_coverage_mask(chain_intervals(rec, alen, blen), contigs, bounds[0], bounds[1])
    if improper_bounds is not None:
        m = m + bad_coverage_mask(chain_intervals(rec, alen, blen, True, allowance), contigs, improper_bounds[0], improper_bounds[1])
    out = []
    for c, b, e in sorted(m):
        if out and out[-1][0] == c and b <= out[-1][2]:
            out[-1] = (c, out[-1][1], max(out[-1][2], e))
        else:
            out.append((c, b, e))
    return out


def propagate_mask(rec, traces, tspace, mask_by_contig, blen):
    """MaskPropagator (commands/propagateMask.d:109-300): every local alignment carries the parts of the A-contig mask
    it overlaps over to its B read -- interval begin translated with RoundingMode.floor, end with ceil
    (Trace.translateTracePoint, base.d:185-242, pinned by the 21-tile KAT in tests/golden/las_golden.json), mirrored for
    complement alignments -- and the per-read union is the output mask.  No unit test in the reference for the
    sweep itself: pinned through the translate KAT and the D source only.
    mask_by_contig: per A contig a sorted list of disjoint (begin, end); returns the same for the B reads."""
    from oracle import las as _las
    out = [[] for _ in blen]
    for i in range(len(rec)):
        r = rec[i]
        a, b = int(r["aread"]), int(r["bread"])
        ab, ae, bb = int(r["abpos"]), int(r["aepos"]), int(r["bbpos"])
        hit = [(mb, me) for mb, me in mask_by_contig[a] if me > ab and mb < ae]      # getIntersectingIntervals :216-231
        for k, (mb, me) in enumerate(hit):
            if k == 0:
                mb = max(mb, ab)
            if k == len(hit) - 1:
                me = min(me, ae)
            pb = _las.translate_trace_point(ab, ae, bb, tspace, traces[i], mb, "floor")[1]
            pe = _las.translate_trace_point(ab, ae, bb, tspace, traces[i], me, "ceil")[1]
            if int(r["flags"]) & COMP:
                pb, pe = int(blen[b]) - pe, int(blen[b]) - pb
            out[b].append((pb, pe))
    res = []
    for iv in out:                                                                    # QueryRegion normalisation (mergeMasks :296-300)
        m = []
        for s, e in sorted(x for x in iv if x[1] > x[0]):
            if m and s <= m[-1][1]:
                m[-1] = (m[-1][0], max(m[-1][1], e))
            else:
                m.append((s, e))
        res.append(m)
    return res
