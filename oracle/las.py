"""LAS wire format + trace-point arithmetic -- ORACLE restatement (test infrastructure only).

Follows the reference's reader/writer and data model:
  * header  int64 novl; int32 tspace            dazzler.d:1665-1688, 6024-6047
  * record  40 bytes = DazzlerOverlap[8..48)      dazzler.d:1717-1725, 1988-2016, 2146-2150
            {int32 tlen, diffs, abpos, bbpos, aepos, bepos; uint32 flags; int32 aread, bread; 4 B pad}
  * trace   tlen elements, uint8 if tspace <= 125 (TRACE_XOVR) else uint16   dazzler.d:2019-2025, 1783-1795
  * flags   COMP 0x1, START 0x4, NEXT 0x8, BEST 0x10, ELIM 0x20  dazzler.d:1991-1998;
            mapping to DENTIST flags  dazzler.d:1738-1755 (read), 2052-2073 / 2100-2114 (write)
  * ids     0-based in the file, 1-based in DENTIST  dazzler.d:1731-1734, 2049-2050
  * `diffs` on write = sum of tile diffs  dazzler.d:2143
Pinned against the reference's own vectors in tests/test_oracle_golden.py
(text dump dazzler.d:965-1026 <-> expected records :1045-1166 <-> chains :502-654;
 trace-point KAT base.d:881-944).
"""
import struct

import numpy as np

COMP, START, NEXT, BEST, ELIM = 0x1, 0x4, 0x8, 0x10, 0x20
TRACE_XOVR = 125

REC_DTYPE = np.dtype([("tlen", "<i4"), ("diffs", "<i4"), ("abpos", "<i4"), ("bbpos", "<i4"),
                      ("aepos", "<i4"), ("bepos", "<i4"), ("flags", "<u4"), ("aread", "<i4"),
                      ("bread", "<i4"), ("pad", "<i4")])
assert REC_DTYPE.itemsize == 40


def is_large_trace(tspace):
    return tspace > TRACE_XOVR


def encode(records, traces, tspace):
    """records: iterable of dicts/np records with the REC fields (aread/bread 0-based);
    traces: list of [(diffs, bbases), ...].  Returns the LAS file bytes."""
    out = [struct.pack("<qi", len(records), tspace)]
    large = is_large_trace(tspace)
    for r, t in zip(records, traces):
        t = np.asarray(t, dtype=np.int64).reshape(-1, 2)
        rec = np.zeros(1, REC_DTYPE)
        for f in ("abpos", "bbpos", "aepos", "bepos", "flags", "aread", "bread"):
            rec[f] = r[f]
        rec["tlen"] = 2 * len(t)
        rec["diffs"] = int(t[:, 0].sum())
        out.append(rec.tobytes())
        out.append(t.astype("<u2" if large else "u1").tobytes())
    return b"".join(out)


def decode(buf):
    """Returns (tspace, records REC_DTYPE array, list of trace arrays [n,2] as uint16)."""
    novl, tspace = struct.unpack_from("<qi", buf, 0)
    pos = 12
    large = is_large_trace(tspace)
    recs = np.zeros(novl, REC_DTYPE)
    traces = []
    for i in range(novl):
        if pos + 40 > len(buf):
            raise ValueError("unexpected end of file; expected overlapHead")
        recs[i] = np.frombuffer(buf, REC_DTYPE, 1, pos)[0]
        pos += 40
        tlen = int(recs[i]["tlen"])
        nbytes = tlen * (2 if large else 1)
        if nbytes % 2:
            raise ValueError("illegal value for tlen: must be multiple of 2")
        if pos + nbytes > len(buf):
            raise ValueError("unexpected end of file; expected tracePoints")
        t = np.frombuffer(buf, "<u2" if large else "u1", tlen, pos).astype(np.uint16).reshape(-1, 2)
        traces.append(t)
        pos += nbytes
    return tspace, recs, traces


# --- DENTIST-side flag view (base.d:121-133) -----------------------------------------------

def dentist_flags(las_flags):
    """Set of DENTIST AlignmentFlag names for a LAS flag word (dazzler.d:1738-1755)."""
    f = set()
    if las_flags & ELIM:
        f.add("disabled")
    if las_flags & COMP:
        f.add("complement")
    if (las_flags & START) and not (las_flags & BEST):
        f.add("alternateChain")
    if las_flags & NEXT:
        f.add("chainContinuation")
    if not (las_flags & (START | BEST | NEXT)):
        f.add("unchained")
    return f


def las_flags(dflags):
    """Inverse, as writeFlatLocalAlignment does (dazzler.d:2100-2114)."""
    w = 0
    if "disabled" in dflags:
        w |= ELIM
    if "complement" in dflags:
        w |= COMP
    if "chainContinuation" in dflags:
        w |= NEXT
    elif "unchained" not in dflags:
        w |= START
        if "alternateChain" not in dflags:
            w |= BEST
    return w


def parse_ladump(lines):
    """Parse the `LAdump -cdtl`-style text the reference's unittests feed to dumpLA
    (dazzler.d:965-1026).  Returns (tspace, records, traces) with 0-based read ids."""
    tspace = 100
    recs, traces, cur = [], [], None
    it = iter(lines)
    for ln in it:
        p = ln.split()
        if not p:
            continue
        if p[0] == "X":
            tspace = int(p[1])
        elif p[0] == "P":
            fl = 0
            if p[3] == "c":
                fl |= COMP
            fl |= {">": START | BEST, "+": START, "-": NEXT, ".": 0}[p[4]]
            cur = dict(aread=int(p[1]) - 1, bread=int(p[2]) - 1, flags=fl)
        elif p[0] == "C":
            cur.update(abpos=int(p[1]), aepos=int(p[2]), bbpos=int(p[3]), bepos=int(p[4]))
        elif p[0] == "T":
            n = int(p[1])
            t = [tuple(int(x) for x in next(it).split()) for _ in range(n)]
            recs.append(cur)
            traces.append(t)
    return tspace, recs, traces


def chains(recs):
    """Group flat records into DENTIST AlignmentChains like AlignmentChainPacker
    (dazzler.d:708-743): a START record opens a chain, NEXT records continue it,
    flag-less records are single 'unchained' chains.  Returns list of index lists."""
    out = []
    i = 0
    n = len(recs)
    while i < n:
        f = int(recs[i]["flags"])
        grp = [i]
        i += 1
        if f & NEXT:
            raise ValueError("chain is missing a start")
        if f & (START | BEST):
            while i < n and (int(recs[i]["flags"]) & NEXT):
                grp.append(i)
                i += 1
        out.append(grp)
    return out


# --- trace point arithmetic (base.d:185-242) ------------------------------------------------

def num_tiles(abpos, aepos, tspace):
    """ceil(aepos/ts) - floor(abpos/ts): first tile ends at the next multiple of tspace after
    abpos, the last one at aepos (base.d:196-200, 217-219)."""
    return -(-aepos // tspace) - abpos // tspace


def trace_points_up_to_a(abpos, aepos, tspace, ntp, pos, mode):
    """Trace.tracePointsUpTo!"contigA" (base.d:210-242); mode is 'floor' or 'ceil'."""
    assert abpos <= pos <= aepos
    second = (abpos // tspace) * tspace + tspace
    second_from_last = ((aepos - 1) // tspace) * tspace
    if mode == "floor":
        if pos < second:
            return 0
        if pos < aepos:
            return 1 + (pos - second) // tspace
        return ntp
    if pos == abpos:
        return 0
    if pos <= second:
        return 1
    if pos <= second_from_last:
        return 1 + -(-(pos - second) // tspace)
    return ntp


def translate_trace_point(abpos, aepos, bbpos, tspace, trace, pos, mode):
    """Trace.translateTracePoint!"contigA" (base.d:185-203) -> (contigA pos, contigB pos)."""
    if not (abpos <= pos <= aepos):
        raise ValueError("position outside alignment")
    idx = trace_points_up_to_a(abpos, aepos, tspace, len(trace), pos, mode)
    b = bbpos + int(sum(int(t[1]) for t in trace[:idx]))
    if idx == 0:
        a = abpos
    elif idx < len(trace):
        a = (abpos // tspace) * tspace + idx * tspace
    else:
        a = aepos
    return a, b
