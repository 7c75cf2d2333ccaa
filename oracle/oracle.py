"""ctypes binding of the CPU oracle (oracle/liboracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  The product path (dentist_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return so


class OrcParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("k", "w", "h", "t", "tspace", "minlen", "cdiff", "xdrop", "wmax", "rounds", "self_", "poolmul")]


class OrcBlock(C.Structure):
    _fields_ = [("nreads", C.c_int32), ("off", C.c_void_p), ("bases", C.c_void_p), ("mask", C.c_void_p),
                ("group", C.c_void_p)]


class OrcResult(C.Structure):
    _fields_ = [("nla", C.c_int64), ("la", C.c_void_p), ("ntrace", C.c_int64), ("trace", C.c_void_p),
                ("nhits", C.c_int64), ("nseeds", C.c_int64), ("next", C.c_int64)]


LA_DTYPE = np.dtype([("tlen", "<i4"), ("diffs", "<i4"), ("abpos", "<i4"), ("bbpos", "<i4"),
                     ("aepos", "<i4"), ("bepos", "<i4"), ("flags", "<u4"), ("aread", "<i4"),
                     ("bread", "<i4"), ("toff", "<i4")])

DEFAULTS = dict(k=14, w=6, h=35, t=32, tspace=100, minlen=500, cdiff=20, xdrop=300, wmax=30, rounds=3,
                self_=0, poolmul=64)


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_align_mt.restype = C.c_int
        _LIB.orc_align_mt.argtypes = [C.POINTER(OrcBlock), C.POINTER(OrcBlock), C.POINTER(OrcParams),
                                      C.POINTER(OrcResult), C.c_int]
        _LIB.orc_free.argtypes = [C.POINTER(OrcResult)]
    return _LIB


def _block(off, bases, mask, group=None):
    off = np.ascontiguousarray(off, dtype=np.int64)
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    keep = [off, bases]
    b = OrcBlock(len(off) - 1, off.ctypes.data, bases.ctypes.data, None, None)
    if group is not None:
        group = np.ascontiguousarray(group, dtype=np.int32)
        keep.append(group)
        b.group = group.ctypes.data
    if mask is not None:
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        keep.append(mask)
        b.mask = mask.ctypes.data
    return b, keep


def align(a_off, a_bases, b_off, b_bases, a_mask=None, b_mask=None, a_group=None, b_group=None, threads=1, **params):
    """Run the oracle.  Blocks are (offsets[nreads+1], bases uint8 0..3 concatenated).
    `threads` host threads share the A index and map their own B reads (same result for any count).
    Returns (la structured array, trace uint16 array, stats dict)."""
    p = dict(DEFAULTS)
    if "self" in params:
        params["self_"] = params.pop("self")
    p.update(params)
    P = OrcParams(**p)
    A, ka = _block(a_off, a_bases, a_mask, a_group)
    B, kb = _block(b_off, b_bases, b_mask, b_group)
    R = OrcResult()
    rc = lib().orc_align_mt(C.byref(A), C.byref(B), C.byref(P), C.byref(R), int(threads))
    if rc != 0:
        raise RuntimeError("oracle failed: %d" % rc)
    la = np.ctypeslib.as_array(C.cast(R.la, C.POINTER(C.c_uint8)), shape=(R.nla * LA_DTYPE.itemsize,)) \
        .view(LA_DTYPE).copy() if R.nla else np.zeros(0, LA_DTYPE)
    tr = np.ctypeslib.as_array(C.cast(R.trace, C.POINTER(C.c_uint16)), shape=(R.ntrace,)).copy() \
        if R.ntrace else np.zeros(0, np.uint16)
    stats = dict(nhits=R.nhits, nseeds=R.nseeds, next=R.next)
    lib().orc_free(C.byref(R))
    del ka, kb
    return la, tr, stats


# ---------------------------------------------------------------------------------------------
# per-pile stages (oracle/pile_oracle.c)

LAS40 = np.dtype([("tlen", "<i4"), ("diffs", "<i4"), ("abpos", "<i4"), ("bbpos", "<i4"), ("aepos", "<i4"),
                  ("bepos", "<i4"), ("flags", "<u4"), ("aread", "<i4"), ("bread", "<i4"), ("pad", "<i4")])


def _las40(rec):
    out = np.zeros(len(rec), LAS40)
    for f in ("tlen", "diffs", "abpos", "bbpos", "aepos", "bepos", "flags", "aread", "bread"):
        out[f] = rec[f]
    return out


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def filter_error(rec, max_err):
    r = _las40(rec); keep = np.zeros(len(r), np.uint8)
    f = lib().orc_filter_error; f.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_void_p]
    f(_p(r), len(r), float(max_err), _p(keep))
    return keep.astype(bool)


def filter_pileup(rec, alen, blen, allowance):
    r = _las40(rec); keep = np.zeros(len(r), np.uint8)
    alen = np.ascontiguousarray(alen, np.int32); blen = np.ascontiguousarray(blen, np.int32)
    f = lib().orc_filter_pileup; f.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    f(_p(r), len(r), _p(alen), _p(blen), int(allowance), _p(keep))
    return keep.astype(bool)


def qv(rlen, rec, toff, trace, tspace, cov):
    """rec sorted by aread.  Returns (qv bytes, qv_off[nreads+1])."""
    rlen = np.ascontiguousarray(rlen, np.int32)
    r = _las40(rec); toff = np.ascontiguousarray(toff, np.int64); trace = np.ascontiguousarray(trace, np.uint16)
    nt = (rlen.astype(np.int64) + tspace - 1) // tspace
    qoff = np.zeros(len(rlen) + 1, np.int64); qoff[1:] = np.cumsum(nt)
    out = np.zeros(int(qoff[-1]), np.uint8)
    f = lib().orc_qv
    f.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    f(_p(rlen), len(rlen), _p(r), len(r), _p(toff), _p(trace), int(tspace), int(cov), _p(qoff), _p(out))
    return out, qoff


def consensus(off, bases, rec, toff, trace, tspace, read):
    off = np.ascontiguousarray(off, np.int64); bases = np.ascontiguousarray(bases, np.uint8)
    r = _las40(rec); toff = np.ascontiguousarray(toff, np.int64); trace = np.ascontiguousarray(trace, np.uint16)
    L = int(off[read + 1] - off[read])
    out = np.zeros(2 * L + 16, np.uint8); n = C.c_int32(0)
    f = lib().orc_consensus
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_int32)]
    f(_p(off), _p(bases), _p(r), len(r), _p(toff), _p(trace), int(tspace), int(read), _p(out), C.byref(n))
    return out[:n.value].copy()


def transpose(a_off, a_bases, b_off, b_bases, rec, toff, trace, tspace):
    """orc_transpose: the records of B.A.las for the records of A.B.las (input order; the caller sorts).
    Returns (records LAS40, toff, trace)."""
    a_off = np.ascontiguousarray(a_off, np.int64); a_bases = np.ascontiguousarray(a_bases, np.uint8)
    b_off = np.ascontiguousarray(b_off, np.int64); b_bases = np.ascontiguousarray(b_bases, np.uint8)
    r = _las40(rec); toff = np.ascontiguousarray(toff, np.int64); trace = np.ascontiguousarray(trace, np.uint16)
    blen = np.diff(b_off)
    comp = (r["flags"] & 1).astype(bool)
    ab2 = np.where(comp, blen[r["bread"]] - r["bepos"], r["bbpos"]); ae2 = np.where(comp, blen[r["bread"]] - r["bbpos"], r["bepos"])
    nt2 = np.where(ae2 > ab2, -(-ae2 // tspace) - ab2 // tspace, 0)
    out = np.zeros(len(r), LAS40); otoff = np.zeros(len(r), np.int64); otr = np.zeros(int(2 * nt2.sum()) + 2, np.uint16)
    f = lib().orc_transpose
    f.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    f(_p(a_off), _p(a_bases), _p(b_off), _p(b_bases), _p(r), len(r), _p(toff), _p(trace), int(tspace), _p(out), _p(otoff), _p(otr))
    assert np.array_equal(out["tlen"], 2 * nt2)
    return out, otoff, otr[:int(2 * nt2.sum())]


def bridge(a_off, a_bases, b_off, b_bases, rec, toff, trace, tspace, cdiff=20):
    """orc_bridge (`daligner -B`): neighbouring records of one (aread, bread, comp) separated by a short gap become one.
    rec in LAsort order.  Returns (records LAS40, toff, trace, number of bridges)."""
    a_off = np.ascontiguousarray(a_off, np.int64); a_bases = np.ascontiguousarray(a_bases, np.uint8)
    b_off = np.ascontiguousarray(b_off, np.int64); b_bases = np.ascontiguousarray(b_bases, np.uint8)
    r = _las40(rec); toff = np.ascontiguousarray(toff, np.int64); trace = np.ascontiguousarray(trace, np.uint16)
    out = np.zeros(len(r) + 1, LAS40); otoff = np.zeros(len(r) + 1, np.int64)
    otr = np.zeros(int(r["tlen"].sum()) + 2 * len(r) * (128 // int(tspace) + 2) + 2, np.uint16)
    nb = C.c_int64(0)
    f = lib().orc_bridge
    f.restype = C.c_int64
    f.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    n = f(_p(a_off), _p(a_bases), _p(b_off), _p(b_bases), _p(r), len(r), _p(toff), _p(trace), int(tspace), int(cdiff),
          _p(out), _p(otoff), _p(otr), C.byref(nb))
    out = out[:n]; otoff = otoff[:n]
    return out, otoff, otr[:int(out["tlen"].sum())], int(nb.value)
