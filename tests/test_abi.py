"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/dentist_b200.h declares, and refuses (loudly) to compute without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from dentist_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dentist_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), "missing export %s" % n
    assert set(_lib.EXPORTS) <= set(names)


def test_version_and_defaults():
    L = _lib.lib()
    assert b"sm_100a" in L.dn_version()
    p = _lib.AlignParams()
    L.dn_align_params_default(C.byref(p))
    assert (p.k, p.w, p.h, p.tspace, p.minlen) == (14, 6, 35, 100, 1000)
    assert abs(p.e - 0.7) < 1e-12


def test_record_struct_is_the_40_byte_las_record():
    assert C.sizeof(_lib.LasRecord) == 40 == _lib.REC_DTYPE.itemsize


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = _lib.lib()
    rlen = np.array([8], np.int32); boff = np.array([0], np.int64); data = np.zeros(8, np.uint8)
    d = _lib.BlockDesc(1, 0, rlen.ctypes.data, boff.ctypes.data, data.ctypes.data, 8, None, None, None)
    h = C.c_void_p()
    rc = L.dn_block_upload(C.byref(d), C.byref(h))
    assert rc == 2 and b"no CUDA device" in L.dn_last_error()
    assert L.dn_init(0, None) == 2
    buf = _lib.LasBuf()
    assert L.dn_align_host(C.byref(d), C.byref(d), None, C.byref(buf)) == 2


def test_las_file_roundtrip_through_the_abi(tmp_path):
    """dn_las_write / dn_las_read against the oracle codec (both trace widths; dazzler.d:1864-1905)."""
    from oracle import las
    from dentist_b200 import dazzler
    recs = [dict(aread=0, bread=1, flags=las.START | las.BEST, abpos=3, aepos=204, bbpos=5, bepos=210),
            dict(aread=18, bread=19, flags=las.COMP, abpos=21, aepos=22, bbpos=23, bepos=24)]
    traces = [[(7, 99), (1, 101), (0, 5)], [(0, 1)]]
    for ts in (100, 126):
        p = str(tmp_path / ("t%d.las" % ts))
        open(p, "wb").write(las.encode(recs, traces, ts))
        tspace, rec, toff, tr = dazzler.read_las(p)
        assert tspace == ts and len(rec) == 2
        assert rec["diffs"].tolist() == [8, 0] and rec["aread"].tolist() == [0, 18]
        assert tr[toff[0]:toff[0] + rec[0]["tlen"]].reshape(-1, 2).tolist() == [list(t) for t in traces[0]]
        # write back through the ABI and compare bytes with the oracle encoder
        buf = _lib.LasBuf()
        _lib.check(_lib.lib().dn_las_read(p.encode(), C.byref(buf)))
        q = str(tmp_path / ("w%d.las" % ts))
        _lib.check(_lib.lib().dn_las_write(q.encode(), C.byref(buf)))
        _lib.lib().dn_las_free(C.byref(buf))
        assert open(q, "rb").read() == open(p, "rb").read()
    with pytest.raises(dazzler.DnError):
        open(str(tmp_path / "bad.las"), "wb").write(open(p, "rb").read()[:-3])
        dazzler.read_las(str(tmp_path / "bad.las"))


def test_reference_read_candidates_matches_python_mirror():
    """dn_reference_read_candidates (host logic, processPileUps/package.d:518-568) == the numpy mirror in pileups.py."""
    from dentist_b200 import dazzler, pileups
    rng = np.random.default_rng(8)
    nreads, npiles = 60, 7
    group = np.sort(rng.integers(0, npiles, nreads)).astype(np.int32)
    ntiles = rng.integers(5, 40, nreads)
    qoff = np.zeros(nreads + 1, np.int64); qoff[1:] = np.cumsum(ntiles)
    qv = rng.integers(0, 51, int(qoff[-1])).astype(np.uint8)
    got = dazzler.findReferenceReadCandidates(qv, qoff, group, npiles)
    for p in range(npiles):
        members = np.flatnonzero(group == p)
        if len(members) == 0:
            assert len(got[p]) == 0
            continue
        assert got[p].tolist() == pileups.find_reference_read_candidates(qv, qoff, members)


def test_reference_read_candidates_hand_derived_vectors():
    """findReferenceReadCandidates worked out BY HAND from the D source (processPileUps/package.d:518-568), independent of
    both implementations in this repo:
      hist over qv < maxQV(50) of the allowed reads (:525-532); badThres = cast(size_t)(badFraction * histTotal) (:534);
      badQV = 49 - (first index of the cumulative sum over the reversed histogram that reaches badThres) (:535-538);
      reads ordered by the tuple (count(qv >= badQV), mean(qv), readNumber) (:546-557)."""
    from dentist_b200 import dazzler

    def run(per_read, group, npiles, bad_fraction):
        qoff = np.zeros(len(per_read) + 1, np.int64); qoff[1:] = np.cumsum([len(q) for q in per_read])
        qv = np.concatenate([np.array(q, np.uint8) for q in per_read])
        return [r.tolist() for r in dazzler.findReferenceReadCandidates(qv, qoff, np.array(group, np.int32), npiles, bad_fraction)]

    # Case 1, badFraction 0.2.  hist: 10 -> 4, 12 -> 4, 30 -> 1, 40 -> 1 (the two 50s are >= maxQV: not counted), total 10,
    # badThres = 2.  Reversed cumulative sum: QV 49..41 -> 0, QV 40 -> 1, QV 39..31 -> 1, QV 30 -> 2 >= 2 at index 19
    # => badQV = 49 - 19 = 30.  (numBad, mean): read 0 -> (1, 17.5), read 1 -> (0, 12.0), read 2 -> (3, 35.0)  => 1, 0, 2.
    r0, r1, r2 = [10, 10, 10, 40], [12, 12, 12, 12], [10, 30, 50, 50]
    assert run([r0, r1, r2], [0, 0, 0], 1, 0.2) == [[1, 0, 2]]
    # Case 2, badFraction 0.08 (the default) on the same reads: badThres = cast(size_t)(0.8) = 0; the cumulative sum is
    # >= 0 already at index 0 => badQV = 49; only the 50s count as bad: (0, 17.5), (0, 12.0), (2, 35.0) => 1, 0, 2.
    assert run([r0, r1, r2], [0, 0, 0], 1, 0.08) == [[1, 0, 2]]
    # Case 3: equal numBad and equal mean -> the read number decides; equal numBad, different mean -> lower mean first.
    # badFraction 0.5: hist 20 -> 6, 30 -> 2, total 8, badThres 4; reversed cumulative: QV 30 -> 2, QV 20 -> 8 >= 4 at index 29
    # => badQV = 20: every value is bad.  (3, 23.33), (3, 20.0), (2, 25.0), (3, 20.0) => read 2, then 1 and 3 (tie: id), then 0.
    assert run([[20, 20, 30], [20, 20, 20], [20, 30], [20, 20, 20]], [0, 0, 0, 0], 1, 0.5) == [[2, 1, 3, 0]]
    # Case 4: reads outside allowedReferenceReadIds (group -1) enter neither the histogram nor the ranking (:520-523); piles
    # are independent.  Pile 0 = reads 0, 2 (read 1 not allowed): hist 10 -> 3, 40 -> 1, 30 -> 1 (50s dropped), total 6,
    # badFraction 0.2 -> badThres 1; reversed cumulative reaches 1 at QV 40 (index 9) => badQV 40: (1, 17.5), (2, 35.0) => 0, 2.
    # Pile 1 = read 3 alone.
    assert run([r0, r1, r2, [5, 5]], [0, -1, 0, 1], 2, 0.2) == [[0, 2], [3]]


def test_cli_stand_ins_exist_and_print_usage():
    """dn-damapper / dn-daligner / dn-dbdust keep the argv contract of the tools the workflow calls (Snakefile:1143-1169)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for tool, args in (("dn-damapper", ["only-one.db"]), ("dn-daligner", []), ("dn-dbdust", ["a.db", "b.db"])):
        r = subprocess.run([os.path.join(root, "bin", tool)] + args, capture_output=True, text=True)
        assert r.returncode == 1 and r.stderr.startswith("Usage: " + tool), (tool, r.stderr)
