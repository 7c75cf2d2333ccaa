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


def test_cli_stand_ins_exist_and_print_usage():
    """dn-damapper / dn-daligner / dn-dbdust keep the argv contract of the tools the workflow calls (Snakefile:1143-1169)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for tool, args in (("dn-damapper", ["only-one.db"]), ("dn-daligner", []), ("dn-dbdust", ["a.db", "b.db"])):
        r = subprocess.run([os.path.join(root, "bin", tool)] + args, capture_output=True, text=True)
        assert r.returncode == 1 and r.stderr.startswith("Usage: " + tool), (tool, r.stderr)
