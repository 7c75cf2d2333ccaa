"""`--compare-ladump`: the LAdump text reader on the reference's own dump (dazzler.d:965-1026) and the
comparison report, end to end through tools/compare_ladump.py with a LAS file written by dn_las_write."""
import json
import os
import subprocess
import sys

import numpy as np

from dentist_b200 import dazzler, ladump

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "las_golden.json")))


def test_parse_matches_the_reference_expectation():
    ts, rec, traces = ladump.parse(GOLD["ladump"])
    assert ts == GOLD["tspace"] and len(rec) == len(GOLD["flat"])
    for r, t, want in zip(rec, traces, GOLD["flat"]):
        assert (int(r["aread"]) + 1, int(r["bread"]) + 1) == (want["contigA"], want["contigB"])
        assert [int(r[f]) for f in ("abpos", "aepos", "bbpos", "bepos")] == [want[f] for f in ("abpos", "aepos", "bbpos", "bepos")]
        assert t.tolist() == want["trace"] and bool(int(r["flags"]) & 1) == ("complement" in want["flags"])


def test_compare_tool_reports_recall(tmp_path):
    ts, rec, traces = ladump.parse(GOLD["ladump"])
    keep = np.ones(len(rec), bool); keep[3] = False                     # "ours" misses one alignment and shifts another
    mine = rec[keep].copy(); tr = [t for t, k in zip(traces, keep) if k]
    mine[0]["aepos"] += 1
    toff = np.concatenate([[0], np.cumsum([2 * len(t) for t in tr])])[:-1].astype(np.int64)
    flat = np.concatenate([t.reshape(-1) for t in tr]).astype(np.uint16)
    las = str(tmp_path / "ours.las"); txt = str(tmp_path / "real.txt")
    dazzler.write_las(las, ts, mine, toff, flat)
    open(txt, "w").write("\n".join(GOLD["ladump"]))
    root = os.path.dirname(HERE)
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "compare_ladump.py"), "--ladump", txt, "--las", las], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout.strip().split("\n")[-1])
    assert out["x"] == len(rec) and out["y"] == len(rec) - 1
    assert out["matched"] == len(rec) - 1 and out["same_coords"] == len(rec) - 2 and out["same_trace"] == len(rec) - 2
    assert out["y_only"] == 0 and 0.8 < out["recall"] < 1.0


def test_abi_las_io_matches_the_oracle_codec_for_both_trace_widths(tmp_path):
    """dn_las_write / dn_las_read (host-only entry points) against the pinned oracle codec: uint8 traces at trace spacing
    100, uint16 at 200 and 1337 (dazzler.d:1864-1905, 1962-1984, 6049-6112)."""
    from oracle import las
    ts, rec, traces = ladump.parse(GOLD["ladump"])
    toff = np.concatenate([[0], np.cumsum([2 * len(t) for t in traces])])[:-1].astype(np.int64)
    flat = np.concatenate([t.reshape(-1) for t in traces]).astype(np.uint16)
    ntp = sum(len(t) for t in traces)
    for tspace in (100, 200, 1337):
        path = str(tmp_path / ("t%d.las" % tspace))
        dazzler.write_las(path, tspace, rec, toff, flat)
        raw = open(path, "rb").read()
        assert len(raw) == 12 + 40 * len(rec) + 2 * ntp * (1 if tspace <= 125 else 2)
        ts2, rec2, tr2 = las.decode(raw)
        assert ts2 == tspace and all(np.array_equal(a, b) for a, b in zip(tr2, traces))
        for f in ("aread", "bread", "abpos", "aepos", "bbpos", "bepos", "flags", "diffs"):
            assert np.array_equal(rec2[f], rec[f]), f
        ts3, rec3, toff3, tr3 = dazzler.read_las(path)
        assert ts3 == tspace and rec3.tobytes() == rec.tobytes() and np.array_equal(tr3, flat) and np.array_equal(toff3, toff)
