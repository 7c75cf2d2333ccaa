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
