#!/usr/bin/env python
"""compare_ladump.py --ladump real.txt --las ours.las   (SURVEY §8c `--compare-ladump`)
real.txt = `LAdump -cdtl <db> <db> <las>` of the real daligner / damapper on the same DBs; ours.las = the file this
engine wrote (dn_dalign / dn_damap / bin/dn-damapper).  Prints one JSON line: how many of the real tool's local
alignments this engine reproduces (by overlap, exactly by coordinates, exactly by trace) and the shared A-base
coverage.  Needs no GPU: both inputs are files."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ladump", required=True)
    ap.add_argument("--las", required=True)
    ap.add_argument("--min-overlap", type=float, default=0.5)
    a = ap.parse_args()
    from dentist_b200 import dazzler, ladump
    ts_x, rec_x, tr_x = ladump.parse(open(a.ladump).read().split("\n"))
    ts_y, rec_y, toff, trace = dazzler.read_las(a.las)
    tr_y = [trace[int(toff[i]):int(toff[i]) + int(rec_y[i]["tlen"])].reshape(-1, 2) for i in range(len(rec_y))]
    out = ladump.compare(rec_x, tr_x, rec_y, tr_y, a.min_overlap)
    out.update(tspace_ladump=ts_x, tspace_las=ts_y,
               recall=out["matched"] / max(out["x"], 1), exact=out["same_coords"] / max(out["x"], 1),
               a_base_recall=out["a_bases_shared"] / max(out["a_bases_x"], 1))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
