"""Cuts a fixture out of the reference's example dataset (BASELINE.json configs[0]): 1.5 Mbp of
example/data/assembly-reference.fasta.gz with the four gaps example/data/gaps.bed places inside it (example/Makefile:13-17
builds its test assembly by cutting exactly these intervals out of the reference sequence).  Run in the survey container
(reads /root/reference); the GPU box only sees the committed .npz.

    python tests/golden/make_example_excerpt.py [--full out.npz]     # --full: the whole scaffold + all 147 gaps (not committed)
"""
import gzip
import os
import sys

import numpy as np

REF = "/root/reference/example/data"
HERE = os.path.dirname(os.path.abspath(__file__))
BEGIN, END = 1700000, 3200000


def load():
    seq = []
    with gzip.open(os.path.join(REF, "assembly-reference.fasta.gz"), "rt") as f:
        for ln in f:
            if not ln.startswith(">"):
                seq.append(ln.strip())
    s = np.frombuffer("".join(seq).lower().encode(), np.uint8)
    codes = np.full(256, 255, np.uint8)
    for i, c in enumerate(b"acgt"):
        codes[c] = i
    b = codes[s]
    assert (b < 4).all(), "non-ACGT base in the reference sequence"
    gaps = [tuple(int(x) for x in ln.split()[1:3]) for ln in open(os.path.join(REF, "gaps.bed"))]
    return b, gaps


def pack(b):
    pad = (-len(b)) % 4
    q = np.concatenate([b, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    return ((q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]).astype(np.uint8)


if __name__ == "__main__":
    b, gaps = load()
    if len(sys.argv) > 2 and sys.argv[1] == "--full":
        np.savez_compressed(sys.argv[2], packed=pack(b), length=len(b), gaps=np.array(gaps, np.int64), source="example/data: whole scaffold, 147 gaps")
        print("wrote", sys.argv[2], len(b), "bp", len(gaps), "gaps")
    else:
        g = [(s - BEGIN, e - BEGIN) for s, e in gaps if BEGIN < s and e < END]
        out = os.path.join(HERE, "example_excerpt.npz")
        np.savez_compressed(out, packed=pack(b[BEGIN:END]), length=END - BEGIN, gaps=np.array(g, np.int64),
                            source="example/data/assembly-reference.fasta.gz [%d, %d) + gaps.bed" % (BEGIN, END))
        print("wrote", out, END - BEGIN, "bp", g)
