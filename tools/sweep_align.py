#!/usr/bin/env python
"""BASELINE.json configs[4]: align-only throughput sweep -- read length 1k..100k x error 5..15 %, one block pair
per cell, device-resident blocks, vs the CPU oracle port on a bounded sample.  Writes JSON lines.
  python tools/sweep_align.py [--ref-mbp 50] [--reads-mbp 50] [--out profiles/r01_sweep_align.jsonl]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dentist_b200 import dazzler, synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-mbp", type=float, default=50); ap.add_argument("--reads-mbp", type=float, default=50)
    ap.add_argument("--out", default="gpurun_out/sweep_align.jsonl"); ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--lens", default="1000,2000,5000,10000,20000,50000,100000"); ap.add_argument("--errors", default="0.05,0.10,0.15")
    ap.add_argument("--k", type=int, default=20, help="k-mer length (damapper default 20)")
    ap.add_argument("--ont", action="store_true", help="ONT-like error mix and a log-normal length spread (configs[2] reads)")
    ap.add_argument("--resident-index", action="store_true", help="build the reference block's k-mer index once (dn_block_index), as a multi-block job does")
    a = ap.parse_args()
    dazzler.init(0)
    n_sc = max(1, int(a.ref_mbp))
    sc = synth.make_scaffolds(n_sc, 1000000, 4001, n_repeats=0)
    ref, _ = synth.contigs_from(sc, [[] for _ in sc])
    ga = dazzler.Block(ref.off, ref.bases)
    if a.resident_index:
        ga.index(a.k)
    out = open(a.out, "w")
    for L in [int(x) for x in a.lens.split(",")]:
        for ei, e in enumerate([float(x) for x in a.errors.split(",")]):
            if a.ont:
                reads, _ = synth.simulate_reads(sc, a.reads_mbp / a.ref_mbp, L, L // 2, e, 2003, mix=(0.25, 0.45, 0.30))
            else:
                reads, _ = synth.simulate_reads(sc, a.reads_mbp / a.ref_mbp, L, 1, e, 4002 + ei + L, min_len=L, lognormal=False)
            gb = dazzler.Block(reads.off, reads.bases)
            minlen = min(1000, L // 2)
            for _ in range(2):
                dazzler.align_blocks(ga, gb, tspace=100, minlen=minlen, k=a.k)
            ms, al = 0.0, 0
            for _ in range(3):
                rec, _, _, st = dazzler.align_blocks(ga, gb, tspace=100, minlen=minlen, k=a.k)
                ms += st["ms_total"]; al += st["aligned_bases"]
            row = dict(k=a.k, read_len=L, error=e, ref_bp=int(ref.total), reads_bp=int(reads.total), reads=int(reads.nreads), las=int(len(rec)),
                       ms_per_step=ms / 3, gbp_aligned_per_s=al / 1e9 / (ms / 1e3), input_gbp_per_s=3 * reads.total / 1e9 / (ms / 1e3),
                       hits=int(st["hits"]), seeds=int(st["seeds"]), ms_extend=float(st["ms_extend"]), resident_index=bool(a.resident_index))
            if a.cpu:
                from oracle import oracle
                nr = max(1, int(np.searchsorted(reads.off, 300000)))
                sub = synth.Block(sc and np.array([0, len(sc[0])], np.int64), sc[0])
                t0 = time.perf_counter()
                la, _, _ = oracle.align(sub.off, sub.bases, reads.off[:nr + 1], reads.bases[:reads.off[nr]], tspace=100, minlen=minlen, k=a.k)
                dt = time.perf_counter() - t0
                row["cpu_port_gbp_aligned_per_s_1core"] = float((la["aepos"] - la["abpos"]).sum()) / 1e9 / dt
                row["cpu_sample"] = "first %d reads vs scaffold 0 (1 Mbp)" % nr
            out.write(json.dumps(row) + "\n"); out.flush()
            print(row, flush=True)
            gb.free()


if __name__ == "__main__":
    main()
