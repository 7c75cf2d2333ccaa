"""Stage trace (DN_TRACE=1) of one C3-scale block pair: 100 Mbp assembly x ~190 Mbp ONT-like reads."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dentist_b200 import dazzler, synth
dazzler.init(0)
sc = synth.make_scaffolds(100, 1000000, 4001, n_repeats=0)
ref, _ = synth.contigs_from(sc, [[] for _ in sc])
reads, _ = synth.simulate_reads(sc, 1.95, 20000, 10000, 0.12, 2003, mix=(0.25, 0.45, 0.30))
ga, gb = dazzler.Block(ref.off, ref.bases), dazzler.Block(reads.off, reads.bases)
for i in range(3):
    if i == 2:
        print("---- with resident index", file=sys.stderr); ga.index(20)
    rec, _, _, st = dazzler.align_blocks(ga, gb, tspace=100, minlen=1000, k=20)
    print("ms_total %.2f extend %.2f LAs %d" % (st["ms_total"], st["ms_extend"], len(rec)), file=sys.stderr)
