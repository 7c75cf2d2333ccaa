"""Times dn_process_pileups on the bench's pile-up batch; DN_TRACE=1 prints the engine's per-stage wall clock."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dentist_b200 import dazzler, synth
dazzler.init(0)
NSC = int(sys.argv[2]) if len(sys.argv) > 2 else 10          # scaffolds of 10 gaps each: the batch holds 10 * NSC pile-ups
sc = synth.make_scaffolds(NSC, 1000000, 1001)
gaps = synth.make_gaps(sc, 10, 1002)
ref, _ = synth.contigs_from(sc, gaps)
preads, pgroup, _ = synth.make_pile_batch(sc, gaps, 1004, depth=20, anchor=1500)
npiles = int(pgroup.max()) + 1
flank_of, c = [], 0
for gl in gaps:
    for _g in gl:
        flank_of.append([c, c + 1]); c += 1
    c += 1
order = np.argsort(pgroup, kind="stable"); bounds = np.searchsorted(pgroup[order], np.arange(npiles + 1))
piles = [dict(reads=[preads.read(int(r)) for r in order[bounds[p]:bounds[p + 1]]], flanks=flank_of[p]) for p in range(npiles)]
ga = dazzler.Block(ref.off, ref.bases)
batch = dazzler.PileupBatch(ga, piles)
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    l0 = dazzler.launch_count(); t = time.perf_counter()
    res = batch.run()
    print("iter %d: %.2f ms, %d launches, %d consensus bases" % (it, (time.perf_counter() - t) * 1e3, dazzler.launch_count() - l0, res.consensus_bases()), flush=True)
