import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
from dentist_b200 import dazzler, pileups, synth
dazzler.init(0)
sc = synth.make_scaffolds(10, 1000000, 1001)
gaps = synth.make_gaps(sc, 10, 1002)
reads, group, _ = synth.make_pile_batch(sc, gaps, 1004, depth=20, anchor=1500)
ref, _ = synth.contigs_from(sc, gaps)
fl = dazzler.Block(ref.off, ref.bases)
lens = np.diff(reads.off).astype(np.int32)
T = lambda: time.perf_counter()
for it in range(4):
    t = [T()]
    g = dazzler.Block(reads.off, reads.bases, group=group); t.append(T())
    las = dazzler.align(g, g, tspace=126, minlen=500, e=0.7, self_block=1); t.append(T())
    n0 = len(las); las.filterLocalAlignments(0.3); t.append(T())
    las.chainLocalAlignments(min_score=126); t.append(T())
    qv, qoff = dazzler.computeQVs(lens, las, np.full(len(lens), 20, np.int32)); t.append(T())
    las.filterPileUpAlignments(lens, lens, 126); las.forceFlat(); t.append(T())
    npiles = int(group.max()) + 1
    order = np.argsort(group, kind="stable"); bounds = np.searchsorted(group[order], np.arange(npiles + 1))
    refs = [pileups.find_reference_read_candidates(qv, qoff, order[bounds[p]:bounds[p + 1]])[0] for p in range(npiles)]; t.append(T())
    cons = dazzler.getConsensus(g, las, refs); t.append(T())
    coff = np.zeros(len(cons) + 1, np.int64); coff[1:] = np.cumsum([len(c) for c in cons])
    cb = dazzler.Block(coff, np.concatenate(cons)); fla = dazzler.align(fl, cb, tspace=126, minlen=126, e=0.7); t.append(T())
    names = ["upload", "align", "filterErr", "chain", "qv", "filterPile", "refread(host)", "consensus", "flank align"]
    print("iter %d total %.1f ms | " % (it, (t[-1] - t[0]) * 1e3) + "  ".join("%s %.1f" % (n, (b - a) * 1e3) for n, a, b in zip(names, t, t[1:])), "| LAs %d->%d" % (n0, len(las)), flush=True)
    g.free(); cb.free()
