import sys, time
sys.path.insert(0, "/root/repo")
import bench, numpy as np
from dentist_b200 import dazzler, synth
dazzler.init(0)
ref, reads = bench.make_workload(1.0, 0)
ga = dazzler.Block(ref.off, ref.bases); gb = dazzler.Block(reads.off, reads.bases)
for i in range(8):
    t=time.perf_counter()
    rec, toff, tr, st = dazzler.align_blocks(ga, gb, **bench.PARAMS)
    dt=time.perf_counter()-t
    print("call %d wall %.2f ms_total %.2f seed %.2f ext %.2f nla %d hits %d seeds %d" % (i, dt*1e3, st["ms_total"], st["ms_seed"], st["ms_extend"], len(rec), st["hits"], st["seeds"]), flush=True)
