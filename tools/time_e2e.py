import sys, time
sys.path.insert(0, "/root/repo")
import bench, numpy as np
from dentist_b200 import dazzler, synth
dazzler.init(0)
ref, reads = bench.make_workload(1.0, 0)
for i in range(6):
    t0 = time.perf_counter()
    ga = dazzler.Block(ref.off, ref.bases); t1 = time.perf_counter()
    gb = dazzler.Block(reads.off, reads.bases); t2 = time.perf_counter()
    rec, toff, tr, st = dazzler.align_blocks(ga, gb, **bench.PARAMS); t3 = time.perf_counter()
    ga.free(); gb.free(); t4 = time.perf_counter()
    print("iter %d upload A %.2f  upload B %.2f  align %.2f (ms_total %.2f)  free %.2f ms" % (i, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, st["ms_total"], (t4-t3)*1e3), flush=True)
