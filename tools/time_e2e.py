"""Where the end-to-end step goes: dn_align_host wall time vs the device-timed alignment, and the upload alone."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from dentist_b200 import dazzler, synth
dazzler.init(0)
ref, reads = bench.make_workload(1.0, 0)
def pinned(a):
    t = torch.from_numpy(a).pin_memory(); return t.numpy()
def bps_of(blk):
    parts, boff, o = [], [], 0
    for r in range(blk.nreads):
        p = synth.pack_2bit_dazz(blk.read(r)); boff.append(o); parts.append(p); o += len(p)
    return pinned(np.concatenate(parts)), np.array(boff, np.int64)
rb, ro = bps_of(ref); qb, qo = bps_of(reads)
P = dict(tspace=100, minlen=1000, e=0.7, k=20)
T = time.perf_counter
for it in range(6):
    t0 = T(); a = dazzler.HostBlock(ref.off, bps=rb, boff=ro); b = dazzler.HostBlock(reads.off, bps=qb, boff=qo); t1 = T()
    rec, toff, tr, st = dazzler.align_host(a, b, **P); t2 = T()
    g = dazzler.Block(reads.off, bps=qb, boff=qo); t3 = T(); g.free()
    print("iter %d  desc %.2f ms  align_host %.2f ms (device ms_total %.2f)  | reads upload alone %.2f ms" % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, st["ms_total"], (t3 - t2) * 1e3), flush=True)
