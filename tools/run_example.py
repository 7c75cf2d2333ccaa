"""BASELINE configs[0]: the reference's example dataset (27.9 Mbp scaffold, 147 gaps; example/Makefile) through
map -> collect filters -> PileUpDb -> dn_process_pileups -> InsertionDb.  The reads come from our generator with the
example's simulator settings (-m25000 -s12500 -e.13 -c20; the DAZZ_DB simulator is absent).
Input: an .npz made by tests/golden/make_example_excerpt.py --full (not committed: 6.9 MB).
    python tools/run_example.py tests/golden/_big/example_full.npz > gpurun_out/example_full.json"""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from dentist_b200 import dazzler, synth
from tests import pipeline_util

z = np.load(sys.argv[1])
n = int(z["length"]); p = z["packed"]
scaffold = np.stack([(p >> 6) & 3, (p >> 4) & 3, (p >> 2) & 3, p & 3], 1).reshape(-1)[:n].astype(np.uint8)
gaps = [tuple(int(x) for x in g) for g in z["gaps"]]
dazzler.init(0)
t0 = time.perf_counter()
ref, meta = synth.contigs_from([scaffold], [gaps])
reads, _ = synth.simulate_reads([scaffold], 20, 25000, 12500, 0.13, 19339)
t1 = time.perf_counter()
with tempfile.TemporaryDirectory() as d:
    l0 = dazzler.launch_count()
    ins, skipped, piles, st = pipeline_util.close_gaps(ref, reads, len(gaps), d, k=20, minlen=1000)
    launches = dazzler.launch_count() - l0
t2 = time.perf_counter()
idents = []
closed = {}
for i in ins:
    g = i["start"][0] - 1
    if i["start"] == (g + 1, "end") and i["end"] == (g + 2, "begin") and len(i["overlaps"]) == 2:
        ident, lc, lt = pipeline_util.gap_identity(i, scaffold, meta, g)
        closed[g] = ident; idents.append(ident)
print(json.dumps({"dataset": str(z["source"]), "assembly_bp": n, "gaps": len(gaps), "contigs": int(ref.nreads), "reads": int(reads.nreads),
                  "read_bp": int(reads.total), "pile_ups": len(piles), "pile_ups_with_>=3_reads": sum(len(pl) >= 3 for pl in piles),
                  "insertions": len(ins), "gaps_closed_as_gap_insertions": len(closed), "skipped": {str(k): v for k, v in skipped.items()},
                  "identity_min": float(min(idents)) if idents else None, "identity_median": float(np.median(idents)) if idents else None,
                  "gaps_closed_at_>=0.97_identity": int(sum(v >= 0.97 for v in idents)), "mapping": {k: int(v) for k, v in st.items()},
                  "seconds": {"generate_reads": t1 - t0, "map_collect_process": t2 - t1}, "gpu_launches": launches}))
