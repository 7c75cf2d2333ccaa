#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2p}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log; tail -3 gpurun_out/${T}_tests.log
for v in 4 6 8; do DN_RADIX_CTAS=$v timeout 300 python bench.py --profile > gpurun_out/${T}_bench_radix$v.json 2> gpurun_out/${T}_bench_radix$v.err; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --profile > gpurun_out/${T}_prof.log 2>&1
for f in bench_radix4 bench_radix6 bench_radix8; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_$f.json")); print("$f", d["value"], d["ms_per_step"], d.get("stage_ms_per_step"), d.get("parity_on_config",{}).get("identical"))
except Exception as e: print("$f failed", e)
PY
done
