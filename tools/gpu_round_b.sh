#!/bin/bash
# tests + bench at both extension register budgets + launch list
mkdir -p gpurun_out
T=${1:-r2b}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
DN_EXT_CTAS=6 timeout 600 python bench.py --profile > gpurun_out/${T}_bench_ctas6.json 2> gpurun_out/${T}_bench_ctas6.err
DN_NO_PACKED=1 timeout 600 python bench.py --profile > gpurun_out/${T}_bench_nopacked.json 2> gpurun_out/${T}_bench_nopacked.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --profile > gpurun_out/${T}_prof.log 2>&1
tail -3 gpurun_out/${T}_tests.log
for f in bench bench_ctas6 bench_nopacked; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_$f.json")); print("$f", d["value"], d["ms_per_step"], d.get("stage_ms_per_step"), d.get("parity_on_config"), d.get("e2e"))
except Exception as e: print("$f failed", e)
PY
done
tail -2 gpurun_out/${T}_bench.err
