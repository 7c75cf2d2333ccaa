#!/bin/bash
# Quick iteration call: the alignment parity tests + the device-resident bench arm (optionally with env variants).
mkdir -p gpurun_out
T=${1:-q}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -3 gpurun_out/${T}_tests.log
timeout 300 python bench.py --profile > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
for v in "$@"; do
  case "$v" in *=*) env $v timeout 300 python bench.py --profile > gpurun_out/${T}_bench_${v%%=*}.json 2> gpurun_out/${T}_bench_${v%%=*}.err;; esac
done
for f in gpurun_out/${T}_bench*.json; do python - <<PY
import json
try:
    d=json.load(open("$f")); print("$f", round(d["value"],3), round(d["ms_per_step"],3), d.get("stage_ms_per_step"), d.get("parity_on_config",{}).get("identical"), d.get("gpu_launches"))
except Exception as e: print("$f failed", e)
PY
done
tail -2 gpurun_out/${T}_bench.err
