#!/bin/bash
# TMA-staged extension variant: parity tests + timing next to the default path
mkdir -p gpurun_out
T=${1:-r2d}
DN_EXT_TMA=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${T}_tests_tma.log 2>&1; echo "tma tests rc=$?"; tail -3 gpurun_out/${T}_tests_tma.log
timeout 300 python bench.py --profile > gpurun_out/${T}_bench_ldg.json 2> gpurun_out/${T}_bench_ldg.err
DN_EXT_TMA=1 timeout 300 python bench.py --profile > gpurun_out/${T}_bench_tma.json 2> gpurun_out/${T}_bench_tma.err
DN_EXT_TMA=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_extend32' -s 2 -c 1 -o gpurun_out/${T}_ext_tma python bench.py --steps 1 --warmup 1 --profile > gpurun_out/${T}_prof_tma.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_extend32' -s 2 -c 1 -o gpurun_out/${T}_ext_ldg python bench.py --steps 1 --warmup 1 --profile > gpurun_out/${T}_prof_ldg.log 2>&1
for f in bench_ldg bench_tma; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_$f.json")); print("$f", "value", d["value"], "ms", d["ms_per_step"], d.get("stage_ms_per_step"), d.get("parity_on_config"))
except Exception as e: print("$f failed", e)
PY
done
