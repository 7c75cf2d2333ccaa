#!/bin/bash
# 2 GPUs: all GPU tests (incl. the in-library NCCL gather), bench at N=1 and N=2
mkdir -p gpurun_out
T=${1:-r2c}; N=${2:-2}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -4 gpurun_out/${T}_tests.log
timeout 600 python bench.py > gpurun_out/${T}_bench1.json 2> gpurun_out/${T}_bench1.err; echo "bench1 rc=$?"
BENCH_DEBUG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${T}_bench$N.json 2> gpurun_out/${T}_bench$N.err; echo "bench$N rc=$?"
for f in bench1 bench$N; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_$f.json")); print("$f", "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d.get("stage_ms_per_step"), d.get("parity_on_config"), "cons", (d.get("consensus") or {}).get("value"))
except Exception as e: print("$f failed", e)
PY
done
grep -E "e2e iter" gpurun_out/${T}_bench$N.err | tail -4
tail -3 gpurun_out/${T}_bench$N.err
