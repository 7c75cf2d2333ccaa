#!/bin/bash
# bench main line, c3 small at N=1, c3 + default at N=2
mkdir -p gpurun_out
T=${1:-r2f}; N=${2:-2}
timeout 600 python bench.py > gpurun_out/${T}_bench1.json 2> gpurun_out/${T}_bench1.err; echo "bench1 rc=$?"; tail -2 gpurun_out/${T}_bench1.err
timeout 900 python bench.py --config c3 --c3-blocks 2 --steps 3 --warmup 2 > gpurun_out/${T}_c3_1.json 2> gpurun_out/${T}_c3_1.err; echo "c3 N=1 rc=$?"; tail -2 gpurun_out/${T}_c3_1.err
if [ $N -gt 1 ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config c3 --c3-blocks $((2*N)) --steps 3 --warmup 2 > gpurun_out/${T}_c3_$N.json 2> gpurun_out/${T}_c3_$N.err; echo "c3 N=$N rc=$?"; tail -2 gpurun_out/${T}_c3_$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${T}_bench$N.json 2> gpurun_out/${T}_bench$N.err; echo "bench$N rc=$?"; tail -2 gpurun_out/${T}_bench$N.err
fi
for f in bench1 c3_1 c3_$N bench$N; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_$f.json")); print("$f", "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "cons", (d.get("consensus") or {}).get("value"), (d.get("consensus") or {}).get("ms_per_batch"), (d.get("consensus") or {}).get("pile_ups_with_both_flanks_aligned"), (d.get("consensus") or {}).get("cpu_baseline"))
except Exception as e: print("$f failed", e)
PY
done
