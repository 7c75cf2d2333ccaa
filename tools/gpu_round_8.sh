#!/bin/bash
# 8 GPUs: the default bench line (configs[1] x 8, weak scaling) and configs[2] block-sharded
mkdir -p gpurun_out
T=${1:-r2k}; N=${2:-8}
BENCH_DEBUG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${T}_bench$N.json 2> gpurun_out/${T}_bench$N.err; echo "bench$N rc=$?"
grep -E "e2e iter (1[0-9]|2[01]) " gpurun_out/${T}_bench$N.err | grep -v "^$" | tail -3 | cut -c1-200
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --config c3 --steps 5 --warmup 2 > gpurun_out/${T}_c3_$N.json 2> gpurun_out/${T}_c3_$N.err; echo "c3 N=$N rc=$?"; tail -2 gpurun_out/${T}_c3_$N.err | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --config c4 --steps 5 --warmup 2 > gpurun_out/${T}_c4_$N.json 2> gpurun_out/${T}_c4_$N.err; echo "c4 N=$N rc=$?"; tail -2 gpurun_out/${T}_c4_$N.err | cut -c1-300
for f in bench$N c3_$N c4_$N; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_$f.json")); print("$f", "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"], "cons", (d.get("consensus") or {}).get("value"))
except Exception as e: print("$f failed", e)
PY
done
df -h /dev/shm | tail -1; nproc; free -g | head -2 | tail -1
