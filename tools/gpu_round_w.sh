#!/bin/bash
# launch list of the step, stage times of the pile-up batch (stepwise + batch entry point), ncu --set full of the lookup kernels
mkdir -p gpurun_out
T=${1:-r2w}
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --profile > gpurun_out/${T}_prof.log 2>&1
timeout 300 python tools/time_pileups.py > gpurun_out/${T}_pileups.log 2>&1
DN_TRACE=1 timeout 300 python tools/time_batch.py 3 > gpurun_out/${T}_batch.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${T}_batch_launches.csv python tools/time_batch.py 2 > gpurun_out/${T}_prof4.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_lookup_(count|emit)_p' -s 3 -c 3 -o gpurun_out/${T}_lookup python bench.py --steps 1 --warmup 1 --profile > gpurun_out/${T}_prof2.log 2>&1
tail -4 gpurun_out/${T}_pileups.log; tail -30 gpurun_out/${T}_batch.log
