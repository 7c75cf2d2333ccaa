#!/bin/bash
# One gpurun call: GPU tests, bench line, launch list, ncu --set full of the final lookup kernels (packed index) and the
# consensus / QV kernels.
mkdir -p gpurun_out
T=${1:-r2t}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --profile > gpurun_out/${T}_prof.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_lookup_(count|emit)_p|k_radix_scatter|k_segsort_radix' -s 12 -c 8 -o gpurun_out/${T}_seed python bench.py --steps 1 --warmup 1 --profile > gpurun_out/${T}_prof2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_cons_vote|k_qv|k_cons_count|k_cons_tasks' -c 4 -o gpurun_out/${T}_cons python tools/time_pileups.py > gpurun_out/${T}_prof3.log 2>&1
tail -3 gpurun_out/${T}_tests.log; head -c 400 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
