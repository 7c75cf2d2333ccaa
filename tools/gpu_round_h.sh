#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2h}
timeout 1500 python tools/run_example.py tests/golden/_big/example_full.npz > gpurun_out/${T}_example_full.json 2> gpurun_out/${T}_example_full.err; echo "example rc=$?"; head -c 1800 gpurun_out/${T}_example_full.json; tail -3 gpurun_out/${T}_example_full.err
timeout 2400 python tools/sweep_align.py --ref-mbp 200 --reads-mbp 200 --resident-index --out gpurun_out/${T}_sweep_200x200_resident.jsonl > gpurun_out/${T}_sweep.log 2>&1; echo "sweep rc=$?"; tail -3 gpurun_out/${T}_sweep.log | cut -c1-400
