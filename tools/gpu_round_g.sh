#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2g}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log; tail -5 gpurun_out/${T}_tests.log
DN_TRACE=1 timeout 300 python tools/time_batch.py 3 > gpurun_out/${T}_batch_trace.log 2>&1; grep "^iter" gpurun_out/${T}_batch_trace.log
timeout 1500 python tools/run_example.py tests/golden/_big/example_full.npz > gpurun_out/${T}_example_full.json 2> gpurun_out/${T}_example_full.err; echo "example rc=$?"; cat gpurun_out/${T}_example_full.json | head -c 1500; tail -3 gpurun_out/${T}_example_full.err
