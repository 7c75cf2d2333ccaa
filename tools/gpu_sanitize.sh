#!/bin/bash
# compute-sanitizer on the kernels and host flows that changed in round 2
mkdir -p gpurun_out
T=${1:-r2s}
DN_NO_ARENA=1 timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_pile.py tests/test_gpu_files.py -m gpu -q -x \
   -k "ref_vs_reads or long_kmers or edge_cases or masked or resident or transposed or batch_entry or oversized or consensus_matches or getDamapping or block_level" > gpurun_out/${T}_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/${T}_memcheck.log | tail -3
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_pile.py -m gpu -q -x \
   -k "ref_vs_reads or transposed or long_kmers" > gpurun_out/${T}_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/${T}_racecheck.log | tail -3
DN_EXT_TMA=1 timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ref_vs_reads or edge_cases" > gpurun_out/${T}_memcheck_tma.log 2>&1
echo "memcheck tma rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/${T}_memcheck_tma.log | tail -3
