#!/bin/bash
# Evidence run for the round's final state: all GPU tests, the bench line, the reference arm, the step's launch list,
# ncu --set full of the hot kernels (alignment step and pile-up batch), compute-sanitizer on the kernels that changed.
mkdir -p gpurun_out
T=${1:-r2f}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "reference arm rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --profile > gpurun_out/${T}_prof.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${T}_batch_launches.csv python tools/time_batch.py 2 > gpurun_out/${T}_prof4.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_extend32|k_lookup_count_p|k_lookup_emit_p' -s 8 -c 4 -o gpurun_out/${T}_hot python bench.py --steps 1 --warmup 1 --profile > gpurun_out/${T}_prof2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_radix_onesweep|k_scan_chained|k_segsort_radix|k_retire' -s 30 -c 8 -o gpurun_out/${T}_seed python bench.py --steps 1 --warmup 1 --profile > gpurun_out/${T}_prof3.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_cons_vote_bv|k_qv|k_cons_count|k_dust_windows' -s 4 -c 4 -o gpurun_out/${T}_cons python tools/time_batch.py 2 > gpurun_out/${T}_prof5.log 2>&1
DN_NO_ARENA=1 timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_pile.py tests/test_abi.py -m gpu -q -x \
   -k "ref_vs_reads or long_kmers or edge_cases or bridging or batch_entry or consensus_matches or qvs_match or align_host or host" > gpurun_out/${T}_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/${T}_memcheck.log | tail -3
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ref_vs_reads or long_kmers" > gpurun_out/${T}_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/${T}_racecheck.log | tail -3
tail -3 gpurun_out/${T}_tests.log; head -c 300 gpurun_out/${T}_bench.json; echo; cat gpurun_out/${T}_bench_reference.json | head -c 600; echo; tail -3 gpurun_out/${T}_bench.err
