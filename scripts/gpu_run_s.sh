#!/bin/bash
# round 2, closing GPU call: whole GPU suite, compute-sanitizer on the cluster case, the bench with all legs (the line kept in
# profiles/r2_bench_final.json)
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q --timeout 200 ) > gpurun_out/r2s_gpu_tests.log 2>&1
tail -4 gpurun_out/r2s_gpu_tests.log
for tool in memcheck racecheck; do
  ( time timeout 300 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 python scripts/sanitize_case.py mp1 2 1 ) > gpurun_out/r2s_${tool}_cluster.log 2>&1
  echo "$tool cluster rc=$?"; grep -h "SUMMARY" gpurun_out/r2s_${tool}_cluster.log
done
( time python bench.py ) > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
tail -c 300 gpurun_out/r2s_bench.json; tail -2 gpurun_out/r2s_bench.err
