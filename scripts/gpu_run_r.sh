#!/bin/bash
# round 2, last GPU call: the final code state - whole GPU suite, compute-sanitizer on the cases whose kernels changed after
# call J (cluster-shared acquisition, AM decoder / interleaver, channeliser epilogue), the bench with all legs, and fresh
# ncu captures of k_am and k_channelize
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2r_smi.txt 2>&1
( time python -m pytest tests -m gpu -x -q --timeout 200 ) > gpurun_out/r2r_gpu_tests.log 2>&1
tail -4 gpurun_out/r2r_gpu_tests.log
san() {
  local name=$1 tool=$2 to=$3; shift 3
  ( time timeout $to compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 "$@" ) > gpurun_out/r2r_${tool}_${name}.log 2>&1
  echo "$tool $name rc=$?"; grep -h "SUMMARY\|gate\|passed\|failed" gpurun_out/r2r_${tool}_${name}.log | tail -2 | cut -c1-200
}
san cluster memcheck 300 python scripts/sanitize_case.py mp1 2 1
san cluster racecheck 420 python scripts/sanitize_case.py mp1 2 1
san am memcheck 300 python scripts/sanitize_am.py
san am racecheck 500 python scripts/sanitize_am.py
san chan memcheck 300 python -m pytest tests/test_channelizer.py -q -m gpu -k "kernel_equals" --timeout 280
( time python bench.py ) > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
tail -c 400 gpurun_out/r2r_bench.json; tail -2 gpurun_out/r2r_bench.err
bash scripts/profile_r2b.sh > gpurun_out/r2r_profile.log 2>&1
tail -4 gpurun_out/r2r_profile.log
