#!/bin/bash
# round 2, GPU call A: the whole GPU suite (incl. the new config-3 / config-5 / edge / RS tests), the bench with all
# legs, and compute-sanitizer on bounded config-5 / config-3 decodes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a_gpu_tests.log 2>&1
tail -5 gpurun_out/r2a_gpu_tests.log
( time python bench.py ) > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 600 gpurun_out/r2a_bench.json
for tool in memcheck racecheck; do
  for k in mp1 mp3; do
    ( time timeout 420 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 python scripts/sanitize_case.py $k ) \
        > gpurun_out/r2a_${tool}_${k}.log 2>&1
    echo "$tool $k rc=$?"; tail -4 gpurun_out/r2a_${tool}_${k}.log
  done
done
