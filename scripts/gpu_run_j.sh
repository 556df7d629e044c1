#!/bin/bash
# round 2, GPU call J: the final state - whole GPU suite, compute-sanitizer on bounded decodes of every kernel family
# (many-stream FM, cluster-per-stream FM, MP3, AM, channeliser), then the bench with all legs
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2j_smi.txt 2>&1
( time python -m pytest tests -m gpu -x -q --timeout 200 ) > gpurun_out/r2j_gpu_tests.log 2>&1
tail -5 gpurun_out/r2j_gpu_tests.log
san() {   # name tool timeout command...
  local name=$1 tool=$2 to=$3; shift 3
  ( time timeout $to compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 "$@" ) > gpurun_out/r2j_${tool}_${name}.log 2>&1
  echo "$tool $name rc=$?"; grep -h "SUMMARY\|gate\|passed\|failed" gpurun_out/r2j_${tool}_${name}.log | tail -3
}
san mp1 memcheck 300 python scripts/sanitize_case.py mp1
san mp1 racecheck 420 python scripts/sanitize_case.py mp1
san mp3 memcheck 300 python scripts/sanitize_case.py mp3
san mp3 racecheck 420 python scripts/sanitize_case.py mp3
san cluster memcheck 300 python scripts/sanitize_case.py mp1 2 1
san cluster racecheck 420 python scripts/sanitize_case.py mp1 2 1
san am memcheck 300 python scripts/sanitize_am.py
san am racecheck 500 python scripts/sanitize_am.py
san chan memcheck 300 python -m pytest tests/test_channelizer.py -q -m gpu -k "kernel_equals" --timeout 280
( time python bench.py ) > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
tail -c 1500 gpurun_out/r2j_bench.json
tail -3 gpurun_out/r2j_bench.err
