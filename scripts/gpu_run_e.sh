#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  ( time timeout 400 compute-sanitizer --tool $tool --print-limit 15 python scripts/sanitize_am.py ma3_clean ) > gpurun_out/r2e_am_${tool}.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error|records" gpurun_out/r2e_am_${tool}.log | head -12
done
xz -dc oracle/_ref/sample.xz > /tmp/sample.cu8
BP=nrsc5_b200/dropin/_build/bench_pipe
LIB=nrsc5_b200/dropin/_build/libnrsc5.so
{
for c in 4 1; do
  echo "== cluster $c"
  NRSC5_B200_CLUSTER=$c NRSC5_B200_TRACE=1 timeout 120 $BP $LIB /tmp/sample.cu8 --reps 3 2>&1 | tail -4
done
echo "== cluster 4, library sincos/atan2 in the Costas loops"
NRSC5_B200_DBG=2 NRSC5_B200_TRACE=1 timeout 120 $BP $LIB /tmp/sample.cu8 --reps 3 2>&1 | tail -3
} > gpurun_out/r2e_pipe.log 2>&1
cat gpurun_out/r2e_pipe.log | cut -c1-700
for dbg in 0 2; do
( timeout 300 python bench.py --no-am --no-l2 --no-mp3 --no-dropin --no-cpu-baseline --no-e2e --dbg $dbg ) > gpurun_out/r2e_bench_quick_dbg$dbg.json 2> gpurun_out/r2e_bench_quick.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2e_bench_quick_dbg$dbg.json') if l.startswith('{')][-1])
print('dbg $dbg', d['value'], d['ms_per_step'], {k: round(v['us_per_call'],2) for k,v in d['roofline']['front_phases_at_1965MHz'].items()})
PY
done
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r2e_gpu_tests.log 2>&1
tail -12 gpurun_out/r2e_gpu_tests.log
