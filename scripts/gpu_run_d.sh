#!/bin/bash
mkdir -p gpurun_out
xz -dc oracle/_ref/sample.xz > /tmp/sample.cu8
BP=nrsc5_b200/dropin/_build/bench_pipe
LIB=nrsc5_b200/dropin/_build/libnrsc5.so
{
for c in 4 2 1; do
  echo "== cluster $c"
  NRSC5_B200_CLUSTER=$c NRSC5_B200_TRACE=1 timeout 120 $BP $LIB /tmp/sample.cu8 --reps 3 2>&1 | tail -4
done
} > gpurun_out/r2d_pipe.log 2>&1
cat gpurun_out/r2d_pipe.log | cut -c1-700
( time timeout 300 python bench.py --no-am --no-l2 --no-mp3 --no-dropin --no-cpu-baseline --no-e2e ) > gpurun_out/r2d_bench_quick.json 2> gpurun_out/r2d_bench_quick.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2d_bench_quick.json') if l.startswith('{')][-1])
print('quick', d['value'], d['ms_per_step'], {k: round(v['us_per_call'],2) for k,v in d['roofline']['front_phases_at_1965MHz'].items()})
PY
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r2d_gpu_tests.log 2>&1
tail -12 gpurun_out/r2d_gpu_tests.log
( time timeout 1200 python bench.py ) > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
tail -c 1500 gpurun_out/r2d_bench.json
