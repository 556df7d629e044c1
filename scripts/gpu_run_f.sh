#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
  ( timeout 300 python -m pytest tests/test_gpu_am.py tests/test_dropin.py -m gpu -q -x --timeout 120 ) > gpurun_out/r2f_am_$i.log 2>&1
  echo "am round $i: $(tail -1 gpurun_out/r2f_am_$i.log)"
done
( time timeout 900 python -m pytest tests -m gpu -q --timeout 150 ) > gpurun_out/r2f_gpu_tests.log 2>&1
tail -12 gpurun_out/r2f_gpu_tests.log
xz -dc oracle/_ref/sample.xz > /tmp/sample.cu8
BP=nrsc5_b200/dropin/_build/bench_pipe
LIB=nrsc5_b200/dropin/_build/libnrsc5.so
{
for c in 4 1; do
  echo "== cluster $c"
  NRSC5_B200_CLUSTER=$c NRSC5_B200_TRACE=1 timeout 120 $BP $LIB /tmp/sample.cu8 --reps 3 2>&1 | tail -4
done
} > gpurun_out/r2f_pipe.log 2>&1
cat gpurun_out/r2f_pipe.log | cut -c1-700
( time timeout 1500 python bench.py ) > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
tail -c 2500 gpurun_out/r2f_bench.json
