#!/bin/bash
mkdir -p gpurun_out
xz -dc oracle/_ref/sample.xz > /tmp/sample.cu8
BP=nrsc5_b200/dropin/_build/bench_pipe
LIB=nrsc5_b200/dropin/_build/libnrsc5.so
{
for c in 4 1 2; do
  echo "== cluster $c"
  NRSC5_B200_CLUSTER=$c NRSC5_B200_TRACE=1 $BP $LIB /tmp/sample.cu8 --reps 3
done
echo "== sync mode"
NRSC5_B200_SYNC=1 NRSC5_B200_TRACE=1 $BP $LIB /tmp/sample.cu8 --reps 2
echo "== reference"
$BP oracle/_ref/libnrsc5_ref.so /tmp/sample.cu8 --reps 2
} > gpurun_out/r2c_pipe.log 2>&1
cat gpurun_out/r2c_pipe.log | cut -c1-600
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2c_gpu_tests.log 2>&1
tail -15 gpurun_out/r2c_gpu_tests.log
( time python bench.py --no-am ) > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
tail -c 1200 gpurun_out/r2c_bench.json
