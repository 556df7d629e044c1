#!/bin/bash
mkdir -p gpurun_out
xz -dc oracle/_ref/sample.xz > /tmp/sample.cu8
BP=nrsc5_b200/dropin/_build/bench_pipe
LIB=nrsc5_b200/dropin/_build/libnrsc5.so
{
for c in 4 1; do
  echo "== cluster $c"
  NRSC5_B200_CLUSTER=$c NRSC5_B200_TRACE=1 timeout 120 $BP $LIB /tmp/sample.cu8 --reps 8 2>&1 | tail -5
done
} > gpurun_out/r2i_pipe.log 2>&1
cat gpurun_out/r2i_pipe.log | cut -c1-600
bash scripts/profile_r2.sh > gpurun_out/r2i_profile.log 2>&1
tail -25 gpurun_out/r2i_profile.log
