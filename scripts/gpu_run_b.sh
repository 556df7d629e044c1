#!/bin/bash
mkdir -p gpurun_out
( time python scripts/diag_dropout.py ) > gpurun_out/r2b_diag.log 2>&1
tail -12 gpurun_out/r2b_diag.log
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2b_gpu_tests.log 2>&1
tail -15 gpurun_out/r2b_gpu_tests.log
( time python bench.py --no-am ) > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -c 1500 gpurun_out/r2b_bench.json
