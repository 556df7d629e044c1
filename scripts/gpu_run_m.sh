#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_stages.py tests/test_gpu_am.py tests/test_zz_gpu_l2.py -q -m gpu --timeout 200 -k "k9 or am or AM" 2>&1 | tail -4
timeout 300 python bench.py --am-leg 2> gpurun_out/r2m_am.err
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 10 python scripts/sanitize_am.py 2>&1 | tail -4
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 10 python scripts/sanitize_am.py 2>&1 | tail -3
} > gpurun_out/r2m.log 2>&1
cut -c1-1500 gpurun_out/r2m.log
