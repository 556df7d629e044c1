#!/bin/bash
mkdir -p gpurun_out
{
timeout 200 python scripts/k9_cycles.py 2>&1 | awk '/^len/ {print; c=0} /^k9 job/ {c++; if (c<=2) print}'
timeout 600 python -m pytest tests/test_gpu_stages.py tests/test_gpu_am.py -q -m gpu --timeout 200 -k "k9 or am or AM" 2>&1 | tail -3
timeout 300 python bench.py --am-leg 2> gpurun_out/r2n_am.err | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(j['value'], j['ms_per_step'], j['phases_us_per_stream_block_at_1965MHz'])"
} > gpurun_out/r2n.log 2>&1
cut -c1-1500 gpurun_out/r2n.log
