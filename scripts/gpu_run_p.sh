#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_edge.py tests/test_zz_reference_cli.py -q -m gpu --timeout 200 2>&1 | tail -3
timeout 300 python bench.py --dropin-leg | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k,v in j.items():
    if isinstance(v,dict):
        for kk,vv in v.items():
            print(k,kk,(vv.get('x_realtime'),vv.get('seconds')) if isinstance(vv,dict) else vv)
"
xz -dc oracle/_ref/sample.xz > /tmp/sample.cu8
NRSC5_B200_TRACE=1 timeout 120 nrsc5_b200/dropin/_build/bench_pipe nrsc5_b200/dropin/_build/libnrsc5.so /tmp/sample.cu8 --reps 4 2>&1 | grep -a "k_stream Mcycles" | tail -1
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-l2 --no-am --no-mp3 --no-dropin --no-chan --steps 10 | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(j['value'], j['ms_per_step'], j['parity_gate']['ok'])"
} > gpurun_out/r2p.log 2>&1
cat gpurun_out/r2p.log | cut -c1-400
