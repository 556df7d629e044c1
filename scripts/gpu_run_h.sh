#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_channelizer.py -m gpu -q --timeout 250 ) > gpurun_out/r2h_chan_tests.log 2>&1
tail -25 gpurun_out/r2h_chan_tests.log | cut -c1-500
( time timeout 600 python -m pytest tests/test_gpu_am.py tests/test_dropin.py -m gpu -q --timeout 150 ) > gpurun_out/r2h_am_tests.log 2>&1
tail -4 gpurun_out/r2h_am_tests.log | cut -c1-300
( time timeout 600 python bench.py --am-leg --am-streams 256 --am-frames 12 --steps 3 ) > gpurun_out/r2h_am_leg.json 2> gpurun_out/r2h_am_leg.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2h_am_leg.json') if l.startswith('{')][-1])
print('AM', d.get('value'), d.get('ms_per_step'), d.get('phases_us_per_stream_block_at_1965MHz'), d.get('parity_gate',{}).get('ok'))
PY
