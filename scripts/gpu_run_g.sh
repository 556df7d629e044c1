#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_channelizer.py -m gpu -q -x --timeout 200 ) > gpurun_out/r2g_chan_tests.log 2>&1
tail -30 gpurun_out/r2g_chan_tests.log | cut -c1-400
( time timeout 600 python bench.py --am-leg --am-streams 256 --am-frames 12 --steps 3 ) > gpurun_out/r2g_am_leg.json 2> gpurun_out/r2g_am_leg.err
tail -c 1800 gpurun_out/r2g_am_leg.json
