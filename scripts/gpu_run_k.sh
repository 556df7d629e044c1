#!/bin/bash
mkdir -p gpurun_out
{
timeout 200 python scripts/single_stream_times.py
timeout 300 python -m pytest tests/test_channelizer.py -q -m gpu --timeout 250 2>&1 | tail -5
timeout 300 python bench.py --chan-leg
} > gpurun_out/r2k.log 2>&1
cat gpurun_out/r2k.log | cut -c1-1500
