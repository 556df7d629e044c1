#!/bin/bash
# Round-2 ncu evidence (one gpurun call, ~10 min): launch list of the headline step, one `--set full` capture each of
# k_stream (one CTA per stream), k_l2, k_am and k_channelize, raw CSV pages extracted on the box.
#   gpurun --timeout 1500 -- 'bash scripts/profile_r2.sh'
# Afterwards here:  python scripts/summarize_r2.py   (writes profiles/r2_*.md / .json / .csv)
set -u
mkdir -p gpurun_out
Q="--no-cpu-baseline --no-e2e --no-l2 --no-am --no-mp3 --no-dropin --no-chan"
# 1. every launch of one headline step with its device time (the gate step, 3 warm-up steps and 1 timed step run under ncu)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py $Q --steps 1 --warmup 3 > gpurun_out/r2_launches_run.log 2>&1
# 2. k_stream: the first two launches of a step are the loaded ones (16 blocks x 128 streams each)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream -c 2 -f -o gpurun_out/prof_r2_k_stream \
    python bench.py $Q --steps 1 --warmup 3 > gpurun_out/r2_prof_k_stream.log 2>&1
ncu -i gpurun_out/prof_r2_k_stream.ncu-rep --page raw --csv > gpurun_out/prof_r2_k_stream_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r2_k_stream.ncu-rep --page source --csv > gpurun_out/prof_r2_k_stream_source.csv 2>/dev/null
# 3. k_l2 (the L2 leg: frames with ~45 packets)
timeout 900 ncu --set full --clock-control none -k regex:k_l2 -c 3 -f -o gpurun_out/prof_r2_k_l2 \
    python bench.py --l2-leg --streams 128 --frames 4 --steps 1 > gpurun_out/r2_prof_k_l2.log 2>&1
ncu -i gpurun_out/prof_r2_k_l2.ncu-rep --page raw --csv > gpurun_out/prof_r2_k_l2_raw.csv 2>/dev/null
rm -f gpurun_out/prof_r2_k_l2.ncu-rep
# 4. k_am (a reduced AM job: 32 channels x 10 frames - ncu replays the kernel ~40 times)
timeout 900 ncu --set full --clock-control none -k regex:k_am -c 1 -f -o gpurun_out/prof_r2_k_am \
    python bench.py --am-leg --am-streams 32 --am-frames 10 --steps 1 > gpurun_out/r2_prof_k_am.log 2>&1
ncu -i gpurun_out/prof_r2_k_am.ncu-rep --page raw --csv > gpurun_out/prof_r2_k_am_raw.csv 2>/dev/null
rm -f gpurun_out/prof_r2_k_am.ncu-rep
# 5. k_channelize
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_channelize -c 1 -s 1 -f -o gpurun_out/prof_r2_k_channelize \
    python bench.py --chan-leg --steps 5 > gpurun_out/r2_prof_k_chan.log 2>&1
ncu -i gpurun_out/prof_r2_k_channelize.ncu-rep --page raw --csv > gpurun_out/prof_r2_k_channelize_raw.csv 2>/dev/null
ls -la gpurun_out | tail -20
