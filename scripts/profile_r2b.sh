#!/bin/bash
# Round-2 ncu evidence, second call: k_am with the new decoder and k_channelize (whose first capture died on the bench
# gate's own indexing mistake).  Afterwards here:  python scripts/summarize_r2.py
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:^k_am$ -c 1 -f -o gpurun_out/prof_r2_k_am \
    python bench.py --am-leg --am-streams 32 --am-frames 10 --steps 1 > gpurun_out/r2_prof_k_am.log 2>&1
ncu -i gpurun_out/prof_r2_k_am.ncu-rep --page raw --csv > gpurun_out/prof_r2_k_am_raw.csv 2>/dev/null
rm -f gpurun_out/prof_r2_k_am.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_channelize -c 1 -s 1 -f -o gpurun_out/prof_r2_k_channelize \
    python bench.py --chan-leg --steps 5 > gpurun_out/r2_prof_k_chan.log 2>&1
ncu -i gpurun_out/prof_r2_k_channelize.ncu-rep --page raw --csv > gpurun_out/prof_r2_k_channelize_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r2_k_channelize.ncu-rep --page source --csv > gpurun_out/prof_r2_k_channelize_source.csv 2>/dev/null
rm -f gpurun_out/prof_r2_k_channelize.ncu-rep
tail -3 gpurun_out/r2_prof_k_am.log gpurun_out/r2_prof_k_chan.log
ls -la gpurun_out | grep prof_r2
