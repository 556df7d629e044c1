#!/bin/bash
# First B200 measurements of the L2 row and of the new bench legs - one gpurun call, about 6 minutes:
#
#   gpurun --timeout 600 -- 'bash scripts/profile_l2.sh'
#
# Writes into gpurun_out/ (merged back by gpurun); copy what is to be judged into profiles/ afterwards
# (python scripts/summarize_profiles.py r2 understands launches_r2.csv / prof_r2.ncu-rep).
set -u
mkdir -p gpurun_out
# 1. the bench line with its three supplementary legs (l2_on_device, am_config4, mp3_config3)
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_l2.json 2> gpurun_out/bench_l2.err
# 2. the L2 leg alone, under ncu: launch list (device time per launch; cold cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --l2-leg --streams 128 --frames 4 --steps 1 > gpurun_out/l2_leg_under_ncu.log 2>&1
# 3. one full capture of k_l2 (and of the kernels around it in that pass)
ncu --set full --clock-control none --import-source on -k regex:k_l2 -c 2 -o gpurun_out/prof_r2_l2 -f \
    python bench.py --l2-leg --streams 128 --frames 4 --steps 1 > gpurun_out/l2_leg_full_ncu.log 2>&1
ncu -i gpurun_out/prof_r2_l2.ncu-rep --page raw --csv > gpurun_out/prof_r2_l2_raw.csv 2>/dev/null
# 4. parity on the box
python -m pytest tests/test_zz_gpu_l2.py tests/test_zz_reference_cli.py -m gpu -q -p no:cacheprovider > gpurun_out/l2_tests.log 2>&1
tail -3 gpurun_out/l2_tests.log
head -c 600 gpurun_out/bench_l2.json
