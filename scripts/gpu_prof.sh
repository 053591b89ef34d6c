#!/bin/bash
set -x
mkdir -p gpurun_out
# launch list of one bench run (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
# full capture of the dominant kernel (third launch = warm)
LOG2N=18 PASSES=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dsm -s 2 -c 1 \
    -o gpurun_out/prof_dsm -f python scripts/prof_dsm.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
