#!/bin/bash
# full -m gpu suite, then the launch list of ScalarBaseMult at n = 4096 (config 1) and a short bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"base_mult|finish_affine" -c 60 --csv \
    --log-file gpurun_out/r2_sbm4096_launches.csv python scripts/bench_small.py > gpurun_out/r2_sbm4096.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_sbm4096_launches.csv
timeout 200 python scripts/bench_small.py 2>&1 | tail -12
