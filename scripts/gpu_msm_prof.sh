#!/bin/bash
# MSM: parity tests, wall-clock per call, then per-kernel durations at 2^20 and 2^17 points
# (ncu launch lists; not bench values)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k msm > gpurun_out/msm_tests.log 2>&1; tail -3 gpurun_out/msm_tests.log
for L in 20 17; do
  LOG2N=$L TIME=1 timeout 300 python scripts/prof_msm.py > gpurun_out/msm_time_$L.log 2>&1; tail -2 gpurun_out/msm_time_$L.log
  LOG2N=$L timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/msm_launches_$L.csv python scripts/prof_msm.py > gpurun_out/msm_prof_$L.log 2>&1
  python scripts/launch_summary.py gpurun_out/msm_launches_$L.csv > gpurun_out/msm_launches_$L.txt 2>&1
  cat gpurun_out/msm_launches_$L.txt
done
