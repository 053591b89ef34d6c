#!/bin/bash
# MSM per-kernel durations at 2^17..2^20 points (ncu launch lists; not bench values)
mkdir -p gpurun_out
for L in ${SIZES:-17 18 19 20}; do
  LOG2N=$L timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"msm|decode|any_invalid|Scan|finish_affine" -c 400 --csv \
    --log-file gpurun_out/r2_msm_launches_$L.csv python scripts/prof_msm.py > gpurun_out/r2_msm_prof_$L.log 2>&1
  python scripts/launch_summary.py gpurun_out/r2_msm_launches_$L.csv > gpurun_out/r2_msm_launches_$L.txt 2>&1
  echo "== 2^$L"; cat gpurun_out/r2_msm_launches_$L.txt
done
