#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 5 --warmup 3 --skip-cpu-baseline --headline-only > gpurun_out/bench_under_ncu.log 2>&1
LOG2N=20 WHICH=verify timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dsm -s 2 -c 1 -o gpurun_out/prof_final_dsm -f python scripts/prof_kernels.py > gpurun_out/ncu1.log 2>&1
LOG2N=18 WHICH=ecdh timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_scalar_mult_ct -s 2 -c 1 -o gpurun_out/prof_final_ct -f python scripts/prof_kernels.py > gpurun_out/ncu2.log 2>&1
LOG2N=18 WHICH=sbm timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_base_mult_ct -s 2 -c 1 -o gpurun_out/prof_final_sbm -f python scripts/prof_kernels.py > gpurun_out/ncu3.log 2>&1
LOG2N=18 WHICH=verify timeout 900 ncu --set full --clock-control none -k regex:k_ecdsa_scalars -s 2 -c 1 -o gpurun_out/prof_final_scalars -f python scripts/prof_kernels.py > gpurun_out/ncu4.log 2>&1
tail -1 gpurun_out/ncu*.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json | cut -c1-400
