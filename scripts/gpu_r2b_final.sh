#!/bin/bash
# Round 2, second session, evidence set: full -m gpu suite (with the constant-time counters), clean bench (N = 1) +
# reference arm, the ncu launch list of the headline steps, ncu --set full of k_dsm (2^20), k_scalar_mult_ct (2^18).
mkdir -p gpurun_out
S256_CT_KEEP=$PWD/gpurun_out/r2b_ct_counters.csv timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest_gpu.log
tail -14 gpurun_out/r2b_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/r2b_bench.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2b_bench_ref.json 2>> gpurun_out/r2b_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_bench.csv \
    python bench.py --steps 5 --warmup 3 --skip-cpu-baseline --headline-only > gpurun_out/r2b_bench_under_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r2b_launches_bench.csv > gpurun_out/r2b_launches_bench.txt; cat gpurun_out/r2b_launches_bench.txt
LOG2N=20 WHICH=verify timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dsm -s 2 -c 1 -o gpurun_out/r2b_prof_dsm -f python scripts/prof_kernels.py > gpurun_out/ncu1.log 2>&1
LOG2N=18 WHICH=ecdh timeout 900 ncu --set full --clock-control none -k regex:k_scalar_mult_ct -s 2 -c 1 -o gpurun_out/r2b_prof_ct -f python scripts/prof_kernels.py > gpurun_out/ncu2.log 2>&1
for f in gpurun_out/ncu*.log; do tail -1 $f; done
