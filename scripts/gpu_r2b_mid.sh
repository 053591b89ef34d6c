#!/bin/bash
# Round 2 (second session), mid-way evidence: bench line of the Jacobian ladder + ncu --set full of k_dsm at 2^20
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
cut -c1-1200 gpurun_out/r2b_bench.json
LOG2N=20 WHICH=verify timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dsm -s 2 -c 1 -o gpurun_out/r2b_prof_dsm -f python scripts/prof_kernels.py > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log
