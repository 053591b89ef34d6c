#!/bin/bash
# usage: gpu_prof_one.sh <kernel regex> <WHICH> <LOG2N>  -> gpurun_out/prof_one.ncu-rep
mkdir -p gpurun_out
LOG2N=${3:-20} WHICH=${2:-sbm} timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s 1 -c 1 \
  -o gpurun_out/prof_one -f python scripts/prof_kernels.py > gpurun_out/ncu_one.log 2>&1
tail -2 gpurun_out/ncu_one.log
