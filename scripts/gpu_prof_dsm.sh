#!/bin/bash
# usage: gpu_prof_dsm.sh <tag> [kernel-regex]   (env passes through, e.g. S256_LADDER)
set -x
TAG=$1; KRE=${2:-k_dsm}
mkdir -p gpurun_out
LOG2N=18 PASSES=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 2 -c 1 \
    -o gpurun_out/prof_$TAG -f python scripts/prof_dsm.py > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
