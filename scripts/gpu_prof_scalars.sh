#!/bin/bash
mkdir -p gpurun_out
LOG2N=20 WHICH=verify timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ecdsa_scalars -s 2 -c 1 \
  -o gpurun_out/prof_scalars -f python scripts/prof_kernels.py > gpurun_out/ncu_scalars.log 2>&1
tail -2 gpurun_out/ncu_scalars.log
