#!/bin/bash
# ncu --set full of k_dsm for two variant builds (S256_LIB selects the .so)
mkdir -p gpurun_out
for v in split xorz2; do
  S256_LIB=$PWD/secp256k1-voi_b200/lib/variants/$v.so LOG2N=18 PASSES=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dsm -s 2 -c 1 \
    -o gpurun_out/prof_r2_$v -f python scripts/prof_dsm.py > gpurun_out/ncu_full_r2_$v.log 2>&1
  tail -2 gpurun_out/ncu_full_r2_$v.log
done
