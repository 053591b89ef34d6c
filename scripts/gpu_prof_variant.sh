#!/bin/bash
# usage: gpu_prof_variant.sh <variant-name>   (profiles k_dsm of a prebuilt variant)
set -x
V=$1
export S256_LIB=$PWD/secp256k1-voi_b200/lib/variants/$V.so
mkdir -p gpurun_out
LOG2N=18 PASSES=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dsm -s 2 -c 1 \
    -o gpurun_out/prof_dsm_$V -f python scripts/prof_dsm.py > gpurun_out/ncu_full_$V.log 2>&1
tail -3 gpurun_out/ncu_full_$V.log
