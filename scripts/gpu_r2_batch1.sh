#!/bin/bash
# Round 2, first GPU session: parity of the split-form multiplier, then the A/B of the k_dsm variants.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
LOG2N=20 timeout 900 python scripts/variant_bench.py > gpurun_out/variants.log 2>&1
cat gpurun_out/variants.log
timeout 200 python scripts/microbench_fe_mul.py > gpurun_out/microbench_fe_mul.log 2>&1; tail -4 gpurun_out/microbench_fe_mul.log
