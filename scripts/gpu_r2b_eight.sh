#!/bin/bash
# N GPUs of one box (gpurun --gpus N): the two-rank NCCL test, then bench.py under torchrun at N ranks
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r2b_pytest_multigpu.log 2>&1; tail -3 gpurun_out/r2b_pytest_multigpu.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2b_bench_${N}gpu.json 2> gpurun_out/r2b_bench_${N}gpu.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/r2b_bench_${N}gpu.json; tail -3 gpurun_out/r2b_bench_${N}gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 scripts/msm_sharded_diag.py 2>&1 | grep rank | head -3
