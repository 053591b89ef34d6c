#!/bin/bash
# two GPUs: the sharded MSM with NCCL inside the C ABI (tests/test_multi_gpu.py) and bench.py at N = 2 under torchrun
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_2gpu.log
tail -15 gpurun_out/pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench rc=$?"
cat gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
