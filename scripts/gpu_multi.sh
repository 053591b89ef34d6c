#!/bin/bash
# usage: gpu_multi.sh N   (under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --skip-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
cat gpurun_out/bench_n$N.json | cut -c1-900; tail -3 gpurun_out/bench_n$N.err
LOG2N=20 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    scripts/msm_nccl.py > gpurun_out/msm_n$N.json 2> gpurun_out/msm_n$N.err; echo "msm rc=$?"
cat gpurun_out/msm_n$N.json; tail -3 gpurun_out/msm_n$N.err
LOG2N=20 timeout 300 python scripts/msm_nccl.py > gpurun_out/msm_n1.json 2>> gpurun_out/msm_n$N.err; cat gpurun_out/msm_n1.json
