#!/bin/bash
# usage: bash tools/gpu_multi.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_gpus.txt 2>&1
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus $N --steps 8 --warmup 2 --no-e2e --no-cpu > gpurun_out/bench_n$N.txt 2>&1
echo "rc=$?"; tail -2 gpurun_out/bench_n$N.txt | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
   bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.txt 2>&1
echo "ref rc=$?"; tail -1 gpurun_out/bench_ref_n$N.txt | cut -c1-300
