#!/bin/bash
# Round 2, call P (gpurun --gpus 4): the default bench at N = 4 as the driver launches it (cfg4, NCCL all-gather, e2e after the
# thread-binding fix), and the 2-GPU part of the suite.
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2q_bench_n4.txt 2> gpurun_out/r2q_bench_n4.err
tail -c 1500 gpurun_out/r2q_bench_n4.txt; tail -3 gpurun_out/r2q_bench_n4.err
