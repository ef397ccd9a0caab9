#!/bin/bash
# Round 2, call H (gpurun --gpus 8): cfg4 at N=8 through pbsgpu_set_allgather; the hit count must equal the single-set run.
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
nvidia-smi -L | wc -l | tee gpurun_out/r2h_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2h_bench_n8.txt 2> gpurun_out/r2h_bench_n8.err; tail -c 1200 gpurun_out/r2h_bench_n8.txt; tail -4 gpurun_out/r2h_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2h_bench_n4.txt 2> gpurun_out/r2h_bench_n4.err; tail -c 600 gpurun_out/r2h_bench_n4.txt
