#!/bin/bash
# Round 2, call P (gpurun --gpus 2): the default bench at N = 2 as the driver launches it (cfg4, NCCL all-gather, e2e after the
# thread-binding fix), and the 2-GPU part of the suite.
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 600 python -m pytest tests/test_gpu_round2.py -q -x -k "allgather" > gpurun_out/r2p_pytest_2gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest_2gpu.txt; tail -3 gpurun_out/r2p_pytest_2gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2p_bench_n2.txt 2> gpurun_out/r2p_bench_n2.err
tail -c 1500 gpurun_out/r2p_bench_n2.txt; tail -3 gpurun_out/r2p_bench_n2.err
