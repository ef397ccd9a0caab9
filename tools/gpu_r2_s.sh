#!/bin/bash
# Round 2, last call: the whole GPU suite + smoke on the final tree.
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r2s_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_pytest_gpu.txt; tail -3 gpurun_out/r2s_pytest_gpu.txt
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/r2s_smoke.txt
