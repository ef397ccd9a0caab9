#!/bin/bash
# Round 2: writer mirror with GPU-rendered DataBlobs + the neighbouring tests, on the final tree.
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 300 python -m pytest tests -q -m gpu -k "dedup or blob or payload or writer or transfer or verify_backed" > gpurun_out/r2t_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest.txt; tail -8 gpurun_out/r2t_pytest.txt
