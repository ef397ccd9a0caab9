#!/bin/bash
# k_scan_lanes: K1 timing (reference checksum for seed 11, 32 GiB: a0096e149910c4db), with and without the SM partition.
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
{
for part in 0 24; do
  for v in ${LANES_VARIANTS:-0 1}; do
    PBSGPU_PARTITION_SMS=$part PBSGPU_SCAN_LANES=$v timeout 120 python tools/scan_bench.py 32
  done
done
} > gpurun_out/lanes_bench.txt 2>&1
cat gpurun_out/lanes_bench.txt
