#!/bin/bash
mkdir -p gpurun_out
for m in 2 20 22 21 23; do PBSGPU_SHA_MODE=$m timeout 300 python tools/sha_bench.py 256 32 2>&1 | tail -1; done | tee gpurun_out/sha_bench3.txt
for m in 22 23; do PBSGPU_SHA_MODE=$m timeout 300 python tools/sha_bench.py 256 32 misaligned 2>&1 | tail -1; done | tee -a gpurun_out/sha_bench3.txt
