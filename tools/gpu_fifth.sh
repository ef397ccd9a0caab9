#!/bin/bash
mkdir -p gpurun_out
for m in 0 1 3 7 10 11 12 13; do PBSGPU_SHA_MODE=$m timeout 300 python tools/sha_bench.py 256 32 2>&1 | tail -1; done | tee gpurun_out/sha_bench.txt
for m in 0 3 13; do PBSGPU_SHA_MODE=$m timeout 300 python tools/sha_bench.py 1024 32 2>&1 | tail -1; done | tee -a gpurun_out/sha_bench.txt
for m in 0 3 13; do PBSGPU_SHA_MODE=$m timeout 300 python tools/sha_bench.py 256 32 misaligned 2>&1 | tail -1; done | tee -a gpurun_out/sha_bench.txt
