#!/bin/bash
mkdir -p gpurun_out
for m in 2 0; do PBSGPU_PARTITION_SMS=0 PBSGPU_SHA_MODE=$m timeout 300 python tools/sha_bench.py 256 32 2>&1 | tail -1; done | tee gpurun_out/sha_bench4.txt
