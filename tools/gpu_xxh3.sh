#!/bin/bash
# K7 (xxh3): parity tests, then the K7 bench.  FULL=1 also runs the scan-lanes / mirror tests.
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
SEL="xxh3"
[ -n "$FULL" ] && SEL="xxh3 or lanes or dedup or cxx or fused or staged"
( timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL" 2>&1 | tail -15 ) > gpurun_out/xxh3_tests.txt
( timeout 200 python tools/xxh3_bench.py 32 2>&1 | tail -8 ) > gpurun_out/xxh3_bench.txt
cat gpurun_out/xxh3_tests.txt gpurun_out/xxh3_bench.txt
