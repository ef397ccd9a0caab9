#!/bin/bash
# Round 2, call G: unrolled scan kernels (k_scan_tuned, k_scan_lanes): parity, K1 alone, ncu --set full of both, consumer CMODE 2, bench.
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2g_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest_gpu.txt; tail -4 gpurun_out/r2g_pytest_gpu.txt
PBSGPU_SCAN_LANES=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -x -k "scan or cfg2 or cfg5 or golden or chunk_digest or structured or streaming or payload or suggested" > gpurun_out/r2g_pytest_lanes.txt 2>&1; echo "pytest(lanes) rc=$?" >> gpurun_out/r2g_pytest_lanes.txt; tail -3 gpurun_out/r2g_pytest_lanes.txt
for l in 0 1; do for part in 24 0; do
  PBSGPU_SCAN_LANES=$l PBSGPU_PARTITION_SMS=$part timeout 200 python tools/scan_bench.py 32 2>&1 | tail -1 | tee -a gpurun_out/r2g_scan_bench.txt
done; done
for l in 0 1; do
  PBSGPU_SCAN_LANES=$l PBSGPU_PARTITION_SMS=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_scan_ --launch-skip 2 -c 1 -f \
      -o gpurun_out/prof_r2_scan_$l python tools/scan_bench.py 32 > gpurun_out/r2g_ncu_scan_$l.log 2>&1; tail -1 gpurun_out/r2g_ncu_scan_$l.log
done
timeout 300 ./tools/sha_lab.bin load 1024 4 > gpurun_out/r2g_sha_lab_load_split.txt 2>&1; grep -E "m=0.25|m=0.50|m=1.00|m=4.00" gpurun_out/r2g_sha_lab_load_split.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2g_bench.txt 2>gpurun_out/r2g_bench.err; tail -c 1500 gpurun_out/r2g_bench.txt; tail -3 gpurun_out/r2g_bench.err
PBSGPU_SCAN_LANES=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-verify --no-distinct > gpurun_out/r2g_bench_lanes.txt 2>&1; tail -c 400 gpurun_out/r2g_bench_lanes.txt
