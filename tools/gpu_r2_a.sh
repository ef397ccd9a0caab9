#!/bin/bash
# Round 2, call A: suite + smoke on HEAD, tensor-map probe, ncu --set full of K7 / K6 / saturated K3 (modes 0,1,3), short bench.
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv | tee gpurun_out/r2a_gpu.txt; nproc >> gpurun_out/r2a_gpu.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2a_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest_gpu.txt; tail -3 gpurun_out/r2a_pytest_gpu.txt
[ -x tools/tma_tensor_probe.bin ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_tensor_probe.bin tools/tma_tensor_probe.cu
timeout 120 ./tools/tma_tensor_probe.bin 2>&1 | tee gpurun_out/r2a_tma_tensor_probe.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_xxh3_ -c 2 -f -o gpurun_out/prof_xxh3 \
    python tools/xxh3_bench.py 8 > gpurun_out/r2a_ncu_xxh3.log 2>&1; tail -2 gpurun_out/r2a_ncu_xxh3.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_crc32 -c 1 -f -o gpurun_out/prof_crc32 \
    python tools/crc_bench.py > gpurun_out/r2a_ncu_crc.log 2>&1; tail -2 gpurun_out/r2a_ncu_crc.log
for m in 0 1 3; do
  PBSGPU_SHA_MODE=$m timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sha_tuned --launch-skip 1 -c 1 -f \
      -o gpurun_out/prof_sha_fullsat_$m python tools/sha_bench.py 256 32 > gpurun_out/r2a_ncu_sha_$m.log 2>&1; tail -1 gpurun_out/r2a_ncu_sha_$m.log
done
timeout 600 python bench.py --steps 16 --warmup 3 --no-cpu > gpurun_out/r2a_bench.txt 2>&1; tail -c 1500 gpurun_out/r2a_bench.txt
