#!/bin/bash
# Round 2, final call: the whole GPU suite + smoke on HEAD, the ncu launch list of a short bench run, the full bench line and the reference arm.
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2k_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest_gpu.txt; tail -4 gpurun_out/r2k_pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/r2k_smoke.txt
# launch list: ncu's kernel-replay serialisation does not get along with launches from several green contexts + the primary
# context (LaunchFailed after the first job), so this pass runs unpartitioned with the hybrid launch forced on (PBSGPU_SHA_HYBRID=2:
# the same kernels as the product configuration); per-launch times are serialised and cold anyway
PBSGPU_PARTITION_SMS=0 PBSGPU_SHA_HYBRID=2 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2k_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-verify --no-distinct --no-prewarm > gpurun_out/r2k_launches_bench.log 2>&1; tail -c 200 gpurun_out/r2k_launches_bench.log; wc -l gpurun_out/r2k_launches.csv
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2k_bench.txt 2>gpurun_out/r2k_bench.err; tail -c 1200 gpurun_out/r2k_bench.txt; tail -3 gpurun_out/r2k_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2k_bench_reference.txt 2>&1; tail -c 600 gpurun_out/r2k_bench_reference.txt
timeout 300 python tools/blob_bench.py 2>&1 | tail -4 | tee gpurun_out/r2k_blob_bench.txt
