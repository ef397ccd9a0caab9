#!/bin/bash
# first GPU visit: host facts, microbench + TMA probe, parity tests, smoke, a short bench
mkdir -p gpurun_out
{ nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm,power.limit --format=csv; lscpu | grep -E "Model name|^CPU\(s\)|sha_ni" | head; } > gpurun_out/host.txt 2>&1
timeout 300 ./tools/microbench.bin > gpurun_out/microbench.txt 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.txt
timeout 900 python bench.py --steps 3 --warmup 1 --files 256 --e2e-files 32 --e2e-steps 2 > gpurun_out/bench_small.txt 2>&1; echo "bench rc=$?" >> gpurun_out/bench_small.txt
tail -5 gpurun_out/pytest_gpu.txt; tail -3 gpurun_out/smoke.txt; tail -2 gpurun_out/bench_small.txt
