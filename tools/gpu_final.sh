#!/bin/bash
# final-of-round evidence: full bench + reference arm + ncu launch list + full captures + cfg3 10 TB run
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt; tail -2 gpurun_out/pytest_gpu.txt
timeout 1500 python bench.py > gpurun_out/bench_full.txt 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_full.txt | cut -c1-3000
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.txt 2>&1; tail -1 gpurun_out/bench_ref.txt | cut -c1-400
# ncu: launch list of the same command shape (2 steps), then one full capture per hot kernel
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 1 --no-prewarm --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
for k in k_scan_tuned k_sha_tuned k_sha_split; do
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/prof_$k -f \
   python bench.py --steps 1 --warmup 1 --no-prewarm --no-e2e --no-cpu > gpurun_out/ncu_$k.log 2>&1
done
timeout 900 python bench.py --workload cfg3 --total-tb 10 > gpurun_out/cfg3.txt 2>&1; tail -1 gpurun_out/cfg3.txt
ls -la gpurun_out | tail -15
