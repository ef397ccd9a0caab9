#!/bin/bash
# Round 2, call E: timeline of the pipelined run, compute-sanitizer over the round-2 kernels, ncu launch list of one bench step +
# full captures of K1/K3 in the real batch (traffic for the r02 roofline line).
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-verify --no-distinct --timeline > gpurun_out/r2e_timeline.txt 2>&1; tail -c 300 gpurun_out/r2e_timeline.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-verify --no-distinct --timeline > gpurun_out/r2e_timeline2.txt 2>&1; tail -c 200 gpurun_out/r2e_timeline2.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -q -x \
    -k "suggested or payload or async_jobs or dense_batch or blob or reserve_commit or single_rank" > gpurun_out/r2e_san_memcheck.txt 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r2e_san_memcheck.txt
PBSGPU_SCAN_LANES=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x \
    -k "scan_lanes" > gpurun_out/r2e_san_lanes.txt 2>&1; echo "memcheck lanes rc=$?" | tee -a gpurun_out/r2e_san_lanes.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -q -x \
    -k "async_jobs or dense_batch or scan_ragged" > gpurun_out/r2e_san_racecheck.txt 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r2e_san_racecheck.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-verify --no-distinct > gpurun_out/r2e_launches_bench.log 2>&1; tail -c 300 gpurun_out/r2e_launches_bench.log
for k in k_scan_tuned k_sha_tuned k_sha_split; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 4 -c 1 -f -o gpurun_out/prof_r2_$k \
      python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-verify --no-distinct > gpurun_out/r2e_ncu_$k.log 2>&1; tail -1 gpurun_out/r2e_ncu_$k.log
done
