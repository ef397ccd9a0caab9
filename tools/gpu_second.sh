#!/bin/bash
# second GPU visit: full parity suite, SHA pipe-balancing A/B, ncu launch list + full capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -3 gpurun_out/pytest_gpu.txt
for m in 0 1 3 7; do
  PBSGPU_SHA_MODE=$m timeout 600 python bench.py --steps 6 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_mode$m.txt 2>&1
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_mode$m.txt").read().strip().splitlines()[-1])
    print("mode $m value", round(d["value"],1), "GiB/s  iso", d["roofline"]["isolated_step_ms"], "clk", d["clocks"]["sm_mhz"], "sha GB/s", round(d["roofline"]["achieved"],1))
except Exception as e:
    print("mode $m failed", e); print(open("gpurun_out/bench_mode$m.txt").read()[-2000:])
PY
done
# launch list (shares) and one full capture per hot kernel; short workload (ncu replays ~40x)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 1 --files 128 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_sha_tuned -s 1 -c 1 -o gpurun_out/prof_sha -f \
   python bench.py --steps 1 --warmup 1 --files 128 --no-e2e --no-cpu > gpurun_out/ncu_sha.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_scan_tuned -s 1 -c 1 -o gpurun_out/prof_scan -f \
   python bench.py --steps 1 --warmup 1 --files 128 --no-e2e --no-cpu > gpurun_out/ncu_scan.log 2>&1
ls -la gpurun_out/
