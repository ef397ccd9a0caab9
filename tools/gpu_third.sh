#!/bin/bash
mkdir -p gpurun_out
for m in 0 3; do
  PBSGPU_SHA_MODE=$m timeout 600 python bench.py --steps 12 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench3_mode$m.txt 2>&1
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench3_mode$m.txt").read().strip().splitlines()[-1])
    print("mode $m value", round(d["value"],1), "GiB/s ms/step", round(d["ms_per_step"],1), " iso", d["roofline"]["isolated_step_ms"], "clk", d["clocks"], "sha GB/s", round(d["roofline"]["achieved"],1))
except Exception as e:
    print("mode $m failed", e); print(open("gpurun_out/bench3_mode$m.txt").read()[-2000:])
PY
done
