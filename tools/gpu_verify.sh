#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt; tail -2 gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 16 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_verify.txt 2>&1
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_verify.txt").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],1), "lat", round(d["single_batch_latency_ms"]), d["clocks"]["reasons"], "launches", d["gpu_launches"])
PY
