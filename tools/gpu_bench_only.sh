#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt; tail -2 gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1500 python bench.py > gpurun_out/bench_full.txt 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_full.txt").read().strip().splitlines()[-1])
d["config"].pop("per_step_sha_interval_ms", None)
print(json.dumps({k: d[k] for k in ("value","ms_per_step","steps","clocks","gpu_launches","single_batch_latency_ms","e2e","cpu_baseline")}, indent=0)[:1800])
print(d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["traffic"])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.txt 2>&1; tail -1 gpurun_out/bench_ref.txt | cut -c1-200
