#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/bench_full.txt 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_full.txt").read().splitlines() if l.startswith("{")][-1])
print({k: d[k] for k in ("value","ms_per_step","steps","gpu_launches","single_batch_latency_ms")}, d["clocks"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["roofline"]["achieved"], d["roofline"]["frac"])
PY
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.txt 2>&1; tail -1 gpurun_out/bench_ref.txt | cut -c1-160
