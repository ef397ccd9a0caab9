#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt; tail -2 gpurun_out/pytest_gpu.txt
show() { python - "$1" "$2" <<'PY'
import json,sys
f,tag=sys.argv[1],sys.argv[2]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(tag, "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],1), "lat", round(d["single_batch_latency_ms"]), "part", d["config"]["sm_partition(long,bulk)"], "sha GB/s", round(d["roofline"]["achieved"]), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(tag, "failed", e); print(open(f).read()[-1500:])
PY
}
run() { tag=$1; shift; timeout 600 python bench.py "$@" --warmup 1 --no-e2e --no-cpu > gpurun_out/b14_$tag.txt 2>&1; show gpurun_out/b14_$tag.txt "$tag"; }
run k16_if12 --steps 16 --inflight 12
run k16_if8 --steps 16 --inflight 8
run k16_if5 --steps 16 --inflight 5
run k16_if3 --steps 16 --inflight 3
run k32_if12 --steps 32 --inflight 12
run k32_if6 --steps 32 --inflight 6
run k8_if8 --steps 8 --inflight 8
