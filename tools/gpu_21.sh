#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "corpus or all_long or golden" > gpurun_out/pytest_21.txt 2>&1; tail -2 gpurun_out/pytest_21.txt
show() { python - "$1" "$2" <<'PY'
import json,sys
f,tag=sys.argv[1],sys.argv[2]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(tag, "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],1), "lat", round(d["single_batch_latency_ms"]), "iso long/bulk", round(d["roofline"]["isolated_step_ms"]["sha_long_ms"]), round(d["roofline"]["isolated_step_ms"]["sha_bulk_ms"]))
except Exception as e:
    print(tag, "failed", e); print(open(f).read()[-1500:])
PY
}
for m in 14 11; do
PBSGPU_SHA_MODE=$m PBSGPU_SHA_HYBRID=2 timeout 600 python bench.py --steps 2 --warmup 1 --no-prewarm --no-e2e --no-cpu > gpurun_out/b21_$m.txt 2>&1; show gpurun_out/b21_$m.txt "splitonly mode$m"
done
timeout 900 python bench.py --workload cfg3 --total-tb 10 > gpurun_out/cfg3.txt 2>&1; tail -1 gpurun_out/cfg3.txt
