#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt; tail -3 gpurun_out/pytest_gpu.txt
for m in 2 0; do PBSGPU_SHA_MODE=$m timeout 300 python tools/sha_bench.py 256 32 2>&1 | tail -1; done | tee gpurun_out/sha_bench2.txt
show() { python - "$1" "$2" <<'PY'
import json,sys
f,tag=sys.argv[1],sys.argv[2]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(tag, "value", round(d["value"],1), "GiB/s ms/step", round(d["ms_per_step"],1), "iso", {k: round(v,1) for k,v in d["roofline"]["isolated_step_ms"].items()}, "lat", round(d["single_batch_latency_ms"]), "sha GB/s", round(d["roofline"]["achieved"],1), "e2e", d.get("e2e"), "cpu", d.get("cpu_baseline"))
except Exception as e:
    print(tag, "failed", e); print(open(f).read()[-1500:])
PY
}
PBSGPU_SHA_MODE=2 timeout 600 python bench.py --steps 16 --warmup 2 --no-e2e --no-cpu > gpurun_out/b6_hyb2.txt 2>&1; show gpurun_out/b6_hyb2.txt "hybrid mode2 K16"
PBSGPU_SHA_MODE=0 timeout 600 python bench.py --steps 16 --warmup 2 --no-e2e --no-cpu > gpurun_out/b6_hyb0.txt 2>&1; show gpurun_out/b6_hyb0.txt "hybrid mode0 K16"
PBSGPU_SHA_HYBRID=0 PBSGPU_SHA_MODE=2 timeout 600 python bench.py --steps 16 --warmup 2 --no-e2e --no-cpu > gpurun_out/b6_nohyb2.txt 2>&1; show gpurun_out/b6_nohyb2.txt "no-hybrid mode2 K16"
timeout 1200 python bench.py --steps 32 --warmup 3 > gpurun_out/b6_full.txt 2>&1; show gpurun_out/b6_full.txt "full default K32"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b6_ref.txt 2>&1; tail -1 gpurun_out/b6_ref.txt | cut -c1-700
