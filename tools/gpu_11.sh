#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt; tail -2 gpurun_out/pytest_gpu.txt
show() { python - "$1" "$2" <<'PY'
import json,sys
f,tag=sys.argv[1],sys.argv[2]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    iv=d['config']['per_step_sha_interval_ms']; base=min(a for a,b in iv)
    print(tag, "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],1), "iso", {k: round(v,1) for k,v in d["roofline"]["isolated_step_ms"].items() if k in ("scan_ms","sha_ms","sha_long_ms","sha_bulk_ms")}, "lat", round(d["single_batch_latency_ms"]), [(round(a-base), round(b-base)) for a,b in iv][:16:3])
except Exception as e:
    print(tag, "failed", e); print(open(f).read()[-1500:])
PY
}
run() { tag=$1; steps=$2; shift; shift; env "$@" timeout 600 python bench.py --steps $steps --warmup 1 --no-e2e --no-cpu > gpurun_out/b11_$tag.txt 2>&1; show gpurun_out/b11_$tag.txt "$tag"; }
run defer 16 PBSGPU_SHA_MODE=2
run defer_p24_s30 16 PBSGPU_PARTITION_SMS=24 PBSGPU_SPLIT_SPREAD_KB=30
run defer_p24_s85 16 PBSGPU_PARTITION_SMS=24 PBSGPU_SPLIT_SPREAD_KB=85
run defer_p32_s30 16 PBSGPU_PARTITION_SMS=32 PBSGPU_SPLIT_SPREAD_KB=30
run defer_p24_s30_thr30 16 PBSGPU_PARTITION_SMS=24 PBSGPU_SPLIT_SPREAD_KB=30 PBSGPU_HYBRID_THR_X10=30
run defer_p16_s30_thr30 16 PBSGPU_PARTITION_SMS=16 PBSGPU_SPLIT_SPREAD_KB=30 PBSGPU_HYBRID_THR_X10=30
run nodefer_p24_s30 16 PBSGPU_DEFER_SHA=0 PBSGPU_PARTITION_SMS=24 PBSGPU_SPLIT_SPREAD_KB=30
