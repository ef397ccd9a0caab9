#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
f,tag=sys.argv[1],sys.argv[2]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    iv=d['config']['per_step_sha_interval_ms']; base=min(a for a,b in iv)
    print(tag, "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],1), "submit_ms", round(d["config"]["host_submit_ms_total"]), "lat", round(d["single_batch_latency_ms"]), [(round(a-base), round(b-base)) for a,b in iv][:16:3])
except Exception as e:
    print(tag, "failed", e); print(open(f).read()[-1500:])
PY
}
run() { tag=$1; steps=$2; shift; shift; env "$@" timeout 600 python bench.py --steps $steps --warmup 1 --no-e2e --no-cpu > gpurun_out/b12_$tag.txt 2>&1; show gpurun_out/b12_$tag.txt "$tag"; }
run plain 16 PBSGPU_DEFER_SHA=0
run plain_conn8 16 PBSGPU_DEFER_SHA=0 CUDA_DEVICE_MAX_CONNECTIONS=8
run defer 16 PBSGPU_DEFER_SHA=1
run p24 16 PBSGPU_DEFER_SHA=0 PBSGPU_PARTITION_SMS=24 PBSGPU_SPLIT_SPREAD_KB=30
run p24_defer 16 PBSGPU_DEFER_SHA=1 PBSGPU_PARTITION_SMS=24 PBSGPU_SPLIT_SPREAD_KB=30
run nohyb 16 PBSGPU_DEFER_SHA=0 PBSGPU_SHA_HYBRID=0
run nohyb_defer 16 PBSGPU_DEFER_SHA=1 PBSGPU_SHA_HYBRID=0
