#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
f,tag=sys.argv[1],sys.argv[2]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(tag, "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],1), "iso", {k: round(v,1) for k,v in d["roofline"]["isolated_step_ms"].items()}, "lat", round(d["single_batch_latency_ms"]))
except Exception as e:
    print(tag, "failed", e); print(open(f).read()[-1500:])
PY
}
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 4 --warmup 1 --no-e2e --no-cpu > gpurun_out/b8_$tag.txt 2>&1; show gpurun_out/b8_$tag.txt "$tag"; }
run base PBSGPU_SHA_MODE=2
run prio PBSGPU_SHA_MODE=2 PBSGPU_HYBRID_PRIO=1
run serial PBSGPU_SHA_MODE=2 PBSGPU_HYBRID_SERIAL=1
run thr35 PBSGPU_SHA_MODE=2 PBSGPU_HYBRID_THR_X10=35
run thr35prio PBSGPU_SHA_MODE=2 PBSGPU_HYBRID_THR_X10=35 PBSGPU_HYBRID_PRIO=1
run conn32 PBSGPU_SHA_MODE=2 CUDA_DEVICE_MAX_CONNECTIONS=32
