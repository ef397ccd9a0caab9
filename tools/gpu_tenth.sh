#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/greenctx_probe.bin 2>&1 | tee gpurun_out/greenctx_probe.txt
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
run() { tag=$1; steps=$2; shift; shift; env "$@" timeout 600 python bench.py --steps $steps --warmup 1 --no-e2e --no-cpu > gpurun_out/b10_$tag.txt 2>&1; show gpurun_out/b10_$tag.txt "$tag"; }
run part24 16 PBSGPU_PARTITION_SMS=24 PBSGPU_SPLIT_SPREAD_KB=85
run part16 16 PBSGPU_PARTITION_SMS=16 PBSGPU_SPLIT_SPREAD_KB=85
run part32 16 PBSGPU_PARTITION_SMS=32 PBSGPU_SPLIT_SPREAD_KB=85
run part24_thr20 16 PBSGPU_PARTITION_SMS=24 PBSGPU_SPLIT_SPREAD_KB=85 PBSGPU_HYBRID_THR_X10=20
run nopart 16 PBSGPU_SPLIT_SPREAD_KB=112
