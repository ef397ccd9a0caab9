#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
f,tag=sys.argv[1],sys.argv[2]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(tag, "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],1), "lat", round(d["single_batch_latency_ms"]), "part", d["config"]["sm_partition(long,bulk)"], "sha GB/s", round(d["roofline"]["achieved"]), "W", d["clocks"].get("power_w_max"))
except Exception as e:
    print(tag, "failed", e); print(open(f).read()[-1500:])
PY
}
run() { tag=$1; shift; env $ENVV timeout 600 python bench.py "$@" --warmup 1 --no-e2e --no-cpu > gpurun_out/b15_$tag.txt 2>&1; show gpurun_out/b15_$tag.txt "$tag"; }
ENVV="" run k16 --steps 16
ENVV="" run k32 --steps 32
ENVV="" run k8 --steps 8
ENVV="PBSGPU_PARTITION_SMS=16" run k32_p16 --steps 32
ENVV="PBSGPU_PARTITION_SMS=32" run k32_p32 --steps 32
ENVV="PBSGPU_SPLIT_SPREAD_KB=85" run k32_s85 --steps 32
ENVV="PBSGPU_HYBRID_THR_X10=30" run k32_thr30 --steps 32
ENVV="PBSGPU_PARTITION_SMS=0" run k32_nopart --steps 32
