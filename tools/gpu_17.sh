#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "didx or golden" > gpurun_out/pytest_didx.txt 2>&1; tail -2 gpurun_out/pytest_didx.txt
show() { python - "$1" "$2" <<'PY'
import json,sys
f,tag=sys.argv[1],sys.argv[2]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(tag, "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],1), "lat", round(d["single_batch_latency_ms"]), "part", d["config"]["sm_partition(long,bulk)"], "iso long/bulk", round(d["roofline"]["isolated_step_ms"]["sha_long_ms"]), round(d["roofline"]["isolated_step_ms"]["sha_bulk_ms"]))
except Exception as e:
    print(tag, "failed", e); print(open(f).read()[-1500:])
PY
}
run() { tag=$1; shift; env $ENVV timeout 600 python bench.py "$@" --warmup 1 --no-e2e --no-cpu > gpurun_out/b17_$tag.txt 2>&1; show gpurun_out/b17_$tag.txt "$tag"; }
ENVV="" run base --steps 32
ENVV="PBSGPU_HYBRID_THR_X10=20" run thr20 --steps 32
ENVV="PBSGPU_HYBRID_THR_X10=15" run thr15 --steps 32
ENVV="PBSGPU_HYBRID_THR_X10=15 PBSGPU_PARTITION_SMS=32" run thr15_p32 --steps 32
ENVV="PBSGPU_HYBRID_THR_X10=20 PBSGPU_PARTITION_SMS=32" run thr20_p32 --steps 32
ENVV="PBSGPU_HYBRID_THR_X10=12 PBSGPU_PARTITION_SMS=40" run thr12_p40 --steps 32
