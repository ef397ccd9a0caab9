#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt; tail -3 gpurun_out/pytest_gpu.txt
show() { python - "$1" "$2" <<'PY'
import json,sys
f,tag=sys.argv[1],sys.argv[2]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(tag, "value", round(d["value"],1), "GiB/s ms/step", round(d["ms_per_step"],1), "iso sha", round(d["roofline"]["isolated_step_ms"]["sha_ms"],1), "W", d["clocks"].get("power_w_max"), "sha GB/s", round(d["roofline"]["achieved"],1), "ivals", d["config"]["per_step_sha_interval_ms"][:6])
except Exception as e:
    print(tag, "failed", e); print(open(f).read()[-1500:])
PY
}
for m in 0 3 13; do
  PBSGPU_SHA_MODE=$m timeout 600 python bench.py --avg-kib 256 --steps 4 --warmup 1 --no-e2e --no-cpu > gpurun_out/b4_small_$m.txt 2>&1; show gpurun_out/b4_small_$m.txt "avg256K mode$m"
done
for m in 13 11 12 10; do
  PBSGPU_SHA_MODE=$m timeout 600 python bench.py --steps 12 --warmup 1 --no-e2e --no-cpu > gpurun_out/b4_$m.txt 2>&1; show gpurun_out/b4_$m.txt "4MiB mode$m"
done
for m in 3 13; do
PBSGPU_SHA_MODE=$m timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sha -s 1 -c 1 -o gpurun_out/prof_sha_sat_$m -f \
   python bench.py --avg-kib 256 --steps 1 --warmup 1 --files 256 --no-e2e --no-cpu > gpurun_out/ncu_sha_sat_$m.log 2>&1
done
ls -la gpurun_out | tail -8
