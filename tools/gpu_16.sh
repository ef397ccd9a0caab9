#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt; tail -3 gpurun_out/pytest_gpu.txt
for th in 1 2 3; do
timeout 900 python bench.py --files 64 --steps 2 --warmup 1 --no-cpu --e2e-threads $th --e2e-steps 6 > gpurun_out/b16_e2e_$th.txt 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/b16_e2e_$th.txt").read().strip().splitlines()[-1]); print("threads $th", d.get("e2e"))
PY
done
