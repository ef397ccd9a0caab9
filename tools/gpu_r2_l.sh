#!/bin/bash
# Round 2, call L: early input release (long-chunk arena): parity, sanitizer, value_distinct with and without it.
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 900 python -m pytest tests/test_gpu_round2.py -q -x -k "early or wait_input or async or dense or knobs" > gpurun_out/r2l_pytest_early.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest_early.txt; tail -15 gpurun_out/r2l_pytest_early.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -q -x -k "overwrite or candidate_overflow" > gpurun_out/r2l_memcheck_early.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2l_memcheck_early.txt; tail -4 gpurun_out/r2l_memcheck_early.txt
for v in "1 6" "0 6" "1 6" "0 8" "1 7"; do set -- $v
  PBSGPU_DEBUG=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-verify --distinct-early $1 --distinct-bufs $2 > gpurun_out/r2l_distinct_e$1_b$2.txt 2>gpurun_out/r2l_distinct_e$1_b$2.err
  python - "$1" "$2" <<'PY' | tee -a gpurun_out/r2l_distinct.txt
import json, sys
e, b = sys.argv[1:3]
try:
    d = json.loads(open(f"gpurun_out/r2l_distinct_e{e}_b{b}.txt").read().strip().splitlines()[-1])
    v = d["value_distinct"]
    print(f"early={e} bufs={b}: value_distinct {v.get('value')} GiB/s in {v.get('seconds')} s, hit {v.get('hit_rate')}, chunks {v.get('chunks')}; value {d['value']:.1f} (K=5)", v.get("error", ""))
except Exception as ex:
    print("early", e, "bufs", b, "failed", repr(ex))
PY
  grep -h "arena" gpurun_out/r2l_distinct_e$1_b$2.err | head -2
done
