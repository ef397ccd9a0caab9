#!/bin/bash
# Round 2, call N: early-input / dense tests again; value_distinct vs batch granularity (smaller batches, more of them in flight).
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 900 python -m pytest tests/test_gpu_round2.py -q -x -k "early or wait_input or async or dense or knobs" > gpurun_out/r2n_pytest_early.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest_early.txt; tail -6 gpurun_out/r2n_pytest_early.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -q -x -k "overwrite or candidate_overflow" > gpurun_out/r2n_memcheck_early.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2n_memcheck_early.txt; tail -4 gpurun_out/r2n_memcheck_early.txt
run() {  # name, files, bufs, early, env...
  name=$1; files=$2; bufs=$3; early=$4; shift 4
  env PBSGPU_DEBUG=1 "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-verify --distinct-early $early --distinct-bufs $bufs --distinct-batch-files $files > gpurun_out/r2n_$name.txt 2>gpurun_out/r2n_$name.err
  python - "$name" <<'PY' | tee -a gpurun_out/r2n_sweep.txt
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2n_{n}.txt").read().strip().splitlines()[-1])
    v = d["value_distinct"]
    print(f"{n}: value_distinct {v.get('value') or 0:.1f} GiB/s ({v.get('batches')} x {v.get('batch_GiB')} GiB, {v.get('buffers')} buffers, early {v.get('early_input')}), hit {v.get('hit_rate')}", v.get("error", ""))
except Exception as ex:
    print(n, "failed", repr(ex))
PY
  grep -h "arena" gpurun_out/r2n_$name.err | head -1
}
run a_256x8_plain 256 8 0
run b_64x32_plain 64 32 0 PBSGPU_SLOTS=32
run c_64x32_early 64 32 1 PBSGPU_SLOTS=32
run d_128x16_early 128 16 1 PBSGPU_SLOTS=24
run e_64x32_early_thr15 64 32 1 PBSGPU_SLOTS=32 PBSGPU_HYBRID_THR_X10=15 PBSGPU_PARTITION_SMS=40 PBSGPU_HYBRID_HEAD_PER_SM=64 PBSGPU_ARENA_FRAC_X16=7
run f_32x64_early 32 64 1 PBSGPU_SLOTS=32
run g_64x32_plain_again 64 32 0 PBSGPU_SLOTS=32
