#!/bin/bash
mkdir -p gpurun_out
# memcheck + racecheck + synccheck on the core parity tests (small inputs; sanitizer is 10-100x slower)
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
K="test_scan_unaligned_offsets or test_sha256_batch_all_small or test_chunk_digest_batch_matches or test_digest_set_tag or test_empty_batch or test_all_long or test_streaming_is_split_invariant"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests -x -q -m gpu -k "$K" > gpurun_out/san_memcheck.txt 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/san_memcheck.txt
tail -4 gpurun_out/san_memcheck.txt
K2="test_chunk_digest_batch_matches or test_all_long or test_scan_unaligned_offsets"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests -x -q -m gpu -k "$K2" > gpurun_out/san_racecheck.txt 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/san_racecheck.txt
tail -4 gpurun_out/san_racecheck.txt
grep -c "ERROR SUMMARY" gpurun_out/san_*.txt; grep -h "ERROR SUMMARY" gpurun_out/san_*.txt | head
