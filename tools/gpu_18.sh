#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "crc32 or didx" > gpurun_out/pytest_crc.txt 2>&1; tail -3 gpurun_out/pytest_crc.txt
timeout 300 python tools/crc_bench.py 2>&1 | tail -1 | tee gpurun_out/crc_bench.txt
