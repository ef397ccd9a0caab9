#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "stream or cfg1 or mirror or cxx" > gpurun_out/pytest_stream.txt 2>&1; tail -2 gpurun_out/pytest_stream.txt
timeout 600 python tools/stream_bench.py 2>&1 | tail -4 | tee gpurun_out/stream_bench.txt
