#!/bin/bash
tag=${1:-t20}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_fullsize_gpu.py -m gpu -x -q -s --no-header 2>&1 | tail -15 | tee gpurun_out/${tag}_tests.log
