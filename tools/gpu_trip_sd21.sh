#!/bin/bash
tag=${1:-sd21}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fullsize_gpu.py -m gpu -x -q -s --no-header -k "sd21" 2>&1 | grep -E "full size|passed|failed|Error" | tee gpurun_out/${tag}_tests.log
timeout 300 python tools/time_unet.py 16 3 --shapes --sd21 > gpurun_out/${tag}_time.log 2>&1
grep -E "forward:|== " gpurun_out/${tag}_time.log; grep -A6 "== sc_attention" gpurun_out/${tag}_time.log | head -8
