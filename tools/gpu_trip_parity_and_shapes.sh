#!/bin/bash
tag=${1:-t27}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_fullsize_gpu.py tests/test_animatediff_gpu.py -m gpu -x -q -s --no-header 2>&1 | grep -E "full size|passed|failed|Error" | tee gpurun_out/${tag}_tests.log
timeout 300 python tools/time_unet.py 16 3 --shapes > gpurun_out/${tag}_time.log 2>&1
grep -E "forward:|== " gpurun_out/${tag}_time.log
timeout 300 python tools/time_unet.py 16 3 --shapes --animatediff > gpurun_out/${tag}_time_ad.log 2>&1
grep -E "forward:|== " gpurun_out/${tag}_time_ad.log
