#!/bin/bash
tag=${1:-xa}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_animatediff_gpu.py -m gpu -x -q --no-header 2>&1 | tail -4 | tee gpurun_out/${tag}_tests.log
timeout 300 python tools/time_unet.py 16 3 --shapes > gpurun_out/${tag}_time.log 2>&1
grep -E "forward:|== " gpurun_out/${tag}_time.log
