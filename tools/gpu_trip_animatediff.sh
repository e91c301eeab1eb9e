#!/bin/bash
# AnimateDiff backbone: kernel + UNet parity, full-size forward timing (per shape)
tag=${1:-t15}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_animatediff_gpu.py -m gpu -x -q --no-header 2>&1 | tail -15 | tee gpurun_out/${tag}_tests.log
timeout 300 python tools/time_unet.py 16 3 --shapes --animatediff > gpurun_out/${tag}_time_ad.log 2>&1
grep -E "forward:|== |Error|error" gpurun_out/${tag}_time_ad.log | head
grep -A8 "== temporal_attention" gpurun_out/${tag}_time_ad.log
