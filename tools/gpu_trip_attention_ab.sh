#!/bin/bash
tag=${1:-t17}
mkdir -p gpurun_out
timeout 300 python tools/attn_variants.py 17 19 > gpurun_out/${tag}_variants.log 2>&1
cat gpurun_out/${tag}_variants.log
UNIVST_ATTN_VARIANT=19 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -x -q --no-header 2>&1 | tail -5 | tee gpurun_out/${tag}_tests.log
timeout 300 python tools/loop_ab.py 17 19 2>&1 | tee gpurun_out/${tag}_loop_ab.log
