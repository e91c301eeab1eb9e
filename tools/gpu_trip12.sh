#!/bin/bash
# attention A/B after the blocking-wait rework + kernel parity + whole-forward timing
tag=${1:-t12}
mkdir -p gpurun_out
timeout 300 python tools/attn_variants.py 0 1 9 10 12 > gpurun_out/${tag}_variants.log 2>&1
cat gpurun_out/${tag}_variants.log
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -x -q --no-header 2>&1 | tail -5 | tee gpurun_out/${tag}_tests.log
for v in 9; do
UNIVST_ATTN_VARIANT=$v timeout 300 python tools/time_unet.py 16 3 --shapes > gpurun_out/${tag}_time_v$v.log 2>&1
grep -E "forward:|== " gpurun_out/${tag}_time_v$v.log
done

