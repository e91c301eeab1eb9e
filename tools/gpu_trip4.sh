#!/bin/bash
tag=${1:-t}
mkdir -p gpurun_out
for v in 0 1 4; do
  echo "=== variant $v"
  UNIVST_ATTN_VARIANT=$v timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -q --no-header -k "attention or unet_forward" 2>&1 | tail -3
  UNIVST_ATTN_VARIANT=$v timeout 300 python tools/time_unet.py 16 3 --shapes 2>&1 | grep -E "forward:|== sc_attention|\(48, 8, 40, 4096, (8192|12288)\)"
done > gpurun_out/${tag}_variants.log 2>&1

cat gpurun_out/${tag}_variants.log
