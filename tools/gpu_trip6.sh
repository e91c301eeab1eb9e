#!/bin/bash
# attention variants: parity tests + per-shape timing, then the full GPU suite with the default variant
tag=${1:-t}
mkdir -p gpurun_out
for v in ${VARIANTS:-0 2}; do
  echo "=== variant $v"
  UNIVST_ATTN_VARIANT=$v timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -q --no-header -k "attention or unet_forward" 2>&1 | tail -3
  UNIVST_ATTN_VARIANT=$v timeout 300 python tools/time_unet.py 16 3 --shapes 2>&1 | grep -E "forward:|== sc_attention|\(48, 8, (40|80|160), "
done > gpurun_out/${tag}_variants.log 2>&1
cat gpurun_out/${tag}_variants.log
timeout 900 python -m pytest tests -m gpu -x -q --no-header 2>&1 | tail -5 | tee gpurun_out/${tag}_tests.log
