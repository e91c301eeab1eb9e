#!/bin/bash
tag=${1:-t}
mkdir -p gpurun_out
for v in 0 1 2 3; do
  echo "=== variant $v"
  UNIVST_ATTN_VARIANT=$v timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -q --no-header -k "attention or unet_forward" 2>&1 | tail -3
  UNIVST_ATTN_VARIANT=$v timeout 300 python tools/time_unet.py 16 3 --shapes 2>&1 | grep -E "forward:|== sc_attention|\(48, 8, 40, 4096, (8192|12288)\)"
done > gpurun_out/${tag}_variants.log 2>&1
UNIVST_ATTN_VARIANT=2 timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"attention_tc_kernel<\(int\)4, \(int\)64" -s 22 -c 1 -o gpurun_out/${tag}_attn4 python tools/time_unet.py 16 1 > gpurun_out/${tag}_ncu_attn.log 2>&1
cat gpurun_out/${tag}_variants.log
