#!/bin/bash
# attention A/B (row sums on the tensor pipe, source dedupe) + kernel parity + per-shape timing + ncu source capture
tag=${1:-t11}
mkdir -p gpurun_out
timeout 300 python tools/attn_variants.py 1 7 8 9 4 > gpurun_out/${tag}_variants.log 2>&1
cat gpurun_out/${tag}_variants.log
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -x -q --no-header 2>&1 | tail -5 | tee gpurun_out/${tag}_tests.log
timeout 300 python tools/time_unet.py 16 3 --shapes > gpurun_out/${tag}_time.log 2>&1
grep -E "forward:|== " gpurun_out/${tag}_time.log
UNIVST_ATTN_VARIANT=${NCU_VARIANT:-7} timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"attention_tc_kernel<\(int\)2, \(int\)128" -s 24 -c 1 -f -o gpurun_out/${tag}_attn python tools/time_unet.py 16 1 > gpurun_out/${tag}_ncu_attn.log 2>&1
ls -la gpurun_out/
