#!/bin/bash
# A/B of the packed-polynomial exp2 variants of the split attention kernel + the torch-eager forward next to ours.
tag=${1:-p2}
mkdir -p gpurun_out
timeout 300 python tools/attn_variants.py 19 20 21 22 19 20 21 > gpurun_out/${tag}_variants.log 2>&1
cat gpurun_out/${tag}_variants.log
for v in 20 21; do
  UNIVST_ATTN_VARIANT=$v timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --no-header -k "attention" 2>&1 | tail -3 | tee gpurun_out/${tag}_tests_v$v.log
done
timeout 300 python tools/loop_ab.py 19 20 21 2>&1 | tee gpurun_out/${tag}_loop_ab.log
timeout 600 python tools/eager_forward.py 16 5 2>&1 | tail -3 | tee gpurun_out/${tag}_eager.log
