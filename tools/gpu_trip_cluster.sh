#!/bin/bash
# 2-CTA cluster GEMM / conv with multicast weight tiles: parity (kernel tests) and per-shape A/B
tag=${1:-cl}
mkdir -p gpurun_out
UNIVST_GEMM_CLUSTER=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --no-header -k "gemm or conv" 2>&1 | tail -4
for c in 0 1; do
  export UNIVST_GEMM_CLUSTER=$c
  echo "--- cluster=$c"
  timeout 60 python tools/conv_shape.py 48 64 64 320 320
  timeout 60 python tools/conv_shape.py 48 32 32 640 640
  timeout 60 python tools/conv_shape.py 48 16 16 1280 1280
  timeout 60 python tools/gemm_shape.py 196608 320 320 residual
  timeout 60 python tools/gemm_shape.py 196608 960 320
  timeout 60 python tools/gemm_shape.py 196608 2560 320 geglu
  timeout 60 python tools/gemm_shape.py 49152 640 640 residual
  timeout 60 python tools/gemm_shape.py 12288 10240 1280 geglu
done 2>&1 | tee gpurun_out/${tag}_ab.log
