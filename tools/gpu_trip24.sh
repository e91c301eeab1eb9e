#!/bin/bash
tag=${1:-t24}
mkdir -p gpurun_out
python tools/gemm_shape.py 196608 320 320 residual
python tools/gemm_shape.py 196608 320 320
python tools/gemm_shape.py 196608 960 320
python tools/gemm_shape.py 196608 2560 320 geglu
python tools/gemm_shape.py 196608 320 1280 residual
python tools/gemm_shape.py 49152 640 640 residual
python tools/gemm_shape.py 12288 1280 1280 residual
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -x -q --no-header 2>&1 | tail -3
timeout 300 python tools/time_unet.py 16 3 2>&1 | grep forward
