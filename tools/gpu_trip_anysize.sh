#!/bin/bash
tag=${1:-any}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --no-header -k "conv3x3" 2>&1 | tail -15 | tee gpurun_out/${tag}_conv.log
timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -x -q --no-header -k "edge_shapes" -s 2>&1 | grep -v "^$" | tail -15 | tee gpurun_out/${tag}_unet.log
