#!/bin/bash
tag=${1:-r2v}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vae_gpu.py tests/test_pipeline_gpu.py tests/test_maskprop_flowwarp_gpu.py tests/test_kernels_gpu.py -m gpu -q --no-header -x 2>&1 | grep -vE "^$" | tail -30 > gpurun_out/${tag}_tests.log
timeout 300 python tools/time_vae.py 16 > gpurun_out/${tag}_vae.json 2> gpurun_out/${tag}_vae.err
tail -12 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_vae.json; tail -3 gpurun_out/${tag}_vae.err
