#!/bin/bash
tag=${1:-t}
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 30 -c 14 -o gpurun_out/${tag}_gemm python tools/time_unet.py 16 1 > gpurun_out/${tag}_ncu_gemm.log 2>&1
tail -2 gpurun_out/${tag}_ncu_gemm.log
