#!/bin/bash
tag=${1:-t21}
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"temporal_attention_mma_kernel<\(int\)3" -s 4 -c 1 -f -o gpurun_out/${tag}_tattn python tools/time_unet.py 16 1 --animatediff > gpurun_out/${tag}_ncu_tattn.log 2>&1
tail -3 gpurun_out/${tag}_ncu_tattn.log
