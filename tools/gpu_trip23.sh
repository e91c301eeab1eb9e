#!/bin/bash
tag=${1:-t23}
mkdir -p gpurun_out
python tools/gemm_shape.py 196608 320 320 residual
python tools/gemm_shape.py 196608 320 320
python tools/gemm_shape.py 196608 960 320
python tools/gemm_shape.py 196608 2560 320 geglu
python tools/gemm_shape.py 196608 320 1280 residual
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"gemm_tc_kernel" -s 5 -c 1 -f -o gpurun_out/${tag}_gemm320 python tools/gemm_shape.py 196608 320 320 residual > gpurun_out/${tag}_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"gemm_tc_kernel" -s 5 -c 1 -f -o gpurun_out/${tag}_gemm960 python tools/gemm_shape.py 196608 960 320 >> gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
