#!/bin/bash
tag=${1:-t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header 2>&1 | grep -vE "^$" | tail -40 > gpurun_out/${tag}_tests.log
timeout 300 python tools/time_unet.py 16 3 --shapes > gpurun_out/${tag}_time.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"attention_tc_kernel<\(int\)2, \(int\)128>" -s 22 -c 2 -o gpurun_out/${tag}_attn python tools/time_unet.py 16 1 > gpurun_out/${tag}_ncu_attn.log 2>&1
tail -4 gpurun_out/${tag}_tests.log; grep "forward:" gpurun_out/${tag}_time.log; grep -A3 "== sc_attention" gpurun_out/${tag}_time.log
