#!/bin/bash
# one ncu --set full capture (with source) of the 64^2-level attention launch
tag=${1:-t}
mkdir -p gpurun_out
UNIVST_ATTN_VARIANT=${VARIANT:-0} timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"attention_tc_kernel<\(int\)2, \(int\)${BKV:-128}" -s ${SKIP:-22} -c 1 -f -o gpurun_out/${tag}_attn python tools/time_unet.py 16 1 > gpurun_out/${tag}_ncu_attn.log 2>&1
tail -3 gpurun_out/${tag}_ncu_attn.log
