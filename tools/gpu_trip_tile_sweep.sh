#!/bin/bash
tag=${1:-t26}
mkdir -p gpurun_out
for ov in "" "320:192" "320:128" "320:256" "320:96" "320:64"; do
  UNIVST_BN_OVERRIDE=$ov python tools/conv_shape.py 48 64 64 320 320
done
for ov in "" "640:256" "640:128" "640:224" "640:192"; do
  UNIVST_BN_OVERRIDE=$ov python tools/conv_shape.py 48 32 32 640 640
done
for ov in "" "1280:192" "1280:128"; do
  UNIVST_BN_OVERRIDE=$ov python tools/conv_shape.py 48 16 16 1280 1280
done
for ov in "" "320:192" "320:128" "320:256"; do
  UNIVST_BN_OVERRIDE=$ov python tools/gemm_shape.py 196608 320 320 residual
done
for ov in "" "960:192" "960:256" "960:240" "960:128"; do
  UNIVST_BN_OVERRIDE=$ov python tools/gemm_shape.py 196608 960 320
done
for ov in "" "640:256" "640:128" "640:224"; do
  UNIVST_BN_OVERRIDE=$ov python tools/gemm_shape.py 49152 640 640 residual
done
