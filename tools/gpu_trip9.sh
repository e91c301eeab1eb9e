#!/bin/bash
# attention timing only (no parity) for experimental variants
tag=${1:-t}
mkdir -p gpurun_out
for cfg in $CFGS; do
  v=${cfg%%:*}; st=${cfg##*:}
  echo "=== variant $v stagger $st"
  export UNIVST_ATTN_VARIANT=$v
  if [ "$st" = "-1" ]; then unset UNIVST_ATTN_STAGGER; else export UNIVST_ATTN_STAGGER=$st; fi
  timeout 300 python tools/time_unet.py 16 3 --shapes 2>&1 | grep -E "forward:|== sc_attention|\(48, 8, (40|80|160), [0-9]+, [0-9][0-9][0-9]+\)"
done > gpurun_out/${tag}_variants.log 2>&1
cat gpurun_out/${tag}_variants.log
