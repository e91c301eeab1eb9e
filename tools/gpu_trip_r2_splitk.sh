#!/bin/bash
# Round 2, one GPU: split-K kernel tests + per-rank shapes of an 8- / 4-GPU frame-sharded run with and without split-K.
tag=${1:-r2k}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --no-header -x -k "split_k or gemm or conv3x3" 2>&1 | grep -vE "^$" | tail -15 > gpurun_out/${tag}_tests.log
for spec in "2 20 --branches 3 --idx 5 --truncate" "2 20 --branches 1 --idx 30" "4 10 --branches 3 --idx 5 --truncate" "4 10 --branches 1 --idx 30"; do
  timeout 300 python tools/time_unet.py $spec --graph 2>&1 | grep UNet >> gpurun_out/${tag}_small.log
  timeout 300 python tools/time_unet.py $spec --graph --splitk 2>&1 | grep UNet | sed 's/$/ splitk/' >> gpurun_out/${tag}_small.log
done
KR="regex:gemm_tc|attention_tc|cross_attention|temporal_attention|gn_|layernorm|colstats|attn_shift|upsample2x|space_to_depth|pack_latents|unpack_latents|timestep_emb|ddim_step|latent_|mask_resize|axpby|set_floats"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" -c 2400 --csv --log-file gpurun_out/${tag}_launches_1x2_splitk.csv python tools/time_unet.py 2 1 --branches 1 --idx 30 --splitk > /dev/null 2>&1
tail -6 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_small.log
