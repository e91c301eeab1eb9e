#!/bin/bash
# One GPU trip: kernel + UNet + pipeline parity, full-size timing, ncu captures.  Usage: tools/gpu_trip.sh <tag>
tag=${1:-t}
mkdir -p gpurun_out
for k in gemm conv3x3 sc_attention cross_attention; do
  timeout 240 python -m pytest tests/test_kernels_gpu.py -m gpu -q --no-header -rA -k "$k" 2>&1 | grep -E "max_abs_err|PASSED|FAILED|passed|failed|Error" | tail -40
done > gpurun_out/${tag}_kernels.log 2>&1
timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_pipeline_gpu.py -m gpu -q --no-header -rA 2>&1 | grep -E "rel=|PASSED|FAILED|passed|failed|Error|error" | tail -40 > gpurun_out/${tag}_unet.log
timeout 300 python tools/time_unet.py 16 3 > gpurun_out/${tag}_time.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_tc|attention_tc|gn_|layernorm|colstats|attn_shift|upsample2x|space_to_depth|pack_latents|unpack_latents|timestep_emb|ddim_step|latent_|mask_resize|axpby" -s 896 -c 448 --csv --log-file gpurun_out/${tag}_launches.csv python tools/time_unet.py 16 1 > gpurun_out/${tag}_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"attention_tc_kernel<2, 128>" -s 22 -c 2 -o gpurun_out/${tag}_attn python tools/time_unet.py 16 1 > gpurun_out/${tag}_ncu_attn.log 2>&1
grep -E "failed|FAILED" gpurun_out/${tag}_kernels.log gpurun_out/${tag}_unet.log | head; tail -2 gpurun_out/${tag}_unet.log; tail -2 gpurun_out/${tag}_time.log
