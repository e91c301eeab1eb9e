#!/bin/bash
# Round 2, one GPU: GPU test suite + bench (both arms) + per-rank shapes of an 8-GPU frame-sharded run (2 frames per branch)
# timed eagerly, as a CUDA graph and kernel by kernel (ncu launch list).  Run: gpurun -- bash tools/gpu_trip_r2_single.sh tag
tag=${1:-r2a}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -x 2>&1 | grep -vE "^$" | tail -25 > gpurun_out/${tag}_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
for spec in "2 20 --branches 3 --idx 5 --truncate" "2 20 --branches 1 --idx 30" "16 5 --branches 3 --idx 5 --truncate" "16 5 --branches 1 --idx 30"; do
  timeout 300 python tools/time_unet.py $spec 2>&1 | grep UNet >> gpurun_out/${tag}_small.log
  timeout 300 python tools/time_unet.py $spec --graph 2>&1 | grep UNet >> gpurun_out/${tag}_small.log
done
KR="regex:gemm_tc|attention_tc|cross_attention|temporal_attention|gn_|layernorm|colstats|attn_shift|upsample2x|space_to_depth|pack_latents|unpack_latents|timestep_emb|ddim_step|latent_|mask_resize|axpby|set_floats"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" -c 2400 --csv --log-file gpurun_out/${tag}_launches_3x2.csv python tools/time_unet.py 2 1 --branches 3 --idx 5 --truncate > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" -c 2400 --csv --log-file gpurun_out/${tag}_launches_1x2.csv python tools/time_unet.py 2 1 --branches 1 --idx 30 > /dev/null 2>&1
tail -8 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_small.log
