#!/bin/bash
# final-build evidence on one GPU: ncu launch list of the bench command, per-shape timings, shard shapes
tag=${1:-r2i}
mkdir -p gpurun_out
KR="regex:gemm_tc|attention_tc|cross_attention|temporal_attention|gn_|layernorm|colstats|attn_shift|upsample2x|space_to_depth|pack_latents|unpack_latents|timestep_emb|ddim_step|latent_|mask_resize|axpby|set_floats"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" -c 3000 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-cuda-graphs > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 300 python tools/time_unet.py 16 3 --shapes > gpurun_out/${tag}_shapes.log 2>&1
for spec in "16 5 --branches 3 --idx 5 --truncate" "16 5 --branches 1 --idx 30" "4 10 --branches 3 --idx 5 --truncate --splitk" "4 10 --branches 1 --idx 30 --splitk" "2 20 --branches 3 --idx 5 --truncate --splitk" "2 20 --branches 1 --idx 30 --splitk"; do
  timeout 300 python tools/time_unet.py $spec --graph 2>&1 | grep UNet >> gpurun_out/${tag}_small.log
done
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"gemm_tc_kernel" -s 300 -c 1 -f -o gpurun_out/${tag}_splitk python tools/time_unet.py 2 1 --branches 1 --idx 30 --splitk > gpurun_out/${tag}_ncu_splitk.log 2>&1
cat gpurun_out/${tag}_small.log; head -3 gpurun_out/${tag}_shapes.log
