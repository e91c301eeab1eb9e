#!/bin/bash
# Round 2, one GPU, the evidence run: GPU test suite, bench (both arms), ncu launch list of the bench command, one
# ncu --set full capture of the dominant kernel (attention, 48 images) and of the K = 320 GEMM, per-rank shard shapes.
tag=${1:-r2f}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header 2>&1 | grep -vE "^$" | tail -12 > gpurun_out/${tag}_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
KR="regex:gemm_tc|attention_tc|cross_attention|temporal_attention|gn_|layernorm|colstats|attn_shift|upsample2x|space_to_depth|pack_latents|unpack_latents|timestep_emb|ddim_step|latent_|mask_resize|axpby|set_floats"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" -c 3000 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-cuda-graphs > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"attention_tc_split_kernel" -s 12 -c 1 -f -o gpurun_out/${tag}_attn python tools/time_unet.py 16 1 > gpurun_out/${tag}_ncu_attn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"gn_stats_kernel" -s 40 -c 1 -f -o gpurun_out/${tag}_gn python tools/time_unet.py 16 1 > gpurun_out/${tag}_ncu_gn.log 2>&1
for spec in "2 20 --branches 3 --idx 5 --truncate" "2 20 --branches 1 --idx 30" "4 10 --branches 3 --idx 5 --truncate" "4 10 --branches 1 --idx 30" "16 5 --branches 3 --idx 5 --truncate" "16 5 --branches 1 --idx 30"; do
  timeout 300 python tools/time_unet.py $spec --graph 2>&1 | grep UNet >> gpurun_out/${tag}_small.log
  timeout 300 python tools/time_unet.py $spec --graph --splitk 2>&1 | grep UNet | sed 's/$/ splitk/' >> gpurun_out/${tag}_small.log
done
timeout 300 python tools/time_unet.py 16 3 --shapes > gpurun_out/${tag}_shapes.log 2>&1
timeout 300 python tools/time_vae.py 16 > gpurun_out/${tag}_vae.json 2>/dev/null
tail -6 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_small.log
