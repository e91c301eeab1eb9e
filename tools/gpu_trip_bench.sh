#!/bin/bash
# tests + bench (both arms) + ncu launch list of the bench command + full capture of the dominant kernel
tag=${1:-t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header 2>&1 | grep -vE "^$" | tail -15 > gpurun_out/${tag}_tests.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
KR="regex:gemm_tc|attention_tc|cross_attention|temporal_attention|gn_|layernorm|colstats|attn_shift|upsample2x|space_to_depth|pack_latents|unpack_latents|timestep_emb|ddim_step|latent_|mask_resize|axpby"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" -c 2700 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-animatediff > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"attention_tc_split_kernel" -s 12 -c 1 -f -o gpurun_out/${tag}_attn python tools/time_unet.py 16 1 > gpurun_out/${tag}_ncu_attn.log 2>&1
tail -3 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
