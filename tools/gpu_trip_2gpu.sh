#!/bin/bash
# 2 GPUs: frame-sharded forward checks (SD and AnimateDiff); then (GPU 0) ncu capture of the split attention kernel
tag=${1:-t29}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29641 tools/check_frame_sharding.py 16 64 --animatediff 2>&1 | grep -E "^\{|Error|error" | tee gpurun_out/${tag}_shard_ad.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29642 tools/check_frame_sharding.py 16 64 2>&1 | grep -E "^\{|Error|error" | tee gpurun_out/${tag}_shard_sd.json
timeout 900 python -m pytest tests/test_frame_sharding_gpu.py -m gpu -x -q --no-header 2>&1 | tail -3 | tee gpurun_out/${tag}_tests.log
CUDA_VISIBLE_DEVICES=0 timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"attention_tc_split_kernel" -s 12 -c 1 -f -o gpurun_out/${tag}_attn python tools/time_unet.py 16 1 > gpurun_out/${tag}_ncu_attn.log 2>&1
tail -2 gpurun_out/${tag}_ncu_attn.log
