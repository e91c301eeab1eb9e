#!/bin/bash
# 2 GPUs: SD frame-sharded forward, NCCL halo exchange vs K/V rows pushed into the peers' symmetric memory
tag=${1:-sdp}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_frame_sharding_gpu.py -m gpu -x -q --no-header 2>&1 | tail -5 | tee gpurun_out/${tag}_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29641 \
  tools/check_frame_sharding.py 16 64 --push 2>&1 | grep -v "^W\|^\*\*\*" | tail -4 | tee gpurun_out/${tag}_full.log
