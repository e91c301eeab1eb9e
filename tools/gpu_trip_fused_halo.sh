#!/bin/bash
# 2 GPUs: frame-sharded SD forward with the K/V halo read from peer memory by the attention kernel (fused) vs NCCL exchange
tag=${1:-fh}
mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29651 tools/check_frame_sharding.py 16 64 --fused > gpurun_out/${tag}_run.log 2>&1
grep -E "^\{" gpurun_out/${tag}_run.log | tee gpurun_out/${tag}.json
tail -25 gpurun_out/${tag}_run.log | grep -v "^\{" | tail -20
