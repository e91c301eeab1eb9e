#!/bin/bash
# P GPUs: micro-benchmark of the cross-rank primitives, then the frame-sharded checks and bench.py at that world size.
tag=${1:-r2x}
P=${2:-4}
K=${3:-3}
W=${4:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$P --master-addr 127.0.0.1 --master-port 29641 \
  tools/xrank_bench.py > gpurun_out/${tag}_xbench.log 2>&1
grep -E "^\{" gpurun_out/${tag}_xbench.log | tee gpurun_out/${tag}_xbench.json; grep -E "Error|error|Traceback" gpurun_out/${tag}_xbench.log | head -5
bash tools/gpu_trip_r2_shard.sh $tag $P $K $W
