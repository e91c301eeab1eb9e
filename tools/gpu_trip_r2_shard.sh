#!/bin/bash
# Round 2, P GPUs (default 2): frame-sharded forward through the xrank transport (peer memory + device-side flags), eager and
# as a replayed CUDA graph, SD and AnimateDiff backbones, next to the NCCL paths; then bench.py at that world size.
# Run: gpurun --gpus 2 -- bash tools/gpu_trip_r2_shard.sh tag 2 [steps] [warmup]
tag=${1:-r2s}
P=${2:-2}
K=${3:-2}
W=${4:-3}
mkdir -p gpurun_out
for flavour in "" "--animatediff"; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$P --master-addr 127.0.0.1 --master-port 29651 \
    tools/check_frame_sharding.py 16 64 --push --xrank $flavour > gpurun_out/${tag}_shard${flavour}.log 2>&1
  grep -E "^\{" gpurun_out/${tag}_shard${flavour}.log | tee -a gpurun_out/${tag}_shard.json
  grep -E "Error|error|Traceback" gpurun_out/${tag}_shard${flavour}.log | head -5
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29655 \
  bench.py --gpus $P --steps $K --warmup $W > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 3000 gpurun_out/${tag}_bench.json; grep -vE "^$|NCCL version" gpurun_out/${tag}_bench.err | tail -15
