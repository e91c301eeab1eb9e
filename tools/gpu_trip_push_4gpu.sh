#!/bin/bash
# 4 (or $2) GPUs: the pushed K/V halo (SD) and the pushed frames <-> pixels exchange (AnimateDiff) beyond the world size they
# were verified on (2), next to the NCCL paths; then bench.py at that world size.  Run: gpurun --gpus 4 -- bash tools/gpu_trip_push_4gpu.sh
tag=${1:-p4}
P=${2:-4}
mkdir -p gpurun_out
for flavour in "" "--animatediff"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$P --master-addr 127.0.0.1 --master-port 29651 \
    tools/check_frame_sharding.py 16 64 --push $flavour 2>&1 | grep -E "^\{|Error|error" | tee -a gpurun_out/${tag}_shard.json
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29655 \
  bench.py --gpus $P --steps 2 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 1200 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
