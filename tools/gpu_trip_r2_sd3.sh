#!/bin/bash
# P GPUs: BASELINE configs[4] (SD-3.5-medium-shaped MMDiT, 16 x 1024 x 1024, 28 steps), one GPU next to frame-sharded.
tag=${1:-r2d}
P=${2:-2}
mkdir -p gpurun_out
if [ "$P" = "1" ]; then
  timeout 900 python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --config5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29657 \
    bench.py --gpus $P --steps 1 --warmup 1 --no-extras --config5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
fi
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
    print(json.dumps(d["config"].get("sd35_medium_rectified_flow"), indent=1)); print(d["value"], d["ms_per_step"])
except Exception as e:
    print("no line:", e)
PY
grep -vE "^$|NCCL version|OMP_NUM|\*\*\*" gpurun_out/${tag}_bench.err | tail -12
