#!/bin/bash
tag=${1:-r2g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header 2>&1 | grep -vE "^$" | tail -8 > gpurun_out/${tag}_tests.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${tag}_bench1.json 2> gpurun_out/${tag}_bench1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/${tag}_bench2.json 2> gpurun_out/${tag}_bench2.err
tail -4 gpurun_out/${tag}_tests.log
python - <<PY
import json
for f in ("bench1", "bench2"):
    try:
        d = json.loads(open("gpurun_out/${tag}_%s.json" % f).read().strip().splitlines()[-1])
        c = d["config"]
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], c.get("all_branches_every_step"), c.get("cuda_graphs"), c.get("frame_sharding"),
              c["animatediff_v2_backbone"].get("frame_sharded"), c["animatediff_v2_backbone"]["ms_per_clip"])
    except Exception as e:
        print(f, "no line:", e)
PY
grep -hE "Error|assert" gpurun_out/${tag}_bench1.err gpurun_out/${tag}_bench2.err | head -5
