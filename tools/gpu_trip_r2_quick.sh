#!/bin/bash
tag=${1:-r2q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -q --no-header -x 2>&1 | grep -vE "^$" | tail -6 > gpurun_out/${tag}_tests.log
for spec in "16 5 --branches 3 --idx 5 --truncate" "16 5 --branches 1 --idx 30" "2 20 --branches 3 --idx 5 --truncate --splitk" "2 20 --branches 1 --idx 30 --splitk"; do
  timeout 300 python tools/time_unet.py $spec --graph 2>&1 | grep UNet >> gpurun_out/${tag}_small.log
done
timeout 300 python tools/time_unet.py 16 3 --shapes > gpurun_out/${tag}_shapes.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -3 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_small.log; grep -A4 "== gemm" gpurun_out/${tag}_shapes.log; python -c "
import json; d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['torch_eager_fp16']['ms_per_forward_univst_b200'], d['roofline']['avg_ms'], d['clocks'])"
