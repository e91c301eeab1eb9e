#!/bin/bash
# Round 2, one GPU: GPU test suite + bench (both arms).  Run: gpurun -- bash tools/gpu_trip_r2_single.sh tag
tag=${1:-r2a}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -x 2>&1 | grep -vE "^$" | tail -25 > gpurun_out/${tag}_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
tail -8 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
