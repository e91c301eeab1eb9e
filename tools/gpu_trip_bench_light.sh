#!/bin/bash
# all GPU tests + bench (both arms), no profiler passes
tag=${1:-t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header 2>&1 | grep -vE "^$" | tail -15 > gpurun_out/${tag}_tests.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
tail -3 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
