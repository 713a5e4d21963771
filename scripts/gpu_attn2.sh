#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout 300 -x -k attention > gpurun_out/pytest_attn.log 2>&1; echo "pytest attn exit $?"
tail -n 4 gpurun_out/pytest_attn.log
timeout 300 python scripts/microbench.py 2>&1 | grep attn | tee gpurun_out/microbench_attn.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; echo "bench exit $?"
cut -c 1-330 gpurun_out/bench_iter.json
