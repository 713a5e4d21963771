#!/bin/bash
# tests + bench, no profiler
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee gpurun_out/summary3.txt
tail -n 8 gpurun_out/pytest_gpu.log | tee -a gpurun_out/summary3.txt
timeout 900 python bench.py --steps 20 --warmup 5 --gemm-breakdown gpurun_out/gemm_breakdown.md > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" | tee -a gpurun_out/summary3.txt
cat gpurun_out/bench.json | tee -a gpurun_out/summary3.txt
tail -n 5 gpurun_out/bench.err | tee -a gpurun_out/summary3.txt
