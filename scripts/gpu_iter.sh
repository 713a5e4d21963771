#!/bin/bash
# one tuning iteration: full GPU test-suite, an optional tuning script ($1, output to gpurun_out/$2), bench without the CPU arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee gpurun_out/summary_iter.txt
tail -n 6 gpurun_out/pytest_gpu.log | tee -a gpurun_out/summary_iter.txt
if [ -n "$1" ]; then
  timeout 400 python "$1" > "gpurun_out/$2" 2>&1; echo "$1 exit $?" | tee -a gpurun_out/summary_iter.txt
  cat "gpurun_out/$2" | tee -a gpurun_out/summary_iter.txt
fi
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --gemm-breakdown gpurun_out/gemm_breakdown.md > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; echo "bench exit $?" | tee -a gpurun_out/summary_iter.txt
cut -c 1-330 gpurun_out/bench_iter.json | tee -a gpurun_out/summary_iter.txt
tail -n 3 gpurun_out/bench_iter.err | tee -a gpurun_out/summary_iter.txt
