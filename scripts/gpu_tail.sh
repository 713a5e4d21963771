#!/bin/bash
# column-slice tail wave: GEMM tests, A/B table, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu -p no:cacheprovider --timeout 500 -x > gpurun_out/pytest_gemm.log 2>&1; echo "pytest gemm exit $?" | tee gpurun_out/summary_tail.txt
tail -n 6 gpurun_out/pytest_gemm.log | tee -a gpurun_out/summary_tail.txt
timeout 300 python scripts/gemm_tail_ab.py > gpurun_out/gemm_tail_ab.md 2>&1; echo "ab exit $?" | tee -a gpurun_out/summary_tail.txt
cat gpurun_out/gemm_tail_ab.md | tee -a gpurun_out/summary_tail.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tail.json 2> gpurun_out/bench_tail.err; echo "bench exit $?" | tee -a gpurun_out/summary_tail.txt
cat gpurun_out/bench_tail.json | cut -c 1-400 | tee -a gpurun_out/summary_tail.txt
