#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout 300 -k "attention" > gpurun_out/pytest_attn.log 2>&1; echo "attn exit $?" | tee gpurun_out/summary_attn.txt
tail -n 30 gpurun_out/pytest_attn.log | tee -a gpurun_out/summary_attn.txt
