#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_multicast.py -q -m gpu -p no:cacheprovider --timeout 300 -s > gpurun_out/pytest_mcast.log 2>&1; echo "mcast exit $?" | tee gpurun_out/summary_mcast.txt
tail -n 25 gpurun_out/pytest_mcast.log | tee -a gpurun_out/summary_mcast.txt
