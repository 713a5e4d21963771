#!/bin/bash
# Runs the GPU test groups in separate processes (a trapped kernel poisons its CUDA context) and logs to gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() {
  name=$1; shift
  timeout 900 python -m pytest "$@" -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/test_$name.log 2>&1
  echo "$name exit $?" | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/test_$name.log | tee -a gpurun_out/summary.txt
}
rm -f gpurun_out/summary.txt
run gemm_plain tests/test_gpu_gemm.py -k "plain"
run gemm_mn tests/test_gpu_gemm.py -k "mn_major"
run gemm_epi tests/test_gpu_gemm.py -k "not plain and not mn_major"
run mel tests/test_gpu_mel.py
run kernels tests/test_gpu_kernels.py
run model tests/test_gpu_model.py
