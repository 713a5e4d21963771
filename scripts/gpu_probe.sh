#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/gemm_epi_probe.py > gpurun_out/probe_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 5 -c 5 -f -o gpurun_out/prof_epi \
   python scripts/gemm_epi_probe.py > gpurun_out/ncu_epi.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_epi.log
