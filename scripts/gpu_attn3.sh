#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/attn_only.py > gpurun_out/attn_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_fused -s 2 -c 1 -f -o gpurun_out/prof_attn_fused \
   python scripts/attn_only.py > gpurun_out/ncu_attn_fused.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_attn_fused.log
