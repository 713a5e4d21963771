#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/microbench_frontback.py > gpurun_out/probe_fb_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mel_fbank|cmvn|ctc_row_stats|ctc_grad" -s 8 -c 4 -f -o gpurun_out/prof_fb2 \
   python scripts/microbench_frontback.py > gpurun_out/ncu_fb2.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_fb2.log
