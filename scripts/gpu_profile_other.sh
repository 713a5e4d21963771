#!/bin/bash
# Full ncu capture of the non-GEMM kernels of one (warm) fine-tune step: the window starts at the final LayerNorm of the
# forward pass and covers the CTC kernels and the first layers of the backward pass (column sums, LayerNorm backward and
# weight gradients, fused attention backward for 1 and 12 heads).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline --no-inference"
timeout 600 $CMD > gpurun_out/plain3.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none -k "regex:mel_fbank|cmvn|ctc_|attn_|layernorm_fwd|layernorm_bwd_kernel|layernorm_wgrad|colsum" -s 225 -c 26 -f -o /tmp/prof_other $CMD > gpurun_out/ncu_other.log 2>&1
echo "ncu other exit $?"
ncu -i /tmp/prof_other.ncu-rep --page raw --csv > gpurun_out/prof_other_raw.csv 2>/dev/null
tail -2 gpurun_out/ncu_other.log
