#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_r2k.txt
rm -f $S
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "$name exit $?" | tee -a $S; }
run t_k 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_fullsize.py tests/test_gpu_fusion.py -q -m gpu -p no:cacheprovider --timeout 900
tail -n 8 gpurun_out/t_k.log | tee -a $S
L="--steps 10 --warmup 3 --no-inference --no-cpu-baseline --no-kernel-rooflines"
run ab_large_fused 900 python bench.py --config large $L
JL_FUSED_WF_TRAIN=0 run ab_large_composed 900 python bench.py --config large $L
run ab_mixed_fused 900 python bench.py --config mixed $L
JL_FUSED_WF_TRAIN=0 run ab_mixed_composed 900 python bench.py --config mixed $L
for f in ab_large_fused ab_large_composed ab_mixed_fused ab_mixed_composed; do python -c "
import json
d=json.load(open('gpurun_out/$f.log'))
print('$f', round(d['value']), round(d['ms_per_step'],3), 'gemm', round(d['roofline']['achieved']), 'enc', round(d['roofline']['encoder_gemms']['achieved']), 'launches', d['gpu_launches_per_step'])
" | tee -a $S; done
