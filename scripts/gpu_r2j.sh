#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_r2j.txt
rm -f $S
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "$name exit $?" | tee -a $S; }
run t_k 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_fullsize.py tests/test_gpu_fusion.py -q -m gpu -p no:cacheprovider --timeout 900
tail -n 6 gpurun_out/t_k.log | tee -a $S
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-inference --no-kernel-rooflines"
JL_GEMM_TAIL=6 run ab_merged_t2 600 $B
run ab_merged_t3 600 $B
JL_MERGED_REDUCE=0 JL_GEMM_TAIL=6 run ab_sep_t2 600 $B
JL_MERGED_REDUCE=0 run ab_sep_t3 600 $B
JL_GEMM_TAIL=6 run ab_large_merged 900 python bench.py --config large --steps 10 --warmup 3 --no-inference --no-cpu-baseline --no-kernel-rooflines
JL_MERGED_REDUCE=0 JL_GEMM_TAIL=6 run ab_large_sep 900 python bench.py --config large --steps 10 --warmup 3 --no-inference --no-cpu-baseline --no-kernel-rooflines
for f in ab_merged_t2 ab_merged_t3 ab_sep_t2 ab_sep_t3 ab_large_merged ab_large_sep; do python -c "
import json
d=json.load(open('gpurun_out/$f.log'))
print('$f', round(d['value']), round(d['ms_per_step'],3), 'gemm', round(d['roofline']['achieved']), 'enc', round(d['roofline']['encoder_gemms']['achieved']), 'launches', d['gpu_launches_per_step'])
" | tee -a $S; done
