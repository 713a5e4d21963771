#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_r2h.txt
rm -f $S
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "$name exit $?" | tee -a $S; }
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-inference --no-kernel-rooflines"
run ab_t3 600 $B
JL_SIDE_PRIORITY=-1 run ab_t3_hi 600 $B
JL_SIDE_PRIORITY=-1 JL_GEMM_TAIL=6 run ab_t2_hi 600 $B
JL_GEMM_TAIL=6 run ab_t2 600 $B
JL_SIDE_PRIORITY=-1 JL_LN_WGRAD=main run ab_t3_hi_lnmain 600 $B
for f in ab_t3 ab_t3_hi ab_t2_hi ab_t2 ab_t3_hi_lnmain; do python -c "
import json
d=json.load(open('gpurun_out/$f.log'))
print('$f', round(d['value']), round(d['ms_per_step'],3), 'gemm', round(d['roofline']['achieved']), 'enc', round(d['roofline']['encoder_gemms']['achieved']), 'launches', d['gpu_launches_per_step'])
" | tee -a $S; done
