#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_lp2.txt
rm -f $S
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "$name exit $?" | tee -a $S; }
L="--steps 20 --warmup 5 --no-inference --no-cpu-baseline --no-kernel-rooflines"
for i in 1 2 3; do
JL_LP_WGRAD=1 run ab_w1_$i 600 python bench.py $L
JL_LP_WGRAD=0 run ab_w0_$i 600 python bench.py $L
done
for f in ab_w1_1 ab_w0_1 ab_w1_2 ab_w0_2 ab_w1_3 ab_w0_3; do python -c "
import json
d=json.load(open('gpurun_out/$f.log'))
print('$f', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches_per_step'])
" | tee -a $S; done
