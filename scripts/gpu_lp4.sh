#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_lp4.txt
rm -f $S
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "$name exit $?" | tee -a $S; }
L="--steps 20 --warmup 5 --no-inference --no-cpu-baseline --no-kernel-rooflines"
for i in 1 2; do
JL_LNPROJ_SPLIT=2 run ab_s2_$i 600 python bench.py $L
JL_LNPROJ_SPLIT=1 run ab_s1_$i 600 python bench.py $L
done
JL_LNPROJ_SPLIT=2 run ab_s2_large 600 python bench.py --config large --steps 10 --warmup 3 --no-inference --no-cpu-baseline --no-kernel-rooflines
JL_LNPROJ_SPLIT=1 run ab_s1_large 600 python bench.py --config large --steps 10 --warmup 3 --no-inference --no-cpu-baseline --no-kernel-rooflines
for f in ab_s2_1 ab_s1_1 ab_s2_2 ab_s1_2 ab_s2_large ab_s1_large; do python -c "
import json
d=json.load(open('gpurun_out/$f.log'))
print('$f', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches_per_step'])
" | tee -a $S; done
