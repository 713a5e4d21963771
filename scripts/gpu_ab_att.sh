#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_ab_att.txt
rm -f $S
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "$name exit $?" | tee -a $S; }
L="--steps 20 --warmup 5 --no-inference --no-cpu-baseline --no-kernel-rooflines"
JL_FUSED_ATT=1 run ab_att_fused 600 python bench.py $L
JL_FUSED_ATT=0 run ab_att_composed 600 python bench.py $L
JL_FUSED_ATT=1 run ab_att_fused2 600 python bench.py $L
JL_FUSED_ATT=0 run ab_att_composed2 600 python bench.py $L
JL_FUSED_ATT=1 run ab_large_att_fused 600 python bench.py --config large --steps 10 --warmup 3 --no-inference --no-cpu-baseline --no-kernel-rooflines
JL_FUSED_ATT=0 run ab_large_att_composed 600 python bench.py --config large --steps 10 --warmup 3 --no-inference --no-cpu-baseline --no-kernel-rooflines
for f in ab_att_fused ab_att_composed ab_att_fused2 ab_att_composed2 ab_large_att_fused ab_large_att_composed; do python -c "
import json
d=json.load(open('gpurun_out/$f.log'))
print('$f', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches_per_step'])
" | tee -a $S; done
