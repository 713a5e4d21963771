#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_att_split.txt
rm -f $S
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "$name exit $?" | tee -a $S; }
run t_att 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout 60 -k "attadapter" -x
tail -n 3 gpurun_out/t_att.log | tee -a $S
if grep -q "failed\|rror" gpurun_out/t_att.log; then grep -n "Error\|assert" gpurun_out/t_att.log | head; exit 1; fi
L="--steps 20 --warmup 5 --no-inference --no-cpu-baseline --no-kernel-rooflines"
for i in 1 2; do
JL_ATT_COL_SPLIT=2 run ab_z2_$i 600 python bench.py $L
JL_ATT_COL_SPLIT=1 run ab_z1_$i 600 python bench.py $L
done
for f in ab_z2_1 ab_z1_1 ab_z2_2 ab_z1_2; do python -c "
import json
d=json.load(open('gpurun_out/$f.log'))
print('$f', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches_per_step'])
" | tee -a $S; done
