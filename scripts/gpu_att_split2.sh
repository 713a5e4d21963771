#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_att_split2.txt
rm -f $S
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "$name exit $?" | tee -a $S; }
L="--steps 10 --warmup 3 --no-inference --no-cpu-baseline --no-kernel-rooflines"
JL_ATT_COL_SPLIT=2 run lz2 600 python bench.py --config large $L
JL_ATT_COL_SPLIT=1 run lz1 600 python bench.py --config large $L
JL_ATT_COL_SPLIT=2 run mz2 600 python bench.py --config mixed $L
JL_ATT_COL_SPLIT=1 run mz1 600 python bench.py --config mixed $L
JL_ATT_COL_SPLIT=2 JL_LNPROJ_COL_SPLIT=2 run mz22 600 python bench.py --config mixed $L
for f in lz2 lz1 mz2 mz1 mz22; do python -c "
import json
d=json.load(open('gpurun_out/$f.log'))
print('$f', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches_per_step'])
" | tee -a $S; done
