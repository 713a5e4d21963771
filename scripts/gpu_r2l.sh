#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_r2l.txt
rm -f $S
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "$name exit $?" | tee -a $S; }
run t_att 400 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout 120 -k attadapter
tail -n 15 gpurun_out/t_att.log | tee -a $S
run t_model 900 python -m pytest tests/test_gpu_model.py -q -m gpu -p no:cacheprovider --timeout 300 -x
tail -n 8 gpurun_out/t_model.log | tee -a $S
run att_bench 300 python scripts/att_bench.py
cat gpurun_out/att_bench.log | tee -a $S
L="--steps 20 --warmup 5 --no-inference --no-cpu-baseline --no-kernel-rooflines"
run ab_att_fused 600 python bench.py $L
JL_FUSED_ATT=0 run ab_att_composed 600 python bench.py $L
run ab_att_fused2 600 python bench.py $L
JL_FUSED_ATT=0 run ab_att_composed2 600 python bench.py $L
run ab_large_att_fused 600 python bench.py --config large --steps 10 --warmup 3 --no-inference --no-cpu-baseline --no-kernel-rooflines
JL_FUSED_ATT=0 run ab_large_att_composed 600 python bench.py --config large --steps 10 --warmup 3 --no-inference --no-cpu-baseline --no-kernel-rooflines
for f in ab_att_fused ab_att_composed ab_att_fused2 ab_att_composed2 ab_large_att_fused ab_large_att_composed; do python -c "
import json
d=json.load(open('gpurun_out/$f.log'))
print('$f', round(d['value']), round(d['ms_per_step'],3), 'gemm', round(d['roofline']['achieved']), 'launches', d['gpu_launches_per_step'])
" | tee -a $S; done
