#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_lp.txt
rm -f $S
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "$name exit $?" | tee -a $S; }
run t_lp 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout 60 -k "lnproj" -x
tail -n 12 gpurun_out/t_lp.log | tee -a $S
if grep -q "failed\|rror" gpurun_out/t_lp.log; then exit 1; fi
run t_model 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout 300 -x
tail -n 4 gpurun_out/t_model.log | tee -a $S
L="--steps 20 --warmup 5 --no-inference --no-cpu-baseline --no-kernel-rooflines"
run ab_lp_on 600 python bench.py $L
JL_FUSED_ATT_BWD=0 run ab_lp_off 600 python bench.py $L
run ab_lp_on2 600 python bench.py $L
JL_FUSED_ATT_BWD=0 run ab_lp_off2 600 python bench.py $L
for f in ab_lp_on ab_lp_off ab_lp_on2 ab_lp_off2; do python -c "
import json
d=json.load(open('gpurun_out/$f.log'))
print('$f', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches_per_step'], 'loss', d['loss'])
" | tee -a $S; done
