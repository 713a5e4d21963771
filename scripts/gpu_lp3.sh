#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_lp3.txt
rm -f $S
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "$name exit $?" | tee -a $S; }
run t_lp 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout 60 -k "lnproj" -x
tail -n 3 gpurun_out/t_lp.log | tee -a $S
JL_LP_WGRAD=2 run t_m2 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py -q -m gpu -p no:cacheprovider --timeout 600 -x
tail -n 2 gpurun_out/t_m2.log | tee -a $S
L="--steps 20 --warmup 5 --no-inference --no-cpu-baseline --no-kernel-rooflines"
for i in 1 2; do
JL_LP_WGRAD=2 run ab_w2_$i 600 python bench.py $L
JL_LP_WGRAD=0 run ab_w0_$i 600 python bench.py $L
done
for f in ab_w2_1 ab_w0_1 ab_w2_2 ab_w0_2; do python -c "
import json
d=json.load(open('gpurun_out/$f.log'))
print('$f', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches_per_step'])
" | tee -a $S; done
