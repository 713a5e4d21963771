#!/bin/bash
# quick check: adapter / model tests + bench lines of the three training configs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_quick.txt
rm -f $S
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "$name exit $?" | tee -a $S; }
run t_q 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -q -m gpu -p no:cacheprovider --timeout 600 -x
tail -n 3 gpurun_out/t_q.log | tee -a $S
L="--steps 20 --warmup 5 --no-inference --no-cpu-baseline --no-kernel-rooflines"
run q1 600 python bench.py $L
run q2 600 python bench.py $L
run q_large 600 python bench.py --config large --steps 10 --warmup 3 --no-inference --no-cpu-baseline --no-kernel-rooflines
run q_mixed 600 python bench.py --config mixed --steps 10 --warmup 3 --no-inference --no-cpu-baseline --no-kernel-rooflines
for f in q1 q2 q_large q_mixed; do python -c "
import json
d=json.load(open('gpurun_out/$f.log'))
print('$f', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches_per_step'], 'gemm', round(d['roofline']['achieved']))
" | tee -a $S; done
