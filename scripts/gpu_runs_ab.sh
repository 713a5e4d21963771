#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_runs.txt
rm -f $S
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "$name exit $?" | tee -a $S; }
run t_runs 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_fullsize.py tests/test_gpu_fusion.py -q -m gpu -p no:cacheprovider --timeout 300 -x
tail -n 4 gpurun_out/t_runs.log | tee -a $S
if grep -q "failed\|rror" gpurun_out/t_runs.log; then grep -n "Error\|assert\|FAILED" gpurun_out/t_runs.log | head -20 | tee -a $S; exit 1; fi
L="--steps 10 --warmup 3 --no-inference --no-cpu-baseline --no-kernel-rooflines"
run mr1 600 python bench.py --config mixed $L
JL_WF_MULTI_RUN=0 run mr0 600 python bench.py --config mixed $L
run mr1b 600 python bench.py --config mixed $L
JL_WF_MULTI_RUN=0 run mr0b 600 python bench.py --config mixed $L
for f in mr1 mr0 mr1b mr0b; do python -c "
import json
d=json.load(open('gpurun_out/$f.log'))
print('$f', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches_per_step'], 'loss', d['loss'])
" | tee -a $S; done
