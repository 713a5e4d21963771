#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_wf.txt
rm -f $S
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "$name exit $?" | tee -a $S; }
run t_wf 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_fullsize.py tests/test_gpu_fusion.py -q -m gpu -p no:cacheprovider --timeout 600 -x
tail -n 12 gpurun_out/t_wf.log | tee -a $S
if grep -q "failed\|rror" gpurun_out/t_wf.log; then exit 1; fi
L="--steps 10 --warmup 3 --no-inference --no-cpu-baseline --no-kernel-rooflines"
run wf_large_on 600 python bench.py --config large $L
JL_FUSED_WF_BWD=0 run wf_large_off 600 python bench.py --config large $L


for f in wf_large_on wf_large_off; do python -c "
import json
d=json.load(open('gpurun_out/$f.log'))
print('$f', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches_per_step'], 'loss', d['loss'])
" | tee -a $S; done
