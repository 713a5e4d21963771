#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_r2n.txt
rm -f $S
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "$name exit $?" | tee -a $S; }
run t_att 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout 60 -k attadapter -x
tail -n 15 gpurun_out/t_att.log | tee -a $S
if grep -q "failed\|error" gpurun_out/t_att.log; then exit 1; fi
JL_B200_LIB=$PWD/jiao-liao_speech_recognition_b200/libjl_b200_aatiming.so run att_timing 200 python scripts/att_one.py
cat gpurun_out/att_timing.log | tee -a $S
run att_bench 300 python scripts/att_bench.py
grep "rows=8000\|rows=32000\|rows=250 " gpurun_out/att_bench.log | tee -a $S
L="--steps 20 --warmup 5 --no-inference --no-cpu-baseline --no-kernel-rooflines"
JL_FUSED_ATT=1 run ab_att_fused 600 python bench.py $L
JL_FUSED_ATT=0 run ab_att_composed 600 python bench.py $L
JL_FUSED_ATT=1 run ab_att_fused2 600 python bench.py $L
JL_FUSED_ATT=0 run ab_att_composed2 600 python bench.py $L
for f in ab_att_fused ab_att_composed ab_att_fused2 ab_att_composed2; do python -c "
import json
d=json.load(open('gpurun_out/$f.log'))
print('$f', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches_per_step'])
" | tee -a $S; done
