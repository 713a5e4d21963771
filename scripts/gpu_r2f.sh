#!/bin/bash
# Round 2, call F: A/B on one box — tail slices 3-way vs 2-way, with and without the weight-gradient side branch (timing only).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_r2f.txt
rm -f $S
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "$name exit $?" | tee -a $S; }
run t_fusion 900 python -m pytest tests/test_gpu_fusion.py -q -m gpu -p no:cacheprovider --timeout 600
tail -n 4 gpurun_out/t_fusion.log | tee -a $S
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-inference --no-kernel-rooflines"
for rep in 1 2; do
  run ab_t3_$rep 600 $B
  JL_GEMM_TAIL=6 run ab_t2_$rep 600 $B
  JL_GEMM_TAIL=0 run ab_t0_$rep 600 $B
done
JL_DEBUG_SKIP_SIDE=1 run ab_noside_t3 600 $B
JL_DEBUG_SKIP_SIDE=1 JL_GEMM_TAIL=6 run ab_noside_t2 600 $B
for f in ab_t3_1 ab_t2_1 ab_t0_1 ab_t3_2 ab_t2_2 ab_t0_2 ab_noside_t3 ab_noside_t2; do python -c "
import json
d=json.load(open('gpurun_out/$f.log'))
print('$f', round(d['value']), round(d['ms_per_step'],3), 'gemm', round(d['roofline']['achieved']), 'enc', round(d['roofline']['encoder_gemms']['achieved']), 'launches', d['gpu_launches_per_step'])
" | tee -a $S; done
