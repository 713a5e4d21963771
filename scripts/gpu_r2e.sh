#!/bin/bash
# Round 2, call E: f4 fusion tests; GEMM tests (three-way tail slices); A/B timing: tail slices 2 vs 3, LN wgrad side vs main, no side branch.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_r2e.txt
rm -f $S
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "$name exit $?" | tee -a $S; }
run t_fusion 900 python -m pytest tests/test_gpu_fusion.py -q -m gpu -p no:cacheprovider --timeout 600
tail -n 12 gpurun_out/t_fusion.log | tee -a $S
run t_gemm 1200 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_model.py -q -m gpu -p no:cacheprovider --timeout 600
tail -n 6 gpurun_out/t_gemm.log | tee -a $S
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-inference --no-kernel-rooflines"
run ab_base 600 $B --gemm-breakdown gpurun_out/gemm_r2e.md
JL_GEMM_TAIL=6 run ab_tail2 600 $B
JL_LN_WGRAD=main run ab_lnmain 600 $B
JL_DEBUG_SKIP_SIDE=1 run ab_noside 600 $B
run ab_base2 600 $B
for f in ab_base ab_tail2 ab_lnmain ab_noside ab_base2; do python -c "
import json
d=json.load(open('gpurun_out/$f.log'))
print('$f', round(d['value']), round(d['ms_per_step'],3), 'gemm', round(d['roofline']['achieved']), 'enc', round(d['roofline']['encoder_gemms']['achieved']), 'launches', d['gpu_launches_per_step'])
" | tee -a $S; done
cat gpurun_out/gemm_r2e.md | head -12 | tee -a $S
