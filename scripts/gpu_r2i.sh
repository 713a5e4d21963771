#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_r2i.txt
rm -f $S
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "$name exit $?" | tee -a $S; }
run t_model 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py tests/test_gpu_fusion.py -q -m gpu -p no:cacheprovider --timeout 900
tail -n 6 gpurun_out/t_model.log | tee -a $S
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-inference --no-kernel-rooflines"
run ab_defer_t3 600 $B
JL_GEMM_TAIL=6 run ab_defer_t2 600 $B
JL_DEFER_WGRADS=0 JL_GEMM_TAIL=6 run ab_nodefer_t2 600 $B
JL_LN_WGRAD=main run ab_defer_t3_lnmain 600 $B
run ab_large 900 python bench.py --config large --steps 10 --warmup 3 --no-inference --no-cpu-baseline --no-kernel-rooflines
JL_DEFER_WGRADS=0 JL_GEMM_TAIL=6 run ab_large_nodefer 900 python bench.py --config large --steps 10 --warmup 3 --no-inference --no-cpu-baseline --no-kernel-rooflines
run ab_mixed 900 python bench.py --config mixed --steps 10 --warmup 3 --no-inference --no-cpu-baseline --no-kernel-rooflines
for f in ab_defer_t3 ab_defer_t2 ab_nodefer_t2 ab_defer_t3_lnmain ab_large ab_large_nodefer ab_mixed; do python -c "
import json
d=json.load(open('gpurun_out/$f.log'))
print('$f', round(d['value']), round(d['ms_per_step'],3), 'gemm', round(d['roofline']['achieved']), 'enc', round(d['roofline']['encoder_gemms']['achieved']), 'launches', d['gpu_launches_per_step'])
" | tee -a $S; done
