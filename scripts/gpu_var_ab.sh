#!/bin/bash
# A/B of library variants: gpu_var_ab.sh <variant> [<variant> ...]  (base = the default library)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_var.txt
rm -f $S
L="--steps 20 --warmup 5 --no-inference --no-cpu-baseline --no-kernel-rooflines"
for rep in 1 2; do
for v in base "$@"; do
  if [ $v = base ]; then unset JL_B200_LIB; else export JL_B200_LIB=$PWD/jiao-liao_speech_recognition_b200/libjl_b200_$v.so; fi
  timeout 600 python bench.py $L > gpurun_out/var_${v}_$rep.log 2> gpurun_out/var_${v}_$rep.err
  python -c "
import json
d=json.load(open('gpurun_out/var_${v}_$rep.log'))
print('$v', $rep, round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))
" | tee -a $S
done
done
