#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_ln_ab.txt
rm -f $S
L="--steps 20 --warmup 5 --no-inference --no-cpu-baseline"
for v in base lnh lnh4 base lnh; do
  if [ $v = base ]; then unset JL_B200_LIB; else export JL_B200_LIB=$PWD/jiao-liao_speech_recognition_b200/libjl_b200_$v.so; fi
  timeout 600 python bench.py $L > gpurun_out/ab_$v.log 2> gpurun_out/ab_$v.err
  python -c "
import json
d=json.load(open('gpurun_out/ab_$v.log'))
hk=d['roofline'].get('hbm_kernels',{})
ks={k:(round(v['us'],2) if isinstance(v,dict) and 'us' in v else None) for k,v in (hk.items() if isinstance(hk,dict) else [])}
print('$v', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), ks)
" | tee -a $S
done
