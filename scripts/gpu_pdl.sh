#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 0 1; do
  JL_PDL=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pdl$v.json 2> gpurun_out/bench_pdl$v.err
  echo "JL_PDL=$v exit $?"; python -c "
import json;d=json.load(open('gpurun_out/bench_pdl$v.json'));print(d['ms_per_step'], d['value'], d['e2e']['value'])"
done
