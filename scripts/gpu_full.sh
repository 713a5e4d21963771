#!/bin/bash
# full GPU check: the -m gpu suite, smoke(), the default bench line and the reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_full.txt
rm -f $S
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "$name exit $?" | tee -a $S; }
run pytest_gpu 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 900 -x
tail -n 6 gpurun_out/pytest_gpu.log | tee -a $S
run smoke 600 python -c "import __graft_entry__ as g; g.smoke()"
tail -n 3 gpurun_out/smoke.log | tee -a $S
run bench_full 900 python bench.py --steps 20 --warmup 5
python -c "
import json
d=json.load(open('gpurun_out/bench_full.log'))
print('bench', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'sync', round(d['e2e']['sync_value']), 'serial', round(d['e2e']['serial_value']), 'gemm frac', d['roofline']['frac'], 'cpu', d['cpu_baseline'])
" | tee -a $S
