#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
L="--no-cpu-baseline --no-inference --no-kernel-rooflines"
timeout 400 $TR bench.py --gpus 2 --config large --steps 10 --warmup 3 $L > gpurun_out/r2y_bench_large_2gpu.json 2> gpurun_out/r2y_large.err; echo "large exit $?"
timeout 400 $TR bench.py --gpus 2 --config mixed --steps 10 --warmup 3 $L > gpurun_out/r2y_bench_mixed_2gpu.json 2> gpurun_out/r2y_mixed.err; echo "mixed exit $?"
for f in r2y_bench_large_2gpu r2y_bench_mixed_2gpu; do python -c "
import json
d=json.load(open('gpurun_out/$f.json'))
print('$f', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))
"; done
