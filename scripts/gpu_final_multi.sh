#!/bin/bash
# Final multi-GPU run (gpurun --gpus N): N-rank == 1-rank gradient check, bench base / large / mixed at N ranks, inference sweep at N ranks.
# usage: gpu_final_multi.sh N [tag]
N=${1:-2}
T=${2:-r2m}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_${T}_multi_$N.txt
rm -f $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR scripts/ddp_gpu_check.py > gpurun_out/${T}_ddp_check_${N}gpu.log 2>&1; echo "ddp_check exit $?" | tee -a $S
grep -E "collective|world|DDP_CHECK" gpurun_out/${T}_ddp_check_${N}gpu.log | tee -a $S
L="--no-cpu-baseline --no-inference --no-kernel-rooflines"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 $L > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err; echo "bench base exit $?" | tee -a $S
timeout 900 $TR bench.py --gpus $N --config large --steps 10 --warmup 3 $L > gpurun_out/${T}_bench_large_${N}gpu.json 2> gpurun_out/${T}_bench_large_${N}gpu.err; echo "bench large exit $?" | tee -a $S
timeout 900 $TR bench.py --gpus $N --config mixed --steps 10 --warmup 3 $L > gpurun_out/${T}_bench_mixed_${N}gpu.json 2> gpurun_out/${T}_bench_mixed_${N}gpu.err; echo "bench mixed exit $?" | tee -a $S
echo "sweep skipped" | tee -a $S
for f in ${T}_bench_${N}gpu ${T}_bench_large_${N}gpu ${T}_bench_mixed_${N}gpu; do python - <<PY | tee -a $S
import json
try:
    d=json.load(open('gpurun_out/$f.json'))
    print('$f', 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'sync', round(d['e2e']['sync_value']), d.get('exchange',{}).get('mode'), d['config'].get('load_imbalance_max_over_mean'))
except Exception as e:
    print('$f', 'FAILED', e)
PY
done

