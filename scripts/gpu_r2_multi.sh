#!/bin/bash
# Multi-GPU run (gpurun --gpus N): N-rank == 1-rank gradient check, bench base + mixed at N ranks (exchange captured in the step graph),
# and the same with the exchange outside the graph (A/B).  usage: gpu_r2_multi.sh N
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_multi_$N.txt
rm -f $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR scripts/ddp_gpu_check.py > gpurun_out/ddp_check_$N.log 2>&1; echo "ddp_check exit $?" | tee -a $S
grep -E "collective|world|DDP_CHECK" gpurun_out/ddp_check_$N.log | tee -a $S
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}.json 2> gpurun_out/bench_${N}.err; echo "bench base exit $?" | tee -a $S
JL_EXCHANGE_IN_GRAPH=0 timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}_serial.json 2> gpurun_out/bench_${N}_serial.err; echo "bench base (exchange after the graph) exit $?" | tee -a $S
timeout 900 $TR bench.py --gpus $N --config mixed --steps 10 --warmup 3 > gpurun_out/bench_mixed_${N}.json 2> gpurun_out/bench_mixed_${N}.err; echo "bench mixed exit $?" | tee -a $S
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-inference --no-kernel-rooflines > gpurun_out/bench_1_same_box.json 2> gpurun_out/bench_1_same_box.err; echo "bench N=1 exit $?" | tee -a $S
for f in bench_${N} bench_${N}_serial bench_mixed_${N} bench_1_same_box; do python - <<PY | tee -a $S
import json
try:
    d=json.load(open('gpurun_out/$f.json'))
    print('$f', 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'serial_e2e', round(d['e2e']['serial_value']), d.get('exchange'), d['config'].get('load_imbalance_max_over_mean'))
except Exception as e:
    print('$f', 'FAILED', e)
PY
done
tail -n 3 gpurun_out/bench_${N}.err | tee -a $S
