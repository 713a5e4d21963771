#!/bin/bash
# usage: gpu_multi.sh N   (run under gpurun --gpus N)
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/ddp_gpu_check.py > gpurun_out/ddp_check_$N.log 2>&1
echo "ddp check exit $?"; grep -E "world|DDP_CHECK" gpurun_out/ddp_check_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err
echo "bench exit $?"; cat gpurun_out/bench_$N.json | cut -c1-400
