#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
