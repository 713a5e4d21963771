#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_r2m.txt
rm -f $S
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "$name exit $?" | tee -a $S; }
run e2e_probe 300 python scripts/e2e_probe.py
cat gpurun_out/e2e_probe.log | tee -a $S
run t_att 400 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout 120 -k attadapter
tail -n 5 gpurun_out/t_att.log | tee -a $S
run att_one 200 python scripts/att_one.py
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attadapter_fwd -f -o /tmp/prof_att python scripts/att_one.py > gpurun_out/ncu_att.log 2>&1
echo "ncu exit $?" | tee -a $S
ncu -i /tmp/prof_att.ncu-rep --page raw --csv > gpurun_out/prof_att_raw.csv 2>/dev/null
ncu -i /tmp/prof_att.ncu-rep --page source --csv --kernel-id :::4 > gpurun_out/prof_att_source.csv 2>/dev/null
ls -la gpurun_out/prof_att_* | tee -a $S
