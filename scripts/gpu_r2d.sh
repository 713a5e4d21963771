#!/bin/bash
# Round 2, call D: ncu launch list of one eager step + one --set full window (last forward layer → CTC → first backward layer).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_r2d.txt
rm -f $S
P="python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline --no-inference --no-kernel-rooflines"
timeout 600 $P > gpurun_out/plain.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r2d.csv $P > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?" | tee -a $S
gzip -f gpurun_out/launches_r2d.csv
timeout 600 $P > gpurun_out/plain2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none -s 1205 -c 75 -o /tmp/prof_r2d $P > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?" | tee -a $S
ls -la /tmp/prof_r2d.ncu-rep | tee -a $S
ncu -i /tmp/prof_r2d.ncu-rep --page raw --csv > gpurun_out/prof_r2d_raw.csv 2>> $S
gzip -f gpurun_out/prof_r2d_raw.csv
sz=$(stat -c %s /tmp/prof_r2d.ncu-rep)
if [ "$sz" -lt 40000000 ]; then cp /tmp/prof_r2d.ncu-rep gpurun_out/; fi
du -sh gpurun_out | tee -a $S
