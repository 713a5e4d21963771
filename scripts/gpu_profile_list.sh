#!/bin/bash
# Launch list (ncu gpu__time_duration) of one eager fine-tune step + a SHORT --set full capture: the first launches of the kernels named
# by the regex in the 4th step.  Usage: scripts/gpu_profile_list.sh <tag>
cd "$(dirname "$0")/.."
T=${1:-r2v}
mkdir -p gpurun_out
S=gpurun_out/summary_$T.txt
rm -f $S
P="python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline --no-inference --no-kernel-rooflines"
timeout 600 $P > gpurun_out/${T}_plain.log 2>&1
echo "plain exit $?" | tee -a $S
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${T}_launches.csv $P > gpurun_out/${T}_ncu_list.log 2>&1
echo "ncu list exit $?" | tee -a $S
python scripts/summarize_launches.py gpurun_out/${T}_launches.csv 3 > gpurun_out/${T}_launches_summary.md 2>> $S
head -14 gpurun_out/${T}_launches_summary.md | tee -a $S
gzip -f gpurun_out/${T}_launches.csv
timeout 900 ncu --set full --clock-control none -k "regex:attadapter_fwd|lnproj_bwd_kernel|lnfold_pack_multi|lnproj_bwd_reduce" -s 30 -c 8 -f -o /tmp/prof_$T $P > gpurun_out/${T}_ncu_sel.log 2>&1
echo "ncu selected exit $?" | tee -a $S
ncu -i /tmp/prof_$T.ncu-rep --page raw --csv > gpurun_out/${T}_ncu_sel_raw.csv 2>> $S
gzip -f gpurun_out/${T}_ncu_sel_raw.csv
