#!/bin/bash
# Final-build profiles: (1) ncu launch list of one eager fine-tune step (gpu__time_duration), (2) one --set full capture of a whole
# warm step (raw page only: the report itself is too large to bring back), (3) --set full + source of the fused AttAdapter forward.
# Usage: scripts/gpu_profile_final.sh <tag>      → gpurun_out/<tag>_launches.csv.gz, <tag>_ncu_full_raw.csv.gz, <tag>_att_*.csv
cd "$(dirname "$0")/.."
T=${1:-r2z}
mkdir -p gpurun_out
S=gpurun_out/summary_$T.txt
rm -f $S
P="python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline --no-inference --no-kernel-rooflines"
timeout 600 $P > gpurun_out/${T}_plain.log 2>&1
echo "plain exit $?" | tee -a $S
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${T}_launches.csv $P > gpurun_out/${T}_ncu_list.log 2>&1
echo "ncu list exit $?" | tee -a $S
python scripts/summarize_launches.py gpurun_out/${T}_launches.csv 3 > gpurun_out/${T}_launches_summary.md 2>> $S
head -12 gpurun_out/${T}_launches_summary.md | tee -a $S
# index of the first launch of the 4th step in the list → the --set full window
SKIP=$(python - <<PY
import csv
rows=list(csv.DictReader(l for l in open('gpurun_out/${T}_launches.csv') if not l.startswith('==')))
idx=[i for i,r in enumerate(rows) if 'adamw_advance' in r['Kernel Name']]
print(idx[3], idx[4]-idx[3])
PY
)
set -- $SKIP
echo "full capture: skip $1 count $2" | tee -a $S
gzip -f gpurun_out/${T}_launches.csv
timeout 2400 ncu --set full --clock-control none -s $1 -c $2 -f -o /tmp/prof_$T $P > gpurun_out/${T}_ncu_full.log 2>&1
echo "ncu full exit $?" | tee -a $S
ncu -i /tmp/prof_$T.ncu-rep --page raw --csv > gpurun_out/${T}_ncu_full_raw.csv 2>> $S
gzip -f gpurun_out/${T}_ncu_full_raw.csv
timeout 200 python scripts/att_one.py > gpurun_out/${T}_att_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attadapter_fwd -f -o /tmp/prof_att_$T python scripts/att_one.py > gpurun_out/${T}_ncu_att.log 2>&1
echo "ncu att exit $?" | tee -a $S
ncu -i /tmp/prof_att_$T.ncu-rep --page raw --csv > gpurun_out/${T}_att_raw.csv 2>> $S
du -sh gpurun_out | tee -a $S
