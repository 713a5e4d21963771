#!/bin/bash
# ncu of the CTC kernels from the front/back micro-benchmark: launch durations, full capture, source page of the lattice.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -p no:cacheprovider -k dialect > gpurun_out/pytest_dialect.log 2>&1; echo "pytest dialect exit $?"; tail -3 gpurun_out/pytest_dialect.log
CMD="python scripts/microbench_frontback.py"
timeout 300 $CMD > gpurun_out/microbench_fb.txt 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:ctc_" -c 5 -o /tmp/prof_fb $CMD > gpurun_out/ncu_fb.log 2>&1
echo "ncu fb exit $?"
ncu -i /tmp/prof_fb.ncu-rep --page raw --csv > gpurun_out/prof_ctc_raw.csv 2>/dev/null
ncu -i /tmp/prof_fb.ncu-rep --page source --csv -k regex:ctc_lattice > gpurun_out/prof_ctc_lattice_source.csv 2>/dev/null
ls -la gpurun_out | tail -5
