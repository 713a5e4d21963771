#!/bin/bash
# Round 2, call B: re-run the failed groups + f1 tests + inference bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_r2b.txt
rm -f $S
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "$name exit $?" | tee -a $S; }
run t_f1 900 python -m pytest tests/test_gpu_f1.py -q -m gpu -p no:cacheprovider --timeout 600
tail -n 12 gpurun_out/t_f1.log | tee -a $S
run t_fullsize 1500 python -m pytest tests/test_gpu_fullsize.py -q -m gpu -p no:cacheprovider --timeout 900
tail -n 12 gpurun_out/t_fullsize.log | tee -a $S
run t_model 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout 900
tail -n 12 gpurun_out/t_model.log | tee -a $S
run smoke 600 python -c "import __graft_entry__ as g; g.smoke()"
tail -n 2 gpurun_out/smoke.log | tee -a $S
run sweep 900 python scripts/sweep_inference.py
tail -n 20 gpurun_out/sweep.log | tee -a $S
