#!/bin/bash
# Round 2, call A: new full-size parity tests, the whole GPU suite, smoke, bench (base / mixed packed+padded / large).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_r2.json gpurun_out/summary_r2a.txt
S=gpurun_out/summary_r2a.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "$name exit $?" | tee -a $S; }
run t_fullsize 1500 python -m pytest tests/test_gpu_fullsize.py -q -m gpu -p no:cacheprovider --timeout 900
tail -n 25 gpurun_out/t_fullsize.log | tee -a $S
run t_rest 1800 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 900 --deselect tests/test_gpu_fullsize.py
tail -n 15 gpurun_out/t_rest.log | tee -a $S
run smoke 600 python -c "import __graft_entry__ as g; g.smoke()"
tail -n 2 gpurun_out/smoke.log | tee -a $S
run bench_base 900 python bench.py --steps 20 --warmup 5 --gemm-breakdown gpurun_out/gemm_breakdown_r2a.md
cat gpurun_out/bench_base.log | cut -c1-3000 | tee -a $S; tail -n 5 gpurun_out/bench_base.err | tee -a $S
run bench_mixed_packed 600 python bench.py --config mixed --steps 10 --warmup 3 --no-inference --no-cpu-baseline --no-kernel-rooflines
cat gpurun_out/bench_mixed_packed.log | cut -c1-1500 | tee -a $S; tail -n 5 gpurun_out/bench_mixed_packed.err | tee -a $S
run bench_mixed_padded 600 python bench.py --config mixed --padded --steps 10 --warmup 3 --no-inference --no-cpu-baseline --no-kernel-rooflines
cat gpurun_out/bench_mixed_padded.log | cut -c1-1500 | tee -a $S; tail -n 5 gpurun_out/bench_mixed_padded.err | tee -a $S
run bench_large 900 python bench.py --config large --steps 10 --warmup 3 --no-inference --no-cpu-baseline --no-kernel-rooflines
cat gpurun_out/bench_large.log | cut -c1-1500 | tee -a $S; tail -n 5 gpurun_out/bench_large.err | tee -a $S
