#!/bin/bash
# Round-end style run: full GPU test-suite, smoke, bench (ours + reference arm), ncu launch list and one full capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee gpurun_out/summary2.txt
tail -n 5 gpurun_out/pytest_gpu.log | tee -a gpurun_out/summary2.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary2.txt
tail -n 2 gpurun_out/smoke.log | tee -a gpurun_out/summary2.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" | tee -a gpurun_out/summary2.txt
cat gpurun_out/bench.json | tee -a gpurun_out/summary2.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?" | tee -a gpurun_out/summary2.txt
cat gpurun_out/bench_ref.json | tee -a gpurun_out/summary2.txt
if [ "$1" == "ncu" ]; then
  timeout 600 python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
  echo "ncu list exit $?" | tee -a gpurun_out/summary2.txt
  timeout 600 python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 470 -c 8 -o gpurun_out/prof_gemm \
     python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?" | tee -a gpurun_out/summary2.txt
fi
