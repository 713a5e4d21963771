#!/bin/bash
# Profiling pass: ncu launch list of one eager fine-tune step + full captures of the GEMM, attention, LayerNorm, mel and CTC kernels.
# Raw pages are exported to CSV on the box; only the GEMM report (with source) is kept as .ncu-rep (gpurun_out <= 64 MiB).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline --no-inference"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
timeout 600 $CMD > gpurun_out/plain2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 330 -c 12 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm exit $?"
ncu -i gpurun_out/prof_gemm.ncu-rep --page raw --csv > gpurun_out/prof_gemm_raw.csv 2>/dev/null
timeout 600 $CMD > gpurun_out/plain3.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none -k "regex:mel_fbank|cmvn|ctc_|attn_|layernorm_fwd|layernorm_bwd_kernel|colsum" -s 170 -c 18 -o /tmp/prof_other $CMD > gpurun_out/ncu_other.log 2>&1
echo "ncu other exit $?"
ncu -i /tmp/prof_other.ncu-rep --page raw --csv > gpurun_out/prof_other_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
du -sh gpurun_out
