#!/bin/bash
# Round 2, call C: tolerance re-check, 12- vs 16-epilogue-warp GEMM A/B, ncu launch list + full captures of selected kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_r2c.txt
rm -f $S
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "$name exit $?" | tee -a $S; }
run t_fullsize 1500 python -m pytest tests/test_gpu_fullsize.py -q -m gpu -p no:cacheprovider --timeout 900
tail -n 6 gpurun_out/t_fullsize.log | tee -a $S
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-inference"
run ab_w12 600 $B --gemm-breakdown gpurun_out/gemm_w12.md
JL_B200_LIB=$PWD/jiao-liao_speech_recognition_b200/libjl_b200_w16.so run ab_w16 600 $B --no-kernel-rooflines --gemm-breakdown gpurun_out/gemm_w16.md
run ab_w12b 600 $B --no-kernel-rooflines
for f in ab_w12 ab_w16 ab_w12b; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.log'))
print('$f', round(d['value']), round(d['ms_per_step'],3), 'gemm', round(d['roofline']['achieved']), 'enc', round(d['roofline']['encoder_gemms']['achieved']))
for k,v in (d['roofline'].get('hbm_kernels') or {}).items(): print('   ',k, round(v['us'],1),'us', round(v['achieved_gbs']),'GB/s', round(v['frac_of_hbm_peak'],3))
" | tee -a $S; done
# ncu: launch list of one eager step (cold-cache, serialised: shares only)
P="python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline --no-inference --no-kernel-rooflines"
timeout 600 $P > gpurun_out/plain.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r2c.csv $P > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?" | tee -a $S
# ncu full: a few launches of the kernels VERDICT names (skip the warm-up steps: ~370 launches each)
timeout 600 $P > gpurun_out/plain2.log 2>&1 &&
timeout 1800 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05_kernel|attn_bwd_fused|ctc_row_stats|layernorm_fwd|layernorm_bwd|attn_fwd_short|colsum|mel_fbank" -s 1500 -c 60 -o gpurun_out/prof_r2c $P > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?" | tee -a $S
ls -la gpurun_out/*.ncu-rep | tee -a $S
