#!/bin/bash
# Build a variant of libjl_b200.so with extra nvcc flags (A/B experiments): scripts/build_variant.sh <name> <flags...>
# → jiao-liao_speech_recognition_b200/libjl_b200_<name>.so; run with JL_B200_LIB=<that path>.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
PKG=jiao-liao_speech_recognition_b200
OBJ=build/obj_$name
mkdir -p $OBJ
SRCS="common gemm_tcgen05 mel_cmvn layernorm ctc elementwise attention attention_tc wfadapter_tc comm w2v_frontend fusion attadapter_tc lnproj_bwd_tc"
for s in $SRCS; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC "$@" -c $PKG/csrc/$s.cu -o $OBJ/$s.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $PKG/libjl_b200_$name.so $(for s in $SRCS; do echo $OBJ/$s.o; done) -lcudart_static -ldl -lrt -lpthread
echo built $PKG/libjl_b200_$name.so
