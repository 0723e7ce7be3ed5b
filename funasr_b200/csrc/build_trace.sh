#!/bin/bash
# Debug variant: libfunasr_b200_trace.so = the normal objects + attention_tc.cu compiled with -DFA_ATT_TRACE (tools/att_trace.py).
set -e
cd "$(dirname "$0")"
bash ./build.sh > /dev/null
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DFA_ATT_TRACE -c attention_tc.cu -o ../_build/attention_tc_trace.o
objs=""
for f in fbank layernorm gemm_f32 gemm_tc attention_f32 fsmn cif decode_ops model offline lstm resample; do objs="$objs ../_build/$f.o"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../libfunasr_b200_trace.so $objs ../_build/attention_tc_trace.o -lcudart
echo "built $(cd ..; pwd)/libfunasr_b200_trace.so"
