#!/bin/bash
# Builds jrystal_b200/csrc/libjrystal_b200.so for sm_100a (in-tree; nvcc cross-compiles
# without a GPU).  Usage: build.sh [-j N]
set -e
cd "$(dirname "$0")"
JOBS=${JOBS:-8}
NVCC=${NVCC:-nvcc}
FLAGS="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -diag-suppress 128 -Xcompiler -fPIC -Xcompiler -O2"
SRCS="plan.cu api.cu comm.cu optim.cu nonlocal.cu fft_dispatch.cu grid_kernels.cu qr.cu fft_passes_g0.cu fft_passes_g1.cu fft_passes_g2.cu fft_passes_g3.cu fft_passes_g4.cu fft_passes_g5.cu fft_fused_g0.cu fft_fused_g1.cu fft_fused_g2.cu fft_fused_g3.cu"
mkdir -p build
pids=()
for s in $SRCS; do
  o=build/${s%.cu}.o
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ] || [ -n "$(find . -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer "$o" 2>/dev/null)" ] || [ ../../include/jrystal_b200.h -nt "$o" ]; then
    ( $NVCC $FLAGS -c "$s" -o "$o" ) &
    pids+=($!)
    while [ "$(jobs -rp | wc -l)" -ge "$JOBS" ]; do sleep 0.2; done
  fi
done
for p in "${pids[@]}"; do wait "$p"; done
OBJS=$(for s in $SRCS; do echo build/${s%.cu}.o; done)
$NVCC -shared -o libjrystal_b200.so $OBJS -gencode arch=compute_100a,code=sm_100a -lcudart
echo "built $(pwd)/libjrystal_b200.so"
