#!/bin/bash
# Builds ffi/libjrb_xla_ffi.so on a machine that has jax >= 0.5.3 (for xla/ffi/api/ffi.h) and the
# CUDA toolkit; libjrystal_b200.so must have been built first (jrystal_b200/csrc/build.sh).
set -e
cd "$(dirname "$0")"
INC=$(python -c "import jax; print(jax.ffi.include_dir())")
CUDA=${CUDA_HOME:-/usr/local/cuda}
g++ -std=c++17 -O2 -shared -fPIC -I"$INC" -I"$CUDA/include" -I../include jrb_xla_ffi.cc \
    -L../jrystal_b200/csrc -ljrystal_b200 -L"$CUDA/lib64" -lcudart \
    -Wl,-rpath,'$ORIGIN/../jrystal_b200/csrc' -o libjrb_xla_ffi.so
echo "built $(pwd)/libjrb_xla_ffi.so"
