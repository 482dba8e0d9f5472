#!/bin/bash
# build + ABI check here, then run "$@" on a B200 through gpurun (usage: tools/gpu.sh TIMEOUT cmd...)
set -e
cd "$(dirname "$0")/.."
bash jrystal_b200/csrc/build.sh > /dev/null
python -m pytest tests/test_abi.py -x -q 2>&1 | tail -1
T=$1; shift
exec gpurun --timeout $T -- "$*"
