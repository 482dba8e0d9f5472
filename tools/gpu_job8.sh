#!/bin/bash
mkdir -p gpurun_out
./tools/fp64_lds_mix.bin | tee gpurun_out/fp64_lds_mix.txt
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_yx" \
    --launch-skip 2 --launch-count 2 -f -o gpurun_out/prof_fused3 \
    python tools/profile_eval.py --config C2 --evals 1 > gpurun_out/prof_fused3.log 2>&1
ncu -i gpurun_out/prof_fused3.ncu-rep --page raw --csv > gpurun_out/prof_fused3_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_fused3.ncu-rep --page source --csv --kernel-name regex:k_yx_density > gpurun_out/prof_fused3_src_density.csv 2>/dev/null
ls -la gpurun_out | head
