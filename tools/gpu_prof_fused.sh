#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_yx" \
    --launch-skip 2 --launch-count 2 -f -o gpurun_out/prof_fused4 \
    python tools/profile_eval.py --config C2 --evals 1 > gpurun_out/prof_fused4.log 2>&1
ncu -i gpurun_out/prof_fused4.ncu-rep --page raw --csv > gpurun_out/prof_fused4_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_fused4.ncu-rep --page source --csv --kernel-name regex:k_yx_density > gpurun_out/prof_fused4_src_density.csv 2>/dev/null
ncu -i gpurun_out/prof_fused4.ncu-rep --page source --csv --kernel-name regex:k_yx_vmul > gpurun_out/prof_fused4_src_vmul.csv 2>/dev/null
rm -f gpurun_out/prof_fused4.ncu-rep
