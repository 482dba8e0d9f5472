#!/bin/bash
# round-1b profiles: launch lists of C2 / C3a with the orbital grid, ncu --set full of one whole C2 evaluation
mkdir -p gpurun_out
bash tools/gpu_launch_list.sh C2 400 | head -24
cp gpurun_out/launches_C2.csv gpurun_out/launches_C2_og.csv
bash tools/gpu_launch_list.sh C3a 400 | head -24
cp gpurun_out/launches_C3a.csv gpurun_out/launches_C3a_og.csv
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 70 --launch-count 70 -f -o gpurun_out/prof_C2_og \
    python tools/profile_eval.py --config C2 --evals 2 > gpurun_out/prof_C2_og.log 2>&1
ncu -i gpurun_out/prof_C2_og.ncu-rep --page raw --csv > gpurun_out/prof_C2_og_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_C2_og_raw.csv > gpurun_out/prof_C2_og_summary.txt
rm -f gpurun_out/prof_C2_og.ncu-rep
cut -c1-200 gpurun_out/prof_C2_og_summary.txt | head -80
