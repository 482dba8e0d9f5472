#!/bin/bash
# ncu --set full of every kernel of the second C3a evaluation + serialised launch list
mkdir -p gpurun_out
CFG=${1:-C3a}
N=${2:-38}
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_" \
    --launch-skip $N --launch-count $N -f -o gpurun_out/prof_${CFG}_full \
    python tools/profile_eval.py --config $CFG --evals 2 > gpurun_out/prof_${CFG}_full.log 2>&1
ncu -i gpurun_out/prof_${CFG}_full.ncu-rep --page raw --csv > gpurun_out/prof_${CFG}_full_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_${CFG}_full_raw.csv > gpurun_out/prof_${CFG}_full_summary.txt
rm -f gpurun_out/prof_${CFG}_full.ncu-rep
tail -45 gpurun_out/prof_${CFG}_full_summary.txt
