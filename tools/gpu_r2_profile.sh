#!/bin/bash
# Round 2 profiles: launch lists (gpu__time_duration) of bench.py for C2 / C3a, ncu --set full of one
# whole C2 (and C3a) evaluation with the raw page exported, per-kernel summary, stamped traffic.
mkdir -p gpurun_out
for cfg in C2 C3a; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_$cfg.csv \
      python bench.py --config $cfg --steps 2 --warmup 1 --no-cpu --no-graph > gpurun_out/r02_launches_$cfg.log 2>&1
done
for cfg in C2 C3a; do
  n=$(python - <<PY
import csv,re
rows=[r for r in csv.reader(l for l in open('gpurun_out/r02_launches_$cfg.csv') if l.startswith('"'))]
k=rows[0].index('Kernel Name')
idx=[i for i,r in enumerate(rows[1:]) if 'k_pack_energies' in r[k]]
print(idx[1]-idx[0] if len(idx)>1 else 70)
PY
)
  timeout 1200 ncu --set full --clock-control none --import-source on --launch-skip $((n+1)) --launch-count $n -f -o gpurun_out/r02_full_$cfg \
      python tools/profile_eval.py --config $cfg --evals 2 > gpurun_out/r02_full_$cfg.log 2>&1
  ncu -i gpurun_out/r02_full_$cfg.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_${cfg}_raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/r02_ncu_full_${cfg}_raw.csv > gpurun_out/r02_ncu_full_${cfg}_summary.txt
  python tools/capture_traffic.py $cfg gpurun_out/r02_ncu_full_${cfg}_raw.csv "profiles/r02_ncu_full_${cfg}_raw.csv (ncu --set full of one whole evaluation, $n launches)"
  cp profiles/traffic.json gpurun_out/r02_traffic.json
  # keep the source-level page of the two dominant kernels, drop the big report
  ncu -i gpurun_out/r02_full_$cfg.ncu-rep --page source --csv --print-source sass -k regex:k_yx 2>/dev/null | head -c 3000000 > gpurun_out/r02_ncu_source_${cfg}_kyx.csv
  rm -f gpurun_out/r02_full_$cfg.ncu-rep
  cut -c1-190 gpurun_out/r02_ncu_full_${cfg}_summary.txt | head -40
done
