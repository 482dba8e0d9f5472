#!/bin/bash
# ncu --set full of the two kernels the psi(r) cache changes (second evaluation of C2).
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_yx_density|k_x_vmul_cached" --launch-skip 6 --launch-count 6 -f -o gpurun_out/r02_psi_C2 \
    python tools/profile_eval.py --config C2 --evals 2 > gpurun_out/r02_psi_C2.log 2>&1
tail -2 gpurun_out/r02_psi_C2.log
ncu -i gpurun_out/r02_psi_C2.ncu-rep --page raw --csv > gpurun_out/r02_psi_ncu_C2_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02_psi_ncu_C2_raw.csv > gpurun_out/r02_psi_ncu_C2_summary.txt
cat gpurun_out/r02_psi_ncu_C2_summary.txt | cut -c1-400
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02_psi_ncu_C2_raw.csv')))
hdr=rows[0]
want=[i for i,h in enumerate(hdr) if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h]
extra=[i for i,h in enumerate(hdr) if h in ('l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__lsuin_requests.avg.pct_of_peak_sustained_active','l1tex__m_xbar2l1tex_read_bytes.sum','l1tex__m_l1tex2xbar_write_bytes.sum','smsp__issue_active.avg.pct_of_peak_sustained_active')]
ki=hdr.index('Kernel Name')
for r in rows[2:]:
  if len(r)<len(hdr): continue
  st=sorted(((float(r[i].replace(',','') or 0),hdr[i].replace('smsp__pcsamp_warps_issue_stalled_','')) for i in want),reverse=True)[:7]
  print(r[ki][:40], [(n,int(v)) for v,n in st])
  print('   ', [(hdr[i][:60], r[i]) for i in extra])
PY
