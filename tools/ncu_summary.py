#!/usr/bin/env python
"""Summarise `ncu --page raw --csv` exports: one row per profiled launch with the metrics that
decide which roof a kernel sits under.  Usage: ncu_summary.py raw1.csv [raw2.csv ...]"""
import csv
import sys

WANT = [
  ('gpu__time_duration.sum', 'us', 1e-3),
  ('dram__bytes_read.sum', 'rdMB', 1e-6),
  ('dram__bytes_write.sum', 'wrMB', 1e-6),
  ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%', 1),
  ('lts__t_bytes.sum', 'l2MB', 1e-6),
  ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%', 1),
  ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smem%', 1),
  ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'fp64%', 1),
  ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'fp64cyc%', 1),
  ('smsp__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active', 'dmma%', 1),
  ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%', 1),
  ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%', 1),
  ('launch__registers_per_thread', 'regs', 1),
  ('launch__grid_size', 'grid', 1),
  ('launch__block_size', 'blk', 1),
  ('launch__occupancy_limit_registers', 'limR', 1),
  ('launch__occupancy_limit_shared_mem', 'limS', 1),
  ('smsp__cycles_active.avg', 'cyc', 1),
  ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'bankconf', 1),
  # the L1 data pipe carries shared-memory AND global/local wavefronts: the resource that binds
  # the plane kernels once psi(r) streams through it
  ('l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed', 'l1pipe%', 1),
]


def num(s):
  try:
    return float(s.replace(',', ''))
  except Exception:
    return None


for path in sys.argv[1:]:
  rows = list(csv.reader(open(path)))
  hdr = rows[0]
  units = rows[1]
  ki = hdr.index('Kernel Name')
  cols = {}
  for name, _, _ in WANT:
    if name in hdr:
      cols[name] = hdr.index(name)
  print('#', path)
  print('kernel'.ljust(34), ' '.join(lbl.rjust(8) for n, lbl, _ in WANT if n in cols))
  for r in rows[2:]:
    if len(r) <= ki:
      continue
    out = []
    for name, lbl, scale in WANT:
      if name not in cols:
        continue
      v = num(r[cols[name]])
      u = units[cols[name]]
      if v is None:
        out.append('-'.rjust(8))
        continue
      if lbl == 'us':
        v = v * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}.get(u, 1e-3)
      elif lbl.endswith('MB'):
        v = v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}.get(u, 1e-6)
      out.append(f'{v:8.1f}')
    print(r[ki].split('(')[0].replace('void jrb::', '')[:34].ljust(34), ' '.join(out))
