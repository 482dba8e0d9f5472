#!/usr/bin/env python
"""Per-kernel roofline table from a tools/ncu_summary.py summary of ONE whole evaluation
(every launch captured with `ncu --set full`): launches, time and share of the evaluation,
achieved DRAM GB/s against the measured peak, the busiest on-chip resource, and what bounds it.
Times under ncu are serialised and cold-cache: the SHARES are what carries over to the bench.

  python tools/kernel_table.py profiles/r01_ncu_full_c2_orbital_grid_summary.txt [peak_GB/s]"""
import collections
import re
import sys


def main(path, peak=6550.4):
  rows = collections.OrderedDict()
  for line in open(path):
    if line.startswith('#') or line.startswith('kernel') or not line.strip():
      continue
    m = re.match(r'(.{1,40}?)\s+([-\d.]+(?:\s+[-\d.]+){16,17})\s*$', line.rstrip())
    if not m:
      continue
    name = re.sub(r'^jrb::', '', m.group(1).strip())
    v = [float(x) for x in m.group(2).split()]
    us, rd, wr, dram, l2, smem, fp64, _, sm = v[:9]
    l1 = v[17] if len(v) > 17 else smem  # l1pipe% (older summaries: shared-memory share only)
    r = rows.setdefault(name, dict(n=0, us=0.0, mb=0.0, smem=0.0, fp64=0.0, sm=0.0, l2=0.0, l1=0.0))
    r['n'] += 1
    r['us'] += us
    r['mb'] += rd + wr
    for k, val in (('smem', smem), ('fp64', fp64), ('sm', sm), ('l2', l2), ('l1', l1)):
      r[k] += val * us
  total = sum(r['us'] for r in rows.values())
  print(f'| kernel | launches | µs | share | DRAM GB/s (frac of {peak:.0f}) | FP64 pipe % | smem % | L1 data pipe % | L2 % | bound by |')
  print('|---|---|---|---|---|---|---|---|---|---|')
  for name, r in sorted(rows.items(), key=lambda kv: -kv[1]['us']):
    if r['us'] < 0.005 * total:
      continue
    gbs = r['mb'] / r['us'] * 1e3 if r['us'] else 0.0
    fp64, smem, l2, l1 = r['fp64'] / r['us'], r['smem'] / r['us'], r['l2'] / r['us'], r['l1'] / r['us']
    cands = {'DRAM': gbs / peak * 100, 'FP64 issue': fp64, 'L1 data pipe (shared + global)': l1, 'L2': l2}
    top = max(cands, key=cands.get)
    if 'gram' in name or 'apply' in name:
      top = 'FP64 tensor (DMMA) pipe'        # sm% is the DMMA pipe for these (fp64% counts DFMA only)
    elif cands[top] < 30:
      top = 'latency (small grid / dependent chain)'
    print(f'| `{name}` | {r["n"]} | {r["us"]:.0f} | {r["us"] / total * 100:.1f} % | {gbs:.0f} ({gbs / peak:.2f}) | '
          f'{fp64:.0f} | {smem:.0f} | {l1:.0f} | {l2:.0f} | {top} |')
  print(f'\ntotal under ncu: {total / 1e3:.2f} ms over {sum(r["n"] for r in rows.values())} launches')


if __name__ == '__main__':
  main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 6550.4)
