#!/usr/bin/env python
"""Instruction mix of the hot kernels from the built objects (cuobjdump -sass): the mnemonics that
show which pipes a kernel uses -- DMMA (FP64 tensor core), DFMA/DADD/DMUL (FP64 pipe), LDGSTS
(cp.async), LDS/STS (shared memory), BAR (barriers).  No GPU needed.
  python tools/sass_mix.py > profiles/r01_sass_mix.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ['k_gram<', 'k_apply<', 'k_yx_vmul<64', 'k_yx_density<64', 'k_yx_vmul<81', 'k_yx_density<81',
        'k_yx128_vmul', 'k_yx128_density', 'k_z_inv_scatter<49', 'k_z_fwd_gather<49',
        'k_z_inv_scatter<64', 'k_z_inv_scatter<81', 'k_gga_local', 'k_near_identity', 'k_resample',
        'k_chol_panel', 'k_nl_project']
COLS = ['DMMA', 'DFMA', 'DADD', 'DMUL', 'LDGSTS', 'LDG', 'STG', 'LDS', 'STS', 'BAR', 'total']
print('kernel'.ljust(58), ' '.join(c.rjust(7) for c in COLS))
for obj in sorted(glob.glob(os.path.join(ROOT, 'jrystal_b200', 'csrc', 'build', '*.o'))):
  out = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
  name, mix = None, None
  rows = []
  for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
      if name:
        rows.append((name, mix))
      name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
      name = name.replace('void jrb::', '').replace('jrb::', '').split('(')[0]
      mix = collections.Counter()
      continue
    m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
    if m and mix is not None:
      mix[m.group(1)] += 1
      mix['total'] += 1
  if name:
    rows.append((name, mix))
  for name, mix in rows:
    if any(name.startswith(w) for w in WANT):
      print(name[:58].ljust(58), ' '.join(str(mix.get(c, 0)).rjust(7) for c in COLS))
