#!/usr/bin/env python
"""Write profiles/traffic.json from an `ncu --set full --page raw --csv` export of ONE whole
evaluation of `bench.py`'s workload: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of
the H-apply sweep kernels (k_x_vmul_cached / k_yx_vmul / k_yx128_vmul + k_z_fwd_gather), stamped with the sha256 of
the kernel sources they were captured with.  bench.py refuses the entry when the stamp differs from
the sources the library is built from.   Usage: capture_traffic.py CONFIG raw.csv [source-note]"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

config, raw = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else raw
rows = list(csv.reader(open(raw)))
hdr = rows[0]
ki = hdr.index('Kernel Name')
ri, wi = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
unit_r, unit_w = rows[1][ri], rows[1][wi]
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def num(s):
  return float(s.replace(',', ''))


kernels = {}
total = 0.0
for r in rows[2:]:
  name = r[ki]
  if ('k_x_vmul_cached' in name or 'k_yx_vmul' in name or 'k_yx128_vmul' in name or
      'k_z_fwd_gather' in name):
    short = name.split('jrb::')[-1].split('(')[0]
    b = num(r[ri]) * scale[unit_r] + num(r[wi]) * scale[unit_w]
    k = kernels.setdefault(short, {'launches': 0, 'dram_bytes': 0.0})
    k['launches'] += 1
    k['dram_bytes'] += b
    total += b
path = os.path.join(ROOT, 'profiles', 'traffic.json')
try:
  prof = json.load(open(path))
except Exception:
  prof = {}
wl = bench.build_workload(config)
ngrid = int(__import__('numpy').prod(wl['grid']))
m = wl['kpts'].shape[0] * wl['nb']
prof[config] = {
  'happly_dram_bytes_per_eval': total, 'kernels': kernels, 'kernels_sha': bench.kernels_stamp(),
  'algorithmic_bytes_per_eval': m * (32.0 * ngrid + 32.0 * wl['ng']), 'source': note,
  'note': 'one whole evaluation captured with ncu --set full (every launch); with the psi(r) cache '
          'the sweep streams psi(r) of the density sweep back from HBM (k_x_vmul_cached), so its '
          'traffic is close to the algorithmic figure by design; without it the sweep works on the '
          'kept z columns and moves ~6x fewer bytes',
}
json.dump(prof, open(path, 'w'), indent=1)
print(config, 'H-apply DRAM bytes per evaluation', total, 'algorithmic', prof[config]['algorithmic_bytes_per_eval'],
      'stamp', prof[config]['kernels_sha'])
