"""The evidence under profiles/ is consistent with the sources it claims to describe (CPU only)."""
import json
import os
import subprocess
import sys

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_traffic_capture_is_stamped_with_the_current_kernel_sources():
  """bench.py refuses profiles/traffic.json when jrystal_b200/csrc changed after the ncu capture;
  this keeps the committed capture current (re-run tools/gpu_r2_box36.sh's capture block after a
  kernel edit)."""
  stamp = bench.kernels_stamp()
  prof = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
  for cfg in ('C2', 'C3a'):
    assert prof[cfg]['kernels_sha'] == stamp, (cfg, prof[cfg]['kernels_sha'], stamp)
    traffic, note = bench.measured_traffic(cfg, 1)
    assert traffic and traffic > 0, note
    # the H-apply sweep with the psi(r) cache moves about what the design must move and less than
    # the SURVEY 8d convention charges on the reference grid
    assert traffic < prof[cfg]['algorithmic_bytes_per_eval']
    assert 'k_x_vmul_cached' in ' '.join(prof[cfg]['kernels'])
  assert bench.measured_traffic('C2', 8)[0] is None   # single-GPU captures only


def test_fp64_flop_count_follows_the_psi_cache():
  wl = bench.build_workload('C1')
  nk, nb, ng = wl['kpts'].shape[0], wl['nb'], wl['ng']
  box = tuple(wl['grid'])
  f3, q3 = bench.fp64_flops(wl['mask'], box, nk * nb, nk, ng, nb, psi_cache=False)
  f2, q2 = bench.fp64_flops(wl['mask'], box, nk * nb, nk, ng, nb, psi_cache=True)
  assert q2 == q3 and 0.6 * f3 < f2 < f3   # one y and one x pass fewer, the z passes stay


def test_ncu_summary_and_kernel_table_read_the_committed_capture(tmp_path):
  raw = os.path.join(ROOT, 'profiles', 'r02_psi_cache_ncu_full_C2_raw.csv')
  summ = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_summary.py'), raw],
                        capture_output=True, text=True, check=True).stdout
  assert 'l1pipe%' in summ and 'k_x_vmul_cached' in summ and 'k_yx_density' in summ
  path = tmp_path / 'summary.txt'
  path.write_text(summ)
  table = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'kernel_table.py'), str(path)],
                         capture_output=True, text=True, check=True).stdout
  row = [l for l in table.splitlines() if 'k_x_vmul_cached' in l][0]
  cells = [c.strip() for c in row.strip('|').split('|')]
  assert int(cells[1]) == 3                      # three launches captured
  assert 60 < float(cells[7]) < 90               # L1 data pipe busy: the binding resource
  assert float(cells[7]) > float(cells[6])       # shared memory is only a part of it
