"""bench.py workloads are the configurations BASELINE.md section 3 / SURVEY.md 8d name: grid, number
of plane waves, k-points, bands (CPU; no CUDA needed), and the JSON contract of the reference arm."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)
import bench  # noqa: E402
from jrystal_b200 import grid  # noqa: E402

# name: (grid, ng, nk, nb, alias-free minimum 4 gmax + 1)
EXPECT = {
  'C1': ([32, 32, 32], 1139, 8, 24, (29, 29, 29)),
  'C2': ([64, 64, 64], 8409, 64, 66, (49, 49, 49)),
  'C3a': ([128, 128, 128], 29423, 1, 208, (77, 77, 77)),
  'C3b': ([128, 128, 128], 29423, 8, 208, (77, 77, 77)),
  'C4': ([64, 64, 64], 4945, 216, 30, (41, 41, 41)),
  'C5': ([48, 48, 48], 1917, 64, 15, (33, 33, 33)),
}


@pytest.mark.parametrize('name', list(EXPECT))
def test_workload_matches_baseline_table(name):
  g, ng, nk, nb, need = EXPECT[name]
  wl = bench.build_workload(name)
  assert wl['grid'] == g and wl['ng'] == ng and wl['nb'] == nb
  assert wl['kpts'].shape == (nk, 3)
  assert wl['occ'].shape == (1, nk, nb)
  assert grid.min_orbital_grid(wl['mask']) == need
  # every benchmark grid is alias free, so the orbital grid is exact for all of them
  assert all(m <= n for m, n in zip(need, g))
  if not wl['band_mode']:
    # occupation.uniform: 2 / nk on the lowest ceil(N_e / 2) bands, zero on the empty ones
    ne = wl['crystal'].num_electron
    assert abs(wl['occ'].sum() - ne) < 1e-9 * ne
    assert (wl['occ'][0, 0, int(np.ceil(ne / 2)):] == 0).all()


def test_auto_orbital_boxes_of_the_benchmarks():
  box = {n: grid.orbital_grid_candidates(EXPECT[n][0], EXPECT[n][4])[0] for n in EXPECT}
  assert box == {'C1': (32, 32, 32), 'C2': (64, 64, 49), 'C3a': (81, 81, 81), 'C3b': (81, 81, 81),
                 'C4': (64, 64, 49), 'C5': (36, 36, 36)}


def test_synthetic_params_are_rank_independent():
  """Every rank draws the same global stream and keeps its k block."""
  full_re, full_im = bench.synthetic_params(50, 4, 3, 0, 4)
  for k0, k1 in ((0, 2), (2, 4), (1, 2)):
    re, im = bench.synthetic_params(50, 4, 3, k0, k1)
    np.testing.assert_array_equal(re, full_re[:, k0:k1])
    np.testing.assert_array_equal(im, full_im[:, k0:k1])
  assert 0.0 <= full_re.min() and full_re.max() < 1.0


def test_reference_arm_json_contract():
  """`bench.py --impl reference` prints ONE JSON line with the keys the driver reads (C1: the
  reference's own CPU-runnable configuration; a bounded sample)."""
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                        '--config', 'C1', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
  assert out.returncode == 0, out.stderr[-2000:]
  lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
  assert len(lines) == 1
  d = json.loads(lines[0])
  for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step',
              'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'e2e',
              'cpu_baseline', 'gpu_launches'):
    assert key in d, key
  assert d['impl'] == 'reference' and d['cpu_baseline']['kind'] == 'port'
  assert d['value'] > 0 and d['e2e']['h2d_bytes_per_step'] == 0
  assert d['metric'] == bench.METRIC and d['unit'] == bench.UNIT
