"""Golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle).

CPU: the oracle still reproduces its committed vectors (guards the checker against drift).
GPU: the CUDA path through the C ABI reproduces them (1e-10 energy, 1e-8 gradients / density).
"""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import reference_port as rp
from tests.common import make_plan, relerr, to_dev

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location('make_golden',
                                               os.path.join(HERE, 'golden', 'make_golden.py'))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)


def _load(key):
  return np.load(os.path.join(HERE, 'golden', key + '.npz'))


@pytest.mark.parametrize('key', list(mg.CASES))
def test_oracle_reproduces_golden(key):
  c = mg.CASES[key]
  g = _load(key)
  s, w_re, w_im, occ = mg.inputs(c)
  assert int(g['num_g']) == s.num_g and float(g['vol']) == s.vol
  assert float(g['w_re_sum']) == w_re.sum() and float(g['w_im_sum']) == w_im.sum()
  np.testing.assert_array_equal(g['occ'], occ)
  ref = rp.energy_and_grad(s, w_re, w_im, occ, xc=c['xc'], occ_grad=True)
  e = np.array([ref['e_kin'], ref['e_ext'], ref['e_har'], ref['e_xc']])
  # thread-count dependent summation order in BLAS/FFT: allow round-off only
  np.testing.assert_allclose(e, g['energies'], rtol=1e-12)
  for k in ('density', 'g_re', 'g_im', 'g_occ'):
    assert relerr(ref[k], g[k]) < 1e-11, k


@pytest.mark.gpu
@pytest.mark.parametrize('key', list(mg.CASES))
def test_cuda_matches_golden(cuda_device, key):
  c = mg.CASES[key]
  g = _load(key)
  s, w_re, w_im, occ = mg.inputs(c)
  plan = make_plan(s, c['nb'])
  occ_d = to_dev(occ)
  rho, e_kin = plan.eval_begin(to_dev(w_re), to_dev(w_im), occ_d)
  en, g_re, g_im, g_occ = plan.eval_finish(occ_d, rho, e_kin, c['xc'], want_occ_grad=True)
  torch.cuda.synchronize()
  en = en.cpu().numpy()
  for i in range(4):
    assert abs(en[i] - g['energies'][i]) / abs(g['energies'][i]) < 1e-10, i
  assert abs(en.sum() - g['energies'].sum()) / abs(g['energies'].sum()) < 1e-10
  assert relerr(rho.cpu().numpy(), g['density']) < 1e-8
  assert relerr(g_re.cpu().numpy(), g['g_re']) < 1e-8
  assert relerr(g_im.cpu().numpy(), g['g_im']) < 1e-8
  assert relerr(g_occ.cpu().numpy(), g['g_occ']) < 1e-8
  # band mode (hamiltonian_matrix_trace value + gradient) on the golden density
  qd, r = plan.qr_fwd(to_dev(w_re), to_dev(w_im))
  _, veff = plan.grid_potential(to_dev(g['density']), c['xc'], True)
  hq = plan.hpsi(qd, veff)
  eps = plan.band_expect(qd, hq).cpu().numpy()
  assert relerr(eps, g['band_per_band']) < 1e-10
  b_re, b_im = plan.qr_bwd(qd, r, hq)
  assert relerr(b_re.cpu().numpy(), g['band_g_re']) < 1e-8
  assert relerr(b_im.cpu().numpy(), g['band_g_im']) < 1e-8
