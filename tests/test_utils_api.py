"""jrystal_b200.utils / entropy / the calc and package namespaces under the reference's names
(jrystal/utils/__init__.py, jrystal/entropy.py, jrystal/calc/__init__.py, jrystal/__init__.py)."""
import numpy as np
import pytest
import torch

from oracle import reference_port as rp
from tests.common import relerr
from tests.conftest import BACKENDS


def test_namespaces_carry_the_reference_names():
  import jrystal_b200 as jb
  for name in ('calc', 'config', 'crystal', 'energy', 'entropy', 'ewald', 'grid', 'hamiltonian',
               'occupation', 'potential', 'pseudopotential', 'pw', 'sbt', 'utils', 'Crystal', 'get_pkg_path'):
    assert hasattr(jb, name), name
  for name in ('energy_all_electrons', 'band_all_electrons', 'energy_normcons', 'band_normcons'):
    assert callable(getattr(jb.calc, name)), name
  assert jb.calc.energy_all_electrons is jb.calc.energy
  assert jb.sbt.sbt_numerical is jb.pseudopotential.beta.sbt_numerical
  import os
  assert os.path.isdir(os.path.join(jb.get_pkg_path(), 'jrystal_b200'))


def test_safe_real_and_entropy():
  from jrystal_b200 import entropy, utils
  assert utils.safe_real(np.array([1.0 + 1e-10j]))[0] == 1.0
  assert utils.safe_real(torch.tensor([2.0 + 1e-10j], dtype=torch.complex128))[0] == 2.0
  x = np.array([1.0, 2.0])
  assert utils.safe_real(x) is not None and not np.iscomplexobj(utils.safe_real(x))
  for bad in (np.array([1.0 + 1.0j]), torch.tensor([1.0 + 1.0j])):
    with pytest.raises(ValueError):
      utils.safe_real(bad)
  with pytest.raises(ValueError):
    utils.check_spin_number(8, 1)
  utils.check_spin_number(13, 1)
  assert utils.fft_factor(11) == 12 and abs(utils.volume(np.diag([1.0, 2.0, 3.0])) - 6.0) < 1e-15
  occ = np.random.default_rng(0).random((1, 3, 4)) / 3 * 2
  s_np = entropy.fermi_dirac(occ)
  t = torch.from_numpy(occ).requires_grad_(True)
  s_t = entropy.fermi_dirac(t)
  assert abs(float(s_t) - s_np) < 1e-14 * abs(s_np)
  fmax = 2.0 / 3
  want = -np.sum(occ * np.log(1e-8 + occ) + (fmax - occ) * np.log(1e-8 + fmax - occ))
  assert abs(s_np - want) < 1e-14 * abs(want)
  (g,) = torch.autograd.grad(s_t, t)                    # keeps its graph (the -T S term of the driver)
  assert g.shape == t.shape and torch.isfinite(g).all()


@pytest.mark.parametrize('backend', BACKENDS, indirect=True)
def test_utils_on_the_current_plan(backend):
  """expand_coefficient / squeeze_coefficient against the mask scatter they stand for
  (utils.py:277-281, 303-308), wave_to_density(_reciprocal) of the dense psi(r) against
  pw.density_grid(_reciprocal) and the oracle."""
  import jrystal_b200 as jb
  kind, Plan, dev = backend
  s = rp.System.from_name('diamond', [8, 9, 12], [1, 1, 2], 8.0)
  nb = 5
  p = rp.param_init(4, nb, s.num_k, s.mask)
  occ = rp.occupation_uniform(s.num_k, s.num_electrons, num_bands=nb).numpy()
  occ = occ * (1.0 + 0.2 * np.random.default_rng(1).random(occ.shape))
  plan = Plan(s.cell, s.mask, s.kpts, nb)
  plan.set_atoms(s.positions, s.charges)
  with jb.use_plan(plan):
    coeff = jb.pw.coeff({'w_re': dev(p['w_re']), 'w_im': dev(p['w_im'])}, s.mask)
    dense = jb.utils.expand_coefficient(coeff.q, s.mask)
    q = coeff.q.cpu().numpy()
    want = np.zeros((1, s.num_k, nb) + tuple(s.mask.shape), dtype=np.complex128)
    want[..., s.mask] = np.swapaxes(q, -1, -2)
    np.testing.assert_array_equal(dense.cpu().numpy(), want)
    back = jb.utils.squeeze_coefficient(dense, s.mask)
    np.testing.assert_array_equal(back.cpu().numpy(), q)
    with pytest.raises(ValueError):
      jb.utils.expand_coefficient(coeff.q, ~s.mask)
    psi = jb.pw.wave_grid(coeff, s.vol)
    rho = jb.utils.wave_to_density(psi, dev(occ))
    rho_fused = jb.pw.density_grid(coeff, s.vol, dev(occ))
    assert relerr(rho.cpu().numpy(), rho_fused.cpu().numpy()) < 1e-12
    per_orbital = jb.utils.wave_to_density(psi)
    assert tuple(per_orbital.shape) == tuple(psi.shape) and not per_orbital.is_complex()
    rho_g = jb.utils.wave_to_density_reciprocal(psi, dev(occ))
    assert relerr(rho_g.cpu().numpy(), jb.pw.density_grid_reciprocal(coeff, s.vol, dev(occ)).cpu().numpy()) < 1e-12
    with pytest.raises(ValueError):
      jb.utils.wave_to_density(psi, dev(occ)[:, :, :-1])
    # pw.py:273-284: without occupation density_grid returns the per-orbital densities
    per = jb.pw.density_grid(coeff, s.vol)
    assert tuple(per.shape) == tuple(psi.shape) and not per.is_complex()
    assert relerr(per.cpu().numpy(), np.abs(psi.cpu().numpy()) ** 2) < 1e-13
    assert relerr((per.cpu().numpy() * occ[..., None, None, None]).sum((1, 2)),
                  rho_fused.cpu().numpy()) < 1e-12
    per_g = jb.pw.density_grid_reciprocal(coeff, s.vol)
    assert relerr(per_g.cpu().numpy(), np.fft.fftn(per.cpu().numpy(), axes=(-3, -2, -1))) < 1e-12
  ref = rp.density_grid(rp.expand_coefficient(torch.from_numpy(q), s.mask), s.vol, torch.from_numpy(occ))
  assert relerr(rho.cpu().numpy(), ref.numpy()) < 1e-10


@pytest.mark.parametrize('backend', BACKENDS, indirect=True)
def test_point_evaluations_equal_the_fft_grids(backend):
  """The reference's own T1 property (pw_test.py:36-49: wave_grid via FFT == the direct plane-wave
  sum wave_r at the grid points, atol 1e-8) through the host API, plus density_r == density_grid at
  the same points and nabla_density_r / nabla_density_grid against central differences."""
  import jrystal_b200 as jb
  kind, Plan, dev = backend
  s = rp.System.from_name('diamond', [7, 8, 9], [1, 1, 1], None, mask_method='cubic')   # pw_test.py:33-34
  nb = 4
  p = rp.param_init(2, nb, s.num_k, s.mask)
  occ = np.random.default_rng(3).random((1, s.num_k, nb))
  plan = Plan(s.cell, s.mask, s.kpts, nb)
  plan.set_atoms(s.positions, s.charges)
  r_grid = jb.grid.r_vectors(s.cell, [7, 8, 9])
  with jb.use_plan(plan):
    coeff = jb.pw.coeff({'w_re': dev(p['w_re']), 'w_im': dev(p['w_im'])}, s.mask)
    psi = jb.pw.wave_grid(coeff, s.vol).cpu().numpy()
    rho = jb.pw.density_grid(coeff, s.vol, dev(occ)).cpu().numpy()
    for idx in [(0, 0, 0), (3, 5, 2), (6, 7, 8)]:
      r = r_grid[idx]
      w = jb.pw.wave_r(r, coeff, s.cell).cpu().numpy()
      np.testing.assert_allclose(w, psi[(slice(None),) * 3 + idx], atol=1e-8)
      d = float(jb.pw.density_r(r, coeff, s.cell, s.g_vec, dev(occ)))
      assert abs(d - rho[(0,) + idx]) < 1e-10 * abs(rho).max()
    r0 = np.array([0.31, -0.47, 1.13])
    g = jb.pw.nabla_density_r(r0, coeff, s.cell, None, dev(occ)).cpu().numpy()
    h = 1e-5
    fd = np.array([(float(jb.pw.density_r(r0 + h * e, coeff, s.cell, None, dev(occ)))
                    - float(jb.pw.density_r(r0 - h * e, coeff, s.cell, None, dev(occ)))) / (2 * h)
                   for e in np.eye(3)])
    np.testing.assert_allclose(g, fd, rtol=1e-6, atol=1e-9)
    g2 = jb.pw.nabla_density_grid(r0, coeff, s.cell, s.g_vec, dev(occ)).cpu().numpy()
    np.testing.assert_allclose(g2, g, rtol=1e-12, atol=1e-14)
    per = jb.pw.nabla_density_grid(r0, coeff, s.cell)
    assert tuple(per.shape) == (1, s.num_k, nb, 3)
    np.testing.assert_allclose((per.cpu().numpy() * occ[..., None]).sum((0, 1, 2)) / s.vol, g, rtol=1e-12)
    assert tuple(jb.pw.density_r(r0, coeff, s.cell).shape) == (1, s.num_k, nb)
    with pytest.raises(ValueError):
      jb.pw.wave_r(np.zeros(2), coeff, s.cell)
