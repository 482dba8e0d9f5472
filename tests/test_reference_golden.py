"""The oracle and the host modules against vectors produced by the REFERENCE'S OWN SOURCE.

tests/golden/reference_*.npz were written by tests/golden/make_reference_golden.py, which executes
the files under /root/reference/jrystal verbatim over a numpy stand-in for jax
(tests/golden/numpy_jax_standin.py).  These tests never read /root/reference: they rebuild the
seeded inputs from the recipe and compare

  CPU: oracle/reference_port.py (the checker of every GPU parity test) and the numpy host
       modules jrystal_b200/{grid,occupation,ewald}.py  -> pins the oracle to the reference;
  CPU: the oracle-made fixtures tests/golden/<case>.npz the GPU tests consume;
  GPU: the CUDA path through the C ABI, directly against the reference's numbers.

Not covered by reference vectors: XC values (jax_xc absent) and gradients (no autodiff in the
stand-in); those stay pinned by known values / identities / finite differences (test_oracle.py).
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

SAMPLE_STRIDE = (2, 3, 5)  # make_reference_golden.py:G


def _ref(key):
  return np.load(os.path.join(HERE, 'golden', f'reference_{key}.npz'))


def _grid_sample(x):
  x = np.asarray(x)
  if x.shape[-3] * x.shape[-2] * x.shape[-1] <= 8192:
    return x
  return x[..., ::SAMPLE_STRIDE[0], ::SAMPLE_STRIDE[1], ::SAMPLE_STRIDE[2]]


def _close(a, b, tol):
  a, b = float(a), float(b)
  return abs(a - b) <= tol * max(abs(b), 1e-300)


@pytest.mark.parametrize('key', list(mg.CASES))
def test_oracle_matches_reference_source(key):
  c, g = mg.CASES[key], _ref(key)
  s, w_re, w_im, occ = mg.inputs(c)
  # geometry / index space: exact
  assert float(g['vol']) == pytest.approx(s.vol, rel=1e-15)
  np.testing.assert_array_equal(g['grid'], s.grid_sizes)
  np.testing.assert_array_equal(g['mask'], s.mask)
  np.testing.assert_allclose(g['kpts'], s.kpts, rtol=0, atol=1e-15)
  np.testing.assert_allclose(g['g_vec_corner'], s.g_vec[1, 2, 3], rtol=1e-15)
  assert _close(np.abs(s.g_vec).sum(), g['g_vec_sum'], 1e-13)
  np.testing.assert_allclose(g['r_vec_corner'], rp.r_vectors(s.cell, s.grid_sizes)[1, 2, 3],
                             rtol=1e-14, atol=1e-15)
  # the recipe regenerates the same inputs the reference run consumed
  assert float(g['w_re_sum']) == w_re.sum() and float(g['w_im_sum']) == w_im.sum()
  np.testing.assert_allclose(g['occ'], occ, rtol=1e-15)

  wr, wi, o = torch.from_numpy(w_re), torch.from_numpy(w_im), torch.from_numpy(occ)
  q = rp.unitary_matrix(wr, wi)                       # unitary_module.py:66-81 (LAPACK both sides)
  assert relerr(q.numpy(), g['q']) < 1e-12
  cg = rp.expand_coefficient(q, s.mask)
  psi = rp.wave_grid(cg, s.vol)
  assert relerr(_grid_sample(psi[0, 0, 0].numpy()), g['psi_band0']) < 1e-12
  rho = rp.density_grid(cg, s.vol, o)
  assert relerr(_grid_sample(rho.numpy()), g['density']) < 1e-12
  assert _close(rho.sum(), g['density_sum'], 1e-12)
  assert _close((rho ** 2).sum(), g['density_abs2_sum'], 1e-12)
  rho_g = rp.density_grid_reciprocal(cg, s.vol, o)
  assert relerr(_grid_sample(rho_g.numpy()), g['density_reciprocal']) < 1e-12

  assert _close(rp.energy_kinetic(s.g_vec, s.kpts, cg, o), g['e_kin'], 1e-12)
  assert _close(rp.energy_hartree(rho_g, s.g_vec, s.vol), g['e_har'], 1e-12)
  assert _close(rp.energy_hartree(rho_g, s.g_vec, s.vol, kohn_sham=True), g['e_har_kohn_sham'], 1e-12)
  assert _close(rp.energy_external(rho_g, s.positions, s.charges, s.g_vec, s.vol), g['e_ext'], 1e-12)
  v_har = rp.hartree_reciprocal(rho_g, s.g_vec)
  assert relerr(_grid_sample(v_har.numpy()), g['v_har_reciprocal']) < 1e-12
  assert _close(v_har.abs().sum(), g['v_har_abs_sum'], 1e-12)
  v_ext = rp.external_reciprocal(s.positions, s.charges, s.g_vec, s.vol)
  assert relerr(_grid_sample(np.asarray(v_ext)), g['v_ext_reciprocal']) < 1e-12
  assert _close(np.abs(np.asarray(v_ext)).sum(), g['v_ext_abs_sum'], 1e-12)
  t_k = rp.kinetic_operator(s.g_vec, s.kpts)
  kin = rp.expectation(cg, t_k, s.vol, diagonal=True, mode='kinetic')
  assert relerr(np.asarray(kin), g['kinetic_per_band']) < 1e-12

  # <psi_i|v|psi_j> for a seeded real potential (the contraction of band mode)
  v_r = torch.from_numpy(np.random.default_rng(c['seed'] + 3).standard_normal(tuple(s.grid_sizes)))
  assert relerr(np.asarray(rp.expectation(psi, v_r, s.vol, diagonal=True, mode='real')),
                g['expect_v_diag']) < 1e-12
  assert relerr(np.asarray(rp.expectation(psi, v_r, s.vol, diagonal=False, mode='real')),
                g['expect_v_full']) < 1e-12

  # |grad rho|^2 as the GGA branch forms it (complex absolute square on the FFT grid)
  assert relerr(_grid_sample(np.asarray(rp.sigma_r_fn(rho, s.g_vec))), g['sigma_r']) < 1e-11

  ne = s.num_electrons
  np.testing.assert_allclose(rp.occupation_uniform(s.num_k, ne, num_bands=c['nb']).numpy(),
                             g['occ_uniform'], rtol=1e-15)
  np.testing.assert_allclose(rp.occupation_gamma(s.num_k, ne, num_bands=c['nb']).numpy(),
                             g['occ_gamma'], rtol=1e-15)
  ent = rp.entropy_fermi_dirac(torch.from_numpy(g['entropy_input']))
  assert _close(ent, g['entropy_fermi_dirac'], 1e-13)


@pytest.mark.parametrize('key', list(mg.CASES))
def test_oracle_made_fixtures_agree_with_reference_source(key):
  """The fixtures the GPU parity tests consume (made by the oracle) carry the reference's own
  E_kin, E_ext, E_H and density."""
  o, g = np.load(os.path.join(HERE, 'golden', key + '.npz')), _ref(key)
  assert _close(o['energies'][0], g['e_kin'], 1e-12)
  assert _close(o['energies'][1], g['e_ext'], 1e-12)
  assert _close(o['energies'][2], g['e_har'], 1e-12)
  assert relerr(_grid_sample(o['density']), g['density']) < 1e-12


@pytest.mark.parametrize('key', list(mg.CASES))
def test_host_modules_match_reference_source(key):
  """jrystal_b200/grid.py, occupation.py, ewald.py (numpy set-up code) against the reference."""
  from jrystal_b200 import ewald, grid, occupation
  c, g = mg.CASES[key], _ref(key)
  s, _, _, _ = mg.inputs(c)
  gs = grid.proper_grid_size(c['grid'])
  np.testing.assert_array_equal(np.asarray(gs), g['grid'])
  gv = grid.g_vectors(s.cell, gs)
  np.testing.assert_allclose(gv[1, 2, 3], g['g_vec_corner'], rtol=1e-15)
  assert _close(np.abs(gv).sum(), g['g_vec_sum'], 1e-13)
  np.testing.assert_allclose(grid.r_vectors(s.cell, gs)[1, 2, 3], g['r_vec_corner'], rtol=1e-14,
                             atol=1e-15)
  np.testing.assert_allclose(grid.k_vectors(s.cell, c['kgrid']), g['kpts'], rtol=0, atol=1e-15)
  if c['mask'] == 'spherical':
    m = grid.spherical_mask(s.cell, gs, c['cutoff'])
  else:
    m = grid.cubic_mask(gs)
  np.testing.assert_array_equal(np.asarray(m), g['mask'])
  ne = s.num_electrons
  np.testing.assert_allclose(np.asarray(occupation.uniform(s.num_k, ne, num_bands=c['nb'])),
                             g['occ_uniform'], rtol=1e-15)
  np.testing.assert_allclose(np.asarray(occupation.gamma(s.num_k, ne, num_bands=c['nb'])),
                             g['occ_gamma'], rtol=1e-15)
  assert _close(occupation.fermi_dirac_entropy(g['entropy_input']), g['entropy_fermi_dirac'], 1e-13)
  # the reference truncates its Ewald sums at (eta = 0.1, cutoff = 2e4): converged to ~1e-8 Ha;
  # the host module sums to 1e-14 regardless of eta
  e_nuc = ewald.ewald_coulomb_repulsion(s.positions, s.charges, s.cell)
  assert abs(e_nuc - float(g['e_nuc'])) < 1e-6 * abs(float(g['e_nuc']))


def test_occupation_maps_match_reference_source():
  """simplex_projector / proj / idempotent (occupation.py:55-80, 281-366): oracle restatement and
  the host module the energy driver differentiates through, against the reference's outputs."""
  from jrystal_b200 import occupation
  g = np.load(os.path.join(HERE, 'golden', 'reference_occupation.npz'))
  nk, nb, ne = int(g['nk']), int(g['nb']), int(g['ne'])
  t = torch.from_numpy
  up, dn = t(g['logits_up']), t(g['logits_down'])
  for impl in ('oracle', 'host'):
    if impl == 'oracle':
      r = rp.occupation_simplex_projector(up, dn, ne)
      u = rp.occupation_simplex_projector(up, dn, ne, spin=2, spin_restricted=False)
      i_u = rp.occupation_idempotent(t(g['w_up']), t(g['w_down']), nk, spin_restricted=False)
      i_r = rp.occupation_idempotent(t(g['w_up']), t(g['w_up']), nk)
    else:
      p = {'param_up': up, 'param_down': dn}
      r = occupation.simplex_projector(p, ne)
      u = occupation.simplex_projector(p, ne, spin=2, spin_restricted=False)
      i_u = occupation.idempotent({'param_up': {'w_re': t(g['w_up'])},
                                   'param_down': {'w_re': t(g['w_down'])}}, nk, spin_restricted=False)
      i_r = occupation.idempotent({'param_up': {'w_re': t(g['w_up'])},
                                   'param_down': {'w_re': t(g['w_up'])}}, nk)
    assert relerr(r.detach().numpy(), g['simplex_restricted']) < 1e-13, impl
    assert relerr(u.detach().numpy(), g['simplex_spin2']) < 1e-13, impl
    assert relerr(i_u.detach().numpy(), g['idempotent_unrestricted']) < 1e-12, impl
    assert relerr(i_r.detach().numpy(), g['idempotent_restricted']) < 1e-12, impl
  init = occupation.simplex_projector_init(nb, nk)
  np.testing.assert_allclose(init['param_up'].detach().cpu().numpy(), g['simplex_init_up'], rtol=1e-15)
  np.testing.assert_allclose(init['param_down'].detach().cpu().numpy(), g['simplex_init_down'], rtol=1e-15)
  cpu_init = {k: v.detach().cpu() for k, v in init.items()}
  assert relerr(occupation.simplex_projector(cpu_init, ne).numpy(), g['simplex_from_init']) < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize('key', list(mg.CASES))
def test_cuda_matches_reference_source(cuda_device, key):
  """The CUDA path against numbers the reference's own code produced (no oracle in between):
  Q (up to the column sign gauge), density, E_kin, E_H, E_ext, V_H(G), per-band kinetic energy."""
  c, g = mg.CASES[key], _ref(key)
  s, w_re, w_im, occ = mg.inputs(c)
  plan = make_plan(s, c['nb'])
  occ_d = to_dev(occ)
  qd, _ = plan.qr_fwd(to_dev(w_re), to_dev(w_im))
  q = qd.cpu().numpy()
  # Cholesky-QR fixes diag(R) > 0, Householder does not: compare up to a sign per column
  sign = np.sign(np.real(np.sum(np.conj(g['q']) * q, axis=-2, keepdims=True)))
  assert relerr(q * sign, g['q']) < 1e-9
  rho, e_kin = plan.eval_begin(to_dev(w_re), to_dev(w_im), occ_d)
  en, _, _, _ = plan.eval_finish(occ_d, rho, e_kin, c['xc'])
  torch.cuda.synchronize()
  en = en.cpu().numpy()
  assert _close(en[0], g['e_kin'], 1e-10)
  assert _close(en[1], g['e_ext'], 1e-10)
  assert _close(en[2], g['e_har'], 1e-10)
  rho = rho.cpu().numpy()
  assert relerr(_grid_sample(rho), g['density']) < 1e-8
  assert _close(rho.sum(), g['density_sum'], 1e-10)
  eps_kin = plan.kinetic(qd).cpu().numpy()        # <c|T|c> per orbital (braket.py:167-207)
  assert relerr(eps_kin, g['kinetic_per_band']) < 1e-10
  # band mode: <q_i|T + v|q_j> with the seeded real potential of the reference run
  v_r = np.random.default_rng(c['seed'] + 3).standard_normal(tuple(s.grid_sizes))
  hq = plan.hpsi(qd, to_dev(v_r[None]))
  eps = plan.band_expect(qd, hq).cpu().numpy()
  assert relerr(eps, g['kinetic_per_band'] + np.real(g['expect_v_diag'])) < 1e-10
  h = plan.overlap(qd, hq).cpu().numpy()
  t_k = np.asarray(rp.kinetic_operator(s.g_vec, s.kpts))[:, s.mask]          # (nk, ng)
  t_full = np.einsum('skgi,kg,skgj->skij', np.conj(q), t_k, q)
  sg = sign[:, :, 0, :]                                                       # (ns, nk, nb)
  v_full = g['expect_v_full'] * sg[..., :, None] * sg[..., None, :]
  assert relerr(h, t_full + v_full) < 1e-10
