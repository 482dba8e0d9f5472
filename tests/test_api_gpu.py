"""The reference's own unit tests, re-stated on the reference-named API of jrystal_b200
(pw / energy / potential / hamiltonian over the CUDA kernels), same fixture: diamond, grid
[7, 8, 9], cubic mask (jrystal/_src/pw_test.py:25-34, energy_test.py:30-49,
hamiltonian_test.py:33-51), plus parity with the oracle."""
import numpy as np
import pytest
import torch

import jrystal_b200 as jb
from oracle import reference_port as rp
from tests.common import relerr

pytestmark = pytest.mark.gpu


@pytest.fixture()
def fixture_789(cuda_device):
  s = rp.System.from_name('diamond', [7, 8, 9], [2, 2, 1], mask_method='cubic')
  nb = s.num_electrons
  plan = jb.Plan(s.cell, s.mask, s.kpts, nb)
  plan.set_atoms(s.positions, s.charges)
  with jb.use_plan(plan):
    params = jb.pw.param_init(123, nb, s.num_k, s.mask)
    yield s, nb, plan, params


def test_param_init_and_coeff(fixture_789):
  s, nb, plan, params = fixture_789
  assert tuple(params['w_re'].shape) == (1, s.num_k, s.num_g, nb)
  assert params['w_re'].dtype == torch.float64 and params['w_re'].is_cuda
  c = jb.pw.coeff(params, s.mask)
  assert c.shape == (1, s.num_k, nb) + tuple(s.grid_sizes)
  dense = c.dense().cpu().numpy()
  # orthonormal columns: sum_G conj(c_i) c_j = delta_ij  (unitary_module.py:66-81)
  flat = dense.reshape(1, s.num_k, nb, -1)
  gram = np.einsum('skig,skjg->skij', flat.conj(), flat)
  assert np.abs(gram - np.eye(nb)).max() < 1e-13
  assert np.abs(dense[..., ~s.mask]).max() == 0.0
  with pytest.raises(ValueError):
    jb.pw.coeff(params, np.ones_like(s.mask))


def test_T1_wave_grid_equals_plane_wave_sum(fixture_789):
  """pw_test.py:36-49."""
  s, nb, plan, params = fixture_789
  c = jb.pw.coeff(params, s.mask)
  wg = jb.pw.wave_grid(c, s.vol).cpu().numpy()
  dense = torch.from_numpy(c.dense().cpu().numpy())
  r_vec = jb.grid.r_vectors(s.cell, s.grid_sizes)
  for idx in [(0, 0, 0), (1, 2, 3), (6, 7, 8), (3, 0, 5)]:
    direct = rp.wave_r(r_vec[idx], dense, s.cell, s.g_vec).numpy()
    np.testing.assert_allclose(wg[(slice(None),) * 3 + idx], direct, atol=1e-8)


def test_T2_potential_brakets_equal_energies(fixture_789):
  """energy_test.py:61-112: sum_i f_i <psi_i|v_X|psi_i> == energy.X."""
  s, nb, plan, params = fixture_789
  c = jb.pw.coeff(params, s.mask)
  occ = torch.from_numpy(rp.occupation_gamma(s.num_k, s.num_electrons, num_bands=nb).numpy()).cuda()
  rho = jb.pw.density_grid(c, s.vol, occ)
  rho_g = jb.pw.density_grid_reciprocal(c, s.vol, occ)
  v_h, v_e, v_xc = jb.potential.effective(rho, s.positions, s.charges, s.g_vec, s.vol, split=True)
  wg = jb.pw.wave_grid(c, s.vol)
  dens_i = (wg.real**2 + wg.imag**2)
  e1 = [float(torch.einsum('skbxyz,sxyz,skb->', dens_i, v, occ)) * s.vol / np.prod(s.grid_sizes)
        for v in (v_h, v_e, v_xc)]
  e2 = [float(jb.energy.hartree(rho_g, s.g_vec, s.vol)),
        float(jb.energy.external(rho_g, s.positions, s.charges, s.g_vec, s.vol)),
        float(jb.energy.xc_energy(rho, s.g_vec, s.vol, 'lda_x'))]
  np.testing.assert_allclose(e1, e2, atol=1e-7)
  # and against the oracle
  dens_o = torch.from_numpy(rho.cpu().numpy())
  dens_g = torch.fft.fftn(dens_o, dim=(-3, -2, -1))
  assert relerr(rho_g.cpu().numpy(), dens_g.numpy()) < 1e-12
  ref = [rp.energy_hartree(dens_g, s.g_vec, s.vol).item(),
         rp.energy_external(dens_g, s.positions, s.charges, s.g_vec, s.vol).item(),
         rp.energy_xc(dens_o, s.vol, 'lda_x').item()]
  np.testing.assert_allclose(e2, ref, rtol=1e-10)
  v_ref = rp.effective(dens_o, s.positions, s.charges, s.g_vec, s.vol, split=True, kohn_sham=False)
  for v, r in zip((v_h, v_e, v_xc), v_ref):
    assert relerr(v.cpu().numpy(), np.broadcast_to(r.real.numpy(), v.shape)) < 1e-11


def test_T3_hamiltonian_trace_and_matrix(fixture_789):
  """hamiltonian_test.py:53-102: trace(H) == total_energy(kohn_sham=True) with f = 1; the
  analytic H_ij equals the oracle's explicit matrix."""
  s, nb, plan, params = fixture_789
  c = jb.pw.coeff(params, s.mask)
  occ = torch.ones((1, s.num_k, nb), dtype=torch.float64, device='cuda')
  rho = jb.pw.density_grid(c, s.vol, occ)
  e1 = float(jb.hamiltonian.hamiltonian_matrix_trace(c, s.positions, s.charges, rho, s.g_vec,
                                                     s.kpts, s.vol, kohn_sham=True))
  e2 = float(jb.energy.total_energy(c, s.positions, s.charges, s.g_vec, s.kpts, s.vol,
                                    kohn_sham=True))
  np.testing.assert_allclose(e1, e2, atol=1e-7)
  h = jb.hamiltonian.hamiltonian_matrix(c, s.positions, s.charges, rho, s.g_vec, s.kpts, s.vol)
  h = h.cpu().numpy()
  assert np.abs(h - np.conj(np.swapaxes(h, -1, -2))).max() < 1e-10
  c_o = torch.from_numpy(c.dense().cpu().numpy())
  h_ref = rp.hamiltonian_matrix_explicit(c_o, s.positions, s.charges,
                                         torch.from_numpy(rho.cpu().numpy()), s.g_vec, s.kpts,
                                         s.vol).numpy()
  assert relerr(h, h_ref) < 1e-10
  per_k = jb.hamiltonian.hamiltonian_matrix_trace(c, s.positions, s.charges, rho, s.g_vec, s.kpts,
                                                  s.vol, keep_kpts_axis=True, keep_spin_axis=True)
  assert tuple(per_k.shape) == (1, s.num_k)
  np.testing.assert_allclose(per_k.cpu().numpy(), np.einsum('skii->sk', h).real, rtol=1e-10)


def test_total_energy_split_matches_oracle(fixture_789):
  s, nb, plan, params = fixture_789
  occ_np = rp.occupation_uniform(s.num_k, s.num_electrons, num_bands=nb).numpy()
  c = jb.pw.coeff(params, s.mask)
  parts = jb.energy.total_energy(c, s.positions, s.charges, s.g_vec, s.kpts, s.vol,
                                 torch.from_numpy(occ_np).cuda(), split=True)
  ref = rp.energy_and_grad(s, params['w_re'].cpu().numpy(), params['w_im'].cpu().numpy(), occ_np)
  for got, key in zip(parts, ['e_kin', 'e_ext', 'e_har', 'e_xc']):
    assert abs(float(got) - ref[key]) / abs(ref[key]) < 1e-10, key
  # kinetic per band without occupation (energy.py:179-180)
  t = jb.energy.kinetic(s.g_vec, s.kpts, c)
  assert tuple(t.shape) == (1, s.num_k, nb)
  with pytest.raises(ValueError):
    jb.pw.density_grid(c, s.vol, torch.ones((1, 1, nb), device='cuda', dtype=torch.float64))
