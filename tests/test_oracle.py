"""CPU tests that PIN THE ORACLE before it is trusted as the checker.

The reference cannot be imported here (no jax), so the oracle is pinned by
  T5  the literal goldens of jrystal/_src/crystal_test.py:25-61,
  T1  pw_test.py:36-49   (FFT wave_grid == direct plane-wave sum, atol 1e-8),
  T2  energy_test.py:61-112 (sum_i f_i <psi_i|v_X|psi_i> == energy.X, atol 1e-7),
  T3  hamiltonian_test.py:53-102 (trace(H) == total_energy(kohn_sham=True), f = 1;
      explicit H matrix Hermitian with the same trace),
  plus mask / k-grid facts stated in the reference's docstrings and SURVEY App. A,
  an independent hand-derived route (oracle/analytic.py) and finite differences.
Fixtures mirror the reference tests: diamond, grid [7, 8, 9], cubic mask.
"""
import numpy as np
import pytest
import torch

from oracle import analytic, structures
from oracle import reference_port as rp


@pytest.fixture(scope='module')
def fixture_789():
  s = rp.System.from_name('diamond', [7, 8, 9], [1, 1, 1], mask_method='cubic')
  nb = s.num_electrons  # pw_test.py:30
  p = rp.param_init(123, nb, s.num_k, s.mask)
  c = rp.coeff(torch.from_numpy(p['w_re']), torch.from_numpy(p['w_im']), s.mask)
  occ = rp.occupation_gamma(s.num_k, s.num_electrons)
  return s, nb, p, c, occ


def test_crystal_goldens():
  """T5: jrystal/_src/crystal_test.py:25-61."""
  cell, pos, chg = structures.load('diamond')
  np.testing.assert_almost_equal(
    pos, [[-0.84251071, -0.84251071, -0.84251071], [0.84251071, 0.84251071, 0.84251071]])
  np.testing.assert_almost_equal(chg, [6, 6])
  np.testing.assert_almost_equal(
    cell, [[0., 3.37004284, 3.37004284], [3.37004284, 0., 3.37004284],
           [3.37004284, 3.37004284, 0.]])
  assert abs(rp.volume(cell) - 76.5484253352856) < 1e-12
  np.testing.assert_almost_equal(pos @ np.linalg.inv(cell).T,
                                 [[-0.125] * 3, [0.125] * 3])
  np.testing.assert_almost_equal(
    rp.reciprocal_vectors(cell),
    [[-0.93221149, 0.93221149, 0.93221149], [0.93221149, -0.93221149, 0.93221149],
     [0.93221149, 0.93221149, -0.93221149]])


def test_grid_facts():
  # cubic_mask([7,8,9]) keeps 4*4*5 points (grid.py:315-325)
  assert int(rp.cubic_mask([7, 8, 9]).sum()) == 80
  # 7-smooth sizing (utils.py:227-254, const.py:265)
  assert list(rp.proper_grid_size([11, 13, 17])) == [12, 14, 18]
  assert list(rp.proper_grid_size(47)) == [48, 48, 48]
  with pytest.raises(ValueError):
    rp.fft_factor(2049)
  # g_vectors uses integer fftfreq ordering
  cell, _, _ = structures.load('diamond')
  g = rp.g_vectors(cell, [4, 4, 4])
  b = rp.reciprocal_vectors(cell)
  np.testing.assert_allclose(g[1, 0, 3], b[0] - b[2], atol=1e-14)
  np.testing.assert_allclose(g[2, 0, 0], -2 * b[0], atol=1e-14)
  # Monkhorst-Pack: (i + 1/2)/n - 1/2
  np.testing.assert_allclose(rp.monkhorst_pack([2, 1, 1]), [[-0.25, 0, 0], [0.25, 0, 0]])
  # BASELINE config sphere sizes computed in the survey from the shipped cells
  for name, grid, ecut, ng in [('si', 32, 20, 1139), ('si8', 64, 30, 8409)]:
    c, _, _ = structures.load(name)
    assert int(rp.spherical_mask(c, [grid] * 3, ecut).sum()) == ng


def test_T1_wave_grid_equals_plane_wave_sum(fixture_789):
  s, nb, p, c, occ = fixture_789
  wg = rp.wave_grid(c, s.vol).numpy()
  r_vec = rp.r_vectors(s.cell, s.grid_sizes)
  for idx in [(0, 0, 0), (1, 2, 3), (6, 7, 8), (3, 0, 5), (4, 4, 4)]:
    direct = rp.wave_r(r_vec[idx], c, s.cell, s.g_vec).numpy()
    np.testing.assert_allclose(wg[(slice(None),) * 3 + idx], direct, atol=1e-8)


def test_T2_potential_brakets_equal_energies(fixture_789):
  s, nb, p, c, occ = fixture_789
  wg = rp.wave_grid(c, s.vol)
  dens = rp.density_grid(c, s.vol, occ)
  dens_g = torch.fft.fftn(dens, dim=(-3, -2, -1))
  v_h, v_e, v_xc = rp.effective(dens, s.positions, s.charges, s.g_vec, s.vol, split=True,
                                kohn_sham=False)
  e1 = [torch.sum(rp.expectation(wg, v, s.vol, diagonal=True, mode='real') * occ).real.item()
        for v in (v_h, v_e, v_xc)]
  e2 = [rp.energy_hartree(dens_g, s.g_vec, s.vol, False).item(),
        rp.energy_external(dens_g, s.positions, s.charges, s.g_vec, s.vol).item(),
        rp.energy_xc(dens, s.vol, 'lda_x', False).item()]
  np.testing.assert_allclose(e1, e2, atol=1e-7)
  # electron count: sum_r rho Omega/N = N_e  
  assert abs(dens.sum().item() * s.vol / np.prod(s.grid_sizes) - s.num_electrons) < 1e-10


def test_T3_hamiltonian_trace_equals_ks_energy():
  s = rp.System.from_name('diamond', [7, 8, 9], [2, 2, 1], mask_method='cubic')
  nb = s.num_electrons
  p = rp.param_init(123, nb, s.num_k, s.mask)
  c = rp.coeff(torch.from_numpy(p['w_re']), torch.from_numpy(p['w_im']), s.mask)
  occ = torch.ones((1, s.num_k, nb), dtype=torch.float64)
  dens = rp.density_grid(c, s.vol, occ)
  e1 = rp.hamiltonian_matrix_trace(c, s.positions, s.charges, dens, s.g_vec, s.kpts, s.vol,
                                   kohn_sham=True).item()
  e2 = rp.total_energy(c, s.positions, s.charges, s.g_vec, s.kpts, s.vol, kohn_sham=True).item()
  np.testing.assert_allclose(e1, e2, atol=1e-7)
  h = rp.hamiltonian_matrix_explicit(c, s.positions, s.charges, dens, s.g_vec, s.kpts, s.vol)
  assert torch.abs(h - h.conj().transpose(-1, -2)).max().item() < 1e-10
  np.testing.assert_allclose(torch.einsum('skii->', h).real.item(), e1, atol=1e-8)


def test_lda_x_known_values():
  # eps_x(rho) = -3/4 (3/pi)^(1/3) rho^(1/3); rho = pi/3 -> -3/4 exactly
  rho = torch.tensor([[np.pi / 3, 1.0, 0.0, 1e-20]], dtype=torch.float64)
  e = rp.xc_density(rho, False, 'lda_x').numpy()
  np.testing.assert_allclose(e[:2], [-0.75, -0.7385587663820224], rtol=1e-14)
  assert e[2] == 0.0 and e[3] == 0.0
  v = rp.xc_density(rho, True, 'lda_x').numpy()[0]
  np.testing.assert_allclose(v[:2], 4.0 / 3.0 * e[:2], rtol=1e-13)


def test_analytic_route_matches_autograd():
  s = rp.System.from_name('diamond', [7, 8, 9], [2, 2, 1], mask_method='cubic')
  nb = 12
  p = rp.param_init(123, nb, s.num_k, s.mask)
  occ = rp.occupation_uniform(s.num_k, 12, num_bands=nb).numpy()
  occ = occ * (1 + 0.1 * np.random.default_rng(0).random(occ.shape))
  a = rp.energy_and_grad(s, p['w_re'], p['w_im'], occ, occ_grad=True)
  b = analytic.energy_and_grad(s, p['w_re'], p['w_im'], occ)
  for k in ['e_kin', 'e_ext', 'e_har', 'e_xc', 'e_tot']:
    assert abs(a[k] - b[k]) / abs(a[k]) < 1e-13, k
  for k in ['density', 'g_re', 'g_im', 'g_occ']:
    assert np.abs(a[k] - b[k]).max() / np.abs(a[k]).max() < 1e-12, k


def test_gradient_finite_difference():
  s = rp.System.from_name('diamond', [7, 8, 9], [1, 1, 1], mask_method='cubic')
  nb = 8
  p = rp.param_init(7, nb, s.num_k, s.mask)
  occ = rp.occupation_uniform(s.num_k, 12, num_bands=nb).numpy()
  base = rp.energy_and_grad(s, p['w_re'], p['w_im'], occ)
  rng = np.random.default_rng(3)
  for key, gkey in (('w_re', 'g_re'), ('w_im', 'g_im')):
    d = rng.standard_normal(p[key].shape)
    h = 1e-5
    ep = dict(p); em = dict(p)
    ep[key] = p[key] + h * d
    em[key] = p[key] - h * d
    fp = rp.energy_and_grad(s, ep['w_re'], ep['w_im'], occ)['e_tot']
    fm = rp.energy_and_grad(s, em['w_re'], em['w_im'], occ)['e_tot']
    fd = (fp - fm) / (2 * h)
    an = float(np.sum(base[gkey] * d))
    assert abs(fd - an) / abs(an) < 1e-7


def test_occupation_and_entropy():
  o = rp.occupation_uniform(4, 12, num_bands=10)
  assert o.shape == (1, 4, 10) and abs(o.sum().item() - 12) < 1e-14
  assert (o[0, :, 6:] == 0).all() and np.allclose(o[0, :, :6], 0.5)
  g = rp.occupation_gamma(4, 12, num_bands=10)
  assert abs(g.sum().item() - 12) < 1e-14 and (g[0, 1:] == 0).all()
  # entropy.py:46-54 on a 1-k, 2-band toy case: f = (2, 1), fmax = 2, eps = 1e-8
  f = torch.tensor([[[2.0, 1.0]]], dtype=torch.float64)
  expect = -(2 * np.log(2 + 1e-8) + 0 * np.log(1e-8) + 1 * np.log(1 + 1e-8) + 1 * np.log(1 + 1e-8))
  assert abs(rp.entropy_fermi_dirac(f).item() - expect) < 1e-12


def test_pbe_known_values():
  """Pins of the GGA restatement (the arithmetic lives in jax_xc / LibXC, not vendored): the
  published enhancement factor F_x(s) = 1 + kappa - kappa / (1 + mu s^2 / kappa) (Perdew, Burke,
  Ernzerhof 1996, eq. 14), the uniform-gas limits (sigma = 0: PBE x -> LDA x, PBE c -> PW92), and
  the PW92 correlation energy at r_s = 2 (-0.0448 Ha, Perdew-Wang 1992 table I)."""
  rho = torch.tensor([0.3, 1e-3, 2.5], dtype=torch.float64)
  zero = torch.zeros_like(rho)
  assert torch.allclose(rp._eps_gga_x_pbe(rho, zero), rp._eps_lda_x(rho), rtol=1e-14)
  ec0 = rp._eps_gga_c_pbe(rho, zero)
  pw = rp._eps_lda_c_pw(rho)  # PW92 with the 1992 parameter digits: differs from pw_mod by ~1e-6
  assert torch.allclose(ec0, pw, rtol=2e-5)
  kf = (3 * np.pi**2 * rho)**(1 / 3)
  for s_ in (0.5, 1.0, 2.0):
    sigma = (2 * kf * rho * s_)**2
    fx = rp._eps_gga_x_pbe(rho, sigma) / rp._eps_lda_x(rho)
    want = 1 + 0.804 - 0.804 / (1 + 0.2195149727645171 * s_**2 / 0.804)
    assert torch.allclose(fx, torch.full_like(fx, want), rtol=1e-13)
  rs2 = torch.tensor([3.0 / (4 * np.pi * 2.0**3)], dtype=torch.float64)
  assert abs(rp._eps_gga_c_pbe(rs2, torch.zeros(1, dtype=torch.float64)).item() + 0.0448) < 1e-4
  # gradient correction H >= 0 and -> -eps_c^PW for t -> infinity (correlation vanishes)
  big = rp._eps_gga_c_pbe(rho, torch.full_like(rho, 1e12))
  assert torch.all(big.abs() < 1e-3 * ec0.abs())


def test_gga_energy_gradient_finite_difference():
  s = rp.System.from_name('diamond', 9, [1, 1, 1], 8.0)
  p = rp.param_init(3, 4, s.num_k, s.mask)
  occ = rp.occupation_uniform(s.num_k, s.num_electrons, num_bands=4).numpy()
  xc = 'gga_x_pbe+gga_c_pbe'
  ref = rp.energy_and_grad(s, p['w_re'], p['w_im'], occ, xc=xc)
  d = np.random.default_rng(1).standard_normal(p['w_re'].shape)
  h = 1e-5
  ep = rp.energy_and_grad(s, p['w_re'] + h * d, p['w_im'], occ, xc=xc)['e_tot']
  em = rp.energy_and_grad(s, p['w_re'] - h * d, p['w_im'], occ, xc=xc)['e_tot']
  assert abs((ep - em) / (2 * h) - (ref['g_re'] * d).sum()) < 1e-6 * abs((ref['g_re'] * d).sum())


def test_blocked_evaluation_equals_the_one_graph_evaluation():
  """reference_port.energy_and_grad_blocked (k-blocks / band chunks with bounded memory; the
  checker of the BASELINE-sized GPU tests and the full-workload CPU arm of bench.py) against
  energy_and_grad, which runs the reference's loss through ONE autograd graph."""
  s = rp.System.from_name('diamond', 16, [1, 2, 2], 20.0)
  nb = 10
  p = rp.param_init(123, nb, s.num_k, s.mask)
  occ = rp.occupation_uniform(s.num_k, s.num_electrons, num_bands=nb).numpy()
  occ = occ * (1 + 0.3 * np.random.default_rng(0).random(occ.shape))
  for xc in ('lda_x', 'lda_x+lda_c_pw'):
    a = rp.energy_and_grad(s, p['w_re'], p['w_im'], occ, xc=xc)
    for b in (rp.energy_and_grad_chunked(s, p['w_re'], p['w_im'], occ, xc, band_chunk=3),
              rp.energy_and_grad_kblocks(s, p['w_re'], p['w_im'], occ, xc, kblock=3),
              rp.energy_and_grad_blocked(s, p['w_re'], p['w_im'], occ, xc, 2, 4)):
      for k in ('e_kin', 'e_ext', 'e_har', 'e_xc', 'e_tot'):
        assert abs(a[k] - b[k]) < 1e-13 * abs(a['e_tot']), k
      for k in ('density', 'g_re', 'g_im'):
        assert np.abs(a[k] - b[k]).max() < 1e-13 * np.abs(a[k]).max(), k
