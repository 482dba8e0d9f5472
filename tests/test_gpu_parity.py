"""GPU parity: every C-ABI entry point against the oracle on seeded inputs.

Tolerances (BASELINE.json north_star): total energy 1e-10 relative, gradients and density
1e-8 relative, FP64.
"""
import numpy as np
import pytest
import torch

from oracle import analytic
from oracle import reference_port as rp
from tests.common import make_inputs, make_plan, make_system, relerr, to_dev

pytestmark = pytest.mark.gpu

E_TOL = 1e-10
G_TOL = 1e-8

CASES = {
  # 64^3 (the C2 benchmark grid): band-limited sphere -> sparse radix-8 butterflies in the fused
  # kernels; the high cut-off case is NOT band-limited and takes the dense butterflies
  'si8_64': dict(name='si8', grid=64, kgrid=[1, 1, 1], mask='spherical', cutoff=30, nb=11),
  'si8_64_hicut': dict(name='si8', grid=64, kgrid=[1, 1, 1], mask='spherical', cutoff=75, nb=9),
  # cubic grids with radix-3 lengths and several band groups (fused y+x kernels, chunked density)
  'si_48_cubic': dict(name='si', grid=48, kgrid=[1, 1, 2], mask='spherical', cutoff=15, nb=21),
  'diamond_12': dict(name='diamond', grid=12, kgrid=[2, 1, 1], mask='spherical', cutoff=10, nb=5),
  # the reference's own test fixture: diamond, grid [7,8,9], cubic mask (pw_test.py:33-34)
  'diamond_789_cubic': dict(name='diamond', grid=[7, 8, 9], kgrid=[2, 2, 1], mask='cubic',
                            cutoff=None, nb=12),
  'diamond_16': dict(name='diamond', grid=16, kgrid=[1, 1, 1], mask='spherical', cutoff=20, nb=9),
  'si_32': dict(name='si', grid=32, kgrid=[2, 1, 1], mask='spherical', cutoff=12, nb=18),
  # 128^3 (the C3 benchmark grid): table twiddles, two resident CTAs per SM, unfused passes
  'si8_128': dict(name='si8', grid=128, kgrid=[1, 1, 1], mask='spherical', cutoff=40, nb=6),
  # more than 96 bands: multi-CTA panel Cholesky + blocked triangular inverse (C3-style QR)
  'si8_32_nb130': dict(name='si8', grid=32, kgrid=[1, 1, 1], mask='spherical', cutoff=8, nb=130),
  'diamond_24x32x48': dict(name='diamond', grid=[24, 32, 48], kgrid=[1, 1, 2], mask='spherical',
                           cutoff=30, nb=10),
}


def _setup(case, **plan_kw):
  c = CASES[case]
  s = make_system(c['name'], c['grid'], c['kgrid'], c['cutoff'], c['mask'])
  w_re, w_im, occ = make_inputs(s, c['nb'], jitter=0.1)
  plan = make_plan(s, c['nb'], **plan_kw)
  return s, plan, w_re, w_im, occ


@pytest.mark.parametrize('case', list(CASES))
def test_fft3d_dense(cuda_device, case):
  s, plan, *_ = _setup(case)
  rng = np.random.default_rng(5)
  x = rng.standard_normal((3,) + tuple(s.grid_sizes)) + 1j * rng.standard_normal((3,) + tuple(s.grid_sizes))
  xd = to_dev(x)
  f = plan.fft3d(xd, inverse=False).cpu().numpy()
  i = plan.fft3d(xd, inverse=True).cpu().numpy()
  assert relerr(f, np.fft.fftn(x, axes=(-3, -2, -1))) < 1e-13
  assert relerr(i, np.fft.ifftn(x, axes=(-3, -2, -1))) < 1e-13
  # in place
  y = xd.clone()
  plan.fft3d(y, inverse=False, out=y)
  assert relerr(y.cpu().numpy(), f) == 0.0


@pytest.mark.parametrize('case', list(CASES))
def test_qr(cuda_device, case):
  s, plan, w_re, w_im, occ = _setup(case)
  q, r = plan.qr_fwd(to_dev(w_re), to_dev(w_im))
  q = q.cpu().numpy()
  r = r.cpu().numpy()
  w = w_re + 1j * w_im
  nb = w.shape[-1]
  for k in range(w.shape[1]):
    assert np.abs(q[0, k].conj().T @ q[0, k] - np.eye(nb)).max() < 1e-13
    assert relerr(q[0, k] @ r[0, k], w[0, k]) < 1e-13
    assert np.abs(np.tril(r[0, k], -1)).max() == 0.0
    assert (np.diag(r[0, k]).real > 0).all() and np.abs(np.diag(r[0, k]).imag).max() < 1e-14
    # reference (Householder) Q up to the documented per-column sign
    q_ref = np.linalg.qr(w[0, k])[0]
    d = np.sign(np.real(np.sum(q_ref.conj() * q[0, k], axis=0)))
    assert relerr(q[0, k] * d[None, :], q_ref) < 1e-12


@pytest.mark.parametrize('case', list(CASES))
def test_qr_bwd(cuda_device, case):
  s, plan, w_re, w_im, occ = _setup(case)
  rng = np.random.default_rng(11)
  gq = rng.standard_normal(w_re.shape) + 1j * rng.standard_normal(w_re.shape)
  q, r = plan.qr_fwd(to_dev(w_re), to_dev(w_im))
  g_re, g_im = plan.qr_bwd(q, r, to_dev(gq))
  for k in range(w_re.shape[1]):
    gw = analytic.qr_backward(q[0, k].cpu().numpy(), r[0, k].cpu().numpy(), gq[0, k])
    assert relerr(g_re[0, k].cpu().numpy(), 2 * gw.real) < 1e-11
    assert relerr(g_im[0, k].cpu().numpy(), 2 * gw.imag) < 1e-11


@pytest.mark.parametrize('case', list(CASES))
def test_density_kinetic(cuda_device, case):
  s, plan, w_re, w_im, occ = _setup(case)
  q = rp.unitary_matrix(torch.from_numpy(w_re), torch.from_numpy(w_im))
  c = rp.expand_coefficient(q, s.mask)
  rho_ref = rp.density_grid(c, s.vol, torch.from_numpy(occ)).numpy()
  t_ref = rp.energy_kinetic(s.g_vec, s.kpts, c).numpy()
  qd = q.cuda().contiguous()
  rho = plan.density(qd, to_dev(occ)).cpu().numpy()
  t = plan.kinetic(qd).cpu().numpy()
  assert relerr(rho, rho_ref) < G_TOL
  assert relerr(rho, rho_ref) < 1e-12
  assert relerr(t, t_ref) < 1e-12
  # expand / squeeze round trip and parity with utils.expand_coefficient
  dense = plan.expand(qd)
  assert relerr(dense.cpu().numpy(), c.numpy()) == 0.0
  assert relerr(plan.squeeze(dense).cpu().numpy(), q.numpy()) == 0.0


@pytest.mark.parametrize('case', list(CASES))
@pytest.mark.parametrize('kohn_sham', [False, True])
@pytest.mark.parametrize('xc', ['lda_x', 'lda_x+lda_c_pw'])
def test_grid_potential(cuda_device, case, kohn_sham, xc):
  s, plan, w_re, w_im, occ = _setup(case)
  q = rp.unitary_matrix(torch.from_numpy(w_re), torch.from_numpy(w_im))
  c = rp.expand_coefficient(q, s.mask)
  rho = rp.density_grid(c, s.vol, torch.from_numpy(occ))
  rho_g = torch.fft.fftn(rho, dim=(-3, -2, -1))
  e_h = rp.energy_hartree(rho_g, s.g_vec, s.vol, kohn_sham).item()
  e_e = rp.energy_external(rho_g, s.positions, s.charges, s.g_vec, s.vol).item()
  e_x = rp.energy_xc(rho, s.vol, xc, kohn_sham).item()
  en, veff = plan.grid_potential(rho.cuda().contiguous(), xc, kohn_sham)
  en = en.cpu().numpy()
  assert abs(en[0] - e_h) / abs(e_h) < E_TOL
  assert abs(en[1] - e_e) / abs(e_e) < E_TOL
  assert abs(en[2] - e_x) / abs(e_x) < E_TOL
  v_ref = rp.effective(rho, s.positions, s.charges, s.g_vec, s.vol, False, xc, True)
  # ifftn(V_ext) is complex on even grids (Nyquist planes are not Hermitian); the reference
  # keeps only the real part of <psi|v|psi> (hamiltonian.py:164) and d E/d rho is its real part.
  assert relerr(veff.cpu().numpy(), v_ref.real.numpy()) < 1e-11


@pytest.mark.parametrize('case', list(CASES))
def test_hpsi_and_band_trace(cuda_device, case):
  """hamiltonian_matrix_trace value + gradient (band mode) through jrb_hpsi."""
  s, plan, w_re, w_im, occ = _setup(case)
  q = rp.unitary_matrix(torch.from_numpy(w_re), torch.from_numpy(w_im))
  c = rp.expand_coefficient(q, s.mask)
  rho = rp.density_grid(c, s.vol, torch.from_numpy(occ))
  ref = rp.band_trace_and_grad(s, w_re, w_im, rho.numpy())
  qd, r = plan.qr_fwd(to_dev(w_re), to_dev(w_im))
  _, veff = plan.grid_potential(rho.cuda().contiguous(), 'lda_x', True)
  hq = plan.hpsi(qd, veff)
  eps = plan.band_expect(qd, hq).cpu().numpy()
  assert relerr(eps, ref['per_band']) < 1e-11
  assert abs(eps.sum() - ref['trace']) / abs(ref['trace']) < E_TOL
  g_re, g_im = plan.qr_bwd(qd, r, hq)
  assert relerr(g_re.cpu().numpy(), ref['g_re']) < G_TOL
  assert relerr(g_im.cpu().numpy(), ref['g_im']) < G_TOL


@pytest.mark.parametrize('case', list(CASES))
@pytest.mark.parametrize('batch_groups', [0, 1, 3])
@pytest.mark.parametrize('fuse', ['fused', 'fused_dense_butterflies', 'fused_no_psi_cache', 'unfused'])
def test_energy_and_grad(cuda_device, case, batch_groups, fuse, monkeypatch):
  # 'unfused' forces the single-pass pencil kernels (the path non-cubic grids and 128^3 take);
  # 'fused_dense_butterflies' disables the band-limited (sparse radix-8) variant;
  # 'fused_no_psi_cache' makes the H-apply repeat the inverse transforms on the kept z columns
  # instead of reading psi(r) of the density sweep back (the path a plan over budget takes)
  monkeypatch.setenv('JRB_NO_FUSE', '1' if fuse == 'unfused' else '0')
  monkeypatch.setenv('JRB_NO_SPARSE', '1' if fuse == 'fused_dense_butterflies' else '0')
  monkeypatch.setenv('JRB_PSI_CACHE_MB', '0' if fuse == 'fused_no_psi_cache' else '65536')
  if fuse == 'fused_dense_butterflies' and case != 'si8_64':
    pytest.skip('only the band-limited 64^3 case has a sparse variant to switch off')
  s, plan, w_re, w_im, occ = _setup(case, batch_groups=batch_groups)
  if plan.lib.jrb_plan_orbital_fused(plan._h) == 1:
    assert (plan.psi_cache_bytes > 0) == (fuse != 'fused_no_psi_cache')
  else:
    assert plan.psi_cache_bytes == 0
  ref = rp.energy_and_grad(s, w_re, w_im, occ, occ_grad=True)
  occ_d = to_dev(occ)
  rho, e_kin = plan.eval_begin(to_dev(w_re), to_dev(w_im), occ_d)
  en, g_re, g_im, g_occ = plan.eval_finish(occ_d, rho, e_kin, 'lda_x', want_occ_grad=True)
  torch.cuda.synchronize()
  en = en.cpu().numpy()
  for i, key in enumerate(['e_kin', 'e_ext', 'e_har', 'e_xc']):
    assert abs(en[i] - ref[key]) / abs(ref[key]) < E_TOL, key
  assert abs(en.sum() - ref['e_tot']) / abs(ref['e_tot']) < E_TOL
  assert relerr(rho.cpu().numpy(), ref['density']) < G_TOL
  assert relerr(g_re.cpu().numpy(), ref['g_re']) < G_TOL
  assert relerr(g_im.cpu().numpy(), ref['g_im']) < G_TOL
  assert relerr(g_occ.cpu().numpy(), ref['g_occ']) < G_TOL


def test_energy_grad_host(cuda_device):
  s, plan, w_re, w_im, occ = _setup('diamond_16')
  ref = rp.energy_and_grad(s, w_re, w_im, occ)
  en, g_re, g_im, rho = plan.energy_grad_host(w_re, w_im, occ, want_rho=True)
  assert abs(en.sum() - ref['e_tot']) / abs(ref['e_tot']) < E_TOL
  assert relerr(g_re, ref['g_re']) < G_TOL
  assert relerr(g_im, ref['g_im']) < G_TOL
  assert relerr(rho, ref['density']) < G_TOL


@pytest.mark.parametrize('chunks', ['1', '2', '3'])
def test_energy_grad_host_chunked(cuda_device, chunks, monkeypatch):
  """jrb_energy_grad_host with k-point chunks (copies overlapped with kernels)."""
  monkeypatch.setenv('JRB_HOST_CHUNKS', chunks)
  s = make_system('diamond', 12, [2, 2, 1], 10, 'spherical')
  w_re, w_im, occ = make_inputs(s, 7, jitter=0.1)
  plan = make_plan(s, 7)
  ref = rp.energy_and_grad(s, w_re, w_im, occ)
  for _ in range(2):  # second call: reused buffers / events
    en, g_re, g_im, rho = plan.energy_grad_host(w_re, w_im, occ, want_rho=True)
    assert abs(en.sum() - ref['e_tot']) / abs(ref['e_tot']) < E_TOL
    assert relerr(g_re, ref['g_re']) < G_TOL
    assert relerr(g_im, ref['g_im']) < G_TOL
    assert relerr(rho, ref['density']) < G_TOL


def test_rank_deficient_parameters_are_reported(cuda_device):
  """Cholesky-QR cannot orthonormalise linearly dependent columns: the host call says so."""
  from jrystal_b200._lib import JrbError
  s = make_system('diamond', 12, [1, 1, 1], 10, 'spherical')
  w_re, w_im, occ = make_inputs(s, 4)
  w_re[..., 1] = w_re[..., 0]
  w_im[..., 1] = w_im[..., 0]
  plan = make_plan(s, 4)
  with pytest.raises(JrbError):
    plan.energy_grad_host(w_re, w_im, occ)
  # the flag does not stick: a well-posed call on the same plan succeeds afterwards
  w_re2, w_im2, _ = make_inputs(s, 4)
  plan.energy_grad_host(w_re2, w_im2, occ)


def test_large_band_count_qr(cuda_device):
  """nb > 72: several Gram super-tiles, several Cholesky panels, several apply column tiles."""
  s = make_system('si', 16, [1, 1, 1], 30, 'spherical')
  nb = 100
  assert s.num_g > 2 * nb
  rng = np.random.default_rng(3)
  w_re = rng.random((1, 1, s.num_g, nb))
  w_im = rng.random((1, 1, s.num_g, nb))
  plan = make_plan(s, nb)
  q, r = plan.qr_fwd(to_dev(w_re), to_dev(w_im))
  qn, rn = q.cpu().numpy()[0, 0], r.cpu().numpy()[0, 0]
  w = (w_re + 1j * w_im)[0, 0]
  assert np.abs(qn.conj().T @ qn - np.eye(nb)).max() < 1e-12
  assert relerr(qn @ rn, w) < 1e-12
  assert np.abs(np.tril(rn, -1)).max() == 0.0
  gq = rng.standard_normal(w_re.shape) + 1j * rng.standard_normal(w_re.shape)
  g_re, g_im = plan.qr_bwd(q, r, to_dev(gq))
  gw = analytic.qr_backward(qn, rn, gq[0, 0])
  assert relerr(g_re[0, 0].cpu().numpy(), 2 * gw.real) < 1e-10
  assert relerr(g_im[0, 0].cpu().numpy(), 2 * gw.imag) < 1e-10


@pytest.mark.parametrize('case', ['diamond_16', 'si8_32_nb130'])
def test_row_sharded_evaluator_single_rank(cuda_device, case):
  """The split-phase QR (jrb_qr_rows_*) + band plan of the Gamma-only multi-GPU layout, world 1:
  must equal the oracle like the fused evaluation does."""
  from jrystal_b200.parallel import RowShardedEvaluator
  c = CASES[case]
  s = make_system(c['name'], c['grid'], c['kgrid'], c['cutoff'], c['mask'])
  w_re, w_im, occ = make_inputs(s, c['nb'], jitter=0.1)
  ref = rp.energy_and_grad(s, w_re, w_im, occ)
  ev = RowShardedEvaluator(s.cell, s.mask, s.kpts, c['nb'], s.positions, s.charges)
  en, g_re, g_im, rho = ev.evaluate(to_dev(w_re), to_dev(w_im), to_dev(occ))
  ev.rows.check_status()
  en = en.cpu().numpy()
  assert abs(en.sum() - ref['e_tot']) / abs(ref['e_tot']) < E_TOL
  assert relerr(rho.cpu().numpy(), ref['density']) < G_TOL
  assert relerr(g_re.cpu().numpy(), ref['g_re']) < G_TOL
  assert relerr(g_im.cpu().numpy(), ref['g_im']) < G_TOL


def test_rows_plan_rejects_grid_calls(cuda_device):
  from jrystal_b200._lib import JrbError, load
  from jrystal_b200.plan import RowsPlan
  rows = RowsPlan(100, 1, 4)
  with pytest.raises(JrbError):
    _check = load().jrb_density(rows._h, None, None, None, None)
    from jrystal_b200._lib import check
    check(_check)


@pytest.mark.parametrize('xc', ['lda_x', 'lda_x+lda_c_pw'])
@pytest.mark.parametrize('case', ['diamond_12', 'si8_64'])
def test_spin_polarised_energy_and_grad(cuda_device, case, xc):
  """Two spin channels (spin_restricted=False, pw.py:88-91 ns = 2): spin-scaled LDA exchange
  (xc.py:54-59) and the functional's own polarised form for the correlation (xc.py:60-61, PW92
  with the spin interpolation), per-spin densities and potentials, gradients of both channels."""
  import jrystal_b200 as jb
  c = CASES[case]
  s = make_system(c['name'], c['grid'], c['kgrid'], c['cutoff'], c['mask'])
  nb = c['nb'] if case != 'diamond_12' else 8
  p = rp.param_init(21, nb, s.num_k, s.mask, spin_restricted=False)
  occ = rp.occupation_uniform(s.num_k, s.num_electrons, spin=2, num_bands=nb,
                              spin_restricted=False).numpy()
  occ = occ * (1.0 + 0.1 * np.random.default_rng(5).random(occ.shape))
  ref = rp.energy_and_grad(s, p['w_re'], p['w_im'], occ, xc=xc, occ_grad=True)
  plan = jb.Plan(s.cell, s.mask, s.kpts, nb, num_spin=2)
  plan.set_atoms(s.positions, s.charges)
  occ_d = to_dev(occ)
  rho, e_kin = plan.eval_begin(to_dev(p['w_re']), to_dev(p['w_im']), occ_d)
  en, g_re, g_im, g_occ = plan.eval_finish(occ_d, rho, e_kin, xc, want_occ_grad=True)
  en = en.cpu().numpy()
  assert abs(en.sum() - ref['e_tot']) / abs(ref['e_tot']) < E_TOL
  assert rho.shape[0] == 2 and relerr(rho.cpu().numpy(), ref['density']) < G_TOL
  assert relerr(g_re.cpu().numpy(), ref['g_re']) < G_TOL
  assert relerr(g_im.cpu().numpy(), ref['g_im']) < G_TOL
  assert relerr(g_occ.cpu().numpy(), ref['g_occ']) < G_TOL


@pytest.mark.parametrize('case,nproj', [('diamond_16', 5), ('si_32', 18), ('diamond_789_cubic', 3)])
def test_nonlocal_pseudopotential_term(cuda_device, case, nproj):
  """energy_nonlocal / hamiltonian_nonlocal (pseudopotential/nloc.py:143-158, 217-236) with
  synthetic projectors on the sphere: E_nl, the total energy (kinetic slot = kinetic + non-local),
  the gradients and dE/d occ against the oracle contracting over the whole box."""
  s, plan, w_re, w_im, occ = _setup(case)
  rng = np.random.default_rng(17)
  phi = 0.3 * (rng.standard_normal((s.num_k, nproj, s.num_g)) +
               1j * rng.standard_normal((s.num_k, nproj, s.num_g)))
  dense = np.zeros((s.num_k, nproj) + tuple(s.mask.shape), dtype=np.complex128)
  dense[:, :, s.mask] = phi
  ref = rp.energy_and_grad(s, w_re, w_im, occ, occ_grad=True, nonlocal_phi=dense)
  plan.set_nonlocal(to_dev(phi))
  occ_d = to_dev(occ)
  rho, e_kin = plan.eval_begin(to_dev(w_re), to_dev(w_im), occ_d)
  en, g_re, g_im, g_occ = plan.eval_finish(occ_d, rho, e_kin, 'lda_x', want_occ_grad=True)
  en = en.cpu().numpy()
  assert abs(en[0] - (ref['e_kin'] + ref['e_nl'])) < E_TOL * abs(ref['e_kin'] + ref['e_nl'])
  assert abs(en.sum() - ref['e_tot']) / abs(ref['e_tot']) < E_TOL
  assert relerr(g_re.cpu().numpy(), ref['g_re']) < G_TOL
  assert relerr(g_im.cpu().numpy(), ref['g_im']) < G_TOL
  assert relerr(g_occ.cpu().numpy(), ref['g_occ']) < G_TOL
  q, _ = plan.qr_fwd(to_dev(w_re), to_dev(w_im))
  e_nl = plan.nonlocal_energy(q, occ_d).item()
  assert abs(e_nl - ref['e_nl']) < 1e-11 * abs(ref['e_nl'])
  # host path (falls back to the unchunked variant) and removal of the projectors
  en_h, g_re_h, _, _ = plan.energy_grad_host(w_re, w_im, occ)
  assert abs(en_h.sum() - ref['e_tot']) / abs(ref['e_tot']) < E_TOL
  assert relerr(g_re_h, ref['g_re']) < G_TOL
  plan.set_nonlocal(None)
  ref0 = rp.energy_and_grad(s, w_re, w_im, occ)
  rho, e_kin = plan.eval_begin(to_dev(w_re), to_dev(w_im), occ_d)
  en0, *_ = plan.eval_finish(occ_d, rho, e_kin, 'lda_x')
  assert abs(en0.sum().item() - ref0['e_tot']) / abs(ref0['e_tot']) < E_TOL


def test_external_position_gradient(cuda_device):
  """dE_ext/dR from the kernel against central differences of the oracle's energy.external."""
  s, plan, w_re, w_im, occ = _setup('diamond_16')
  q = rp.unitary_matrix(torch.from_numpy(w_re), torch.from_numpy(w_im))
  rho = rp.density_grid(rp.expand_coefficient(q, s.mask), s.vol, torch.from_numpy(occ))
  rho_g = torch.fft.fftn(rho, dim=(-3, -2, -1))
  g = plan.external_position_gradient(rho.cuda().contiguous()).cpu().numpy()
  assert g.shape == (len(s.charges), 3)
  h = 1e-5
  for a, c in [(0, 0), (1, 2), (0, 1)]:
    e = []
    for sgn in (+1, -1):
      pos = s.positions.copy()
      pos[a, c] += sgn * h
      e.append(rp.energy_external(rho_g, pos, s.charges, s.g_vec, s.vol).item())
    fd = (e[0] - e[1]) / (2 * h)
    assert abs(g[a, c] - fd) < 1e-6 * max(1.0, abs(fd)), (a, c, g[a, c], fd)
  # translation invariance of the density-potential system is broken only by rho: sum_a grad_a
  # equals minus the force on the electrons; at least it must be finite and real
  assert np.isfinite(g).all()


def test_external_potential_from_the_caller(cuda_device):
  """jrb_set_external_potential with the oracle's V_ext(G) reproduces jrb_set_atoms; a scaled
  V(G) scales E_ext (the local-pseudopotential use: any V_loc(G) the host prepares)."""
  s, plan, w_re, w_im, occ = _setup('diamond_16')
  q = rp.unitary_matrix(torch.from_numpy(w_re), torch.from_numpy(w_im))
  rho = rp.density_grid(rp.expand_coefficient(q, s.mask), s.vol, torch.from_numpy(occ))
  rho_d = rho.cuda().contiguous()
  en0, v0 = plan.grid_potential(rho_d, 'lda_x', False)
  vhat = rp.external_reciprocal(s.positions, s.charges, s.g_vec, s.vol)
  import jrystal_b200 as jb
  plan2 = jb.Plan(s.cell, s.mask, s.kpts, CASES['diamond_16']['nb'])
  with pytest.raises(RuntimeError):
    plan2.grid_potential(rho_d)
  plan2.set_external_potential(torch.as_tensor(vhat).to(torch.complex128).cuda().contiguous())
  en1, v1 = plan2.grid_potential(rho_d, 'lda_x', False)
  assert relerr(en1.cpu().numpy(), en0.cpu().numpy()) < 1e-12
  assert relerr(v1.cpu().numpy(), v0.cpu().numpy()) < 1e-12
  plan2.set_external_potential(0.5 * torch.as_tensor(vhat).to(torch.complex128).cuda().contiguous())
  en2, _ = plan2.grid_potential(rho_d, 'lda_x', False)
  assert abs(en2[1].item() - 0.5 * en0[1].item()) < 1e-12 * abs(en0[1].item())
  assert abs(en2[0].item() - en0[0].item()) < 1e-12 * abs(en0[0].item())


def test_mid_band_count_few_kpoints_qr(cuda_device):
  """33..96 bands with few k-points take the multi-CTA panel Cholesky (two panels, the second one
  partial) and the blocked triangular inverse."""
  s = make_system('si', 16, [1, 1, 2], 30, 'spherical')
  nb = 40
  rng = np.random.default_rng(4)
  w_re = rng.random((1, 2, s.num_g, nb))
  w_im = rng.random((1, 2, s.num_g, nb))
  plan = make_plan(s, nb)
  q, r = plan.qr_fwd(to_dev(w_re), to_dev(w_im))
  gq = rng.standard_normal(w_re.shape) + 1j * rng.standard_normal(w_re.shape)
  g_re, g_im = plan.qr_bwd(q, r, to_dev(gq))
  for k in range(2):
    qn, rn = q.cpu().numpy()[0, k], r.cpu().numpy()[0, k]
    w = (w_re + 1j * w_im)[0, k]
    assert np.abs(qn.conj().T @ qn - np.eye(nb)).max() < 1e-12
    assert relerr(qn @ rn, w) < 1e-12
    assert np.abs(np.tril(rn, -1)).max() == 0.0
    gw = analytic.qr_backward(qn, rn, gq[0, k])
    assert relerr(g_re[0, k].cpu().numpy(), 2 * gw.real) < 1e-10
    assert relerr(g_im[0, k].cpu().numpy(), 2 * gw.imag) < 1e-10


def test_errors(cuda_device):
  import jrystal_b200 as jb
  from jrystal_b200._lib import JrbError
  s = make_system('diamond', [7, 8, 9], [1, 1, 1], None, 'cubic')
  with pytest.raises(JrbError):
    jb.Plan(s.cell, np.ones((11, 8, 9), dtype=bool), s.kpts, 4)  # 11 is not 7-smooth
  with pytest.raises(JrbError):
    jb.Plan(s.cell, np.zeros((7, 8, 9), dtype=bool), s.kpts, 4)  # empty mask
  plan = jb.Plan(s.cell, s.mask, s.kpts, 4)
  with pytest.raises(RuntimeError):
    plan.grid_potential(torch.zeros((1, 7, 8, 9), dtype=torch.float64, device='cuda'))
  with pytest.raises(ValueError):
    plan.density(torch.zeros((1, 1, 3, 4), dtype=torch.complex128, device='cuda'),
                 torch.zeros((1, 1, 4), dtype=torch.float64, device='cuda'))


@pytest.mark.parametrize('shortcut', ['on', 'off'])
@pytest.mark.parametrize('kind', ['random', 'near_dependent'])
@pytest.mark.parametrize('case', ['si_32', 'si8_32_nb130'])
def test_qr_second_pass_paths(cuda_device, case, kind, shortcut, monkeypatch):
  """Second Cholesky-QR pass: the closed-form factor of I + E (|E| < 1e-10: well-conditioned
  parameters) and the regular factorisation (nearly dependent columns push |E| above the
  threshold; JRB_NO_QR_SHORTCUT=1 forces it) must both give W = Q R with orthonormal Q."""
  monkeypatch.setenv('JRB_NO_QR_SHORTCUT', '1' if shortcut == 'off' else '0')
  s, plan, w_re, w_im, occ = _setup(case)
  if kind == 'near_dependent':
    rng = np.random.default_rng(9)
    w_re = w_re.copy()
    w_im = w_im.copy()
    w_re[..., 1] = w_re[..., 0] + 3e-5 * rng.standard_normal(w_re[..., 0].shape)
    w_im[..., 1] = w_im[..., 0] + 3e-5 * rng.standard_normal(w_im[..., 0].shape)
  q, r = plan.qr_fwd(to_dev(w_re), to_dev(w_im))
  plan.check_status()
  q = q.cpu()
  r = r.cpu()
  w = torch.from_numpy(w_re + 1j * w_im)
  nb = w.shape[-1]
  eye = torch.eye(nb, dtype=torch.complex128)
  orth = (q.conj().transpose(-1, -2) @ q - eye).abs().max().item()
  rec = ((q @ r) - w).abs().max().item() / w.abs().max().item()
  assert orth < 5e-13, orth
  # the composed R = R2 R1 of a matrix with kappa ~ 1e5 carries eps * kappa-sized relative errors
  assert rec < (1e-13 if kind == 'random' else 1e-11), rec
  assert (torch.diagonal(r, dim1=-2, dim2=-1).real > 0).all()
  assert torch.tril(r, -1).abs().max().item() == 0.0
