"""GPU parity of the orbital grid (jrb_plan_set_orbital_grid): the per-orbital transforms run on a
smaller alias-free box, every result must still be the reference's on the plan's own grid
(oracle = reference dataflow on the full box; tolerances of BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import reference_port as rp
from tests.common import make_inputs, make_plan, make_system, relerr, to_dev

pytestmark = pytest.mark.gpu

E_TOL = 1e-10
G_TOL = 1e-8

# (system, cutoff, nb, grid) -> orbital boxes to try; 4 gmax + 1 is 49 (si8 at 30 Ha), 57 (si8 at
# 40 Ha), 21 (diamond 24x32x48 at 30 Ha)
CASES = {
  'si8_64': dict(name='si8', grid=64, kgrid=[1, 1, 1], cutoff=30, nb=11,
                 boxes=[(64, 64, 49), (64, 64, 50), (64, 64, 54), (64, 64, 56), (64, 64, 60),
                        (56, 56, 56), (64, 60, 49), (49, 49, 49)]),
  'si8_128': dict(name='si8', grid=128, kgrid=[1, 1, 1], cutoff=40, nb=6,
                  boxes=[(128, 128, 60), (128, 128, 64), (128, 128, 80), (128, 128, 81),
                         (128, 128, 90), (64, 64, 64), (72, 72, 60), (81, 81, 81)]),
  'diamond_24x32x48': dict(name='diamond', grid=[24, 32, 48], kgrid=[1, 1, 2], cutoff=30, nb=10,
                           boxes=[(24, 24, 32), (24, 32, 40), (24, 24, 45)]),
  # 48^3 (the C5 band-mode grid): 36 x 36 planes (6 x 6 line plan) instead of 48 x 48
  'si_48': dict(name='si', grid=48, kgrid=[1, 1, 2], cutoff=15, nb=21,
                boxes=[(36, 36, 36), (48, 48, 36)]),
  # x and y under-resolved by the caller's own grid (4 gmax + 1 = 13 > 12: the reference aliases
  # there, and so must we, Nyquist planes included) while z shrinks exactly
  'diamond_12x12x32_aliasing': dict(name='diamond', grid=[12, 12, 32], kgrid=[1, 1, 1], cutoff=10,
                                    nb=6, boxes=[(12, 12, 16), (12, 12, 24)]),
}
PARAMS = [(c, b) for c, v in CASES.items() for b in v['boxes']]


def _setup(case, **plan_kw):
  c = CASES[case]
  s = make_system(c['name'], c['grid'], c['kgrid'], c['cutoff'], 'spherical')
  w_re, w_im, occ = make_inputs(s, c['nb'], jitter=0.1)
  plan = make_plan(s, c['nb'], **plan_kw)
  return s, plan, w_re, w_im, occ


_REF = {}


def _ref(case, s, w_re, w_im, occ):
  if case not in _REF:
    _REF[case] = rp.energy_and_grad(s, w_re, w_im, occ, occ_grad=True)
  return _REF[case]


@pytest.mark.parametrize('case,box', PARAMS)
def test_energy_and_grad_on_orbital_grid(cuda_device, case, box):
  s, plan, w_re, w_im, occ = _setup(case, orbital_grid=box)
  assert plan.orbital_grid == tuple(box)
  need = plan.min_orbital_grid
  assert all(n >= m or n == f for n, m, f in zip(box, need, s.mask.shape))
  ref = _ref(case, s, w_re, w_im, occ)
  occ_d = to_dev(occ)
  rho, e_kin = plan.eval_begin(to_dev(w_re), to_dev(w_im), occ_d)
  en, g_re, g_im, g_occ = plan.eval_finish(occ_d, rho, e_kin, 'lda_x', want_occ_grad=True)
  torch.cuda.synchronize()
  en = en.cpu().numpy()
  for i, key in enumerate(['e_kin', 'e_ext', 'e_har', 'e_xc']):
    assert abs(en[i] - ref[key]) / abs(ref[key]) < E_TOL, key
  assert abs(en.sum() - ref['e_tot']) / abs(ref['e_tot']) < E_TOL
  assert tuple(rho.shape[1:]) == tuple(s.mask.shape)
  assert relerr(rho.cpu().numpy(), ref['density']) < G_TOL
  assert relerr(g_re.cpu().numpy(), ref['g_re']) < G_TOL
  assert relerr(g_im.cpu().numpy(), ref['g_im']) < G_TOL
  assert relerr(g_occ.cpu().numpy(), ref['g_occ']) < G_TOL


def test_auto_orbital_grid(cuda_device):
  s, plan, w_re, w_im, occ = _setup('si8_64', orbital_grid='auto')
  assert plan.min_orbital_grid == (49, 49, 49)
  assert plan.orbital_grid == (64, 64, 49)
  ref = _ref('si8_64', s, w_re, w_im, occ)
  occ_d = to_dev(occ)
  rho, e_kin = plan.eval_begin(to_dev(w_re), to_dev(w_im), occ_d)
  en, g_re, _, _ = plan.eval_finish(occ_d, rho, e_kin, 'lda_x')
  assert abs(en.sum().item() - ref['e_tot']) / abs(ref['e_tot']) < E_TOL
  assert relerr(g_re.cpu().numpy(), ref['g_re']) < G_TOL


def test_orbital_grid_host_path_chunked(cuda_device, monkeypatch):
  """jrb_energy_grad_host: k-chunked pipeline accumulating the density on the orbital grid."""
  monkeypatch.setenv('JRB_HOST_CHUNKS', '2')
  s = make_system('diamond', 32, [2, 2, 1], 20, 'spherical')
  w_re, w_im, occ = make_inputs(s, 7, jitter=0.1)
  plan = make_plan(s, 7)
  assert plan.min_orbital_grid == (17, 17, 17)
  plan.set_orbital_grid((32, 32, 24))
  ref = rp.energy_and_grad(s, w_re, w_im, occ)
  for _ in range(2):
    en, g_re, g_im, rho = plan.energy_grad_host(w_re, w_im, occ, want_rho=True)
    assert abs(en.sum() - ref['e_tot']) / abs(ref['e_tot']) < E_TOL
    assert relerr(g_re, ref['g_re']) < G_TOL
    assert relerr(g_im, ref['g_im']) < G_TOL
    assert relerr(rho, ref['density']) < G_TOL


@pytest.mark.parametrize('box', [(64, 64, 49), (56, 56, 50)])
def test_band_mode_on_orbital_grid(cuda_device, box):
  """hamiltonian_matrix_trace value + gradient: jrb_hpsi resamples the caller's v_eff; a k-point
  move (jrb_set_kpoints) reaches the orbital grid's kinetic table."""
  s, plan, w_re, w_im, occ = _setup('si8_64', orbital_grid=box)
  q = rp.unitary_matrix(torch.from_numpy(w_re), torch.from_numpy(w_im))
  c = rp.expand_coefficient(q, s.mask)
  rho = rp.density_grid(c, s.vol, torch.from_numpy(occ))
  kpts2 = s.kpts + np.array([[0.11, -0.07, 0.05]])
  s2 = make_system('si8', 64, [1, 1, 1], 30, 'spherical')
  s2.kpts = kpts2
  for system, kp in ((s, s.kpts), (s2, kpts2)):
    plan.set_kpoints(kp)
    ref = rp.band_trace_and_grad(system, w_re, w_im, rho.numpy())
    qd, r = plan.qr_fwd(to_dev(w_re), to_dev(w_im))
    _, veff = plan.grid_potential(rho.cuda().contiguous(), 'lda_x', True)
    hq = plan.hpsi(qd, veff)
    eps = plan.band_expect(qd, hq).cpu().numpy()
    assert relerr(eps, ref['per_band']) < 1e-11
    g_re, g_im = plan.qr_bwd(qd, r, hq)
    assert relerr(g_re.cpu().numpy(), ref['g_re']) < G_TOL
    assert relerr(g_im.cpu().numpy(), ref['g_im']) < G_TOL
    # the prepared potential (jrb_hpsi_prepare + veff = NULL) gives the same H-apply
    plan.prepare_potential(veff)
    veff.zero_()   # the plan must not read the caller's buffer any more
    hq2 = plan.hpsi(qd, None)
    assert (hq2 - hq).abs().max().item() <= 1e-14 * hq.abs().max().item()


def test_hpsi_on_36_box_equals_full_grid(cuda_device):
  """Band mode calls jrb_hpsi alone (no psi(r) cache): the recomputing plane kernel of the 36 x 36
  box against the same H-apply on the plan's own 48^3 grid."""
  s, plan, w_re, w_im, occ = _setup('si_48')
  rng = np.random.default_rng(3)
  veff = to_dev(rng.standard_normal((1,) + tuple(s.mask.shape)))
  qd, _ = plan.qr_fwd(to_dev(w_re), to_dev(w_im))
  hq_full = plan.hpsi(qd, veff).clone()
  plan.set_orbital_grid((36, 36, 36))
  assert plan.lib.jrb_plan_orbital_fused(plan._h) == 1
  hq = plan.hpsi(qd, veff)
  assert (hq - hq_full).abs().max().item() <= 1e-12 * hq_full.abs().max().item()


def test_prepared_potential_rules(cuda_device):
  """veff = NULL without a prepared potential is an error; an explicit veff or an evaluation
  replaces the prepared one; works on a plan without an orbital grid too (48^3 -> 36^3 with)."""
  from jrystal_b200._lib import JrbError
  s = make_system('si', 48, [1, 1, 2], 15, 'spherical')
  nb = 9
  w_re, w_im, occ = make_inputs(s, nb, jitter=0.1)
  for og in (None, 'auto'):
    plan = make_plan(s, nb, orbital_grid=og)
    if og == 'auto':
      assert plan.orbital_grid[:2] == (36, 36) and plan.orbital_grid[2] < 48, plan.orbital_grid
    q, r = plan.qr_fwd(to_dev(w_re), to_dev(w_im))
    with pytest.raises(JrbError):
      plan.hpsi(q, None)
    rho = plan.density(q, to_dev(occ))
    _, veff = plan.grid_potential(rho, 'lda_x', True)
    ref = plan.hpsi(q, veff).clone()
    with pytest.raises(JrbError):   # an explicit veff does not leave a prepared potential behind
      plan.hpsi(q, None)
    plan.prepare_potential(veff)
    assert (plan.hpsi(q, None) - ref).abs().max().item() <= 1e-14 * ref.abs().max().item()
    occ_d = to_dev(occ)
    rho2, e_kin = plan.eval_begin(to_dev(w_re), to_dev(w_im), occ_d)
    plan.eval_finish(occ_d, rho2, e_kin, 'lda_x')
    with pytest.raises(JrbError):   # the evaluation overwrote the plan's potential
      plan.hpsi(q, None)


def test_spin_polarised_on_orbital_grid(cuda_device):
  import jrystal_b200 as jb
  s = make_system('si8', 64, [1, 1, 1], 30, 'spherical')
  nb = 11
  p = rp.param_init(21, nb, s.num_k, s.mask, spin_restricted=False)
  occ = rp.occupation_uniform(s.num_k, s.num_electrons, spin=2, num_bands=nb,
                              spin_restricted=False).numpy()
  occ = occ * (1.0 + 0.1 * np.random.default_rng(5).random(occ.shape))
  ref = rp.energy_and_grad(s, p['w_re'], p['w_im'], occ)
  plan = jb.Plan(s.cell, s.mask, s.kpts, nb, num_spin=2, orbital_grid=(64, 64, 50))
  plan.set_atoms(s.positions, s.charges)
  occ_d = to_dev(occ)
  rho, e_kin = plan.eval_begin(to_dev(p['w_re']), to_dev(p['w_im']), occ_d)
  en, g_re, g_im, _ = plan.eval_finish(occ_d, rho, e_kin, 'lda_x')
  assert abs(en.sum().item() - ref['e_tot']) / abs(ref['e_tot']) < E_TOL
  assert relerr(rho.cpu().numpy(), ref['density']) < G_TOL
  assert relerr(g_re.cpu().numpy(), ref['g_re']) < G_TOL
  assert relerr(g_im.cpu().numpy(), ref['g_im']) < G_TOL


def test_nonlocal_on_orbital_grid(cuda_device):
  s = make_system('si', 32, [2, 1, 1], 12, 'spherical')
  nb, nproj = 18, 7
  w_re, w_im, occ = make_inputs(s, nb, jitter=0.1)
  plan = make_plan(s, nb)
  need = plan.min_orbital_grid
  box = tuple(min(n for n in (16, 24, 32) if n >= m) for m in need)
  plan.set_orbital_grid((32, 32, box[2]))
  rng = np.random.default_rng(17)
  phi = 0.3 * (rng.standard_normal((s.num_k, nproj, s.num_g)) +
               1j * rng.standard_normal((s.num_k, nproj, s.num_g)))
  dense = np.zeros((s.num_k, nproj) + tuple(s.mask.shape), dtype=np.complex128)
  dense[:, :, s.mask] = phi
  ref = rp.energy_and_grad(s, w_re, w_im, occ, nonlocal_phi=dense)
  plan.set_nonlocal(to_dev(phi))
  occ_d = to_dev(occ)
  rho, e_kin = plan.eval_begin(to_dev(w_re), to_dev(w_im), occ_d)
  en, g_re, g_im, _ = plan.eval_finish(occ_d, rho, e_kin, 'lda_x')
  assert abs(en.sum().item() - ref['e_tot']) / abs(ref['e_tot']) < E_TOL
  assert relerr(g_re.cpu().numpy(), ref['g_re']) < G_TOL
  assert relerr(g_im.cpu().numpy(), ref['g_im']) < G_TOL


def test_orbital_grid_errors(cuda_device):
  from jrystal_b200._lib import JrbError
  s, plan, *_ = _setup('si8_64')
  with pytest.raises(JrbError):   # 48 < 4 gmax + 1 = 49: would alias
    plan.set_orbital_grid((64, 64, 48))
  with pytest.raises(JrbError):   # larger than the plan's grid
    plan.set_orbital_grid((64, 64, 72))
  with pytest.raises(JrbError):   # 51 has no compiled line plan
    plan.set_orbital_grid((64, 64, 51))
  assert plan.orbital_grid == (64, 64, 64)
  plan.set_orbital_grid((64, 64, 49))
  plan.set_orbital_grid((64, 64, 56))  # replacing the child plan is allowed
  assert plan.orbital_grid == (64, 64, 56)
