"""GPU parity of the GGA branch (gga_x_pbe, gga_x_pbe+gga_c_pbe; jrystal/_src/xc.py:67-112,
242-253) against the oracle: energies, the potential = exact grid derivative of E_xc, the
evaluation's gradients, band mode."""
import numpy as np
import pytest
import torch

from oracle import reference_port as rp
from tests.common import make_inputs, make_plan, make_system, relerr, to_dev

pytestmark = pytest.mark.gpu

E_TOL = 1e-10
G_TOL = 1e-8
XCS = ['gga_x_pbe', 'gga_x_pbe+gga_c_pbe']
CASES = {
  'diamond_12': dict(name='diamond', grid=12, kgrid=[2, 1, 1], cutoff=10, nb=5),
  'diamond_789': dict(name='diamond', grid=[7, 8, 9], kgrid=[1, 1, 1], cutoff=8, nb=6),
  'si_32': dict(name='si', grid=32, kgrid=[2, 1, 1], cutoff=12, nb=18),
  'si8_64': dict(name='si8', grid=64, kgrid=[1, 1, 1], cutoff=30, nb=11),
  'diamond_24x32x48': dict(name='diamond', grid=[24, 32, 48], kgrid=[1, 1, 2], cutoff=30, nb=10),
}


def _setup(case, **kw):
  c = CASES[case]
  s = make_system(c['name'], c['grid'], c['kgrid'], c['cutoff'], 'spherical')
  w_re, w_im, occ = make_inputs(s, c['nb'], jitter=0.1)
  return s, make_plan(s, c['nb'], **kw), w_re, w_im, occ


def _density(s, w_re, w_im, occ):
  q = rp.unitary_matrix(torch.from_numpy(w_re), torch.from_numpy(w_im))
  return rp.density_grid(rp.expand_coefficient(q, s.mask), s.vol, torch.from_numpy(occ))


@pytest.mark.parametrize('case', list(CASES))
@pytest.mark.parametrize('xc', XCS)
@pytest.mark.parametrize('kohn_sham', [False, True])
def test_grid_potential_gga(cuda_device, case, xc, kohn_sham):
  s, plan, w_re, w_im, occ = _setup(case)
  rho = _density(s, w_re, w_im, occ)
  e_x = rp.energy_xc(rho, s.vol, xc, kohn_sham, s.g_vec).item()
  en, veff = plan.grid_potential(rho.cuda().contiguous(), xc, kohn_sham)
  assert abs(en[2].item() - e_x) / abs(e_x) < E_TOL
  # v_eff = d (E_H + E_ext + E_xc) / d rho on the grid, in units of Omega / N
  r = rho.clone().requires_grad_(True)
  r_g = torch.fft.fftn(r, dim=(-3, -2, -1))
  e = (rp.energy_hartree(r_g, s.g_vec, s.vol) +
       rp.energy_external(r_g, s.positions, s.charges, s.g_vec, s.vol) +
       rp.energy_xc(r, s.vol, xc, False, s.g_vec))
  (g,) = torch.autograd.grad(e, r)
  v_ref = g.numpy() * rho[0].numel() / s.vol
  assert relerr(veff.cpu().numpy(), v_ref) < 1e-10


@pytest.mark.parametrize('case', list(CASES))
@pytest.mark.parametrize('xc', XCS)
def test_energy_and_grad_gga(cuda_device, case, xc):
  s, plan, w_re, w_im, occ = _setup(case, orbital_grid='auto')
  ref = rp.energy_and_grad(s, w_re, w_im, occ, xc=xc, occ_grad=True)
  occ_d = to_dev(occ)
  rho, e_kin = plan.eval_begin(to_dev(w_re), to_dev(w_im), occ_d)
  en, g_re, g_im, g_occ = plan.eval_finish(occ_d, rho, e_kin, xc, want_occ_grad=True)
  en = en.cpu().numpy()
  for i, key in enumerate(['e_kin', 'e_ext', 'e_har', 'e_xc']):
    assert abs(en[i] - ref[key]) / abs(ref[key]) < E_TOL, key
  assert abs(en.sum() - ref['e_tot']) / abs(ref['e_tot']) < E_TOL
  assert relerr(g_re.cpu().numpy(), ref['g_re']) < G_TOL
  assert relerr(g_im.cpu().numpy(), ref['g_im']) < G_TOL
  assert relerr(g_occ.cpu().numpy(), ref['g_occ']) < G_TOL


@pytest.mark.parametrize('xc', XCS)
def test_band_trace_gga(cuda_device, xc):
  s, plan, w_re, w_im, occ = _setup('si_32')
  rho = _density(s, w_re, w_im, occ)
  ref = rp.band_trace_and_grad(s, w_re, w_im, rho.numpy(), xc=xc)
  qd, r = plan.qr_fwd(to_dev(w_re), to_dev(w_im))
  _, veff = plan.grid_potential(rho.cuda().contiguous(), xc, True)
  hq = plan.hpsi(qd, veff)
  eps = plan.band_expect(qd, hq).cpu().numpy()
  assert relerr(eps, ref['per_band']) < 1e-10
  g_re, g_im = plan.qr_bwd(qd, r, hq)
  assert relerr(g_re.cpu().numpy(), ref['g_re']) < G_TOL
  assert relerr(g_im.cpu().numpy(), ref['g_im']) < G_TOL


def test_gga_potential_api_and_errors(cuda_device):
  """jrb_potential with the reference's semantics (eps_xc unless kohn_sham); polarised GGA is
  rejected loudly."""
  import jrystal_b200 as jb
  from jrystal_b200._lib import JrbError
  s, plan, w_re, w_im, occ = _setup('diamond_12')
  rho = _density(s, w_re, w_im, occ)
  xc = 'gga_x_pbe+gga_c_pbe'
  v = plan.potential(rho.cuda().contiguous(), xc, False, parts=4)
  ref = rp.xc_density(rho, False, xc, s.g_vec)
  assert relerr(v.cpu().numpy()[0], ref.numpy()) < 1e-11
  plan2 = jb.Plan(s.cell, s.mask, s.kpts, 5, num_spin=2)
  plan2.set_atoms(s.positions, s.charges)
  rho2 = torch.cat([rho, rho]).cuda().contiguous() * 0.5
  with pytest.raises(JrbError):
    plan2.grid_potential(rho2, xc, False)
