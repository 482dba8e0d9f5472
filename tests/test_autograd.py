"""jrystal_b200.autograd: the drivers' losses as torch.autograd functions (the in-container
counterpart of the jax.custom_vjp binding).  Same bodies on the GPU product and on the CPU stand-in."""
import numpy as np
import pytest
import torch

from oracle import reference_port as rp
from tests.common import relerr
from tests.conftest import BACKENDS


def _case():
  s = rp.System.from_name('diamond', [12, 12, 12], [1, 1, 2], 10.0)
  nb = 8
  p = rp.param_init(9, nb, s.num_k, s.mask)
  return s, nb, p


@pytest.mark.parametrize('backend', BACKENDS, indirect=True)
def test_total_energy_is_differentiable_through_a_trainable_occupation(backend):
  """d/d(w_re, w_im, occupation parameters) of total_energy(param_pw, idempotent(params_occ)) by
  torch.autograd == the oracle's autograd of the same composition (what jax.value_and_grad gives
  the reference for its free-energy loss at T = 0)."""
  import jrystal_b200 as jb
  kind, Plan, dev = backend
  s, nb, p = _case()
  ne = s.num_electrons
  w_occ = np.random.default_rng(2).random((nb * s.num_k, (ne // 2) * s.num_k))
  plan = Plan(s.cell, s.mask, s.kpts, nb)
  plan.set_atoms(s.positions, s.charges)
  w_re = dev(p['w_re']).requires_grad_(True)
  w_im = dev(p['w_im']).requires_grad_(True)
  leaf = dev(w_occ).requires_grad_(True)
  with jb.use_plan(plan):
    occ = jb.occupation.idempotent({'param_up': {'w_re': leaf}, 'param_down': {'w_re': leaf}}, s.num_k)
    e, energies, rho = jb.autograd.total_energy({'w_re': w_re, 'w_im': w_im}, occ, 'lda_x', split=True)
    g_re, g_im, g_leaf = torch.autograd.grad(3.0 * e, [w_re, w_im, leaf])     # a cotangent of 3
  assert not energies.requires_grad and not rho.requires_grad
  # oracle: the same composition
  o_leaf = torch.from_numpy(w_occ).requires_grad_(True)
  o_occ = rp.occupation_idempotent(o_leaf, o_leaf, s.num_k)
  ref = rp.energy_and_grad(s, p['w_re'], p['w_im'], o_occ.detach().numpy(), occ_grad=True)
  (o_g_leaf,) = torch.autograd.grad((torch.from_numpy(ref['g_occ']) * o_occ).sum(), o_leaf)
  assert abs(float(e) - ref['e_tot']) < 1e-10 * abs(ref['e_tot'])
  assert relerr(rho.cpu().numpy(), ref['density']) < 1e-8
  assert relerr(g_re.cpu().numpy(), 3.0 * ref['g_re']) < 1e-8
  assert relerr(g_im.cpu().numpy(), 3.0 * ref['g_im']) < 1e-8
  assert relerr(g_leaf.cpu().numpy(), 3.0 * o_g_leaf.numpy()) < 1e-8


@pytest.mark.parametrize('backend', BACKENDS, indirect=True)
def test_hamiltonian_trace_is_differentiable(backend):
  """Band-mode loss with a fixed potential: value, per-band values and d/d(w_re, w_im) against the
  oracle's hamiltonian_matrix_trace and its autograd."""
  import jrystal_b200 as jb
  kind, Plan, dev = backend
  s, nb, p = _case()
  occ = rp.occupation_uniform(s.num_k, s.num_electrons, num_bands=nb)
  q = rp.unitary_matrix(torch.from_numpy(p['w_re']), torch.from_numpy(p['w_im']))
  rho = rp.density_grid(rp.expand_coefficient(q, s.mask), s.vol, occ)
  plan = Plan(s.cell, s.mask, s.kpts, nb)
  plan.set_atoms(s.positions, s.charges)
  w_re = dev(p['w_re']).requires_grad_(True)
  w_im = dev(p['w_im']).requires_grad_(True)
  veff = plan.potential(dev(rho.numpy()), 'lda_x', True, 7)
  tr, eps = jb.autograd.hamiltonian_trace((w_re, w_im), veff, plan=plan, per_band=True)
  g_re, g_im = torch.autograd.grad(tr, [w_re, w_im])
  ref = rp.band_trace_and_grad(s, p['w_re'], p['w_im'], rho.numpy(), 'lda_x')
  assert abs(float(tr) - ref['trace']) < 1e-10 * abs(ref['trace'])
  assert relerr(eps.cpu().numpy(), ref['per_band']) < 1e-10
  assert relerr(g_re.cpu().numpy(), ref['g_re']) < 1e-8
  assert relerr(g_im.cpu().numpy(), ref['g_im']) < 1e-8
  # prepared potential: veff=None
  plan.prepare_potential(veff)
  tr2 = jb.autograd.hamiltonian_trace({'w_re': w_re, 'w_im': w_im}, None, plan=plan)
  assert abs(float(tr2) - float(tr)) < 1e-12 * abs(float(tr))


@pytest.mark.parametrize('backend', BACKENDS, indirect=True)
def test_reference_call_sequence_under_autograd(backend):
  """The tutorial's closure, call for call (docs/examples/dft100lines.rst: occupation.idempotent ->
  pw.coeff -> energy.total_energy), differentiated by torch.autograd through the fine-grained
  custom backwards (QR adjoint; H-apply) == the fused evaluation == the oracle."""
  import jrystal_b200 as jb
  kind, Plan, dev = backend
  s, nb, p = _case()
  ne = s.num_electrons
  w_occ = np.random.default_rng(2).random((nb * s.num_k, (ne // 2) * s.num_k))
  plan = Plan(s.cell, s.mask, s.kpts, nb)
  w_re = dev(p['w_re']).requires_grad_(True)
  w_im = dev(p['w_im']).requires_grad_(True)
  leaf = dev(w_occ).requires_grad_(True)
  with jb.use_plan(plan):
    occ = jb.occupation.idempotent({'param_up': {'w_re': leaf}, 'param_down': {'w_re': leaf}}, s.num_k)
    coeff = jb.pw.coeff({'w_re': w_re, 'w_im': w_im}, s.mask)
    assert coeff.q.requires_grad
    e = jb.energy.total_energy(coeff, s.positions, s.charges, s.g_vec, s.kpts, s.vol, occ, xc='lda_x')
    g_re, g_im, g_leaf = torch.autograd.grad(e, [w_re, w_im, leaf])
    parts = jb.energy.total_energy(coeff, s.positions, s.charges, s.g_vec, s.kpts, s.vol, occ,
                                   xc='lda_x', split=True)
    with torch.no_grad():   # outside autograd nothing changes: plain values
      e_plain = jb.energy.total_energy(jb.pw.coeff({'w_re': w_re, 'w_im': w_im}, s.mask), s.positions,
                                       s.charges, s.g_vec, s.kpts, s.vol, occ.detach())
  o_leaf = torch.from_numpy(w_occ).requires_grad_(True)
  o_occ = rp.occupation_idempotent(o_leaf, o_leaf, s.num_k)
  ref = rp.energy_and_grad(s, p['w_re'], p['w_im'], o_occ.detach().numpy(), occ_grad=True)
  (o_g_leaf,) = torch.autograd.grad((torch.from_numpy(ref['g_occ']) * o_occ).sum(), o_leaf)
  assert abs(e.item() - ref['e_tot']) < 1e-10 * abs(ref['e_tot'])
  assert abs(float(e_plain) - ref['e_tot']) < 1e-10 * abs(ref['e_tot'])
  for got, key in zip(parts, ['e_kin', 'e_ext', 'e_har', 'e_xc']):
    assert not got.requires_grad and abs(float(got) - ref[key]) < 1e-10 * abs(ref[key])
  assert relerr(g_re.cpu().numpy(), ref['g_re']) < 1e-8
  assert relerr(g_im.cpu().numpy(), ref['g_im']) < 1e-8
  assert relerr(g_leaf.cpu().numpy(), o_g_leaf.numpy()) < 1e-8


@pytest.mark.parametrize('backend', BACKENDS, indirect=True)
def test_band_mode_loss_call_under_autograd(backend):
  """hamiltonian.hamiltonian_matrix_trace(pw.coeff(params), ..., density) differentiated by
  torch.autograd, the call the reference's band driver differentiates."""
  import jrystal_b200 as jb
  kind, Plan, dev = backend
  s, nb, p = _case()
  occ = rp.occupation_uniform(s.num_k, s.num_electrons, num_bands=nb)
  q = rp.unitary_matrix(torch.from_numpy(p['w_re']), torch.from_numpy(p['w_im']))
  rho = rp.density_grid(rp.expand_coefficient(q, s.mask), s.vol, occ)
  plan = Plan(s.cell, s.mask, s.kpts, nb)
  w_re = dev(p['w_re']).requires_grad_(True)
  w_im = dev(p['w_im']).requires_grad_(True)
  with jb.use_plan(plan):
    coeff = jb.pw.coeff({'w_re': w_re, 'w_im': w_im}, s.mask)
    tr = jb.hamiltonian.hamiltonian_matrix_trace(coeff, s.positions, s.charges, dev(rho.numpy()),
                                                 s.g_vec, s.kpts, s.vol, xc='lda_x', kohn_sham=True)
    g_re, g_im = torch.autograd.grad(tr, [w_re, w_im])
  ref = rp.band_trace_and_grad(s, p['w_re'], p['w_im'], rho.numpy(), 'lda_x')
  assert abs(tr.item() - ref['trace']) < 1e-10 * abs(ref['trace'])
  assert relerr(g_re.cpu().numpy(), ref['g_re']) < 1e-8
  assert relerr(g_im.cpu().numpy(), ref['g_im']) < 1e-8
