"""Differentiable entry points: the two losses of the reference's drivers as `torch.autograd`
functions whose backward is the hand-written reverse pass of the CUDA path.

This is what `jax.custom_vjp` is in the XLA-FFI binding, spelled for the
host framework this repo runs on: `total_energy(param_pw, occupation)` behaves like the closure
`jax.value_and_grad` differentiates in calc_ground_state_energy_all_electrons.py:119-181 --
gradients flow to `w_re`, `w_im` AND to the occupation numbers, so an occupation map written in
torch (occupation.simplex_projector / idempotent) chains in front of it -- and `hamiltonian_trace`
like the band-mode loss of calc_band_structure_all_electrons.py:139-152.

Forward = jrb_eval_begin (+ all-reduce of rho, E_kin over a k mesh) + jrb_eval_finish, which
already produces the gradients; backward only scales them by the incoming cotangent.  No graph of
intermediate tensors is kept: psi(r) is recomputed inside the H-apply instead of being saved."""
from typing import Optional

import torch

from . import parallel
from .context import current_plan


class _TotalEnergy(torch.autograd.Function):

  @staticmethod
  def forward(ctx, w_re, w_im, occ, plan, xc, k_mesh):
    rho, e_kin = plan.eval_begin(w_re.contiguous(), w_im.contiguous(), occ.contiguous())
    if k_mesh:
      parallel.allreduce_density(rho, e_kin)
    energies, g_re, g_im, g_occ = plan.eval_finish(occ.contiguous(), rho, e_kin, xc,
                                                   want_occ_grad=True)
    ctx.save_for_backward(g_re, g_im, g_occ)
    ctx.mark_non_differentiable(energies, rho)
    return energies.sum(), energies, rho

  @staticmethod
  def backward(ctx, ct, _ct_energies, _ct_rho):
    g_re, g_im, g_occ = ctx.saved_tensors
    return ct * g_re, ct * g_im, ct * g_occ, None, None, None


def total_energy(param_pw, occupation, xc: str = 'lda_x', plan=None, split: bool = False,
                 k_mesh: bool = False):
  """E_kin + E_ext + E_har + E_xc of the current plan's crystal as a differentiable scalar
  (energy.total_energy of the reference composed with pw.coeff, pw.density_grid, ...).

  param_pw: {'w_re', 'w_im'} float64 tensors (ns, nk, ng, nb) -- leaves or results of other torch
  ops; occupation: (ns, nk, nb) float64, may carry a graph (a trainable occupation map).
  split=True also returns the four terms (kinetic [+ non-local], external / local, Hartree, xc) and
  the density, detached.  k_mesh=True all-reduces rho and E_kin over torch.distributed between the
  two halves (the plan holds this rank's k-points)."""
  plan = plan or current_plan()
  if isinstance(param_pw, dict):
    w_re, w_im = param_pw['w_re'], param_pw['w_im']
  else:
    w_re, w_im = param_pw
  e, energies, rho = _TotalEnergy.apply(w_re, w_im, occupation, plan, xc, bool(k_mesh))
  return (e, energies, rho) if split else e


class _HamiltonianTrace(torch.autograd.Function):

  @staticmethod
  def forward(ctx, w_re, w_im, plan, veff):
    q, r = plan.qr_fwd(w_re.contiguous(), w_im.contiguous())
    hq = plan.hpsi(q, veff)
    eps = plan.band_expect(q, hq)
    g_re, g_im = plan.qr_bwd(q, r, hq)
    ctx.save_for_backward(g_re, g_im)
    ctx.mark_non_differentiable(eps)
    return eps.sum(), eps

  @staticmethod
  def backward(ctx, ct, _ct_eps):
    g_re, g_im = ctx.saved_tensors
    return ct * g_re, ct * g_im, None, None


def hamiltonian_trace(param_pw, veff: Optional[torch.Tensor] = None, plan=None,
                      per_band: bool = False):
  """sum_i <psi_i| T + v_eff (+ V_nl) |psi_i> for orthonormalised parameters, differentiable in
  w_re / w_im: the band-mode loss (hamiltonian.hamiltonian_matrix_trace with a FIXED potential,
  hamiltonian.py:105-168).  veff: (ns, x, y, z) real, e.g. Plan.potential(rho_gs, xc, True);
  None uses the potential of the last Plan.prepare_potential.  per_band=True also returns the
  (ns, nk, nb) expectation values, detached."""
  plan = plan or current_plan()
  if isinstance(param_pw, dict):
    w_re, w_im = param_pw['w_re'], param_pw['w_im']
  else:
    w_re, w_im = param_pw
  tr, eps = _HamiltonianTrace.apply(w_re, w_im, plan, veff)
  return (tr, eps) if per_band else tr


# ---------------------------------------------------------------------------------------------
# Fine-grained pieces: the reference's own call sequence (pw.coeff -> energy.total_energy) under
# torch.autograd.  Gradients of complex tensors follow torch's convention: for a real loss L the
# gradient of z = x + i y is dL/dx + i dL/dy = 2 dL/dz*.
# ---------------------------------------------------------------------------------------------

class _OrthonormalCoefficients(torch.autograd.Function):
  """pw.coeff: Q of the QR of w_re + i w_im (jrb_qr_fwd); backward = the closed-form QR adjoint
  (jrb_qr_bwd), which maps dE/dQ* to (dE/dw_re, dE/dw_im)."""

  @staticmethod
  def forward(ctx, w_re, w_im, plan):
    q, r = plan.qr_fwd(w_re.contiguous(), w_im.contiguous())
    ctx.plan = plan
    ctx.save_for_backward(q, r)
    ctx.mark_non_differentiable(r)
    return q, r

  @staticmethod
  def backward(ctx, grad_q, _grad_r):
    q, r = ctx.saved_tensors
    g_re, g_im = ctx.plan.qr_bwd(q, r, (0.5 * grad_q).contiguous())   # torch grad = 2 dL/dQ*
    return g_re, g_im, None


def orthonormal_coefficients(w_re, w_im, plan):
  return _OrthonormalCoefficients.apply(w_re, w_im, plan)


class _EnergyOfCoefficients(torch.autograd.Function):
  """E_kin + E_ext + E_har + E_xc of orthonormal coefficients Q and occupations f (jrb_density,
  jrb_kinetic, jrb_grid_potential); backward: dE/dQ* = f H Q (one jrb_hpsi with
  v_eff = dE/d rho) and dE/df = <q|H|q> (jrb_band_expect)."""

  @staticmethod
  def forward(ctx, q, occ, plan, xc):
    q, occ = q.contiguous(), occ.contiguous()
    rho = plan.density(q, occ)
    t = plan.kinetic(q)
    en, veff = plan.grid_potential(rho, xc, False)
    e_kin = (t * occ).sum()
    if getattr(plan, 'nproj', 0):
      e_kin = e_kin + plan.nonlocal_energy(q, occ)[0]
    energies = torch.stack([e_kin, en[1], en[0], en[2]])
    ctx.plan = plan
    ctx.save_for_backward(q, occ, veff)
    ctx.mark_non_differentiable(energies, rho)
    return energies.sum(), energies, rho

  @staticmethod
  def backward(ctx, ct, _ct_energies, _ct_rho):
    q, occ, veff = ctx.saved_tensors
    hq = ctx.plan.hpsi(q, veff)
    g_occ = ctx.plan.band_expect(q, hq)
    g_q = 2.0 * hq * occ[:, :, None, :].to(hq.dtype)
    return ct * g_q, ct * g_occ, None, None


def energy_of_coefficients(q, occupation, plan, xc: str = 'lda_x'):
  """(E, energies[4] detached, rho detached) for Q (ns, nk, ng, nb) complex and f (ns, nk, nb)."""
  return _EnergyOfCoefficients.apply(q, occupation, plan, xc)


class _BandExpectation(torch.autograd.Function):
  """eps[s, k, b] = <q_b| T + v (+ V_nl) |q_b> for a FIXED potential (jrb_hpsi + jrb_band_expect);
  H is Hermitian, so d(sum ct eps)/dQ* = ct H Q."""

  @staticmethod
  def forward(ctx, q, plan, veff):
    q = q.contiguous()
    hq = plan.hpsi(q, veff)
    ctx.save_for_backward(hq)
    return plan.band_expect(q, hq)

  @staticmethod
  def backward(ctx, ct):
    (hq,) = ctx.saved_tensors
    return 2.0 * hq * ct[:, :, None, :].to(hq.dtype), None, None


def band_expectation(q, plan, veff):
  return _BandExpectation.apply(q, plan, veff)
