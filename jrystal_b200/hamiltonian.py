"""jrystal.hamiltonian (jrystal/_src/hamiltonian.py) on the CUDA H-apply.

`hamiltonian_matrix_trace` is the band-mode loss (hamiltonian.py:105-168); `hamiltonian_matrix`
(171-240, nb Hessian-vector products through AD in the reference) is evaluated analytically as
C^H (T C + FFT(v IFFT C)) from ONE H-apply."""
import torch

from . import pw as _pw
from .energy import _plan_with_atoms


def _apply(band_coefficient, positions, charges, effictive_density_grid, xc, kohn_sham):
  c = _pw._as_coeff(band_coefficient)
  plan = _plan_with_atoms(positions, charges)
  if c.plan is not plan:
    raise ValueError('coefficients belong to a different plan')
  rho = effictive_density_grid
  if rho.ndim == 3:
    rho = rho[None]
  # v_eff of hamiltonian.py:147-156 = potential.effective(kohn_sham) = dE/drho for kohn_sham
  v = plan.potential(rho.contiguous(), xc, kohn_sham, 7)
  return c, plan, plan.hpsi(c.q, v)


def hamiltonian_matrix_trace(band_coefficient, positions, charges, effictive_density_grid,
                             g_vector_grid, kpts, vol, xc: str = 'lda_x', kohn_sham: bool = True,
                             keep_kpts_axis: bool = False, keep_spin_axis: bool = False):
  """Sum_i <psi_i| T + v_eff |psi_i> (jrystal/_src/hamiltonian.py:105-168).
  keep_kpts_axis=False: the scalar the band-mode loss uses (summed over spin, kpt, band), as the
  reference.  keep_kpts_axis=True: shape [spin, kpt], what the reference's docstring promises; its
  code sums axes (1, 2) there and returns [spin] (hamiltonian.py:165-166) -- `.sum(1)` of this
  result.  keep_spin_axis (not a reference argument) only matters with keep_kpts_axis=False:
  True keeps [spin]."""
  del g_vector_grid, kpts, vol
  c = _pw._as_coeff(band_coefficient)
  if torch.is_grad_enabled() and c.q.requires_grad:
    # the band-mode loss under torch.autograd (calc_band_structure_all_electrons.py:139-152
    # differentiates this call): the potential of the given density is a constant
    from .autograd import band_expectation
    plan = _plan_with_atoms(positions, charges)
    rho = effictive_density_grid if effictive_density_grid.ndim == 4 else effictive_density_grid[None]
    eps = band_expectation(c.q, plan, plan.potential(rho.detach().contiguous(), xc, kohn_sham, 7))
  else:
    c, plan, hq = _apply(c, positions, charges, effictive_density_grid, xc, kohn_sham)
    eps = plan.band_expect(c.q, hq)        # (ns, nk, nb)
  out = eps.sum(dim=-1)                    # (ns, nk)
  if keep_kpts_axis:
    return out
  out = out.sum(dim=1)
  return out if keep_spin_axis else out.sum(dim=0)


def hamiltonian_matrix(band_coefficient, positions, charges, effictive_density_grid,
                       g_vector_grid, kpts, vol, xc: str = 'lda_x', kohn_sham: bool = True):
  """H_ij = <psi_i| T + v_eff |psi_j>, shape (spin, kpt, band, band)
  (jrystal/_src/hamiltonian.py:171-240)."""
  del g_vector_grid, kpts, vol
  c, plan, hq = _apply(band_coefficient, positions, charges, effictive_density_grid, xc, kohn_sham)
  return plan.overlap(c.q, hq)
