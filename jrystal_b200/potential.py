"""Potential module: same names as jrystal.potential (jrystal/_src/potential.py), real-space
results from the fused grid kernels, with the reference's semantics (Hartree potential halved
and xc = eps_xc unless kohn_sham; see include/jrystal_b200.h jrb_potential)."""
from .context import current_plan
from .energy import _plan_with_atoms

HARTREE, EXTERNAL, XC = 1, 2, 4


def effective(density_grid, position, charge, g_vector_grid, vol, split: bool = False,
              xc_type: str = 'lda_x', kohn_sham: bool = False):
  """jrystal/_src/potential.py:203-279 (real part; the reference returns a complex array whose
  imaginary part comes from non-Hermitian Nyquist planes and is discarded downstream,
  hamiltonian.py:164)."""
  del g_vector_grid, vol
  plan = _plan_with_atoms(position, charge)
  rho = density_grid.contiguous()
  if split:
    return tuple(plan.potential(rho, xc_type, kohn_sham, part) for part in (HARTREE, EXTERNAL, XC))
  return plan.potential(rho, xc_type, kohn_sham, HARTREE | EXTERNAL | XC)


def hartree(density_grid, position, charge, kohn_sham: bool = False):
  """Real-space Hartree potential of a density (ifftn of potential.hartree_reciprocal,
  jrystal/_src/potential.py:24-77)."""
  plan = _plan_with_atoms(position, charge)
  return plan.potential(density_grid.contiguous(), 'lda_x', kohn_sham, HARTREE)


def external(position, charge, g_vector_grid=None, vol=None):
  """jrystal/_src/potential.py:169-200: real-space external potential (real part)."""
  import torch
  del g_vector_grid, vol
  plan = _plan_with_atoms(position, charge)
  zero = torch.zeros((plan.ns, plan.nx, plan.ny, plan.nz), dtype=torch.float64, device=plan.tdev)
  return plan.potential(zero, 'lda_x', False, EXTERNAL)[0]
