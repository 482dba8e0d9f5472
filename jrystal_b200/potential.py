"""Potential module: same names as jrystal.potential (jrystal/_src/potential.py), real-space
results from the fused grid kernels, with the reference's semantics (Hartree potential halved
and xc = eps_xc unless kohn_sham; see include/jrystal_b200.h jrb_potential)."""
from .context import current_plan
from .energy import _plan_with_atoms

import numpy as np

HARTREE, EXTERNAL, XC = 1, 2, 4


def hartree_reciprocal(density_grid_reciprocal, g_vector_grid, kohn_sham: bool = False):
  """jrystal/_src/potential.py:24-77: V_H(G) = 4 pi n(G) / |G|^2 (halved unless kohn_sham), V_H(0)
  = 0, n = sum over the spin axis.  A stand-alone G-space helper of the reference's API (the
  evaluation itself fuses this into `Plan.grid_potential`): elementwise on whatever holds the
  density, a torch tensor on any device or a numpy array."""
  g = np.asarray(g_vector_grid, dtype=np.float64)
  if density_grid_reciprocal.ndim != g.shape[-1] + 1:
    raise ValueError('density_grid_reciprocal must contains spin axis')
  g2 = np.sum(g * g, axis=-1)
  g2[(0,) * g2.ndim] = 1.0
  factor = (4 * np.pi if kohn_sham else 2 * np.pi) / g2
  factor[(0,) * g2.ndim] = 0.0
  rho = density_grid_reciprocal.sum(0)
  if isinstance(rho, np.ndarray):
    return rho * factor
  import torch
  return rho * torch.from_numpy(factor).to(rho.device)


def external_reciprocal(position, charge, g_vector_grid, vol):
  """jrystal/_src/potential.py:121-166: V_ext(G) = -(N / Omega) 4 pi sum_a Z_a e^{-i G.R_a} /
  (|G|^2 + 1e-10), V_ext(0) = 0 (numpy, set-up time; jrb_set_atoms builds the same table on the
  device)."""
  g = np.asarray(g_vector_grid, dtype=np.float64)
  pos = np.asarray(position, dtype=np.float64).reshape(-1, 3)
  z = np.asarray(charge, dtype=np.float64).reshape(-1)
  g2 = np.sum(g * g, axis=-1)
  radial = 4 * np.pi / (g2 + 1e-10)
  radial[(0,) * g2.ndim] = 0.0
  out = np.zeros(g2.shape, dtype=np.complex128)
  for r_a, z_a in zip(pos, z):
    out += z_a * np.exp(-1j * (g @ r_a))
  return -out * radial * (g2.size / vol)


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


def hartree(density_grid_reciprocal, g_vector_grid=None, kohn_sham: bool = False):
  """jrystal/_src/potential.py:80-118: real-space Hartree potential (x, y, z) of a reciprocal-space
  density that carries its spin axis (real part of ifftn(hartree_reciprocal(sum over spin)))."""
  from .energy import _real_density
  del g_vector_grid
  if density_grid_reciprocal.ndim != 4:
    raise ValueError('density_grid_reciprocal must contains spin axis')
  plan = current_plan()
  if not plan._atoms:
    raise RuntimeError('call Plan.set_atoms (or potential.external) before potential.hartree')
  return plan.potential(_real_density(plan, density_grid_reciprocal), 'lda_x', kohn_sham, HARTREE)[0]


def external(position, charge, g_vector_grid=None, vol=None):
  """jrystal/_src/potential.py:169-200: real-space external potential (real part)."""
  import torch
  del g_vector_grid, vol
  plan = _plan_with_atoms(position, charge)
  zero = torch.zeros((plan.ns, plan.nx, plan.ny, plan.nz), dtype=torch.float64, device=plan.tdev)
  return plan.potential(zero, 'lda_x', False, EXTERNAL)[0]
