"""The norm-conserving bundle the drivers use (jrystal/pseudopotential/normcons.py:15-37 re-exports
the local and non-local pieces under one name) plus `attach`, which puts both parts on a plan."""
import numpy as np
import torch

from .beta import beta_sbt_grid
from .local import energy_local, potential_local_reciprocal
from .nloc import energy_nonlocal, hamiltonian_nonlocal, potential_nonlocal_psi_sphere

__all__ = ['beta_sbt_grid', 'potential_local_reciprocal', 'energy_local',
           'potential_nonlocal_psi_sphere', 'hamiltonian_nonlocal', 'energy_nonlocal', 'attach',
           'projectors', 'set_projectors']


def projectors(plan, pseudopot, g_vector_grid, kpts=None, positions=None, kmax=None) -> np.ndarray:
  """Sphere projectors (kpt, proj, g) for the plan's mask and the given k-points."""
  kpts = plan.kpts if kpts is None else np.asarray(kpts, dtype=np.float64).reshape(-1, 3)
  pos = pseudopot.positions if positions is None else positions
  return potential_nonlocal_psi_sphere(pos, g_vector_grid, kpts, plan.mask.astype(bool),
                                       pseudopot.r_grid, pseudopot.nonlocal_beta_grid,
                                       pseudopot.nonlocal_angular_momentum,
                                       pseudopot.nonlocal_d_matrix, kmax=kmax)


def set_projectors(plan, phi: np.ndarray) -> None:
  plan.set_nonlocal(torch.from_numpy(np.ascontiguousarray(phi)).to(plan.tdev)
                    if phi.shape[1] else None)


def attach(plan, pseudopot, g_vector_grid, kpts=None, positions=None, kmax=None):
  """Compute V_loc(G) and the sphere projectors for the plan's k-points and hand them to the device:
  afterwards the external slot of the evaluation is E_loc and E_nl joins the kinetic slot, and
  gradients / H-apply / band expectations carry both.
  `kpts`: the plan's own k-points (its shard under k-sharding).  Returns (v_loc, phi) as numpy."""
  kpts = plan.kpts if kpts is None else np.asarray(kpts, dtype=np.float64).reshape(-1, 3)
  pos = pseudopot.positions if positions is None else positions
  v_loc = potential_local_reciprocal(pos, g_vector_grid, pseudopot.r_grid,
                                     pseudopot.local_potential_grid,
                                     pseudopot.local_potential_charge, plan.vol)
  phi = projectors(plan, pseudopot, g_vector_grid, kpts, pos, kmax)
  plan.set_external_potential(torch.from_numpy(np.ascontiguousarray(v_loc)).to(plan.tdev))
  set_projectors(plan, phi)
  return v_loc, phi
