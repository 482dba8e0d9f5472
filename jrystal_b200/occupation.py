"""Occupation numbers f[spin, kpt, band] (inputs of the hot path).

Same names/semantics as jrystal.occupation (jrystal/_src/occupation.py): `uniform` (181-190)
and `gamma` (229-237) are parameter-free and are passed through `occupation()` unchanged
(stop_gradient, 271-274).  They are nk*nb scalars, built on the host in numpy FP64.
"""
from typing import Optional

import numpy as np


def _check_spin(num_electrons: int, spin: int):
  # jrystal/_src/utils.py:327-328
  if num_electrons % 2 != spin % 2:
    raise ValueError("spin number is not valid for the system. ")


def uniform(num_k: int, num_electrons: int, spin: int = 0, num_bands: Optional[int] = None,
            spin_restricted: bool = True) -> np.ndarray:
  num_electrons = int(num_electrons)
  num_bands = num_electrons if num_bands is None else int(num_bands)
  occ = np.zeros([2, num_k, num_bands])
  occ[0, :, :(num_electrons + spin) // 2] = 1 / num_k
  occ[1, :, :(num_electrons - spin) // 2] = 1 / num_k
  return occ.sum(axis=0, keepdims=True) if spin_restricted else occ


def gamma(num_k: int, num_electrons: int, spin: int = 0, num_bands: Optional[int] = None,
          spin_restricted: bool = True) -> np.ndarray:
  num_electrons = int(num_electrons)
  num_bands = num_electrons if num_bands is None else int(num_bands)
  occ = np.zeros([2, num_k, num_bands])
  occ[0, 0, :(num_electrons + spin) // 2] = 1
  occ[1, 0, :(num_electrons - spin) // 2] = 1
  return occ.sum(axis=0, keepdims=True) if spin_restricted else occ


def param_init(key, num_bands: int, num_electrons: int, num_kpts: int, spin: int = 0,
               method: str = "uniform", spin_restricted: bool = True):
  """jrystal/_src/occupation.py:240-258 (parameter-free methods)."""
  del key
  if method == "uniform":
    return uniform(num_kpts, num_electrons, spin, num_bands, spin_restricted)
  if method == "gamma":
    return gamma(num_kpts, num_electrons, spin, num_bands, spin_restricted)
  if method in ("idempotent", "simplex-projector"):
    raise NotImplementedError(
      f'occupation method "{method}" (trainable occupations) is not implemented yet; the '
      'fused evaluation already returns dE/d occupation for it (Plan.eval_finish)')
  raise ValueError(f"Invalid method: {method}")


def occupation(params, num_kpts: int, num_electrons: Optional[int] = None, spin: int = 0,
               method: str = "uniform", spin_restricted: bool = True):
  """jrystal/_src/occupation.py:261-278."""
  if method in ("uniform", "gamma"):
    return params
  if method in ("idempotent", "simplex-projector"):
    raise NotImplementedError(f'occupation method "{method}" is not implemented yet')
  raise ValueError(f"Invalid method: {method}")


def fermi_dirac_entropy(occ: np.ndarray, eps: float = 1e-8) -> float:
  """entropy.fermi_dirac (jrystal/_src/entropy.py:46-54)."""
  occ = np.asarray(occ)
  num_spin, num_k, _ = occ.shape
  fmax = (3 - num_spin) / num_k
  return float(-np.sum(occ * np.log(eps + occ) + (fmax - occ) * np.log(eps + fmax - occ)))
