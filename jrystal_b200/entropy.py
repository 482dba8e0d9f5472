"""jrystal.entropy (jrystal/_src/entropy.py:19-54): Fermi-Dirac entropy of the occupations
f[spin, kpt, band], with f_max = (3 - num_spin) / num_k per state.  nk*nb scalars: evaluated where
the occupations live (numpy array -> float, torch tensor -> 0-d tensor that keeps its graph, which
is what the energy driver differentiates for the -T S term of the free energy)."""
from .occupation import fermi_dirac_entropy, fermi_dirac_entropy_torch

__all__ = ['fermi_dirac']


def fermi_dirac(occupation, eps: float = 1e-8):
  try:
    import torch
    if isinstance(occupation, torch.Tensor):
      return fermi_dirac_entropy_torch(occupation, eps)
  except ImportError:  # pragma: no cover
    pass
  return fermi_dirac_entropy(occupation, eps)
