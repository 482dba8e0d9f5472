"""jrystal.kinetic (jrystal/_src/kinetic.py:21-45): the 1/2 |G + k|^2 operator grid (host,
set-up time; the kernels use the plan's sphere-only |G+k|^2 table)."""
import numpy as np


def kinetic_operator(g_vector_grid, kpts=None):
  g = np.asarray(g_vector_grid, dtype=np.float64)
  # kpts=None means the single k-point Gamma: the result keeps its leading kpt axis
  k = np.zeros((1, 3)) if kpts is None else np.asarray(kpts, dtype=np.float64)
  k = k.reshape(-1, 1, 1, 1, 3)
  return np.sum((g[None] + k)**2, axis=-1) / 2
