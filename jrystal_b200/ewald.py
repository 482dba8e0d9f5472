"""Nuclear repulsion energy of the periodic point charges (Ewald summation): the constant the
energy-mode driver adds to the electronic energy when it reports totals
(jrystal/_src/ewald.py:22-86, called once from calc/opt_utils.py:187-201).  Set-up time host code
(numpy FP64).  Same splitting as the reference (Martin, Electronic Structure, App. F.2):

  E = 1/2 sum_ij Z_i Z_j [ sum_T' erfc(eta |tau_ij + T|) / |tau_ij + T|
                           + 4 pi / Omega sum_{G != 0} exp(-G^2 / 4 eta^2) / G^2 cos(G . tau_ij) ]
      - eta / sqrt(pi) sum_i Z_i^2 - pi (sum_i Z_i)^2 / (2 eta^2 Omega)

Unlike the reference, which truncates both sums at the caller's grids, the lattice sums here run
until the neglected terms are below `tol`, so the result does not depend on eta."""
import math

import numpy as np
from scipy.special import erfc


def _shells(vectors, rmax):
  """All integer combinations n . vectors with |n . vectors| <= rmax (plus a safety shell)."""
  inv = np.linalg.inv(vectors)
  # |n_i| <= rmax * |row i of inv^T|
  nmax = np.ceil(rmax * np.linalg.norm(inv, axis=0)).astype(int) + 1
  rng = [np.arange(-n, n + 1) for n in nmax]
  n = np.stack(np.meshgrid(*rng, indexing='ij'), axis=-1).reshape(-1, 3)
  return n @ vectors


def ewald_coulomb_repulsion(positions, charges, g_vector_grid, vol=None, ewald_eta=None,
                            ewald_grid=None, tol: float = 1e-14) -> float:
  """Two call forms.
  Reference form (jrystal/_src/ewald.py:21-86): (positions, charges, g_vector_grid (x, y, z, 3), vol,
  ewald_eta, ewald_grid (num, 3) from grid.translation_vectors) -> the sums truncated at exactly
  those two grids, as the reference computes them.
  Converged form: (positions, charges, cell_vectors (3, 3), ewald_eta=None, tol=...) -> lattice
  sums run until the neglected terms are below `tol`; what the drivers use."""
  third = np.asarray(g_vector_grid, dtype=np.float64)
  if third.ndim == 4:
    if vol is None or ewald_eta is None or ewald_grid is None:
      raise TypeError('reference form needs vol, ewald_eta and ewald_grid')
    return _ewald_on_given_grids(positions, charges, third, float(vol), float(ewald_eta), ewald_grid)
  if third.shape != (3, 3):
    raise ValueError('third argument: g_vector_grid (x, y, z, 3) or cell_vectors (3, 3)')
  if ewald_grid is not None:
    raise TypeError('ewald_grid belongs to the reference form (g_vector_grid as third argument)')
  eta = vol if (vol is not None and ewald_eta is None) else ewald_eta  # positional eta after the cell
  return ewald_converged(positions, charges, third, eta, tol)


def _ewald_on_given_grids(positions, charges, g_vector_grid, vol, eta, ewald_grid) -> float:
  """The reference's truncation: real-space sum over the given translations (a term with
  |tau - T| <= 1e-9 is dropped), reciprocal sum over the given G grid without G = 0."""
  pos = np.asarray(positions, dtype=np.float64).reshape(-1, 3)
  z = np.asarray(charges, dtype=np.float64).reshape(-1)
  t = np.asarray(ewald_grid, dtype=np.float64).reshape(-1, 3)
  g = g_vector_grid.reshape(-1, 3)[1:]                           # C order: element 0 is G = 0
  g2 = np.sum(g * g, axis=1)
  tau = pos[None, :, :] - pos[:, None, :]                        # tau[i, j] = R_j - R_i
  real = np.zeros(tau.shape[:2])
  for i in range(pos.shape[0]):                                   # bounded memory: na x nt per row
    d = np.sqrt(np.sum((tau[i][:, None, :] - t[None]) ** 2, axis=-1) + 1e-20)
    d = np.where(d <= 1e-9, 1e20, d)
    real[i] = np.sum(erfc(eta * d) / d, axis=1)
  w = np.exp(-g2 / (4 * eta * eta)) / g2
  recip = np.einsum('g,ijg->ij', w, np.cos(np.einsum('ijd,gd->ijg', tau, g))) * 4 * np.pi / vol
  pair = 0.5 * z @ (real + recip) @ z
  single = -np.sum(z * z) * eta / math.sqrt(math.pi) - np.sum(z) ** 2 * math.pi / (eta * eta * vol * 2)
  return float(pair + single)


def ewald_converged(positions, charges, cell_vectors, ewald_eta: float = None,
                    tol: float = 1e-14) -> float:
  pos = np.asarray(positions, dtype=np.float64).reshape(-1, 3)
  z = np.asarray(charges, dtype=np.float64).reshape(-1)
  a = np.asarray(cell_vectors, dtype=np.float64).reshape(3, 3)
  vol = abs(np.linalg.det(a))
  b = 2.0 * np.pi * np.linalg.inv(a).T
  eta = float(ewald_eta) if ewald_eta else math.sqrt(math.pi) / vol ** (1.0 / 3.0)
  x = math.sqrt(-math.log(tol))           # erfc(x) ~ exp(-x^2) < tol
  t = _shells(a, x / eta + np.linalg.norm(a, axis=1).max())
  g = _shells(b, 2.0 * eta * x)
  g2 = np.sum(g * g, axis=1)
  g, g2 = g[g2 > 1e-12], g2[g2 > 1e-12]
  tau = pos[:, None, :] - pos[None, :, :]                         # [na, na, 3]
  d = np.linalg.norm(tau[:, :, None, :] + t[None, None], axis=-1)  # [na, na, nt]
  with np.errstate(divide='ignore', invalid='ignore'):
    real = np.where(d > 1e-9, erfc(eta * d) / d, 0.0).sum(axis=-1)
  recip = (np.exp(-g2 / (4 * eta * eta)) / g2 * np.cos(np.einsum('ijd,gd->ijg', tau, g))).sum(-1)
  pair = 0.5 * z @ (real + 4.0 * np.pi / vol * recip) @ z
  self_term = -eta / math.sqrt(math.pi) * np.sum(z * z)
  background = -math.pi * np.sum(z) ** 2 / (2.0 * eta * eta * vol)
  return float(pair + self_term + background)
