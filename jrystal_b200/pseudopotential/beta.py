"""Radial Fourier (spherical Bessel) transforms of the pseudopotential's radial functions, with the
reference's quadrature and interpolation so that the numbers agree with it:

  F_l(k) = sum_i f(r_i) r_i^2 j_l(k r_i) (r_{i+1} - r_i)         (jrystal/sbt/sbt_numerical.py:40-77:
           left Riemann sum on the UPF mesh, last point dropped)
  on k = linspace(1e-4, kmax, 2 * len(r)), then a not-a-knot cubic spline in k
           (jrystal/pseudopotential/beta.py:69-76, local.py:80-86).

Unlike the reference, which evaluates the spline on the whole (kpt, x, y, z) box, the projector
transform here is evaluated only where it is used: on the cut-off sphere of each k-point."""
from typing import Sequence, Tuple

import numpy as np
from scipy.interpolate import CubicSpline
from scipy.special import spherical_jn

K_MIN = 1e-4  # sbt_numerical.py:45


def sbt_numerical(r_grid, f_grid, l, kmax: float, delta_r=None) -> Tuple[np.ndarray, np.ndarray]:
  """(k grid (2 nr,), F (nf, 2 nr)); l is one int or one int per row of f_grid.  `delta_r`
  overrides the quadrature weights (e.g. the UPF's PP_RAB, as the reference's sbt_test does);
  default: forward differences of r with the last point dropped."""
  r = np.asarray(r_grid, dtype=np.float64)
  f = np.atleast_2d(np.asarray(f_grid, dtype=np.float64))
  ls = [int(l)] * f.shape[0] if np.ndim(l) == 0 else [int(v) for v in l]
  if len(ls) != f.shape[0]:
    raise ValueError('The length of l must be the same as the batch dimension of f_grid')
  if delta_r is None:
    dr = np.zeros_like(r)
    dr[:-1] = r[1:] - r[:-1]
  else:
    dr = np.asarray(delta_r, dtype=np.float64)
  k = np.linspace(K_MIN, kmax, 2 * r.shape[0])
  kr = k[:, None] * r[None, :]
  w = r * r * dr
  out = np.empty((f.shape[0], k.shape[0]))
  cache = {}
  for i, li in enumerate(ls):
    if li not in cache:
      cache[li] = spherical_jn(li, kr)
    out[i] = cache[li] @ (f[i] * w)
  return k, out


def max_radius(g_vector_grid, kpts=None) -> float:
  """max |G + k| over the WHOLE box and all k-points: the reference's spline knots run up to this
  value (beta.py:66-71), so it is part of the arithmetic."""
  g = np.asarray(g_vector_grid, dtype=np.float64).reshape(-1, 3)
  if kpts is None:
    return float(np.sqrt((g * g).sum(-1).max()))
  best = 0.0
  for k in np.asarray(kpts, dtype=np.float64).reshape(-1, 3):
    d = g + k
    best = max(best, float((d * d).sum(-1).max()))
  return float(np.sqrt(best))


def beta_sbt_sphere(r_grid, nonlocal_beta_grid, nonlocal_angular_momentum, radius, kmax: float):
  """beta_l(|G + k|) of ONE atom's projectors at the given radii.
  radius: (kpt, g) = |G + k| on the sphere; returns (kpt, beta, g) (beta.py:27-77 restricted to
  the sphere)."""
  r = np.asarray(r_grid, dtype=np.float64)
  b = np.asarray(nonlocal_beta_grid, dtype=np.float64)
  if r[0] == 0:
    r, b = r[1:], b[:, 1:]
  k, beta_k = sbt_numerical(r, b, list(nonlocal_angular_momentum), kmax)
  vals = CubicSpline(k, beta_k, axis=1)(np.asarray(radius, dtype=np.float64))  # (beta, kpt, g)
  return np.swapaxes(vals, 0, 1)


def beta_sbt_grid(r_grid: Sequence, nonlocal_beta_grid: Sequence, nonlocal_angular_momentum: Sequence,
                  g_vector_grid, kpts=None):
  """Dense drop-in for the reference's beta_sbt_grid (beta.py:80-120): a list (one entry per atom)
  of (kpt, beta, x, y, z) arrays.  Kept for API parity and tests; the drivers use
  `beta_sbt_sphere`."""
  g = np.asarray(g_vector_grid, dtype=np.float64)
  ks = np.zeros((1, 3)) if kpts is None else np.asarray(kpts, dtype=np.float64).reshape(-1, 3)
  kmax = max_radius(g, None if kpts is None else ks)
  radius = np.sqrt(((g[None] + ks[:, None, None, None, :]) ** 2).sum(-1))
  shape = radius.shape
  return [beta_sbt_sphere(r, b, l, radius.reshape(shape[0], -1), kmax).reshape(shape[0], -1, *shape[1:])
          for r, b, l in zip(r_grid, nonlocal_beta_grid, nonlocal_angular_momentum)]
