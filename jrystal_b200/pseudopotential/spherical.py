"""Directions and real spherical harmonics of the projector set-up, with the conventions of
jrystal/pseudopotential/spherical.py:33-89 (theta = azimuth in [0, 2 pi), phi = polar angle; the
zero vector maps to r = eps, theta = 0, phi = pi / 2, i.e. the +x direction).  numpy/scipy, set-up
time only."""
import numpy as np
from scipy.special import sph_harm_y


def cartesian_to_spherical(x, eps: float = 1e-10) -> np.ndarray:
  """(..., 3) Cartesian -> (..., 3) (r, theta, phi) (spherical.py:33-63)."""
  x = np.asarray(x, dtype=np.float64)
  r = np.linalg.norm(x, axis=-1)
  r = np.where(r == 0., eps, r)
  phi = np.arccos(np.clip(x[..., 2] / r, -1.0, 1.0))
  theta = np.mod(np.arctan2(x[..., 1], x[..., 0]) + 2 * np.pi, 2 * np.pi)
  return np.stack((r, theta, phi), axis=-1)


def batch_sph_harm(l: int, theta, phi) -> np.ndarray:
  """Complex Y_l^m(polar = phi, azimuth = theta), m = -l..l on the last axis (spherical.py:92-121)."""
  m = np.arange(-int(l), int(l) + 1)
  return sph_harm_y(int(l), m, np.asarray(phi)[..., None], np.asarray(theta)[..., None])


def batch_sph_harm_real(l: int, theta, phi) -> np.ndarray:
  """Real harmonics, m = -l..l on the last axis (spherical.py:66-89):
     m > 0: sqrt(2) (-1)^m Re Y_l^m ;  m = 0: Re Y_l^0 ;  m < 0: sqrt(2) (-1)^m Im Y_l^|m|."""
  y = batch_sph_harm(l, theta, phi)
  m = np.arange(-int(l), int(l) + 1)
  par = np.where(m % 2 == 0, 1.0, -1.0)
  out = np.where(m >= 0, np.sqrt(2.0) * par * y.real, np.sqrt(2.0) * par * (np.conj(y) * par).imag)
  out[..., int(l)] = y[..., int(l)].real
  return out
