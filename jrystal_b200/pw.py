"""Plane-wave module: same names and argument meaning as jrystal.pw (jrystal/_src/pw.py), on the
CUDA kernels.  Coefficients travel as a `Coefficients` handle (compact sphere layout + plan)
instead of the reference's zero-padded (spin, kpt, band, x, y, z) array, which is only built on
request (`.dense()`); every function also accepts that dense array where the reference does.
"""
from typing import Optional, Union

import numpy as np
import torch

from .context import current_plan


class Coefficients:
  """Orthonormal plane-wave coefficients of pw.coeff: Q[s, k, g, b] on the cut-off sphere."""

  def __init__(self, plan, q, r=None):
    self.plan, self.q, self.r = plan, q, r

  @property
  def shape(self):  # the shape the reference's array has
    p = self.plan
    return (p.ns, p.nk, p.nb, p.nx, p.ny, p.nz)

  def dense(self) -> torch.Tensor:
    """utils.expand_coefficient (jrystal/_src/utils.py:277-281)."""
    return self.plan.expand(self.q)


def _as_coeff(coeff) -> Coefficients:
  if isinstance(coeff, Coefficients):
    return coeff
  plan = current_plan()
  if not isinstance(coeff, torch.Tensor):
    coeff = torch.as_tensor(np.asarray(coeff), dtype=torch.complex128).to(plan.tdev)
  if coeff.ndim != 6:
    raise ValueError(f'coeff must have 6 axes (spin, kpt, band, x, y, z), got {coeff.ndim}')
  return Coefficients(plan, plan.squeeze(coeff.contiguous()))


def param_init(key, num_bands: int, num_kpts: int, freq_mask, spin_restricted: bool = True,
               sharding=None) -> dict:
  """jrystal/_src/pw.py:29-91: {'w_re','w_im'} ~ U[0,1), shape (ns, nk, ng, nb) float64.
  `key` is an integer seed or a numpy Generator (JAX's threefry stream is not reproducible
  without jax; parity is defined on given parameters)."""
  del sharding
  ns = 1 if spin_restricted else 2
  ng = int(np.sum(np.asarray(freq_mask)))
  rng = key if isinstance(key, np.random.Generator) else np.random.default_rng(key)
  shape = (ns, int(num_kpts), ng, int(num_bands))
  dev = torch.device('cuda', torch.cuda.current_device())
  return {'w_re': torch.from_numpy(rng.random(shape)).to(dev),
          'w_im': torch.from_numpy(rng.random(shape)).to(dev)}


def coeff(pw_param: Union[dict, tuple, list], freq_mask, sharding=None) -> Coefficients:
  """jrystal/_src/pw.py:94-137: QR-orthonormalise w_re + i w_im (Cholesky-QR2 on DMMA)."""
  del sharding
  plan = current_plan()
  if not np.array_equal(np.asarray(freq_mask).astype(bool), plan.mask.astype(bool)):
    raise ValueError('freq_mask differs from the mask of the current plan')
  if isinstance(pw_param, dict):
    w_re, w_im = pw_param['w_re'], pw_param['w_im']
  else:
    w_re, w_im = pw_param
  if torch.is_grad_enabled() and (getattr(w_re, 'requires_grad', False)
                                  or getattr(w_im, 'requires_grad', False)):
    # under torch.autograd the handle carries a differentiable Q (backward = the QR adjoint)
    from .autograd import orthonormal_coefficients
    q, r = orthonormal_coefficients(w_re, w_im, plan)
  else:
    q, r = plan.qr_fwd(w_re, w_im)
  return Coefficients(plan, q, r)


def wave_grid(coeff, vol: float) -> torch.Tensor:
  """jrystal/_src/pw.py:140-211: psi = ifftn(coeff) * N / sqrt(vol), dense."""
  c = _as_coeff(coeff)
  _check_vol(c.plan, vol)
  return c.plan.wave_grid(c.q)


def density_grid(coeff, vol: float, occupation: Optional[torch.Tensor] = None) -> torch.Tensor:
  """jrystal/_src/pw.py:214-284: rho[s, x, y, z] = sum_kb occ |psi|^2 (fused scatter + IFFT +
  accumulation; psi(r) never reaches global memory)."""
  c = _as_coeff(coeff)
  _check_vol(c.plan, vol)
  if occupation is None:
    # pw.py:273-284: without occupation the reference returns the per-orbital densities
    # |psi_skb(r)|^2, shape (spin, kpt, band, x, y, z) -- the dense diagnostic, M*N numbers
    psi = torch.view_as_real(c.plan.wave_grid(c.q))
    return psi[..., 0] ** 2 + psi[..., 1] ** 2
  p = c.plan
  occ = _occ(p, occupation)
  return p.density(c.q, occ)


def density_grid_reciprocal(coeff, vol: float, occupation=None) -> torch.Tensor:
  """jrystal/_src/pw.py:287-334: fftn(density_grid)."""
  c = _as_coeff(coeff)
  dens = density_grid(c, vol, occupation)
  if occupation is None:  # per-orbital: batched dense transform over (spin, kpt, band)
    return c.plan.fft3d(torch.complex(dens, torch.zeros_like(dens)), inverse=False)
  return c.plan.density_reciprocal(dens)


def _occ(plan, occupation):
  if not isinstance(occupation, torch.Tensor):
    occupation = torch.as_tensor(np.asarray(occupation), dtype=torch.float64)
  occupation = occupation.to(plan.tdev, torch.float64).contiguous()
  if tuple(occupation.shape) != (plan.ns, plan.nk, plan.nb):
    # pw.py:279-283
    raise ValueError(
      f'Occupation should have shape {(plan.ns, plan.nk, plan.nb)}, got {tuple(occupation.shape)}')
  return occupation


def _check_vol(plan, vol):
  if abs(float(vol) - plan.vol) > 1e-9 * plan.vol:
    raise ValueError(f'vol={vol} differs from the cell volume of the current plan ({plan.vol})')


# -- point evaluations (pw.py:337-511 of the reference: diagnostics / test helpers) ---------------
# Direct plane-wave sums over the cut-off sphere at ONE position r (the reference sums over the
# whole zero-padded box); O(M ng) elementwise torch work on the device that holds the coefficients.

def _point_sums(r, coeff, cell_vectors, g_vector_grid, want_gradient):
  from .grid import g_vectors
  c = _as_coeff(coeff)
  p = c.plan
  r = np.asarray(r.detach().cpu() if isinstance(r, torch.Tensor) else r, dtype=np.float64).reshape(-1)
  if r.shape != (3,):
    raise ValueError('r must have shape (3,)')
  _check_vol(p, abs(np.linalg.det(np.asarray(cell_vectors, dtype=np.float64))))
  if g_vector_grid is None:
    g_vector_grid = g_vectors(cell_vectors, [p.nx, p.ny, p.nz])
  g = np.asarray(g_vector_grid, dtype=np.float64)[p.mask.astype(bool)]            # (ng, 3)
  q = c.q
  phase = torch.from_numpy(np.exp(1j * (g @ r))).to(q.device)                     # (ng,)
  a = torch.einsum('skgb,g->skb', q, phase)                                       # sqrt(vol) psi(r)
  if not want_gradient:
    return p, a, None
  ig = torch.from_numpy(1j * g).to(q.device) * phase[:, None]                     # (ng, 3)
  da = torch.einsum('skgb,gd->skbd', q, ig)                                       # sqrt(vol) grad psi
  return p, a, 2.0 * (a.conj()[..., None] * da).real                              # grad |a|^2


def wave_r(r, coeff, cell_vectors, g_vector_grid=None) -> torch.Tensor:
  """jrystal/_src/pw.py:337-381: psi[s, k, b](r) = vol^-1/2 sum_G c_G exp(i G.r) (no Bloch phase,
  as the reference)."""
  p, a, _ = _point_sums(r, coeff, cell_vectors, g_vector_grid, False)
  return a / np.sqrt(p.vol)


def density_r(r, coeff, cell_vectors, g_vector_grid=None, occupation=None) -> torch.Tensor:
  """jrystal/_src/pw.py:384-415: |psi(r)|^2 per orbital, or its occupation-weighted sum."""
  p, a, _ = _point_sums(r, coeff, cell_vectors, g_vector_grid, False)
  dens = (a.real * a.real + a.imag * a.imag) / p.vol
  return dens if occupation is None else torch.sum(dens * _occ(p, occupation).to(dens.device))


def nabla_density_r(r, coeff, cell_vectors, g_vector_grid=None, occupation=None) -> torch.Tensor:
  """jrystal/_src/pw.py:418-447 (jax.grad of density_r there): d density_r / d r, shape (3,) with
  occupations, (spin, kpt, band, 3) without."""
  p, _, grad = _point_sums(r, coeff, cell_vectors, g_vector_grid, True)
  grad = grad / p.vol
  if occupation is None:
    return grad
  return torch.sum(grad * _occ(p, occupation).to(grad.device)[..., None], dim=(0, 1, 2))


def nabla_density_grid(r, coeff, cell_vectors, g_vector_grid=None, occupation=None) -> torch.Tensor:
  """jrystal/_src/pw.py:450-511, the closed-form twin of nabla_density_r with the reference's
  normalisation: the occupation-weighted sum is divided by the volume, the per-orbital result
  (spin, kpt, band, 3) is not."""
  p, _, grad = _point_sums(r, coeff, cell_vectors, g_vector_grid, True)
  if occupation is None:
    return grad
  return torch.sum(grad * _occ(p, occupation).to(grad.device)[..., None], dim=(0, 1, 2)) / p.vol
