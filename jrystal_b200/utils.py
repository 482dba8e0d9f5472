"""jrystal.utils (jrystal/utils/__init__.py: safe_real, wave_to_density,
wave_to_density_reciprocal, expand_coefficient, squeeze_coefficient) on the current plan.

The reference's functions act on dense (spin, kpt, band, x, y, z) arrays; the hot path never
builds those, so these are the API-parity routes: jrb_expand / jrb_squeeze for
the sphere <-> box maps, and elementwise torch on whatever dense psi(r) the caller already holds.
`vmapstack` (a jax.vmap helper) has no counterpart."""
import numpy as np
import torch

from .context import current_plan
from .grid import fft_factor  # noqa: F401  (utils.py:227-254)

__all__ = ['safe_real', 'absolute_square', 'wave_to_density', 'wave_to_density_reciprocal',
           'expand_coefficient', 'squeeze_coefficient', 'volume', 'fft_factor', 'check_spin_number']


def check_spin_number(num_electrons: int, spin: int) -> None:
  """utils.py:311-328."""
  if num_electrons % 2 != spin % 2:
    raise ValueError("spin number is not valid for the system. ")


def safe_real(array, tol: float = 1e-8):
  """utils.py:25-59: the real part if every imaginary part is below `tol`, else ValueError."""
  if isinstance(array, torch.Tensor):
    if not array.is_complex():
      return array
    if bool((array.imag.abs() <= tol).all()):
      return array.real
    raise ValueError("Array has non-zero imaginary part")
  array = np.asarray(array)
  if not np.iscomplexobj(array):
    return array
  if np.allclose(array.imag, 0, atol=tol):
    return array.real
  raise ValueError("Array has non-zero imaginary part")


def absolute_square(array):
  """utils.py:108-130."""
  if isinstance(array, torch.Tensor):
    return (array.real * array.real + array.imag * array.imag) if array.is_complex() else array * array
  return np.abs(np.asarray(array)) ** 2


def volume(cell_vectors) -> float:
  """utils.py:133-158."""
  return float(abs(np.linalg.det(np.asarray(cell_vectors, dtype=np.float64))))


def wave_to_density(wave_grid, occupation=None):
  """utils.py:161-197 on a dense psi(r) the caller holds (diagnostics; the evaluation itself
  accumulates rho inside the last FFT pass, pw.density_grid)."""
  dens = absolute_square(wave_grid)
  if occupation is None:
    return dens
  if not isinstance(occupation, torch.Tensor):
    occupation = torch.as_tensor(np.asarray(occupation), dtype=torch.float64)
  if tuple(occupation.shape) != tuple(dens.shape[:3]):
    raise ValueError(f"wave_grid's shape ({tuple(wave_grid.shape)}) and occupation's shape "
                     f"({tuple(occupation.shape)}) cannot align.")
  occ = occupation.to(dens.device, dens.dtype)[..., None, None, None]
  return (dens * occ).sum(dim=(1, 2))


def wave_to_density_reciprocal(wave_grid, occupation=None):
  """utils.py:200-224: fftn of wave_to_density over the last three axes."""
  dens = wave_to_density(wave_grid, occupation)
  plan = current_plan()
  if not dens.is_cuda:
    dens = dens.to(plan.tdev)
  return plan.fft3d(torch.complex(dens, torch.zeros_like(dens)).contiguous(), inverse=False)


def _check_mask(plan, mask):
  if not np.array_equal(np.asarray(mask).astype(bool), plan.mask.astype(bool)):
    raise ValueError('mask differs from the mask of the current plan')


def expand_coefficient(coeff_compact, mask):
  """utils.py:257-281: (spin, kpt, gpt, band) on the sphere -> zero-padded
  (spin, kpt, band, x, y, z) (jrb_expand)."""
  plan = current_plan()
  _check_mask(plan, mask)
  return plan.expand(coeff_compact.contiguous())


def squeeze_coefficient(coeff, mask):
  """utils.py:284-308: the inverse map (jrb_squeeze)."""
  plan = current_plan()
  _check_mask(plan, mask)
  return plan.squeeze(coeff.contiguous())
