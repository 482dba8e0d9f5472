"""Occupation numbers f[spin, kpt, band] (inputs of the hot path).

Same names/semantics as jrystal.occupation (jrystal/_src/occupation.py): `uniform` (181-190)
and `gamma` (229-237) are parameter-free and are passed through `occupation()` unchanged
(stop_gradient, 271-274): nk*nb scalars, built on the host in numpy FP64.  The trainable schemes
`simplex-projector` (281-366, the default of the shipped config.yaml) and `idempotent` (28-80) are
nk*nb-sized torch computations whose autograd graph the energy driver chains with dE/d occupation
from jrb_eval_finish (they are not on the timed path).
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


def _capped_simplex(x, total: float):
  """Euclidean projection of x in [0, 1]^n onto {0 <= y <= 1, sum(y) = total}, the `proj` of
  jrystal/_src/occupation.py:299-338.  Like the reference it shifts by one multiplier and clips on
  ONE side: y = max(x - lam, 0) when sum(x) > total, y = min(x + lam, 1) otherwise (the other bound
  holds because x already lies in [0, 1]).  lam comes from the sort/cumsum rule of simplex
  projection; torch autograd differentiates through it (piecewise linear in x)."""
  import torch

  def push_down(v, t):
    u, _ = torch.sort(v, descending=True)
    css = torch.cumsum(u, 0) - t
    j = torch.arange(1, v.numel() + 1, dtype=v.dtype, device=v.device)
    k = int(torch.nonzero(u - css / j > 0).max().item())
    lam = css[k] / (k + 1)
    return torch.clamp(v - lam, min=0.0)

  n = x.numel()
  # edges the sort/cumsum rule has no index for: an empty channel and a completely filled one
  # (empty_bands = 0, or num_electrons == spin); the projection is the constant vertex
  if total <= 0:
    return x * 0.0
  if total >= n:
    return x * 0.0 + 1.0
  if float(x.detach().sum()) > total:
    return push_down(x, total)
  return 1.0 - push_down(1.0 - x, n - total)


def proj(x, sum):  # noqa: A002 (the reference's argument name)
  """jrystal/_src/occupation.py:295-338 under its own name: projection of x (1-D torch tensor in
  [0, 1]) onto {0 <= y <= 1, sum(y) = sum}."""
  return _capped_simplex(x, float(sum))


def simplex_projector_init(num_bands: int, num_kpts: int) -> dict:
  """jrystal/_src/occupation.py:281-296: logits (arange(n) - n // 2) * 0.1, both spins alike."""
  import torch
  n = num_bands * num_kpts
  dev = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else 'cpu'
  v = ((torch.arange(n, dtype=torch.float64) - n // 2) * 0.1).reshape(num_kpts, num_bands)
  return {'param_up': v.clone().to(dev).requires_grad_(True),
          'param_down': v.clone().to(dev).requires_grad_(True)}


def simplex_projector(params: dict, num_electrons: int, spin: int = 0,
                      spin_restricted: bool = True):
  """jrystal/_src/occupation.py:341-366: sigmoid of the logits, projected onto the capped simplex
  with sum = electrons-of-that-spin * nk, divided by nk.  Returns a torch tensor (spin, kpt, band)
  that carries its autograd graph."""
  import torch
  num_kpts, num_bands = params['param_up'].shape

  def one(p, m):
    x = torch.sigmoid(p).reshape(-1)
    return (_capped_simplex(x, float(m)) / num_kpts).reshape(num_kpts, num_bands)

  up = one(params['param_up'], (num_electrons + spin) // 2 * num_kpts)
  down = one(params['param_down'], (num_electrons - spin) // 2 * num_kpts)
  occ = torch.stack([up, down], dim=0)
  return occ.sum(dim=0, keepdim=True) if spin_restricted else occ


def idempotent_param_init(key, num_bands: int, num_electrons: int, num_kpts: int, spin: int = 0,
                          spin_restricted: bool = True) -> dict:
  """jrystal/_src/occupation.py:28-52: real (nb nk, n_elec nk) matrices ~ U[0, 1)."""
  import torch
  _check_spin(num_electrons, spin)
  rng = key if isinstance(key, np.random.Generator) else np.random.default_rng(key)
  dev = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else 'cpu'

  def mat(ne):
    w = torch.from_numpy(rng.random((num_bands * num_kpts, ne * num_kpts))).to(dev)
    return {'w_re': w.requires_grad_(True)}

  up = mat((num_electrons + spin) // 2)
  if spin_restricted:
    return {'param_up': up, 'param_down': up}
  return {'param_up': up, 'param_down': mat((num_electrons - spin) // 2)}


def idempotent(params: dict, num_kpts: int, spin_restricted: bool = True):
  """jrystal/_src/occupation.py:55-80: occ = diag(U U^T) / nk with U the Q factor of the real
  parameter matrix (an idempotent density matrix in the band x k basis)."""
  import torch

  def o(p):
    u = torch.linalg.qr(p['w_re'], mode='reduced')[0]
    nb = u.shape[0] // num_kpts
    return (u * u).sum(dim=1).reshape(num_kpts, nb)

  up, down = o(params['param_up']), o(params['param_down'])
  if spin_restricted:
    return (up + down)[None] / num_kpts
  return torch.stack([up, down], dim=0) / num_kpts


def param_init(key, num_bands: int, num_electrons: int, num_kpts: int, spin: int = 0,
               method: str = "simplex-projector", spin_restricted: bool = True):
  """jrystal/_src/occupation.py:240-258.  Parameter-free methods return the occupations as a numpy
  array; the trainable ones a dict of torch leaves (requires_grad) on the current device."""
  if method == "uniform":
    return uniform(num_kpts, num_electrons, spin, num_bands, spin_restricted)
  if method == "gamma":
    return gamma(num_kpts, num_electrons, spin, num_bands, spin_restricted)
  if method == "idempotent":
    return idempotent_param_init(key, num_bands, num_electrons, num_kpts, spin, spin_restricted)
  if method == "simplex-projector":
    return simplex_projector_init(num_bands, num_kpts)
  raise ValueError(f"Invalid method: {method}")


def occupation(params, num_kpts: int, num_electrons: Optional[int] = None, spin: int = 0,
               method: str = "simplex-projector", spin_restricted: bool = True):
  """jrystal/_src/occupation.py:261-278."""
  if method in ("uniform", "gamma"):
    return params
  if method == "idempotent":
    return idempotent(params, num_kpts, spin_restricted)
  if method == "simplex-projector":
    return simplex_projector(params, num_electrons, spin, spin_restricted)
  raise ValueError(f"Invalid method: {method}")


def trainable(method: str) -> bool:
  return method in ("idempotent", "simplex-projector")


def fermi_dirac_entropy_torch(occ, eps: float = 1e-8):
  """entropy.fermi_dirac (jrystal/_src/entropy.py:46-54) on a torch tensor (keeps the graph)."""
  import torch
  num_spin, num_k, _ = occ.shape
  fmax = (3 - num_spin) / num_k
  return -torch.sum(occ * torch.log(eps + occ) + (fmax - occ) * torch.log(eps + fmax - occ))


def fermi_dirac_entropy(occ: np.ndarray, eps: float = 1e-8) -> float:
  """entropy.fermi_dirac (jrystal/_src/entropy.py:46-54)."""
  occ = np.asarray(occ)
  num_spin, num_k, _ = occ.shape
  fmax = (3 - num_spin) / num_k
  return float(-np.sum(occ * np.log(eps + occ) + (fmax - occ) * np.log(eps + fmax - occ)))
