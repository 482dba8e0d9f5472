"""Local part of a norm-conserving pseudopotential in reciprocal space
(jrystal/pseudopotential/local.py:32-136, 'sbt' method).

  V_loc(G) = N / Omega * sum_a e^{-i G.R_a} * 4 pi [ SBT_0{V_a(r) + Z_a / r}(|G|) - Z_a / |G|^2 ],
  V_loc(0) = 0

in the same convention as potential.external_reciprocal, so it goes to the device through
`Plan.set_external_potential` (jrb_set_external_potential) and the external-energy slot of the
evaluation is E_loc (`energy_local` = reciprocal_braket(V_loc, rho_hat), local.py:166-187).  Atoms
of one species share one radial transform (the reference recomputes it per atom)."""
import numpy as np
from scipy.interpolate import CubicSpline

from .beta import max_radius, sbt_numerical


def _atomic_form_factors(g, r_grid, local_potential_grid, local_potential_charge):
  """u_a(G) = 4 pi [SBT_0{V_a + Z_a / r}(|G|) - Z_a / |G|^2], u_a(0) = 0: one real (x, y, z) array per
  atom; atoms of one species share it."""
  g_radius = np.sqrt((g * g).sum(-1))
  kmax = max_radius(g)
  g2 = g_radius ** 2
  g2[0, 0, 0] = 1e20                      # the reference's 1e10 ** 2 guard; the bin is zeroed below
  radial, out = {}, []
  for r, v_r, z in zip(r_grid, local_potential_grid, local_potential_charge):
    r, v_r = np.asarray(r), np.asarray(v_r)
    key = (r.shape[0], float(z), float(r[0]), float(r[-1]), float(v_r[0]), float(v_r.sum()))
    if key not in radial:
      kk, f_k = sbt_numerical(r, (v_r + z / r)[None], 0, kmax)
      short = 4 * np.pi * CubicSpline(kk, f_k[0])(g_radius)     # V + Z/r: short ranged
      atom = short - 4 * np.pi * z / g2                          # minus the Coulomb tail
      atom[0, 0, 0] = 0.0
      radial[key] = atom
    out.append(radial[key])
  return out


def potential_local_reciprocal(positions, g_vector_grid, r_grid, local_potential_grid,
                               local_potential_charge, vol: float,
                               fourier_transform_method: str = 'sbt') -> np.ndarray:
  """(x, y, z) complex128.  Lists hold one entry per atom, as in the reference."""
  if fourier_transform_method != 'sbt':
    raise ValueError(f'Invalid fourier transform method: {fourier_transform_method}. '
                     "Only 'sbt' (the reference's default) is implemented.")
  g = np.asarray(g_vector_grid, dtype=np.float64)
  pos = np.asarray(positions, dtype=np.float64).reshape(-1, 3)
  v_g = np.zeros(g.shape[:-1], dtype=np.complex128)
  u = _atomic_form_factors(g, r_grid, local_potential_grid, local_potential_charge)
  for a in range(pos.shape[0]):
    v_g += u[a] * np.exp(-1j * (g @ pos[a]))
  return v_g * (v_g.size / vol)


def energy_local_position_gradient(reciprocal_density_grid, positions, g_vector_grid, r_grid,
                                   local_potential_grid, local_potential_charge) -> np.ndarray:
  """dE_loc / dR_a, (atom, 3): what jax.grad of energy_local gives the reference through the
  structure factor of local.py:120-126.  E_loc = Re sum_G conj(V_loc) rho_hat Omega / N^2 and
  V_loc = (N / Omega) sum_a u_a(G) e^{-i G.R_a}  =>  dE/dR_a = Re sum_G i G u_a e^{+i G.R_a} rho_hat / N."""
  g = np.asarray(g_vector_grid, dtype=np.float64)
  pos = np.asarray(positions, dtype=np.float64).reshape(-1, 3)
  rho = np.asarray(reciprocal_density_grid)
  if rho.ndim == 4:
    rho = rho.sum(0)
  u = _atomic_form_factors(g, r_grid, local_potential_grid, local_potential_charge)
  out = np.zeros((pos.shape[0], 3))
  for a in range(pos.shape[0]):
    w = u[a] * np.exp(1j * (g @ pos[a])) * rho
    out[a] = np.real(1j * np.einsum('xyzd,xyz->d', g, w)) / rho.size
  return out


def energy_local(reciprocal_density_grid, potential_local_grid_reciprocal, vol: float) -> float:
  """Host-side twin of local.py:166-187 (reciprocal_braket, braket.py:58-72): for checks; in the
  drivers the device computes it as the external-energy slot."""
  rho = np.asarray(reciprocal_density_grid)
  v = np.asarray(potential_local_grid_reciprocal)
  n = v.size
  if rho.ndim == v.ndim + 1:
    rho = rho.sum(0)
  return float(np.real(np.sum(np.conj(v) * rho)) * vol / n / n)


def hamiltonian_local(wave_grid, potential_local_grid_reciprocal, vol: float):
  """jrystal/pseudopotential/local.py:134-161 on a dense psi(r) the caller holds (diagnostic; the
  evaluation applies V_loc inside the H-apply): <psi_i| v_loc(r) |psi_j> vol / N with
  v_loc(r) = ifftn(V_loc(G)), (spin, kpt, band, band).  torch tensor in -> torch tensor out."""
  import torch
  is_torch = isinstance(wave_grid, torch.Tensor)
  w = wave_grid if is_torch else torch.from_numpy(np.asarray(wave_grid))
  v = potential_local_grid_reciprocal
  v = v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v, dtype=np.complex128))
  v_r = torch.fft.ifftn(v.to(w.device), dim=(-3, -2, -1))
  n = w.shape[-3] * w.shape[-2] * w.shape[-1]
  h = torch.einsum('skaxyz,xyz,skbxyz->skab', w.conj(), v_r.to(w.dtype), w) * (float(vol) / n)
  return h if is_torch else h.numpy()
