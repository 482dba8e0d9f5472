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


def potential_local_reciprocal(positions, g_vector_grid, r_grid, local_potential_grid,
                               local_potential_charge, vol: float,
                               fourier_transform_method: str = 'sbt') -> np.ndarray:
  """(x, y, z) complex128.  Lists hold one entry per atom, as in the reference."""
  if fourier_transform_method != 'sbt':
    raise ValueError(f'Invalid fourier transform method: {fourier_transform_method}. '
                     "Only 'sbt' (the reference's default) is implemented.")
  g = np.asarray(g_vector_grid, dtype=np.float64)
  pos = np.asarray(positions, dtype=np.float64).reshape(-1, 3)
  g_radius = np.sqrt((g * g).sum(-1))
  kmax = max_radius(g)
  g2 = g_radius ** 2
  g2[0, 0, 0] = 1e20                      # the reference's 1e10 ** 2 guard; the bin is zeroed below
  v_g = np.zeros(g.shape[:-1], dtype=np.complex128)
  radial = {}
  for a in range(pos.shape[0]):
    r, v_r, z = np.asarray(r_grid[a]), np.asarray(local_potential_grid[a]), local_potential_charge[a]
    key = (r.shape[0], float(z), float(r[0]), float(r[-1]), float(v_r[0]), float(v_r.sum()))
    if key not in radial:
      kk, f_k = sbt_numerical(r, (v_r + z / r)[None], 0, kmax)
      short = 4 * np.pi * CubicSpline(kk, f_k[0])(g_radius)     # V + Z/r: short ranged
      atom = short - 4 * np.pi * z / g2                          # minus the Coulomb tail
      atom[0, 0, 0] = 0.0
      radial[key] = atom
    v_g += radial[key] * np.exp(-1j * (g @ pos[a]))
  return v_g * (g_radius.size / vol)


def energy_local(reciprocal_density_grid, potential_local_grid_reciprocal, vol: float) -> float:
  """Host-side twin of local.py:166-187 (reciprocal_braket, braket.py:58-72): for checks; in the
  drivers the device computes it as the external-energy slot."""
  rho = np.asarray(reciprocal_density_grid)
  v = np.asarray(potential_local_grid_reciprocal)
  n = v.size
  if rho.ndim == v.ndim + 1:
    rho = rho.sum(0)
  return float(np.real(np.sum(np.conj(v) * rho)) * vol / n / n)
