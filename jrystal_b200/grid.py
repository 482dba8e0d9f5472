"""Grid constructors (host side, setup time): G / R / k vector grids, cut-off masks, FFT
sizing.  Same names, arguments and results as jrystal.grid (jrystal/_src/grid.py), numpy
FP64; the index maps the CUDA kernels consume are derived from these masks by the plan.
"""
import numpy as np

_MAX_FFT = 2048  # jrystal/_src/utils.py:250


def _smooth7(n):
  for p in (2, 3, 5, 7):
    while n % p == 0:
      n //= p
  return n == 1


def fft_factor(n: int) -> int:
  """Smallest 7-smooth integer >= n (jrystal/_src/utils.py:227-254, const.CUFFT_FACTORS)."""
  if n > _MAX_FFT:
    raise ValueError(f"The grid number {n} is too large!")
  n = max(int(n), 1)
  while not _smooth7(n):
    n += 1
  return n


def proper_grid_size(grid_sizes):
  """jrystal/_src/grid.py:181-210."""
  if hasattr(grid_sizes, '__len__'):
    sizes = np.array(grid_sizes)
  else:
    try:
      sizes = np.ones(3, dtype=int) * int(grid_sizes)
    except Exception:
      raise TypeError('mesh should be a scalar, tuple, list or np.array.')
  return np.array([fft_factor(int(i)) for i in sizes])


def _frequency_grid(rows, grid_sizes, fractional):
  rows = np.asarray(rows, dtype=np.float64)
  out = 0.0
  for axis, n in enumerate(grid_sizes):
    n = int(n)
    f = np.fft.fftfreq(n, 1.0 if fractional else 1.0 / n)
    shape = [1, 1, 1, 3]
    shape[axis] = n
    out = out + (f[:, None] * rows[axis][None, :]).reshape(shape)
  return out


def g_vectors(cell_vectors, grid_sizes):
  """[x, y, z, 3] reciprocal lattice vectors, integer fftfreq order (grid.py:122-149)."""
  b = 2 * np.pi * np.linalg.inv(np.asarray(cell_vectors, dtype=np.float64)).T
  return _frequency_grid(b, grid_sizes, False)


def r_vectors(cell_vectors, grid_sizes):
  """[x, y, z, 3] real-space sampling points of the FFT grid (grid.py:152-178)."""
  return _frequency_grid(cell_vectors, grid_sizes, True)


def translation_vectors(cell_vectors, cutoff=1e4):
  """Real-space translations of the reference's Ewald sum (grid.py:213-236): the r-vector lattice
  of n^3 cells, n = ceil(cutoff / |a1 + a2 + a3|^2), flattened to (n^3, 3)."""
  a = np.asarray(cell_vectors, dtype=np.float64)
  n = int(np.ceil(cutoff / np.linalg.norm(a.sum(axis=0)) ** 2))
  return _frequency_grid(a, [n] * a.shape[0], False).reshape(-1, a.shape[0])


def g2cell_vectors(g_vector_grid):
  """Cell vectors back from a G grid: least squares of G against the integer frequencies
  (grid.py:428-448)."""
  g = np.asarray(g_vector_grid, dtype=np.float64)
  a = g_vectors(np.eye(3), g.shape[:-1]).reshape(-1, 3)
  return np.linalg.inv(np.linalg.solve(a.T @ a, a.T @ g.reshape(-1, 3)))


def r2cell_vectors(r_vector_grid):
  """Cell vectors back from an r grid (grid.py:451-470)."""
  r = np.asarray(r_vector_grid, dtype=np.float64)
  d = r_vectors(np.eye(3), r.shape[:-1]).reshape(-1, 3)
  return np.linalg.solve(d.T @ d, d.T @ r.reshape(-1, 3))


def g2r_vector_grid(g_vector_grid, cell_vectors=None):
  """grid.py:377-401."""
  g = np.asarray(g_vector_grid)
  cell = g2cell_vectors(g) if cell_vectors is None else cell_vectors
  return r_vectors(cell, g.shape[:-1])


def r2g_vector_grid(r_vector_grid, cell_vectors=None):
  """grid.py:404-425."""
  r = np.asarray(r_vector_grid)
  cell = r2cell_vectors(r) if cell_vectors is None else cell_vectors
  return g_vectors(cell, r.shape[:-1])


def grid_vector_radius(grid_vector):
  """|v| at every grid point (grid.py:352-374)."""
  v = np.asarray(grid_vector, dtype=np.float64)
  return np.sqrt(np.sum(v * v, axis=-1))


def half_frequency_shape(grid_sizes):
  """Shape of the block of frequencies |f| <= ((n - 1) // 2) // 2 per axis: what cubic_mask keeps
  (grid.py:57-69)."""
  out = []
  for n in (int(v) for v in grid_sizes):
    half = ((n - 1) // 2) // 2
    lower = -((n - 1) // 2) // 2        # floor division of the negative bound, as the reference
    out.append((half + 1) + (-lower))
  return tuple(out)


def monkhorst_pack(grid_sizes):
  sizes = [int(s) for s in grid_sizes]
  idx = np.stack(np.meshgrid(*[np.arange(s) for s in sizes], indexing='ij'), axis=-1)
  return (idx.reshape(-1, 3) + 0.5) / np.array(sizes) - 0.5


def k_vectors(cell_vectors, grid_sizes):
  """Monkhorst-Pack k-points in Cartesian 1/Bohr (grid.py:239-262)."""
  b = 2 * np.pi * np.linalg.inv(np.asarray(cell_vectors, dtype=np.float64)).T
  return monkhorst_pack(grid_sizes) @ b


def spherical_mask(cell_vectors, grid_sizes, cutoff_energy: float):
  """|G|^2 <= 2 E_cut (grid.py:265-293)."""
  g = g_vectors(cell_vectors, grid_sizes)
  return np.linalg.norm(g, axis=-1)**2 <= cutoff_energy * 2


def cubic_mask(grid_sizes):
  """Per-axis half-frequency box (grid.py:296-325)."""
  per_axis = []
  for n in grid_sizes:
    n = int(n)
    g_max = (n - 1) // 2
    keep = np.ones(n, dtype=bool)
    keep[g_max // 2 + 1:(-g_max // 2)] = False
    per_axis.append(keep)
  return per_axis[0][:, None, None] & per_axis[1][None, :, None] & per_axis[2][None, None, :]


def estimate_max_cutoff_energy(cell_vectors, mask):
  """grid.py:328-349."""
  g = g_vectors(cell_vectors, mask.shape)
  return float(np.max(np.linalg.norm(g, axis=-1)**2 / 2 * mask))


# axis lengths whose two-stage line plan is square-ish (few elements per thread, 3-4 resident CTAs
# in the z passes); measured on B200, profiles/r01_orbital_grid.md
ORBITAL_Z_LENGTHS = (7, 8, 9, 12, 16, 24, 32, 36, 49, 64, 81, 100, 128)


def min_orbital_grid(freq_mask) -> tuple:
  """Smallest alias-free box of the per-orbital FFTs, 4 gmax + 1 per axis: psi carries the
  frequencies of the mask (|f| <= gmax), |psi|^2 and the sphere part of v_eff psi twice that."""
  mask = np.asarray(freq_mask).astype(bool)
  out = []
  for ax in range(3):
    n = mask.shape[ax]
    idx = np.nonzero(mask.any(axis=tuple(a for a in range(3) if a != ax)))[0]
    f = np.where(idx < (n + 1) // 2, idx, idx - n)
    out.append(4 * int(np.abs(f).max()) + 1)
  return tuple(out)


def orbital_grid_candidates(full, need) -> list:
  """Boxes `orbital_grid='auto'` tries, best first; the last one never depends on the fused plane
  kernels being available.  z (the axis the fused y+x kernels loop over) shrinks to the smallest
  light line length >= need; x and y keep the grid's lengths except 128 -> 81 (fused 81 x 81
  kernels) and 48 -> 36 (fused 36 x 36 kernels: C5 48.2 k -> 68.6 k band-mode evaluations/s,
  profiles/r02_bench_C5*.json); grids below 48^3 are launch-latency bound and left alone."""
  full = tuple(int(v) for v in full)
  need = tuple(int(v) for v in need)
  # an axis the caller's own grid under-resolves (need > n) is left as it is: the reference
  # aliases there and so does this library; the other axes can still shrink exactly
  if full[0] * full[1] * full[2] < 48**3:
    return [full]
  nz = min([n for n in ORBITAL_Z_LENGTHS if need[2] <= n <= full[2]] or [full[2]])
  out = []
  if full[0] == full[1] == 128 and max(need[:2]) <= 81:
    out.append((81, 81, nz))
  if full[0] == full[1] == 48 and max(need[:2]) <= 36:
    out.append((36, 36, nz))
  out.append((full[0], full[1], nz))
  return out
