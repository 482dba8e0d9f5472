"""k-path through the Brillouin zone for the band-structure driver (jrystal/_src/band/k_path.py:
14-51, which delegates to ase.dft.kpoints.bandpath).  ase is not a dependency here: the special
points of the lattices the shipped geometries use (FCC and BCC primitive cells, cubic, tetragonal,
orthorhombic and hexagonal cells) are tabulated in ase's conventions (fractions of b1, b2, b3 of
ase's standard cell of that lattice; a cell with the same metric is a rigid motion of it, so the
fractions carry over), and the `num` points are spread along the path in proportion to the segment
lengths with every special point on a sample.  Cells in another setting (a 60 degree hexagonal cell,
a BCC cell with mixed angles, centred / monoclinic / triclinic lattices) are refused: pass explicit
points with `k_path_file`."""
import numpy as np

T = 1.0 / 3.0
SPECIAL_POINTS = {
  'fcc': {'G': (0, 0, 0), 'X': (0.5, 0, 0.5), 'L': (0.5, 0.5, 0.5), 'W': (0.5, 0.25, 0.75),
          'K': (0.375, 0.375, 0.75), 'U': (0.625, 0.25, 0.625)},
  'cubic': {'G': (0, 0, 0), 'X': (0, 0.5, 0), 'M': (0.5, 0.5, 0), 'R': (0.5, 0.5, 0.5)},
  # primitive cell a/2 [[-1, 1, 1], [1, -1, 1], [1, 1, -1]] (all angles acos(-1/3))
  'bcc': {'G': (0, 0, 0), 'H': (0.5, -0.5, 0.5), 'N': (0, 0, 0.5), 'P': (0.25, 0.25, 0.25)},
  # a1 = a2, 120 degrees between them, a3 perpendicular
  'hex': {'G': (0, 0, 0), 'A': (0, 0, 0.5), 'H': (T, T, 0.5), 'K': (T, T, 0), 'L': (0.5, 0, 0.5),
          'M': (0.5, 0, 0)},
  # orthogonal axes, a1 = a2 != a3
  'tet': {'G': (0, 0, 0), 'A': (0.5, 0.5, 0.5), 'M': (0.5, 0.5, 0), 'R': (0, 0.5, 0.5),
          'X': (0, 0.5, 0), 'Z': (0, 0, 0.5)},
  # orthogonal axes, three different lengths
  'orc': {'G': (0, 0, 0), 'R': (0.5, 0.5, 0.5), 'S': (0.5, 0.5, 0), 'T': (0, 0.5, 0.5),
          'U': (0.5, 0, 0.5), 'X': (0.5, 0, 0), 'Y': (0, 0.5, 0), 'Z': (0, 0, 0.5)},
}
DEFAULT_PATH = {'fcc': 'GXWKGLUWLK', 'cubic': 'GXMGRX', 'bcc': 'GHNGPH', 'hex': 'GMKGALHA',
                'tet': 'GXMGZRAZ', 'orc': 'GXSYGZURTZ'}


def lattice_type(cell_vectors) -> str:
  a = np.asarray(cell_vectors, dtype=np.float64)
  n = np.linalg.norm(a, axis=1)
  cosang = np.array([(a[i] @ a[j]) / (n[i] * n[j]) for i, j in ((0, 1), (0, 2), (1, 2))])
  same = lambda x, y: abs(x - y) <= 1e-6 * max(x, y)
  if same(n[0], n[1]) and same(n[0], n[2]):
    if np.allclose(cosang, 0.0, atol=1e-8):
      return 'cubic'
    if np.allclose(cosang, 0.5, atol=1e-8):
      return 'fcc'
    if np.allclose(cosang, -1.0 / 3.0, atol=1e-8):
      return 'bcc'
  if np.allclose(cosang, 0.0, atol=1e-8):
    if same(n[0], n[1]):
      return 'tet'
    if not same(n[0], n[2]) and not same(n[1], n[2]):
      return 'orc'
  if same(n[0], n[1]) and np.allclose(cosang, [-0.5, 0.0, 0.0], atol=1e-8):
    return 'hex'
  raise NotImplementedError(
    'k-path tables exist for cubic, FCC / BCC primitive, tetragonal (a1 = a2), orthorhombic and '
    'hexagonal (120 degrees between a1 and a2) cells in their standard setting; pass explicit '
    'points (k_path_file) for other lattices or settings')


def get_k_path(cell_vectors, path=None, num: int = 60, fractional: bool = False) -> np.ndarray:
  """`path`: special-point letters, ',' separates disconnected pieces ('GXWK,GL').  Returns (num, 3)
  fractional coordinates or Cartesian k-vectors in 1/Bohr."""
  a = np.asarray(cell_vectors, dtype=np.float64)
  b = 2.0 * np.pi * np.linalg.inv(a).T
  kind = lattice_type(a)
  table = SPECIAL_POINTS[kind]
  pieces = [p for p in (path or DEFAULT_PATH[kind]).replace(' ', '').split(',') if p]
  segs = []
  for piece in pieces:
    for s, e in zip(piece[:-1], piece[1:]):
      if s not in table or e not in table:
        raise ValueError(f'unknown special point in "{piece}" for a {kind} lattice')
      segs.append((np.array(table[s], float), np.array(table[e], float)))
  if not segs:
    raise ValueError('the k-path needs at least two special points')
  lengths = np.array([np.linalg.norm((e - s) @ b) for s, e in segs])
  # special points sit on the sample whose index is proportional to the path length so far
  # (as ase's paths2kpts places them); linear interpolation in between
  total = lengths.sum()
  edges = np.concatenate([[0.0], np.cumsum(lengths)])
  num = int(num)
  idx = np.rint(edges / total * (num - 1)).astype(int)
  out = np.empty((num, 3))
  for j, (s0, e0) in enumerate(segs):
    n = idx[j + 1] - idx[j]
    for i in range(idx[j], idx[j + 1] + 1):
      f = 0.0 if n == 0 else (i - idx[j]) / n
      out[i] = s0 + f * (e0 - s0)
  return out if fractional else out @ b
