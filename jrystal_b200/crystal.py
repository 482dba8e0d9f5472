"""Crystal structure container (host side).

Mirrors jrystal.Crystal (jrystal/_src/crystal.py:33-189): positions / cell in Bohr, charges =
atomic numbers, `vol`, `A`, `B`, `reciprocal_vectors`, `scaled_positions`, `num_electron`.
`ase` (the reference's xyz reader) is not a dependency here: `create_from_file` reads the
extended-xyz subset the reference ships (Lattice="..." header + "symbol x y z" rows).
"""
import re
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

ANGSTROM2BOHR = 1.8897259886   # jrystal/_src/const.py:18
BOHR2ANGSTROM = 0.529177249    # jrystal/_src/const.py:17
HARTREE2EV = 27.211407953      # jrystal/_src/const.py:21

_SYMBOLS = (
  'H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As '
  'Se Br Kr Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd '
  'Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn'
).split()
ATOMIC_NUMBER = {s: i + 1 for i, s in enumerate(_SYMBOLS)}

# Built-in cells (Angstrom) for the benchmark configurations of BASELINE.md; the numbers are the
# published lattice constants also used by the reference's geometry/*.xyz inputs.
_A_SI, _A_C, _A_AL = 5.430957859235396, 3.5667, 4.0495
_DIAMOND_BASIS = np.array([[1, 1, 1], [1, 5, 5], [5, 1, 5], [5, 5, 1],
                           [7, 7, 7], [7, 3, 3], [3, 7, 3], [3, 3, 7]]) / 8.0


def _fcc(a):
  return np.array([[0.0, a / 2, a / 2], [a / 2, 0.0, a / 2], [a / 2, a / 2, 0.0]])


BUILTIN = {
  'diamond': (_fcc(_A_C), ['C', 'C'], np.array([[-1, -1, -1], [1, 1, 1]]) * _A_C / 8),
  'si': (_fcc(2 * 2.71547892), ['Si', 'Si'],
         np.array([[-1, -1, -1], [1, 1, 1]]) * 2.71547892 / 4),
  'si8': (np.eye(3) * _A_SI, ['Si'] * 8,
          np.round(_DIAMOND_BASIS * _A_SI, 8)),
  'diamond8': (np.eye(3) * _A_C, ['C'] * 8, _DIAMOND_BASIS * _A_C),
  # cubic perovskite, a = 3.899 A (geometry/srtio3.xyz of the reference)
  'srtio3': (np.eye(3) * 3.899, ['Sr', 'Ti', 'O', 'O', 'O'],
             np.array([[0.5, 0.5, 0.5], [0, 0, 0], [0.5, 0, 0], [0, 0.5, 0], [0, 0, 0.5]]) * 3.899),
  'al_primitive': (np.array([[0.0, 2.02475, 2.02475], [2.02475, 2.02475, 0.0],
                             [2.02475, 0.0, 2.02475]]), ['Al'], np.zeros((1, 3))),
}


@dataclass
class Crystal:
  charges: np.ndarray
  positions: np.ndarray
  cell_vectors: np.ndarray
  spin: Optional[int] = None
  symbols: Optional[Sequence[str]] = None

  @property
  def scaled_positions(self):
    return self.positions @ np.linalg.inv(self.cell_vectors).T

  @property
  def vol(self):
    return float(np.abs(np.linalg.det(self.cell_vectors)))

  @property
  def num_atom(self):
    return self.positions.shape[0]

  @property
  def num_electron(self):
    return int(np.sum(self.charges))

  @property
  def A(self):
    return self.cell_vectors

  @property
  def reciprocal_vectors(self):
    return 2 * np.pi * np.linalg.inv(self.cell_vectors).T

  @property
  def B(self):
    return self.reciprocal_vectors

  @staticmethod
  def _make(lattice_ang, symbols, positions_ang, spin):
    charges = np.array([ATOMIC_NUMBER[s] for s in symbols])
    if spin is None:
      spin = int(np.sum(charges) % 2)
    return Crystal(
      charges=charges,
      positions=np.asarray(positions_ang, dtype=np.float64) * ANGSTROM2BOHR,
      cell_vectors=np.asarray(lattice_ang, dtype=np.float64).reshape(3, 3) * ANGSTROM2BOHR,
      spin=spin,
      symbols=list(symbols),
    )

  @staticmethod
  def create_from_file(file_path: str, spin: Optional[int] = None):
    with open(file_path) as f:
      lines = [l for l in f.read().splitlines()]
    n = int(lines[0].split()[0])
    m = re.search(r'Lattice="([^"]+)"', lines[1])
    if not m:
      raise ValueError(f'{file_path}: no Lattice="..." in the extended-xyz header')
    lattice = np.array([float(v) for v in m.group(1).split()]).reshape(3, 3)
    symbols, pos = [], []
    for l in lines[2:2 + n]:
      t = l.split()
      symbols.append(t[0])
      pos.append([float(t[1]), float(t[2]), float(t[3])])
    return Crystal._make(lattice, symbols, np.array(pos), spin)

  @staticmethod
  def create_from_symbols(symbols, positions, cell_vectors, spin: Optional[int] = None):
    if isinstance(symbols, str):
      symbols = re.findall(r'[A-Z][a-z]?', symbols)
    return Crystal._make(cell_vectors, symbols, positions, spin)

  @staticmethod
  def create_builtin(name: str, repeat=None, spin: Optional[int] = None):
    if name not in BUILTIN:
      raise ValueError(f'crystal "{name}" is not built in (built in: {sorted(BUILTIN)}); point '
                       'crystal_file_path_path at its extended-xyz file (the reference ships its '
                       'structures as geometry/<name>.xyz)')
    lattice, symbols, pos = BUILTIN[name]
    lattice = np.asarray(lattice, dtype=np.float64)
    pos = np.asarray(pos, dtype=np.float64)
    if repeat is not None:
      reps = [int(r) for r in repeat]
      shifts = np.array([[i, j, k] for i in range(reps[0]) for j in range(reps[1])
                         for k in range(reps[2])], dtype=np.float64) @ lattice
      pos = (pos[None] + shifts[:, None]).reshape(-1, 3)
      symbols = list(symbols) * len(shifts)
      lattice = lattice * np.asarray(reps, dtype=np.float64)[:, None]
    return Crystal._make(lattice, symbols, pos, spin)
