"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Crystal structures used by the parity tests and the benchmark, restated as
literals because ``ase`` (the reference's xyz reader, jrystal/_src/crystal.py:138)
is not installed and ``/root/reference`` does not exist on the GPU box.

Every entry cites the reference geometry file it restates.  Lattice rows are the
cell vectors in Angstrom, exactly as in the ``Lattice="..."`` header of the
extended-xyz file; positions are Cartesian Angstrom.
"""

import numpy as np

# jrystal/_src/const.py:18
ANGSTROM2BOHR = 1.8897259886
# jrystal/_src/const.py:21
HARTREE2EV = 27.211407953

_Z = {'H': 1, 'C': 6, 'O': 8, 'Al': 13, 'Si': 14, 'Ti': 22, 'Sr': 38}

STRUCTURES = {
  # geometry/diamond.xyz
  'diamond': dict(
    lattice=[[0.0, 1.78335, 1.78335], [1.78335, 0.0, 1.78335],
             [1.78335, 1.78335, 0.0]],
    symbols=['C', 'C'],
    positions=[[-0.4458375, -0.4458375, -0.4458375],
               [0.4458375, 0.4458375, 0.4458375]],
  ),
  # geometry/si.xyz
  'si': dict(
    lattice=[[0.0, 2.71547892, 2.71547892], [2.71547892, 0.0, 2.71547892],
             [2.71547892, 2.71547892, 0.0]],
    symbols=['Si', 'Si'],
    positions=[[-0.67886973, -0.67886973, -0.67886973],
               [0.67886973, 0.67886973, 0.67886973]],
  ),
  # geometry/si8.xyz
  'si8': dict(
    lattice=[[5.430957859235396, 0.0, 0.0], [0.0, 5.430957859235396, 0.0],
             [0.0, 0.0, 5.430957859235396]],
    symbols=['Si'] * 8,
    positions=[[0.67886973, 0.67886973, 0.67886973],
               [0.67886973, 3.39434866, 3.39434866],
               [3.39434866, 0.67886973, 3.39434866],
               [3.39434866, 3.39434866, 0.67886973],
               [4.75208813, 4.75208813, 4.75208813],
               [4.75208813, 2.03660920, 2.03660920],
               [2.03660920, 4.75208813, 2.03660920],
               [2.03660920, 2.03660920, 4.75208813]],
  ),
  # geometry/diamond8.xyz
  'diamond8': dict(
    lattice=[[3.5667, 0.0, 0.0], [0.0, 3.5667, 0.0], [0.0, 0.0, 3.5667]],
    symbols=['C'] * 8,
    positions=[[0.44583750, 0.44583750, 0.44583750],
               [0.44583750, 2.22918750, 2.22918750],
               [2.22918750, 0.44583750, 2.22918750],
               [2.22918750, 2.22918750, 0.44583750],
               [3.12086250, 3.12086250, 3.12086250],
               [3.12086250, 1.33751250, 1.33751250],
               [1.33751250, 3.12086250, 1.33751250],
               [1.33751250, 1.33751250, 3.12086250]],
  ),
  # geometry/al_primitive.xyz
  'al_primitive': dict(
    lattice=[[0.0, 2.02475, 2.02475], [2.02475, 2.02475, 0.0],
             [2.02475, 0.0, 2.02475]],
    symbols=['Al'],
    positions=[[0.0, 0.0, 0.0]],
  ),
  # geometry/srtio3.xyz
  'srtio3': dict(
    lattice=[[3.899, 0.0, 0.0], [0.0, 3.899, 0.0], [0.0, 0.0, 3.899]],
    symbols=['Sr', 'Ti', 'O', 'O', 'O'],
    positions=[[1.9495, 1.9495, 1.9495], [0.0, 0.0, 0.0], [1.9495, 0.0, 0.0],
               [0.0, 1.9495, 0.0], [0.0, 0.0, 1.9495]],
  ),
}


def supercell(name: str, reps):
  """Replicate a structure ``reps = (a, b, c)`` times along its cell vectors.

  BASELINE config C3 (diamond-64) is geometry/diamond8.xyz replicated 2x2x2
  (SURVEY.md 8d); the reference has no such file, so it is generated here.
  """
  s = STRUCTURES[name]
  lat = np.asarray(s['lattice'], dtype=np.float64)
  pos = np.asarray(s['positions'], dtype=np.float64)
  reps = tuple(int(r) for r in reps)
  shifts = np.array(
    [[i, j, k] for i in range(reps[0]) for j in range(reps[1])
     for k in range(reps[2])], dtype=np.float64
  ) @ lat
  new_pos = (pos[None, :, :] + shifts[:, None, :]).reshape(-1, 3)
  return dict(
    lattice=(lat * np.asarray(reps, dtype=np.float64)[:, None]).tolist(),
    symbols=list(s['symbols']) * len(shifts),
    positions=new_pos.tolist(),
  )


def load(name: str, reps=None):
  """Return ``(cell_vectors[3,3], positions[na,3], charges[na])`` in Bohr.

  Follows Crystal.create_from_file (jrystal/_src/crystal.py:138-141): positions
  and cell are multiplied by ANGSTROM2BOHR; charges are atomic numbers.
  """
  s = supercell(name, reps) if reps is not None else STRUCTURES[name]
  cell = np.asarray(s['lattice'], dtype=np.float64) * ANGSTROM2BOHR
  pos = np.asarray(s['positions'], dtype=np.float64) * ANGSTROM2BOHR
  chg = np.asarray([_Z[x] for x in s['symbols']], dtype=np.float64)
  return cell, pos, chg
