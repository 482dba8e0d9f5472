"""Norm-conserving pseudopotential set-up (host side): UPF files -> V_loc(G) on the grid and the
projectors <beta|G+k> on the cut-off sphere, which `Plan.set_external_potential` /
`Plan.set_nonlocal` hand to the CUDA path.  Module and function names follow
jrystal/pseudopotential/ (load, dataclass, spherical, beta, local, nloc, normcons)."""
from . import beta, dataclass, load, local, nloc, normcons, spherical
from .beta import beta_sbt_grid, beta_sbt_sphere, sbt_numerical
from .dataclass import NormConservingPseudopotential, Pseudopotential
from .load import find_upf, parse_upf
from .local import energy_local, energy_local_position_gradient, potential_local_reciprocal
from .nloc import (energy_nonlocal, energy_nonlocal_position_gradient, hamiltonian_nonlocal,
                   potential_nonlocal_psi_reciprocal, potential_nonlocal_psi_sphere, projector_rows)

__all__ = [
  'beta', 'dataclass', 'load', 'local', 'nloc', 'normcons', 'spherical',
  'beta_sbt_grid', 'beta_sbt_sphere', 'sbt_numerical', 'NormConservingPseudopotential',
  'Pseudopotential', 'find_upf', 'parse_upf', 'energy_local', 'potential_local_reciprocal',
  'energy_nonlocal', 'hamiltonian_nonlocal', 'potential_nonlocal_psi_sphere',
  'potential_nonlocal_psi_reciprocal', 'projector_rows', 'energy_local_position_gradient',
  'energy_nonlocal_position_gradient',
]
