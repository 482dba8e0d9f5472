"""Drivers on top of the hot path, same module names as jrystal/calc: `energy` (direct
minimisation of the total energy) and `band` (band structure along a k-path); both branch on
`use_pseudopotential` like the reference's main.py:44-54 (`energy_normcons`, `band_normcons`)."""
from .calc_band_structure_all_electrons import calc as band  # noqa: F401
from .calc_band_structure_normcons import calc as band_normcons  # noqa: F401
from .calc_ground_state_energy_all_electrons import calc as energy  # noqa: F401
from .calc_ground_state_energy_normcons import calc as energy_normcons  # noqa: F401
from .convergence import ConvergenceChecker  # noqa: F401

# the names of jrystal/calc/__init__.py
energy_all_electrons = energy
band_all_electrons = band
