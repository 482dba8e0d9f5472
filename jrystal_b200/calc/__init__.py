"""Drivers on top of the hot path, same module names as jrystal/calc: `energy` (direct
minimisation of the total energy) and `band` (band structure along a k-path)."""
from .calc_band_structure_all_electrons import calc as band  # noqa: F401
from .calc_ground_state_energy_all_electrons import calc as energy  # noqa: F401
from .convergence import ConvergenceChecker  # noqa: F401
