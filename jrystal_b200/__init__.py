"""jrystal_b200: B200-native (sm_100a CUDA) energy+gradient path of sail-sg/jrystal.

Public surface mirrors the reference's Python API for this path (pw, grid, energy,
potential, kinetic, hamiltonian, occupation, entropy) over the C ABI of
include/jrystal_b200.h.  No CPU fallback: compute calls need the built
jrystal_b200/csrc/libjrystal_b200.so and a CUDA device.
"""
from . import _lib  # noqa: F401
from . import (autograd, calc, config, crystal, energy, entropy, ewald, grid, hamiltonian,  # noqa: F401
               kinetic, occupation, potential, pseudopotential, pw, sbt, utils)
from .context import current_plan, use_plan  # noqa: F401
from .crystal import Crystal  # noqa: F401
from .plan import Plan  # noqa: F401



def get_pkg_path() -> str:
  """jrystal.get_pkg_path (jrystal/__init__.py): the directory that holds the package."""
  import os
  return os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


__all__ = ['Plan', 'Crystal', 'use_plan', 'current_plan', 'pw', 'grid', 'energy', 'potential',
           'kinetic', 'hamiltonian', 'occupation', 'crystal', 'autograd', 'calc', 'config',
           'entropy', 'ewald', 'pseudopotential', 'sbt', 'utils', 'get_pkg_path']
