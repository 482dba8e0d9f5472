"""Set-up helpers of the drivers, same names as jrystal/calc/opt_utils.py (crystal, grids, masks,
optimiser, Ewald constant), on numpy + the device Adam."""
import numpy as np

from .. import grid
from ..crystal import Crystal
from ..ewald import ewald_coulomb_repulsion
from ..optim import Adam
from ..utils import check_spin_number


def create_crystal(config) -> Crystal:
  """opt_utils.py:105-114: a built-in `crystal` name wins over `crystal_file_path_path`, and the
  spin number must have the parity of the electron count (check_spin_number raises ValueError
  otherwise: with the default `spin: 0` an odd-electron cell would silently lose an electron)."""
  if config.get('crystal') is not None:
    crystal = Crystal.create_builtin(config.crystal, spin=config.get('spin'))
  else:
    crystal = Crystal.create_from_file(config.get('crystal_file_path_path'), spin=config.get('spin'))
  check_spin_number(crystal.num_electron, crystal.spin)
  return crystal


def create_freq_mask(config, crystal=None) -> np.ndarray:
  crystal = crystal or create_crystal(config)
  gs = grid.proper_grid_size(config.grid_sizes)
  if config.freq_mask_method == 'spherical':
    return grid.spherical_mask(crystal.cell_vectors, gs, config.cutoff_energy)
  if config.freq_mask_method == 'cubic':
    return grid.cubic_mask(gs)
  raise ValueError(f'freq_mask_method "{config.freq_mask_method}" is not supported')


def create_pseudopotential(config, crystal=None):
  """opt_utils.py:116-139: the container for `pseudopotential_type` ('nc' family only; ultrasoft
  is not on the B200 path) from the UPF files in `pseudopotential_file_dir` (required here: the
  reference's default points into its own checkout, and no UPF files ship with jrystal_b200)."""
  from ..pseudopotential import NormConservingPseudopotential
  assert config.use_pseudopotential
  kind = config.get('pseudopotential_type', 'nc')
  if kind not in ('normcons', 'normconserving', 'nc'):
    raise ValueError(f'Pseudopotential type {kind} is not supported.')
  crystal = crystal or create_crystal(config)
  return NormConservingPseudopotential.create(crystal, config.get('pseudopotential_file_dir'))


def create_grids(config, crystal=None):
  crystal = crystal or create_crystal(config)
  gs = grid.proper_grid_size(config.grid_sizes)
  ks = grid.proper_grid_size(config.k_grid_sizes)
  return (grid.g_vectors(crystal.cell_vectors, gs), grid.r_vectors(crystal.cell_vectors, gs),
          grid.k_vectors(crystal.cell_vectors, ks))


def create_optimizer(config, params) -> Adam:
  if config.optimizer != 'adam':
    raise NotImplementedError(f'optimizer "{config.optimizer}" is not implemented (adam only)')
  if config.get('scheduler'):
    raise NotImplementedError('Scheduler is not implemented yet.')  # as in the reference
  args = dict(config.optimizer_args)
  # the reference forwards optimizer_args to optax.adam: keys it omits take optax's defaults
  # (b1 0.9, b2 0.999), not the b2 = 0.99 of the shipped config
  args.setdefault('b1', 0.9)
  args.setdefault('b2', 0.999)
  return Adam(params, learning_rate=args.pop('learning_rate'), **args)


def get_ewald_coulomb_repulsion(config, crystal=None) -> float:
  crystal = crystal or create_crystal(config)
  return ewald_coulomb_repulsion(crystal.positions, crystal.charges, crystal.cell_vectors,
                                 ewald_eta=config.ewald_args.get('ewald_eta'))
