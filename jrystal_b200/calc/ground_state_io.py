"""Ground state on disk: what `jrystal -m energy` leaves for `jrystal -m band -l FILE`.

The reference's command line declares `-l/--load` ("Load pickled output from energy calculation
for band structure calculation", main.py:31-38) and a `save_dir` key (config.py:25,64) but wires
neither: its band drivers always redo the energy minimisation first
(calc_band_structure_all_electrons.py:60-66).  Here the energy mode writes `ground_state.npz`
into `save_dir` and the band mode takes it back.  A plain .npz (no pickle): density, parameters,
occupations, the energy split and the settings the density depends on, which `load` checks
against the configuration of the band run."""
import os

import numpy as np
import torch

FILE_NAME = 'ground_state.npz'
_ENERGY_KEYS = ('kinetic', 'external', 'hartree', 'xc', 'ewald', 'entropy', 'external_local',
                'external_nonlocal')


def _grid_sizes(config):
  g = config.grid_sizes
  return [int(g)] * 3 if np.isscalar(g) else [int(v) for v in g]


def _k_grid_sizes(config):
  g = config.k_grid_sizes
  return [int(g)] * 3 if np.isscalar(g) else [int(v) for v in g]


# settings the density depends on beyond cell / grid / functional: stored by `save`, compared by
# `load` (a density from another cut-off or k-mesh must not be accepted silently)
_SETTING_KEYS = ('cutoff_energy', 'freq_mask_method', 'spin_restricted', 'occupation', 'smearing',
                 'spin', 'empty_bands')


def _setting(config, key):
  v = config.get(key)
  return 'None' if v is None else str(float(v)) if isinstance(v, (int, float)) and not isinstance(
    v, bool) else str(v)


def save(out, path: str) -> str:
  """Write a GroundStateEnergyOutput; `path` is a directory (-> path/ground_state.npz) or a file."""
  if os.path.isdir(path) or not path.endswith('.npz'):
    os.makedirs(path, exist_ok=True)
    path = os.path.join(path, FILE_NAME)
  c = out.config
  arrays = dict(
    density=out.density.cpu().numpy(), w_re=out.params_pw['w_re'].cpu().numpy(),
    w_im=out.params_pw['w_im'].cpu().numpy(), occupation=out.occupation.cpu().numpy(),
    total_energy=np.float64(out.total_energy),
    total_energy_history=np.asarray(out.total_energy_history, dtype=np.float64),
    converged=np.bool_(out.converged), steps=np.int64(out.steps),
    crystal=np.str_(str(c.crystal)), xc=np.str_(str(c.xc)),
    use_pseudopotential=np.bool_(c.use_pseudopotential),
    grid_sizes=np.asarray(_grid_sizes(c), dtype=np.int64),
    cell_vectors=np.asarray(out.crystal.cell_vectors, dtype=np.float64),
    k_grid_sizes=np.asarray(_k_grid_sizes(c), dtype=np.int64),
    # rows [k0, k1) of the k-mesh this file holds (a k-sharded run saves its own block)
    k_range=np.asarray(getattr(out, 'k_range', None) or (0, out.occupation.shape[1]),
                       dtype=np.int64))
  for k in _SETTING_KEYS:
    arrays['setting_' + k] = np.str_(_setting(c, k))
  for k in _ENERGY_KEYS:
    if k in out.energies:
      arrays['energy_' + k] = np.float64(out.energies[k])
  np.savez(path, **arrays)
  return path


def load(path: str, config):
  """GroundStateEnergyOutput for calc.band(config, ground_state=...).  Raises ValueError when
  the file was made for another crystal / grid / functional than `config` describes."""
  from .calc_ground_state_energy_all_electrons import GroundStateEnergyOutput
  from .opt_utils import create_crystal
  if os.path.isdir(path):
    path = os.path.join(path, FILE_NAME)
  with np.load(path, allow_pickle=False) as z:
    crystal = create_crystal(config)
    if not np.allclose(z['cell_vectors'], crystal.cell_vectors, rtol=0, atol=1e-10):
      raise ValueError(f'{path}: saved for another cell than crystal={config.crystal!r}')
    if list(z['grid_sizes']) != _grid_sizes(config):
      raise ValueError(f'{path}: saved on grid {list(z["grid_sizes"])}, config has '
                       f'{_grid_sizes(config)}')
    if str(z['xc']) != str(config.xc) or bool(z['use_pseudopotential']) != bool(
        config.use_pseudopotential):
      raise ValueError(f'{path}: saved with xc={str(z["xc"])!r}, use_pseudopotential='
                       f'{bool(z["use_pseudopotential"])}; config differs')
    if 'k_grid_sizes' in z.files and list(z['k_grid_sizes']) != _k_grid_sizes(config):
      raise ValueError(f'{path}: saved with k-mesh {list(z["k_grid_sizes"])}, config has '
                       f'{_k_grid_sizes(config)}')
    for k in _SETTING_KEYS:
      if 'setting_' + k in z.files and str(z['setting_' + k]) != _setting(config, k):
        raise ValueError(f'{path}: saved with {k}={str(z["setting_" + k])}, config has '
                         f'{_setting(config, k)}')
    if 'k_range' in z.files:
      nk_mesh = int(np.prod(_k_grid_sizes(config)))
      if tuple(int(v) for v in z['k_range']) != (0, nk_mesh):
        raise ValueError(f'{path}: holds k-points {tuple(z["k_range"])} of {nk_mesh} (the block of '
                         'one rank of a k-sharded run); gather before loading')
    energies = {k: float(z['energy_' + k]) for k in _ENERGY_KEYS if 'energy_' + k in z.files}
    return GroundStateEnergyOutput(
      config=config, crystal=crystal,
      params_pw={'w_re': torch.from_numpy(z['w_re']), 'w_im': torch.from_numpy(z['w_im'])},
      occupation=torch.from_numpy(z['occupation']), density=torch.from_numpy(z['density']),
      total_energy=float(z['total_energy']), energies=energies,
      total_energy_history=[float(v) for v in z['total_energy_history']],
      converged=bool(z['converged']), steps=int(z['steps']), seconds_per_step=0.0)
