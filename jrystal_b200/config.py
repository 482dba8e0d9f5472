"""Configuration of the drivers: the keys and defaults of the reference's config.yaml
(jrystal/config.py:61-104), as a plain attribute dict (ml_collections is not needed)."""
from typing import Optional

import yaml

default_config = {
  "crystal": "diamond", "crystal_file_path_path": None, "save_dir": None, "spin": 0,
  "xc": "lda_x", "use_pseudopotential": False, "pseudopotential_type": "nc",
  "pseudopotential_file_dir": None,
  "freq_mask_method": "spherical", "cutoff_energy": 100, "grid_sizes": 64, "k_grid_sizes": 3,
  "occupation": "uniform", "smearing": 0.001, "empty_bands": 8, "spin_restricted": True,
  "ewald_args": {'ewald_eta': 0.1, 'ewald_cutoff': 2e4}, "epoch": 5000, "optimizer": "adam",
  "optimizer_args": {"learning_rate": 0.01, "b1": 0.9, "b2": 0.99}, "scheduler": None,
  "convergence_window_size": 20, "convergence_condition": 1e-4,
  "band_structure_empty_bands": 8, "k_path_special_points": None, "num_kpoints": 60,
  "k_path_file": None, "band_structure_epoch": 5000, "k_path_fine_tuning": True,
  "k_path_fine_tuning_epoch": 300, "seed": 123, "parallel_over_k_mesh": False,
  "parallel_over_k_path": True, "xla_preallocate": True, "jax_enable_x64": True,
  "jax_debug_nans": False, "verbose": True, "eps": 1e-8,
  # not a reference key: box of the per-orbital FFTs, 'auto' | 'full' | [nx, ny, nz]
  # (jrb_plan_set_orbital_grid; results do not depend on it)
  "orbital_grid": "auto",
}


class JrystalConfigDict(dict):
  """dict with attribute access (the subset of ml_collections.ConfigDict the drivers use)."""

  def __getattr__(self, k):
    try:
      return self[k]
    except KeyError as e:
      raise AttributeError(k) from e

  def __setattr__(self, k, v):
    self[k] = v


def get_config(config_file: Optional[str] = None, **overrides) -> JrystalConfigDict:
  cfg = JrystalConfigDict(default_config)
  if config_file is not None:
    with open(config_file, 'r') as f:
      cfg.update(yaml.safe_load(f) or {})
  cfg.update(overrides)
  if cfg.get('band_structure_empty_bands') is None:
    cfg.band_structure_empty_bands = cfg.empty_bands
  return cfg
