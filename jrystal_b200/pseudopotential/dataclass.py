"""Containers of the pseudopotential data of a crystal, one list entry PER ATOM, field for field
the reference's jrystal/pseudopotential/dataclass.py:38-196 (units converted Rydberg -> Hartree,
beta functions divided by r, points with r = 0 dropped)."""
import dataclasses
from typing import List, Optional

import numpy as np

from .load import find_upf, parse_upf


@dataclasses.dataclass
class Pseudopotential:
  num_atom: int
  positions: np.ndarray            # (atom, 3) Bohr
  charges: np.ndarray              # (atom,) atomic numbers
  atomic_symbols: List[str]
  valence_charges: List[int]


@dataclasses.dataclass
class NormConservingPseudopotential(Pseudopotential):
  r_grid: List[np.ndarray]                       # radial mesh, r > 0
  r_ab: List[np.ndarray]                         # dr/di of the mesh
  r_cutoff: List[Optional[float]]
  l_max: List[int]
  l_max_rho: List[Optional[int]]
  local_potential_grid: List[np.ndarray]         # V_loc(r) in Hartree
  local_potential_charge: List[int]              # Z_valence (the -Z/r tail)
  num_beta: List[int]
  nonlocal_beta_grid: List[np.ndarray]           # (beta, r): beta(r), NOT r * beta(r)
  nonlocal_beta_cutoff_radius: List
  nonlocal_d_matrix: List[np.ndarray]            # (beta, beta) Hartree
  nonlocal_angular_momentum: List[np.ndarray]    # (beta,) l of each projector
  nonlocal_valence_configuration: List[List[dict]]

  @staticmethod
  def from_upf_dicts(positions, charges, symbols, upf_dicts) -> 'NormConservingPseudopotential':
    """One parsed UPF dict (load.parse_upf layout) per atom."""
    cols = {f.name: [] for f in dataclasses.fields(NormConservingPseudopotential)
            if f.name not in ('num_atom', 'positions', 'charges', 'atomic_symbols')}
    for pp in upf_dicts:
      head, nl = pp['PP_HEADER'], pp['PP_NONLOCAL']
      r_all = np.asarray(pp['PP_MESH']['PP_R'], dtype=np.float64)
      keep = r_all > 0
      z_val = int(float(head['z_valence']))
      cols['valence_charges'].append(z_val)
      cols['r_grid'].append(r_all[keep])
      cols['r_ab'].append(np.asarray(pp['PP_MESH']['PP_RAB'], dtype=np.float64)[keep])
      cols['r_cutoff'].append(None)
      cols['l_max'].append(int(head['l_max']))
      cols['l_max_rho'].append(int(head['l_max_rho']) if 'l_max_rho' in head else None)
      # UPF stores Rydberg: V_loc / 2 and D_ij / 2 are Hartree
      cols['local_potential_grid'].append(np.asarray(pp['PP_LOCAL'], dtype=np.float64)[keep] / 2)
      cols['local_potential_charge'].append(z_val)
      betas = nl.get('PP_BETA')
      if betas:
        nbeta = len(betas)
        # the file holds r * beta(r)
        rb = np.asarray([b['values'] for b in betas], dtype=np.float64)
        cols['nonlocal_beta_grid'].append(rb[:, keep] / r_all[keep])
        cols['nonlocal_beta_cutoff_radius'].append(betas[-1]['cutoff_radius'])
        cols['nonlocal_d_matrix'].append(
          np.asarray(nl['PP_DIJ'], dtype=np.float64).reshape(nbeta, nbeta) / 2)
        ls = [int(b['angular_momentum']) for b in betas]
      else:  # purely local pseudopotential: one null projector, as the reference pads it
        nbeta, ls = 1, [0]
        cols['nonlocal_beta_grid'].append(np.zeros((1, int(keep.sum()))))
        cols['nonlocal_beta_cutoff_radius'].append(0)
        cols['nonlocal_d_matrix'].append(np.zeros((1, 1)))
      cols['num_beta'].append(nbeta)
      cols['nonlocal_angular_momentum'].append(np.asarray(ls))
      cols['nonlocal_valence_configuration'].append(
        pp.get('PP_INFO', {}).get('Valence configuration', []))
    return NormConservingPseudopotential(
      num_atom=len(charges), positions=np.asarray(positions, dtype=np.float64),
      charges=np.asarray(charges), atomic_symbols=list(symbols), **cols)

  @staticmethod
  def create(crystal, dir: Optional[str] = None) -> 'NormConservingPseudopotential':
    """From a crystal (positions, charges, symbols) and a directory holding one UPF per element
    (dataclass.py:92-196).  Each element's file is parsed once."""
    if dir is None:
      raise ValueError('pseudopotential_file_dir is required: no UPF files ship with jrystal_b200')
    parsed = {}
    for sym in crystal.symbols:
      if sym not in parsed:
        parsed[sym] = parse_upf(find_upf(dir, sym))
    return NormConservingPseudopotential.from_upf_dicts(
      crystal.positions, crystal.charges, crystal.symbols, [parsed[s] for s in crystal.symbols])

  @property
  def num_valence_electrons(self) -> int:
    return int(np.sum(self.valence_charges))
