"""Band-mode driver with norm-conserving pseudopotentials (jrystal/calc/
calc_band_structure_normcons.py:66-299): the same k-path walk as the all-electron driver
(`calc_band_structure_all_electrons.calc`, which branches on `use_pseudopotential`), kept as a
module of its own so that `calc.band_normcons` exists under the reference's name."""
from typing import Optional

from ..config import JrystalConfigDict, get_config
from .calc_band_structure_all_electrons import BandStructureOutput
from .calc_band_structure_all_electrons import calc as _band


def calc(config: Optional[JrystalConfigDict] = None, ground_state=None, k_path=None,
         log=None) -> BandStructureOutput:
  config = config or get_config()
  if not config.use_pseudopotential:
    raise ValueError('calc_band_structure_normcons needs use_pseudopotential: true')
  return _band(config, ground_state, k_path, log)
