"""Energy-mode driver with norm-conserving pseudopotentials, the caller of the hot path that the
shipped config.yaml selects (`use_pseudopotential: true`; jrystal/calc/
calc_ground_state_energy_normcons.py:63-308).

Set-up (host, once): UPF files -> NormConservingPseudopotential -> V_loc(G) on the grid and the
projectors <beta|G+k> on the cut-off sphere (jrystal_b200/pseudopotential/), handed to the plan
with jrb_set_external_potential / jrb_set_nonlocal.  The optimisation loop is the all-electron
one (`minimise`): the loss kinetic + hartree + external_local + external_nonlocal + xc of the
reference (lines 175-192) is what jrb_eval_begin / jrb_eval_finish then evaluate and
differentiate.  As in the reference the electron count and the band count come from the valence
charges, while the Ewald constant still uses the crystal's atomic numbers (opt_utils.py:187-201)."""
from typing import Optional

from ..config import JrystalConfigDict, get_config
from ..plan import Plan
from .calc_ground_state_energy_all_electrons import GroundStateEnergyOutput, minimise
from .opt_utils import create_pseudopotential


def calc(config: Optional[JrystalConfigDict] = None, plan: Optional[Plan] = None,
         use_cuda_graph: bool = True, log=None) -> GroundStateEnergyOutput:
  config = config or get_config()
  if not config.use_pseudopotential:
    raise ValueError('calc_ground_state_energy_normcons needs use_pseudopotential: true')
  pseudopot = create_pseudopotential(config)
  if log is not None:
    log(f'norm-conserving pseudopotentials: valence charges {pseudopot.valence_charges}, '
        f'projectors per atom {pseudopot.num_beta}')
  return minimise(config, plan, use_cuda_graph, log, pseudopot=pseudopot)
