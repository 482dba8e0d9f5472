"""Band-mode driver (jrystal/calc/calc_band_structure_all_electrons.py:37-216): ground-state
density from the energy driver, then per k-point of the path a direct minimisation of
trace(C^H H C) over orthonormal C (hamiltonian_matrix_trace, hamiltonian.py:105-168) followed by
the diagonalisation of the nb x nb subspace matrix H_ij (hamiltonian.py:171-240).

What differs from the reference, and why:
  * v_eff[rho_gs] is computed ONCE (the reference recomputes it inside every step although the
    ground-state density is constant, hamiltonian.py:147-156) and handed to the plan with
    jrb_hpsi_prepare, which also resamples it onto the orbital grid once;
  * H_ij comes from one H-apply + one FP64 tensor-core Gram (jrb_hpsi + jrb_hamiltonian_matrix)
    instead of nb Hessian-vector products (hessian.py:21-55);
  * one plan walks the path (jrb_set_kpoints), warm-starting every k-point from its predecessor
    as the reference's fine-tuning scan does; with several ranks the path is split in contiguous
    chunks, one per GPU, without communication (the reference's pmap over the k-path, 115, 184-193).
"""
import dataclasses
from math import ceil
from typing import Optional

import numpy as np
import torch

from .. import parallel
from ..config import JrystalConfigDict, get_config
from ..k_path import get_k_path
from ..optim import Adam
from ..plan import Plan
from ..utils import check_spin_number
from .calc_ground_state_energy_all_electrons import calc as energy_calc
from .opt_utils import create_crystal, create_freq_mask, create_grids, create_pseudopotential


@dataclasses.dataclass
class BandStructureOutput:
  config: JrystalConfigDict
  k_path: np.ndarray          # (num_kpoints, 3) Cartesian, 1/Bohr
  eigenvalues: np.ndarray     # (spin, num_kpoints, band), Hartree
  ground_state: object


def calc(config: Optional[JrystalConfigDict] = None, ground_state=None, k_path=None,
         log=None) -> BandStructureOutput:
  """With `use_pseudopotential: true` this is the norm-conserving band driver
  (calc_band_structure_normcons.py:66-299 of the reference): the fixed potential is
  v_H[rho_gs] + v_xc[rho_gs] + V_loc, the projectors <beta|G+k> are rebuilt on the sphere for every
  k-point of the path (radial transforms on the grid of the WHOLE path, as the reference's
  pre_calc_beta_sbt does) and join the H-apply and H_ij."""
  config = config or get_config()
  crystal = create_crystal(config)
  pseudopot = create_pseudopotential(config, crystal) if config.use_pseudopotential else None
  freq_mask = create_freq_mask(config, crystal)
  if k_path is None:
    if config.get('k_path_file'):
      frac = np.load(config.k_path_file).reshape(-1, 3)
      k_path = frac @ (2.0 * np.pi * np.linalg.inv(crystal.cell_vectors).T)
    else:
      k_path = get_k_path(crystal.cell_vectors, config.k_path_special_points, config.num_kpoints)
  k_path = np.asarray(k_path, dtype=np.float64).reshape(-1, 3)
  if ground_state is None:
    ground_state = energy_calc(config, log=log)
  rho_gs = ground_state.density

  world, rank = parallel._world()
  if not config.parallel_over_k_path:
    world, rank = 1, 0
  lo, hi = parallel.shard_bands(k_path.shape[0], world, rank)  # contiguous chunk of the path
  num_electron = pseudopot.num_valence_electrons if pseudopot is not None else crystal.num_electron
  check_spin_number(int(num_electron), crystal.spin)
  num_bands = ceil(num_electron / 2) + config.band_structure_empty_bands
  og = config.get('orbital_grid', 'auto')
  plan = Plan(crystal.cell_vectors, freq_mask, k_path[lo:lo + 1], num_bands,
              orbital_grid=tuple(og) if isinstance(og, (list, tuple)) else og)
  if pseudopot is not None:
    from ..pseudopotential import normcons
    from ..pseudopotential.beta import max_radius
    g_vec, _, _ = create_grids(config, crystal)
    kmax = max_radius(g_vec, k_path)
    normcons.attach(plan, pseudopot, g_vec, kpts=k_path[lo:lo + 1], positions=crystal.positions,
                    kmax=kmax)
  else:
    plan.set_atoms(crystal.positions, crystal.charges)
  dev = plan.tdev
  veff = plan.potential(rho_gs.to(dev).contiguous(), config.xc, True, 7)
  plan.prepare_potential(veff)  # fixed over the whole scan: copied / resampled once
  rng = np.random.default_rng(config.seed)
  shape = (1, 1, plan.ng, num_bands)
  w_re = torch.from_numpy(rng.random(shape)).to(dev)
  w_im = torch.from_numpy(rng.random(shape)).to(dev)
  args = dict(config.optimizer_args)
  opt = Adam([w_re, w_im], learning_rate=args.pop('learning_rate'), **args)
  q = torch.empty(plan.sphere_shape, dtype=torch.complex128, device=dev)
  r = torch.empty((1, 1, num_bands, num_bands), dtype=torch.complex128, device=dev)
  hq = torch.empty_like(q)
  grads = (torch.empty_like(w_re), torch.empty_like(w_im))

  def update():
    plan.qr_fwd(w_re, w_im, out=(q, r))
    plan.hpsi(q, None, out=hq)
    plan.qr_bwd(q, r, hq, out=grads)
    opt.step(grads)

  eig = np.zeros((1, hi - lo, num_bands))
  for i in range(lo, hi):
    plan.set_kpoints(k_path[i:i + 1])
    if pseudopot is not None and i > lo:
      normcons.set_projectors(plan, normcons.projectors(
        plan, pseudopot, g_vec, k_path[i:i + 1], crystal.positions, kmax))
    epochs = config.band_structure_epoch if i == lo else (
      config.k_path_fine_tuning_epoch if config.k_path_fine_tuning else config.band_structure_epoch)
    for _ in range(int(epochs)):
      update()
    plan.qr_fwd(w_re, w_im, out=(q, r))
    plan.hpsi(q, None, out=hq)
    h = plan.overlap(q, hq)[0, 0].cpu().numpy()
    eig[0, i - lo] = np.linalg.eigvalsh(0.5 * (h + h.conj().T))
    if log is not None:
      log(f'k-point {i + 1}/{k_path.shape[0]}: lowest eigenvalues {eig[0, i - lo, :4]}')
  plan.check_status()
  if world > 1:
    parts = [None] * world
    torch.distributed.all_gather_object(parts, eig)
    eig = np.concatenate(parts, axis=1)
  return BandStructureOutput(config=config, k_path=k_path, eigenvalues=eig,
                             ground_state=ground_state)
