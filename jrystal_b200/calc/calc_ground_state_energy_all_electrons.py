"""Energy-mode driver: direct minimisation of the all-electron total energy over the plane-wave
parameters, the caller of the hot path (jrystal/calc/calc_ground_state_energy_all_electrons.py:63-249).

Per step the reference does jax.value_and_grad(free_energy) + optax.adam inside one jitted
`update`; here a step is jrb_eval_begin -> (all-reduce of rho, E_kin over the k mesh) ->
jrb_eval_finish -> jrb_adam_tick/apply, all asynchronous on one stream and allocation free, so on a
single GPU the whole step is captured once in a CUDA graph and replayed (launch latency is what
bounds the small configurations).  Like the reference the host reads the total energy every step
for the convergence window.

Occupations: the parameter-free schemes ('uniform', 'gamma'), for which the entropy term of the
free energy is a constant, and the trainable ones ('simplex-projector', 'idempotent'): there the
evaluation also returns dE/d occupation, which is chained through the (tiny, torch autograd)
occupation map together with -T dS/d occupation under the reference's annealing schedule; those
steps are not graph-captured."""
import dataclasses
import time
from math import ceil
from typing import List, Optional

import numpy as np
import torch

from .. import occupation, parallel
from ..config import JrystalConfigDict, get_config
from ..plan import Plan
from ..utils import check_spin_number
from .convergence import create_convergence_checker
from .opt_utils import (create_crystal, create_freq_mask, create_grids, create_optimizer,
                        get_ewald_coulomb_repulsion)


@dataclasses.dataclass
class GroundStateEnergyOutput:
  """Mirrors the reference's output container (lines 49-60) plus what the band driver needs."""
  config: JrystalConfigDict
  crystal: object
  params_pw: dict
  occupation: torch.Tensor
  density: torch.Tensor            # (spin, x, y, z), all-reduced
  total_energy: float              # electronic + Ewald
  energies: dict                   # kinetic, external, hartree, xc, ewald, entropy
  total_energy_history: List[float]
  converged: bool
  steps: int
  seconds_per_step: float
  # rows [k0, k1) of the k-mesh that params_pw / occupation hold (a k-sharded run keeps its block)
  k_range: Optional[tuple] = None


def temperature_scheduler(config):
  """optax.exponential_decay(100, epoch // 2, smearing / 100, end_value=smearing) (lines 190-199)."""
  if config.smearing > 0.:
    half = max(config.epoch // 2, 1)
    rate = config.smearing / 100.

    def sched(i):
      return max(100. * rate ** (i / half), config.smearing)
    return sched
  return lambda i: 0.


def calc(config: Optional[JrystalConfigDict] = None, plan: Optional[Plan] = None,
         use_cuda_graph: bool = True, log=None) -> GroundStateEnergyOutput:
  """All-electron energy mode; a config with `use_pseudopotential: true` (the shipped config.yaml)
  goes to the norm-conserving driver, as `jrystal -m energy` dispatches."""
  config = config or get_config()
  if config.use_pseudopotential:
    from . import calc_ground_state_energy_normcons
    return calc_ground_state_energy_normcons.calc(config, plan, use_cuda_graph, log)
  return minimise(config, plan, use_cuda_graph, log)


def minimise(config: JrystalConfigDict, plan: Optional[Plan] = None, use_cuda_graph: bool = True,
             log=None, pseudopot=None) -> GroundStateEnergyOutput:
  """The optimisation loop shared by the all-electron and the norm-conserving drivers.  With
  `pseudopot` (a NormConservingPseudopotential) the electron count is the valence charge, the
  external term is the local pseudopotential and the non-local term joins the kinetic slot
  (calc_ground_state_energy_normcons.py:84-192 of the reference)."""
  crystal = create_crystal(config)
  g_vec, _, k_vec = create_grids(config, crystal)
  freq_mask = create_freq_mask(config, crystal)
  num_kpts = k_vec.shape[0]
  num_electron = pseudopot.num_valence_electrons if pseudopot is not None else crystal.num_electron
  # the count the occupations are built from: a parity mismatch would drop an electron silently
  check_spin_number(int(num_electron), crystal.spin)
  num_bands = ceil(num_electron / 2) + config.empty_bands
  world, rank = parallel._world()
  use_k_mesh = bool(config.parallel_over_k_mesh) and world > 1
  k0, k1 = parallel.shard_kpoints(num_kpts, world, rank) if use_k_mesh else (0, num_kpts)
  ew = get_ewald_coulomb_repulsion(config, crystal)

  method = config.occupation
  trainable = occupation.trainable(method)
  # spin_restricted: false -> two spin channels (pw.py:88-91), per-spin densities and potentials;
  # the device implements the spin-scaled exchange only and refuses other functionals loudly
  num_spin = 1 if config.spin_restricted else 2
  if plan is None:
    og = config.get('orbital_grid', 'auto')
    plan = Plan(crystal.cell_vectors, freq_mask, k_vec[k0:k1], num_bands, num_spin=num_spin,
                orbital_grid=tuple(og) if isinstance(og, (list, tuple)) else og)
  elif plan.ns != num_spin:
    raise ValueError(f'the plan has {plan.ns} spin channel(s), spin_restricted={config.spin_restricted}')
  if pseudopot is not None:
    from ..pseudopotential import normcons
    normcons.attach(plan, pseudopot, g_vec, kpts=k_vec[k0:k1], positions=crystal.positions)
  else:
    plan.set_atoms(crystal.positions, crystal.charges)
  dev = plan.tdev
  rng = np.random.default_rng(config.seed)
  params_occ = occupation.param_init(rng, num_bands, num_electron, num_kpts, crystal.spin,
                                     method, config.spin_restricted)

  def get_occupation():
    """(spin, kpt, band) on the device; a torch graph for the trainable schemes (lines 109-117)."""
    o = occupation.occupation(params_occ, num_kpts, num_electron, crystal.spin, method,
                              config.spin_restricted)
    return o if trainable else torch.from_numpy(np.ascontiguousarray(o)).to(dev)

  occ_t = get_occupation()
  entropy = float(occupation.fermi_dirac_entropy_torch(occ_t.detach(), config.eps))
  # parameters: the same global stream on every rank, each keeps its k block
  shape = (num_spin, num_kpts, plan.ng, num_bands)
  w_re = torch.from_numpy(np.ascontiguousarray(rng.random(shape)[:, k0:k1])).to(dev)
  w_im = torch.from_numpy(np.ascontiguousarray(rng.random(shape)[:, k0:k1])).to(dev)
  occ = occ_t.detach()[:, k0:k1].contiguous()
  optimizer = create_optimizer(config, [w_re, w_im])
  occ_leaves, opt_occ = [], None
  if trainable:
    seen = set()
    for v in params_occ.values():
      for leaf in (v.values() if isinstance(v, dict) else [v]):
        if id(leaf) not in seen:
          seen.add(id(leaf))
          occ_leaves.append(leaf)
    opt_occ = create_optimizer(config, [leaf.detach() for leaf in occ_leaves])
    # the optimiser updates the storage the leaves view
    for leaf, p in zip(occ_leaves, opt_occ.params):
      assert p.data_ptr() == leaf.data_ptr()
  dbuf, rho, e_kin = parallel.density_buffers((num_spin, plan.nx, plan.ny, plan.nz), dev)
  out = (torch.empty(4, dtype=torch.float64, device=dev), torch.empty_like(w_re),
         torch.empty_like(w_im))
  sched = temperature_scheduler(config)
  state = {'i': 0, 'entropy': entropy}

  # On a k mesh the partial densities are all-reduced inside the library over NVLink peer memory
  # (jrb_eval with the plan's communicator) when that can be set up, else by torch.distributed
  # between jrb_eval_begin and jrb_eval_finish; one GPU: the one-call evaluation.
  peer = bool(use_k_mesh and hasattr(plan, 'comm_init') and plan.comm_init())
  one_call = hasattr(plan, 'eval') and (peer or not use_k_mesh)

  def evaluate(want_occ_grad):
    if one_call:
      return plan.eval(w_re, w_im, occ, config.xc, want_occ_grad=want_occ_grad, out=out, rho=rho)[3]
    plan.eval_begin(w_re, w_im, occ, rho, e_kin)
    if use_k_mesh:
      parallel.allreduce_density(rho, e_kin, dbuf)
    return plan.eval_finish(occ, rho, e_kin, config.xc, want_occ_grad=want_occ_grad, out=out)[3]

  def step():
    if trainable:
      o = get_occupation()
      occ.copy_(o.detach()[:, k0:k1])
    g_occ = evaluate(trainable)
    optimizer.step([out[1], out[2]])
    if trainable:
      # d(free energy)/d(occupation parameters): dE/d occ from the evaluation, chained through the
      # occupation map; minus T dS/d(parameters) (lines 139-147: free = total - temp * entropy)
      if use_k_mesh:
        parts = [torch.empty_like(g_occ) for _ in range(world)]
        torch.distributed.all_gather(parts, g_occ.contiguous())
        g_occ = torch.cat(parts, dim=1)
      s_t = occupation.fermi_dirac_entropy_torch(o, config.eps)
      surrogate = (g_occ.detach() * o).sum() - sched(state['i']) * s_t
      grads = torch.autograd.grad(surrogate, occ_leaves)
      opt_occ.step([g.contiguous() for g in grads])
      state['entropy'] = float(s_t.detach())
    state['i'] += 1

  graph = None
  step()  # warm-up outside the capture (one-time attribute calls); counts as step 0
  first_energy = float(out[0].sum().item())
  if use_cuda_graph and dev.type == 'cuda' and (peer or not use_k_mesh) and not trainable:
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
      step()
    # the capture does not execute: parameters are those after step 0

  checker = create_convergence_checker(config)
  history = [first_energy]
  converged = checker.check(first_energy)
  t0 = time.perf_counter()
  steps = 1
  for i in range(1, config.epoch):
    if converged:
      break
    graph.replay() if graph is not None else step()
    etot = float(out[0].sum().item())  # value BEFORE this step's update, as value_and_grad returns
    history.append(etot)
    steps += 1
    converged = checker.check(etot)
    if log is not None and (i % 50 == 0 or converged):
      temp = sched(i)
      log(f'step {i}: Loss {etot - temp * state["entropy"]:.6f} | Energy {etot + ew:.6f} | '
          f'Entropy {state["entropy"]:.4f} | T {temp:.2E}')
  torch.cuda.synchronize()
  dt = (time.perf_counter() - t0) / max(steps - 1, 1)
  plan.check_status()
  # final energies at the final parameters (lines 223-247 of the reference)
  if trainable:
    occ_t = get_occupation().detach()
    occ.copy_(occ_t[:, k0:k1])
    state['entropy'] = float(occupation.fermi_dirac_entropy_torch(occ_t, config.eps))
  evaluate(False)
  en = out[0].cpu().numpy()
  energies = dict(kinetic=float(en[0]), external=float(en[1]), hartree=float(en[2]),
                  xc=float(en[3]), ewald=float(ew), entropy=float(state['entropy']))
  if pseudopot is not None:
    # split the kinetic slot (kinetic + non-local) and name the terms as the reference logs them
    e_nl = torch.zeros(1, dtype=torch.float64, device=dev)
    if getattr(plan, 'nproj', 0):
      q, _ = plan.qr_fwd(w_re, w_im)
      e_nl = plan.nonlocal_energy(q, occ)
      if use_k_mesh:
        parallel.allreduce_sum(e_nl)
    energies.update(kinetic=float(en[0]) - float(e_nl), external_local=float(en[1]),
                    external_nonlocal=float(e_nl))
  return GroundStateEnergyOutput(
    config=config, crystal=crystal, params_pw={'w_re': w_re, 'w_im': w_im}, occupation=occ,
    density=rho.clone(), total_energy=float(en.sum() + ew), energies=energies,
    total_energy_history=history, converged=bool(converged), steps=steps, seconds_per_step=dt,
    k_range=(int(k0), int(k1)))
