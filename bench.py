#!/usr/bin/env python
"""Benchmark of the hot path: energy+gradient evaluations / second (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C1|C2|C3a|C3b] [--impl reference]

One "step" = one evaluation = value_and_grad(total_energy) of jrystal's energy-mode driver
(calc/calc_ground_state_energy_all_electrons.py:119-137,175-181), optimiser excluded:
(w_re, w_im, occ) -> (E_kin, E_ext, E_har, E_xc, dE/dw_re, dE/dw_im, rho).
Default workload: BASELINE config C2 (Si8, 64^3, 30 Ha, 4x4x4 k, 66 bands, M = 4224 orbitals).
Inputs are synthetic (U[0,1) parameters from numpy default_rng(123), uniform occupations).
Multi-GPU: one process per GPU (torchrun), k-points sharded over ranks, partial densities
all-reduced with NCCL; total work is fixed => "strong" scaling.

Prints ONE JSON line (rank 0).  The default run (no --config) measures C2, the configuration
BASELINE.json's metric is quoted on, and appends the two diamond-64 configurations of the same
metric (C3b, C3a) under the key "diamond64", so that the driver's 1/2/4/8-GPU runs carry both
systems the metric names.  `--impl reference` times the CPU restatement of the reference
(oracle/, torch FP64, all host cores): for C2 all 64 k-points, nothing extrapolated.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

WORKLOADS = {
  # name: (crystal, repeat, grid, cutoff Ha, k-grid, empty bands)            BASELINE.md sec. 3
  'C1': dict(crystal='si', repeat=None, grid=32, cutoff=20.0, kgrid=(2, 2, 2), empty=10,
             text='C1: Si2 primitive, 32^3, 20 Ha, 2x2x2 k, 24 bands'),
  'C2': dict(crystal='si8', repeat=None, grid=64, cutoff=30.0, kgrid=(4, 4, 4), empty=10,
             text='C2: Si8 conventional, 64^3, 30 Ha, 4x4x4 k, 66 bands'),
  'C3a': dict(crystal='diamond8', repeat=(2, 2, 2), grid=128, cutoff=40.0, kgrid=(1, 1, 1),
              empty=16, text='C3a: diamond-64 supercell, 128^3, 40 Ha, Gamma, 208 bands'),
  'C3b': dict(crystal='diamond8', repeat=(2, 2, 2), grid=128, cutoff=40.0, kgrid=(2, 2, 2),
              empty=16, text='C3b: diamond-64 supercell, 128^3, 40 Ha, 2x2x2 k, 208 bands'),
  # BASELINE config 4.  The Sr/Ti UPF files are not shipped with the reference, so the
  # pseudopotential DATA are synthetic: valence point charges 10/12/6 for the local
  # part, 18 random projectors per atom for the non-local part.  The arithmetic and the shapes
  # (40 valence electrons, 30 bands, 90 projectors, 216 k-points) are those of the real run.
  'C4': dict(crystal='srtio3', repeat=None, grid=64, cutoff=40.0, kgrid=(6, 6, 6), empty=10,
             valence=[10, 12, 6, 6, 6], nproj_per_atom=18,
             text='C4: SrTiO3 norm-conserving (synthetic PP data), 64^3, 40 Ha, 6x6x6 k, 30 bands, '
                  '90 projectors'),
  # BASELINE config 5: band mode.  One evaluation = value + gradient of hamiltonian_matrix_trace at
  # ONE k-point (hamiltonian.py:147-168: QR, H-apply with the fixed v_eff[rho_gs], trace, QR
  # adjoint); the 64 path points are independent and run as one batch (the reference scans them
  # one after the other per device, calc_band_structure_all_electrons.py:115-182).
  'C5': dict(crystal='al_primitive', repeat=None, grid=48, cutoff=50.0, kpath=64, empty=8,
             band_mode=True,
             text='C5: Al primitive band mode, 48^3, 50 Ha, 64 k-path points, 15 bands; one '
                  'evaluation = Hamiltonian trace + gradient at one k-point'),
}
METRIC = 'energy+grad evals/sec'
UNIT = 'eval/s'

# stdout carries exactly ONE line (the JSON): libraries that print to stdout (NCCL's version banner,
# torchrun notices) are routed to stderr for the whole run -- done in main(), so that importing
# this module (tests, tools) leaves the importer's stdout alone
_REAL_STDOUT = 1


def _route_stdout_to_stderr():
  global _REAL_STDOUT
  sys.stdout.flush()
  _REAL_STDOUT = os.dup(1)
  os.dup2(2, 1)


def emit(line: dict) -> None:
  sys.stdout.flush()
  os.write(_REAL_STDOUT, (json.dumps(line) + '\n').encode())


def build_workload(name):
  from jrystal_b200 import grid, occupation
  from jrystal_b200.crystal import Crystal
  w = WORKLOADS[name]
  crystal = Crystal.create_builtin(w['crystal'], repeat=w['repeat'])
  gs = grid.proper_grid_size(w['grid'])
  mask = grid.spherical_mask(crystal.cell_vectors, gs, w['cutoff'])
  if w.get('band_mode'):
    from jrystal_b200.k_path import get_k_path
    kpts = get_k_path(crystal.cell_vectors, None, w['kpath'])
  else:
    kpts = grid.k_vectors(crystal.cell_vectors, grid.proper_grid_size(w['kgrid']))
  if 'valence' in w:  # pseudopotential run: the ions carry their valence charge
    crystal.charges = np.asarray(w['valence'])
    crystal.spin = 0
  nb = int(np.ceil(crystal.num_electron / 2)) + w['empty']
  occ = occupation.uniform(kpts.shape[0], crystal.num_electron, crystal.spin, nb)
  nproj = w.get('nproj_per_atom', 0) * crystal.num_atom
  return dict(crystal=crystal, grid=[int(g) for g in gs], mask=mask, kpts=kpts, nb=nb, occ=occ,
              ng=int(mask.sum()), text=w['text'], nproj=nproj, band_mode=bool(w.get('band_mode')))


def synthetic_projectors(ng, nk, nproj, k0, k1, device):
  """Random complex projector table (nk_local, nproj, ng), the same global stream on every rank."""
  import torch
  rng = np.random.default_rng(321)
  out = np.empty((k1 - k0, nproj, ng), dtype=np.complex128)
  for k in range(nk):
    blk = 0.1 * (rng.standard_normal((nproj, ng)) + 1j * rng.standard_normal((nproj, ng)))
    if k0 <= k < k1:
      out[k - k0] = blk
  return torch.from_numpy(out).to(device)


def synthetic_params(ng, nk, nb, k0, k1):
  """U[0,1) FP64 parameters (ns=1, nk, ng, nb) from default_rng(123), rows [k0, k1)."""
  rng = np.random.default_rng(123)
  w_re = np.empty((1, k1 - k0, ng, nb))
  w_im = np.empty((1, k1 - k0, ng, nb))
  # generate k by k so that every rank sees the same global stream
  for which in (w_re, w_im):
    for k in range(nk):
      blk = rng.random((ng, nb))
      if k0 <= k < k1:
        which[0, k - k0] = blk
  return w_re, w_im


class ClockSampler:
  """SM clock and throttle reasons of one GPU during the timed region, through NVML (a few
  microseconds per sample; spawning nvidia-smi from every rank perturbs the run it measures)."""
  BAD = {'hw_slowdown': 0x8, 'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20,
         'sw_power_cap': 0x4}

  def __init__(self, index, enabled=True):
    self.samples, self.reasons = [], set()
    self.max_mhz = None
    self._stop = threading.Event()
    self._t = None
    self._h = None
    if not enabled:
      return
    try:
      import pynvml
      pynvml.nvmlInit()
      vis = os.environ.get('CUDA_VISIBLE_DEVICES')
      phys = int(vis.split(',')[index]) if vis and vis.split(',')[index].isdigit() else index
      self._nv = pynvml
      self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
      self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
      self._t = threading.Thread(target=self._run, daemon=True)
    except Exception:
      self._h = None

  def _sample(self):
    nv = self._nv
    self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)) if hasattr(
      nv, 'nvmlDeviceGetCurrentClocksEventReasons') else int(
        nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
    for name, bit in self.BAD.items():
      if mask & bit:
        self.reasons.add(name)

  def _run(self):
    while not self._stop.is_set():
      try:
        self._sample()
      except Exception:
        pass
      self._stop.wait(0.01)

  def __enter__(self):
    if self._t is not None:
      try:
        self._sample()
      except Exception:
        pass
      self._t.start()
    return self

  def __exit__(self, *a):
    self._stop.set()
    if self._t is not None:
      self._t.join(timeout=2)

  def summary(self):
    return {'sm_mhz': float(np.median(self.samples)) if self.samples else None,
            'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
            'samples': len(self.samples), 'source': 'nvml' if self._h is not None else 'unavailable'}


def measured_peak():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    try:
      return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
      pass
  return 6650.0, 'fallback (B200_PROFILING.md)'


# FP64 ceiling of one B200: DFMA and DMMA share one datapath, their times add (measured with
# tools/fp64_dmma_dfma_mix.cu on this pool's B200s, profiles/r01_fp64_dmma_dfma_mix.txt:
# DFMA alone 32.2-33.0, DMMA alone 36.8-37.0, mixed 35.1-35.5 TFLOP/s)
FP64_PEAK_TFLOPS = 35.0
FP64_PEAK_SOURCE = ('measured: tools/fp64_dmma_dfma_mix.cu, profiles/r01_fp64_dmma_dfma_mix.txt '
                    '(DFMA + DMMA mixed on one datapath, 35.1-35.5 TFLOP/s)')


def fp64_flops(mask, orbital_grid, num_orbitals, num_sk, ng, nb, psi_cache=False):
  """FP64 flops of one evaluation as the kernels execute it, counted with the nominal
  5 n log2 n per length-n line on the PRUNED passes of the orbital box (z passes on the occupied
  (x, y) columns only, y passes on the occupied x planes only): per orbital 2 z passes (scatter
  side of the density sweep, gather side of the H-apply; the H-apply reuses the z-transformed
  columns), 3 y and 3 x passes (density; H-apply inverse and forward) -- 2 y and 2 x with the psi(r)
  cache, where the H-apply reads psi(r) back instead of repeating its inverse passes.  QR products: 8 ng nb^2 per
  full complex tall-skinny product; Hermitian Gram, upper-triangle-only Gram and triangular
  factors charged one half: 2 x 1/2 (Gram) + 2 x 1/2 (apply, triangular R^-1) forward,
  1/2 (upper triangle of Q^H G) + 1/2 + 1 (two-term apply) backward = 4 products per (spin, k)."""
  nxw, nyw, nzw = (int(v) for v in orbital_grid)
  ncol = int(np.asarray(mask).any(axis=2).sum())
  nxo = int(np.asarray(mask).any(axis=(1, 2)).sum())
  line = lambda n: 5.0 * n * np.log2(n)
  nyx = 2 if psi_cache else 3
  per_orbital = (2 * ncol * line(nzw) + nyx * nxo * nzw * line(nyw) + nyx * nyw * nzw * line(nxw))
  fft = per_orbital * num_orbitals
  qr = 4.0 * 8.0 * ng * nb * nb * num_sk
  return fft, qr


def kernels_stamp():
  """sha256 over the CUDA sources the library is built from: profiles/traffic.json carries the
  stamp of the sources its ncu capture was made with and is refused when they have changed."""
  import hashlib
  h = hashlib.sha256()
  d = os.path.join(ROOT, 'jrystal_b200', 'csrc')
  for name in sorted(os.listdir(d)):
    if name.endswith(('.cu', '.cuh', '.h')):
      h.update(name.encode())
      h.update(open(os.path.join(d, name), 'rb').read())
  return h.hexdigest()[:16]


def measured_traffic(config, world):
  """(bytes per evaluation of the H-apply sweep, note): dram__bytes_read + dram__bytes_write from
  the ncu --set full capture committed as profiles/traffic.json (tools/capture_traffic.py writes
  it from the capture of `bench.py --config C`), only if it was taken with these kernel sources."""
  try:
    prof = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
  except Exception as e:
    return None, f'profiles/traffic.json unreadable: {e}'
  ent = prof.get(config)
  if not ent or world != 1:
    return None, 'no capture for this configuration / GPU count'
  stamp = kernels_stamp()
  if ent.get('kernels_sha') != stamp:
    return None, (f'capture refused: made with kernel sources {ent.get("kernels_sha")}, the '
                  f'library is built from {stamp} (re-run tools/gpu_capture_traffic.sh)')
  return ent['happly_dram_bytes_per_eval'], f'ncu capture {ent.get("source")} (kernels {stamp})'


def cpu_sample_eval(wl, nk_sample, steps, warmup):
  """Oracle (CPU restatement of the reference dataflow) on `nk_sample` k-points of the
  workload, all host threads.  Returns (eval/s scaled to the full workload, seconds/sample)."""
  import torch
  from oracle import reference_port as rp
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  c = wl['crystal']
  nk = wl['kpts'].shape[0]
  nk_sample = min(nk_sample, nk)
  s = rp.System(c.cell_vectors, c.positions, c.charges, wl['grid'], kpts=wl['kpts'][:nk_sample],
                cutoff_energy=None, mask_method='cubic')
  s.mask = wl['mask']
  s.num_g = wl['ng']
  w_re, w_im = synthetic_params(wl['ng'], nk, wl['nb'], 0, nk_sample)
  occ = wl['occ'][:, :nk_sample]
  dense_phi = None
  if wl['nproj']:
    # the reference contracts dense (kpt, proj, x, y, z) projectors over the whole box (nloc.py:143-158)
    phi = synthetic_projectors(wl['ng'], nk, wl['nproj'], 0, nk_sample, 'cpu').numpy()
    dense_phi = np.zeros((nk_sample, wl['nproj']) + tuple(wl['grid']), dtype=np.complex128)
    dense_phi[:, :, wl['mask']] = phi
  times = []
  for i in range(warmup + steps):
    t0 = time.perf_counter()
    rp.energy_and_grad(s, w_re, w_im, occ, nonlocal_phi=dense_phi)
    dt = time.perf_counter() - t0
    if i >= warmup:
      times.append(dt)
  t = float(np.mean(times))
  return 1.0 / (t * nk / nk_sample), t, cores, nk_sample


GRADIENT_PIN = ('energies / density / Q pinned by the reference source executed over a numpy stand-in '
                '(tests/golden/reference_*.npz); gradients have no reference-derived pin (no AD in '
                'the stand-in): torch-autograd oracle + finite differences + oracle/analytic.py')


def cpu_full_eval(wl, kblock):
  """Oracle evaluation of the WHOLE workload on the host cores with bounded memory: the k-points
  couple only through rho, so (1) rho and E_kin are accumulated over k-blocks, (2) the grid
  energies and v = dE/drho come from the total rho, (3) the gradients of every k-block are
  those of E_kin + <v, rho_block> -- reference_port.energy_and_grad_chunked run per k-block.
  Returns (eval/s, seconds, cores, nk)."""
  import torch
  from oracle import reference_port as rp
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  c = wl['crystal']
  nk = wl['kpts'].shape[0]
  s = rp.System(c.cell_vectors, c.positions, c.charges, wl['grid'], kpts=wl['kpts'],
                cutoff_energy=None, mask_method='cubic')
  s.mask = wl['mask']
  s.num_g = wl['ng']
  w_re, w_im = synthetic_params(wl['ng'], nk, wl['nb'], 0, nk)
  t0 = time.perf_counter()
  rp.energy_and_grad_kblocks(s, w_re, w_im, wl['occ'], kblock=kblock)
  t = time.perf_counter() - t0
  return 1.0 / t, t, cores, nk


def run_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  wl = build_workload(args.config)
  nk = wl['kpts'].shape[0]
  steps = max(1, min(args.steps, 3))
  if wl['band_mode']:
    value, t, cores, nks = cpu_sample_band(wl, 4, steps, min(args.warmup, 1))
    sample = (f'{nks} of {nk} path points x {wl["nb"]} bands at {wl["grid"]} per step, one after '
              f'the other; oracle port (torch FP64 autograd), {cores} threads')
  else:
    # C2 (the headline): every one of the 64 k-points, in k-blocks of 8 (the dense (k, band, x, y, z)
    # tensor of all 64 would need 17.7 GB per copy) -- nothing extrapolated, one step = 25-30 s
    nk_sample = {'C1': 8, 'C2': nk, 'C3a': 1, 'C3b': 1, 'C4': 8}[args.config]
    if args.config == 'C2':
      steps = 1
      value, t, cores, nks = cpu_full_eval(wl, 8)
      sample = (f'all {nk} k-points x {wl["nb"]} bands at {wl["grid"]}, one step of {t:.1f} s in '
                f'k-blocks of 8 (density pass, grid potential, gradient pass); nothing scaled; '
                f'oracle port (torch FP64 autograd), {cores} threads')
    else:
      value, t, cores, nks = cpu_sample_eval(wl, nk_sample, steps, min(args.warmup, 1))
      sample = (f'{nks} of {nk} k-points x {wl["nb"]} bands at {wl["grid"]} per step, time scaled '
                f'by {nk / nks:g}; oracle port (torch FP64 autograd), {cores} threads')
  line = {
    'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
    'steps': steps, 'warmup': 0 if args.config == 'C2' else min(args.warmup, 1),
    'ms_per_step': 1e3 / value,
    'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
    'data': 'synthetic',
    'config': {'workload': wl['text'], 'note': 'CPU restatement of the reference (JAX absent)',
               'gradient_pin': GRADIENT_PIN},
    'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                     'sample': sample},
    'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    'gpu_launches': 0,
  }
  emit(line)


def run_b200_rows(args, wl, world, rank, local_rank, steps=None):
  """Fewer k-points than GPUs (C3a, Gamma only): rows of the parameters sharded for the QR, bands
  for the FFTs (jrystal_b200.parallel.RowShardedEvaluator); strong scaling.  Returns the line."""
  import torch
  import torch.distributed as dist
  from jrystal_b200 import _lib
  from jrystal_b200.parallel import RowShardedEvaluator

  c = wl['crystal']
  nk, nb, ng = wl['kpts'].shape[0], wl['nb'], wl['ng']
  ngrid = int(np.prod(wl['grid']))
  ev = RowShardedEvaluator(c.cell_vectors, wl['mask'], wl['kpts'], nb, c.positions, c.charges,
                           device=local_rank, orbital_grid=parse_orbital_grid(args.orbital_grid))
  w_re_f, w_im_f = synthetic_params(ng, nk, nb, 0, nk)
  w_re_h = np.ascontiguousarray(w_re_f[:, :, ev.g0:ev.g1])
  w_im_h = np.ascontiguousarray(w_im_f[:, :, ev.g0:ev.g1])
  del w_re_f, w_im_f
  occ_h = np.ascontiguousarray(wl['occ'])
  w_re, w_im, occ = (torch.from_numpy(a).cuda() for a in (w_re_h, w_im_h, occ_h))
  result = {}

  def step():
    result['out'] = ev.evaluate(w_re, w_im, occ, 'lda_x')

  def barrier():
    dist.barrier()
    torch.cuda.synchronize()

  def timed(fn, steps):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(steps):
      fn()
    ev1.record()
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device='cuda')
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())

  lib = _lib.load()
  steps = steps or args.steps
  for _ in range(max(args.warmup, 3)):
    step()
  barrier()
  launches0 = lib.jrb_launch_count()
  with ClockSampler(local_rank, enabled=(rank == 0)) as clk:
    total_ms = timed(step, steps)
  launches = int(lib.jrb_launch_count() - launches0)
  ms_per_step = total_ms / steps
  value = 1e3 / ms_per_step
  energies = result['out'][0].cpu().numpy().tolist()

  pin = lambda a: torch.from_numpy(a).pin_memory()
  w_re_p, w_im_p, occ_p = pin(w_re_h), pin(w_im_h), pin(occ_h)
  en_p = torch.empty(4, dtype=torch.float64).pin_memory()
  g_re_p = torch.empty(w_re_h.shape, dtype=torch.float64).pin_memory()
  g_im_p = torch.empty(w_re_h.shape, dtype=torch.float64).pin_memory()

  def e2e_step():
    w_re.copy_(w_re_p, non_blocking=True)
    w_im.copy_(w_im_p, non_blocking=True)
    occ.copy_(occ_p, non_blocking=True)
    step()
    en, g_re, g_im, _ = result['out']
    en_p.copy_(en, non_blocking=True)
    g_re_p.copy_(g_re, non_blocking=True)
    g_im_p.copy_(g_im, non_blocking=True)
    torch.cuda.current_stream().synchronize()

  e2e_steps = max(1, min(steps, 5))
  e2e_step()
  barrier()
  t0 = time.perf_counter()
  for _ in range(e2e_steps):
    e2e_step()
  barrier()
  e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device='cuda')
  dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
  e2e_value = e2e_steps / float(e2e_s.item())
  nw = w_re_h.size
  peak, peak_src = measured_peak()
  m_local = nk * (ev.b1 - ev.b0)
  bytes_alg = 64.0 * nk * nb * (ngrid + ng) / world  # per GPU share of the whole evaluation
  whole_achieved = bytes_alg / (ms_per_step * 1e-3) / 1e9
  fft_fl, qr_fl = fp64_flops(wl['mask'], ev.bands.orbital_grid, nk * nb, nk, ng, nb)
  fp64_tf = (fft_fl + qr_fl) / world / (ms_per_step * 1e-3) / 1e12
  line = None
  if rank == 0:
    line = {
      'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': steps,
      'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True,
      'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
      'config': {'workload': wl['text'], 'orbitals': nk * nb, 'ng': ng, 'grid': wl['grid'],
                 'sharding': f'rows{world} (QR) + bands{world} (FFT), all-to-all between',
                 'reduce_path': ev.reduce_path,
                 'xc': 'lda_x', 'bands_per_rank': ev.b1 - ev.b0, 'rows_per_rank': ev.g1 - ev.g0,
                 'orbital_grid': list(ev.bands.orbital_grid),
                 'l2': 'working set larger than L2 (pencil work space); no flush'},
      'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': 2 * nw * 8 + occ_h.size * 8,
              'd2h_bytes_per_step': 2 * nw * 8 + 32,
              'path': 'pinned H2D of the row block + RowShardedEvaluator.evaluate (NCCL) + D2H'},
      'gpu_launches': launches,
      'roofline': {'bound': 'hbm', 'achieved': whole_achieved, 'peak': peak, 'unit': 'GB/s',
                   'frac': whole_achieved / peak, 'traffic': None, 'peak_source': peak_src,
                   'kernel': 'whole evaluation per GPU (64*M*(N+ng)/n_gpus algorithmic bytes)',
                   'orbitals_per_gpu': m_local,
                   'fp64': {'flops_per_eval': fft_fl + qr_fl, 'achieved': fp64_tf,
                            'peak': FP64_PEAK_TFLOPS, 'unit': 'TFLOP/s',
                            'frac': fp64_tf / FP64_PEAK_TFLOPS, 'peak_source': FP64_PEAK_SOURCE}},
      'clocks': clk.summary(), 'energies_ha': energies,
    }
  del ev, w_re, w_im, occ, result
  torch.cuda.empty_cache()
  return line


def cpu_sample_band(wl, nk_sample, steps, warmup):
  """Oracle band-mode evaluation (hamiltonian_matrix_trace value + gradient) on `nk_sample` path
  points, one after the other as the reference scans them; returns k-point evaluations / s."""
  import torch
  from oracle import reference_port as rp
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  c = wl['crystal']
  rng = np.random.default_rng(7)
  rho = np.abs(rng.standard_normal((1,) + tuple(wl['grid']))) * c.num_electron / c.vol
  w_re, w_im = synthetic_params(wl['ng'], wl['kpts'].shape[0], wl['nb'], 0, nk_sample)
  times = []
  for i in range(warmup + steps):
    t0 = time.perf_counter()
    for k in range(nk_sample):
      s = rp.System(c.cell_vectors, c.positions, c.charges, wl['grid'], kpts=wl['kpts'][k:k + 1],
                    cutoff_energy=None, mask_method='cubic')
      s.mask = wl['mask']
      s.num_g = wl['ng']
      rp.band_trace_and_grad(s, w_re[:, k:k + 1], w_im[:, k:k + 1], rho)
    dt = time.perf_counter() - t0
    if i >= warmup:
      times.append(dt)
  t = float(np.mean(times))
  return nk_sample / t, t, cores, nk_sample


def run_b200_band(args, wl, world, rank, local_rank):
  """C5: band-mode evaluations (Hamiltonian trace + gradient per k-point), the path points split
  in contiguous chunks over the ranks without communication (the reference's pmap over the
  k-path); value = k-point evaluations / s of the whole job."""
  import torch
  import torch.distributed as dist
  import jrystal_b200 as jb
  from jrystal_b200 import _lib, parallel

  c = wl['crystal']
  nk, nb, ng = wl['kpts'].shape[0], wl['nb'], wl['ng']
  ngrid = int(np.prod(wl['grid']))
  k0, k1 = parallel.shard_kpoints(nk, world, rank) if nk % world == 0 else parallel.shard_bands(
    nk, world, rank)
  nkl = k1 - k0
  plan = jb.Plan(c.cell_vectors, wl['mask'], wl['kpts'][k0:k1], nb, device=local_rank,
                 orbital_grid=parse_orbital_grid(args.orbital_grid))
  plan.set_atoms(c.positions, c.charges)
  rng = np.random.default_rng(7)
  rho_h = np.abs(rng.standard_normal((1,) + tuple(wl['grid']))) * c.num_electron / c.vol
  rho = torch.from_numpy(rho_h).cuda()
  _, veff = plan.grid_potential(rho, 'lda_x', True)   # v_eff[rho_gs]: once, not per step
  plan.prepare_potential(veff)                        # ... and resampled onto the orbital grid once
  w_re_h, w_im_h = synthetic_params(ng, nk, nb, k0, k1)
  w_re, w_im = torch.from_numpy(w_re_h).cuda(), torch.from_numpy(w_im_h).cuda()
  cdt = torch.complex128
  q = torch.empty(w_re.shape, dtype=cdt, device='cuda')
  r = torch.empty((1, nkl, nb, nb), dtype=cdt, device='cuda')
  hq = torch.empty_like(q)
  eps = torch.empty((1, nkl, nb), dtype=torch.float64, device='cuda')
  grads = (torch.empty_like(w_re), torch.empty_like(w_im))

  def step():
    plan.qr_fwd(w_re, w_im, out=(q, r))
    plan.hpsi(q, None, out=hq)
    plan.band_expect(q, hq, out=eps)
    plan.qr_bwd(q, r, hq, out=grads)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  flush_buf = torch.zeros(32 * 2**20, dtype=torch.float64, device='cuda')
  lib = _lib.load()
  for _ in range(max(args.warmup, 3)):
    step()
  graph = torch.cuda.CUDAGraph()
  with torch.cuda.graph(graph):
    step()
  graph.replay()
  barrier()

  def timed(fn):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    barrier()
    for e0, e1 in evs:
      flush_buf.add_(1.0)
      e0.record()
      fn()
      e1.record()
    barrier()
    ms = torch.tensor([sum(e0.elapsed_time(e1) for e0, e1 in evs)], dtype=torch.float64,
                      device='cuda')
    if world > 1:
      dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()) / args.steps

  launches0 = lib.jrb_launch_count()
  with ClockSampler(local_rank, enabled=(rank == 0)) as clk:
    ms_eager = timed(step)
  launches = int(lib.jrb_launch_count() - launches0)
  ms_graph = timed(graph.replay)
  ms_per_step = min(ms_eager, ms_graph)
  value = nk / (ms_per_step * 1e-3)

  # end to end: parameters of the chunk from pinned host memory, gradients + band energies back
  pin = lambda a: torch.from_numpy(a).pin_memory()
  w_re_p, w_im_p = pin(w_re_h), pin(w_im_h)
  g_re_p = torch.empty(w_re_h.shape, dtype=torch.float64).pin_memory()
  g_im_p = torch.empty(w_re_h.shape, dtype=torch.float64).pin_memory()
  eps_p = torch.empty((1, nkl, nb), dtype=torch.float64).pin_memory()

  def e2e_step():
    w_re.copy_(w_re_p, non_blocking=True)
    w_im.copy_(w_im_p, non_blocking=True)
    step()
    g_re_p.copy_(grads[0], non_blocking=True)
    g_im_p.copy_(grads[1], non_blocking=True)
    eps_p.copy_(eps, non_blocking=True)
    torch.cuda.current_stream().synchronize()

  e2e_step()
  barrier()
  t0 = time.perf_counter()
  for _ in range(args.steps):
    e2e_step()
  barrier()
  e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device='cuda')
  if world > 1:
    dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
  e2e_value = nk * args.steps / float(e2e_s.item())

  peak, peak_src = measured_peak()
  bytes_alg = nkl * nb * (32.0 * ngrid + 48.0 * ng)  # one dense transform pair + W, Q, HQ, dW
  achieved = bytes_alg / (ms_per_step * 1e-3) / 1e9
  if rank == 0:
    nw = w_re_h.size
    line = {
      'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
      'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True,
      'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
      'config': {'workload': wl['text'], 'k_points': nk, 'bands': nb, 'ng': ng,
                 'grid': wl['grid'], 'sharding': f'kpath{world}' if world > 1 else 'none',
                 'xc': 'lda_x', 'evaluations_per_step': nk,
                 'orbital_grid': list(plan.orbital_grid),
                 'l2': 'working set fits L2: L2 flushed between timed steps (256 MiB rewrite), '
                       'per-step CUDA events',
                 'launch': 'CUDA graph replay' if ms_graph <= ms_eager else 'eager'},
      'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': 2 * nw * 8,
              'd2h_bytes_per_step': 2 * nw * 8 + nkl * nb * 8,
              'path': 'pinned H2D + jrb_qr_fwd/jrb_hpsi/jrb_band_expect/jrb_qr_bwd + D2H'},
      'gpu_launches': launches,
      'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                   'frac': achieved / peak, 'traffic': None, 'peak_source': peak_src,
                   'kernel': 'whole band-mode evaluation batch (launch-latency bound at this size)',
                   'bytes_formula': 'nk*nb*(32*N + 48*ng)'},
      'clocks': clk.summary(),
      'band_step': {'eager_ms': ms_eager, 'graph_ms': ms_graph},
    }
    if world == 1 and not args.no_cpu:
      v, t, cores, nks = cpu_sample_band(wl, 4, 2, 1)
      line['cpu_baseline'] = {
        'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
        'sample': f'{nks} of {nk} path points x {nb} bands ({t:.2f} s), one after the other; '
                  'oracle port (torch FP64 autograd)'}
    return line
  return None


def numa_place_for_gpu(local_rank, enable):
  """Prefer the NUMA node of this rank's GPU for the pinned host buffers allocated next (multi-GPU
  e2e: eight ranks copying from one socket's memory share its controllers and the inter-socket
  link).  set_mempolicy(MPOL_PREFERRED) through libc; a box that exposes one node, or forbids the
  call, is left alone.  Returns what was found / done (reported in the line's e2e object)."""
  info = {'nodes_visible': None, 'gpu_node': None, 'action': 'none'}
  try:
    import ctypes
    import torch
    nodes = sorted(int(d[4:]) for d in os.listdir('/sys/devices/system/node') if d.startswith('node')
                   and d[4:].isdigit())
    info['nodes_visible'] = nodes
    pr = torch.cuda.get_device_properties(local_rank)
    path = f'/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0/numa_node'
    node = int(open(path).read().strip())
    info['gpu_node'] = node
    import platform
    if not enable or node < 0 or node not in nodes or len(nodes) < 2 or platform.machine() != 'x86_64':
      return info   # (the syscall number below is the x86-64 one)
    libc = ctypes.CDLL('libc.so.6', use_errno=True)
    mask = ctypes.c_ulong(1 << node)
    SYS_set_mempolicy, MPOL_PREFERRED = 238, 1   # x86-64
    rc = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(64))
    info['action'] = f'set_mempolicy(MPOL_PREFERRED, node {node})' if rc == 0 else \
        f'set_mempolicy failed (errno {ctypes.get_errno()})'
  except Exception as e:  # noqa: BLE001 - diagnostics only
    info['action'] = f'none ({type(e).__name__}: {e})'[:160]
  return info


def numa_place_default():
  try:
    import ctypes
    import platform
    if platform.machine() != 'x86_64':
      return
    ctypes.CDLL('libc.so.6').syscall(238, 0, None, ctypes.c_ulong(0))   # MPOL_DEFAULT
  except Exception:  # noqa: BLE001
    pass


def parse_orbital_grid(text):
  if text in ('auto', 'full'):
    return text
  return tuple(int(v) for v in text.split(','))


def run_b200(args):
  import torch
  import torch.distributed as dist

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  if world != args.gpus:
    if world == 1 and args.gpus > 1:
      raise SystemExit('launch with torchrun --nproc-per-node N for --gpus N > 1')
  torch.cuda.set_device(local_rank)
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
  try:
    # no --config: the headline configuration C2 in full, then the diamond-64 configurations of the
    # same metric (BASELINE.json: "Si8, diamond64 at 1/2/4/8") as extra objects of the same line
    name = args.config or 'C2'
    line = measure(args, name, world, rank, local_rank, full=True)
    if args.config is None and not args.emulate_ranks > 1:
      extra = {}
      for other in ('C3b', 'C3a'):
        # the headline line must survive a failure of an extra configuration (the same exception
        # on every rank: a layout the shapes do not allow, out of memory); it is reported instead
        try:
          sub = measure(args, other, world, rank, local_rank, full=False,
                        steps=max(3, min(args.steps, 10)))
        except Exception as e:  # noqa: BLE001
          import traceback
          traceback.print_exc(file=sys.stderr)
          sub = {'error': f'{type(e).__name__}: {e}'[:400]} if rank == 0 else None
          torch.cuda.empty_cache()
        if sub is not None:
          extra[other] = sub
      if line is not None:
        line['diamond64'] = extra
    if rank == 0 and line is not None:
      emit(line)
  finally:
    if world > 1:
      dist.destroy_process_group()


def measure(args, name, world, rank, local_rank, full=True, steps=None):
  """One workload on this job's GPUs; returns the JSON object (rank 0) or None."""
  wl = build_workload(name)
  nk = wl['kpts'].shape[0]
  if wl['band_mode']:
    return run_b200_band(args, wl, world, rank, local_rank)
  if nk % world != 0:
    return run_b200_rows(args, wl, world, rank, local_rank, steps=steps)
  return run_b200_kshard(args, name, wl, world, rank, local_rank, full, steps)


def run_b200_kshard(args, name, wl, world, rank, local_rank, full, steps=None):
  """Whole k-points per rank (the reference's k-mesh layout); world == 1 is the single-GPU run."""
  import torch
  import torch.distributed as dist
  from jrystal_b200 import _lib, parallel

  steps = steps or args.steps
  c = wl['crystal']
  nk, nb, ng = wl['kpts'].shape[0], wl['nb'], wl['ng']
  ngrid = int(np.prod(wl['grid']))
  b0, b1 = 0, nb
  kpts = wl['kpts']
  sharding = f'k{world}' if world > 1 else 'none'
  emulated = args.emulate_ranks > 1 and world == 1
  if emulated:
    # tuning aid, NOT a bench line: the per-rank workload of an N-GPU k-sharded run on one GPU
    e0, e1 = parallel.shard_kpoints(nk, args.emulate_ranks, 0)
    kpts = kpts[e0:e1]
    sharding = f'EMULATED rank 0 of k{args.emulate_ranks} (no collectives; not a benchmark value)'
  ev = parallel.KShardedEvaluator(c.cell_vectors, wl['mask'], kpts, nb, c.positions, c.charges,
                                  device=local_rank,
                                  orbital_grid=parse_orbital_grid(args.orbital_grid))
  plan = ev.plan
  k0, k1 = ev.k0, ev.k1
  if wl['nproj']:
    plan.set_nonlocal(synthetic_projectors(ng, nk, wl['nproj'], k0, k1, 'cuda'))
  w_re_h, w_im_h = synthetic_params(ng, nk, nb, k0, k1)
  occ_h = np.ascontiguousarray(wl['occ'][:, k0:k1])
  w_re = torch.from_numpy(w_re_h).cuda()
  w_im = torch.from_numpy(w_im_h).cuda()
  occ = torch.from_numpy(occ_h).cuda()
  out = (torch.empty(4, dtype=torch.float64, device='cuda'), torch.empty_like(w_re),
         torch.empty_like(w_im))

  def step():
    # ONE call of the public API per evaluation: jrb_eval (QR, density sweep, the in-library
    # all-reduce over NVLink peer memory when world > 1, grid potential, H-apply, QR adjoint)
    ev.evaluate(w_re, w_im, occ, 'lda_x', out=out)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def timed(fn, n):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(n):
      fn()
    ev1.record()
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device='cuda')
    if world > 1:
      dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())

  # working sets that fit the 126 MB L2 (C1; C2 on 8 GPUs): flush it between timed steps by
  # rewriting a 256 MiB buffer, and time every step with its own event pair (the flush is outside
  # the timed spans)
  flush_l2 = 2 * w_re_h.size * 8 < 126 * 2**20
  flush_buf = torch.zeros(32 * 2**20, dtype=torch.float64, device='cuda') if flush_l2 else None

  def timed_flushed(fn, n):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(n)]
    barrier()
    for e0, e1 in evs:
      flush_buf.add_(1.0)
      e0.record()
      fn()
      e1.record()
    barrier()
    ms = torch.tensor([sum(e0.elapsed_time(e1) for e0, e1 in evs)], dtype=torch.float64,
                      device='cuda')
    if world > 1:
      dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())

  lib = _lib.load()
  for _ in range(max(args.warmup, 3)):
    step()
  plan.check_status()
  barrier()
  launches0 = lib.jrb_launch_count()
  with ClockSampler(local_rank, enabled=(rank == 0)) as clk:
    total_ms = (timed_flushed if flush_l2 else timed)(step, steps)
  launches = int(lib.jrb_launch_count() - launches0)
  ms_eager = total_ms / steps
  # the same evaluation replayed as ONE CUDA graph (what the energy driver does with its whole
  # optimisation step): the ~60 launches of an evaluation are launch-latency bound once a rank
  # holds few k-points.  jrb_eval is capture safe (no allocation, no synchronisation; the peer
  # all-reduce keeps its epoch in device memory).  The better of the two is the reported value.
  ms_graph = None
  if ev.reduce_path != 'nccl' and not args.no_graph:
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
      step()
      with torch.cuda.graph(graph, stream=side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(2):
      graph.replay()
    barrier()
    with ClockSampler(local_rank, enabled=(rank == 0)) as clk_g:
      ms_graph = (timed_flushed if flush_l2 else timed)(graph.replay, steps) / steps
    if ms_graph < ms_eager:
      clk = clk_g
  ms_per_step = ms_eager if ms_graph is None else min(ms_eager, ms_graph)
  value = 1e3 / ms_per_step
  energies = out[0].cpu().numpy().tolist()

  # ---- end to end through host buffers --------------------------------------------------
  nw = w_re_h.size
  h2d = 2 * nw * 8 + occ_h.size * 8
  d2h = 2 * nw * 8 + 4 * 8
  # N > 1: the pinned buffers of a rank go to the NUMA node of its GPU when the box exposes one
  numa = numa_place_for_gpu(local_rank, enable=world > 1 and os.environ.get('JRB_NO_NUMA') != '1')
  pin = lambda a: torch.from_numpy(a).pin_memory()
  w_re_p, w_im_p, occ_p = pin(w_re_h), pin(w_im_h), pin(occ_h)
  en_p = torch.empty(4, dtype=torch.float64).pin_memory()
  g_re_p = torch.empty(w_re_h.shape, dtype=torch.float64).pin_memory()
  g_im_p = torch.empty(w_re_h.shape, dtype=torch.float64).pin_memory()
  if numa['action'].startswith('set_mempolicy(MPOL_PREFERRED'):
    numa_place_default()
  if ev.reduce_path != 'nccl':
    # the reference-facing call with HOST buffers: k-chunked copies overlapped with the kernels,
    # the density all-reduce (world > 1) inside the library
    def e2e_step():
      plan.energy_grad_host(w_re_p, w_im_p, occ_p, 'lda_x', out=(en_p, g_re_p, g_im_p))
    e2e_path = 'jrb_energy_grad_host (C ABI, pinned host buffers' + (
      '; density all-reduce over NVLink peer memory inside)' if world > 1 else ')')
  else:
    def e2e_step():
      w_re.copy_(w_re_p, non_blocking=True)
      w_im.copy_(w_im_p, non_blocking=True)
      occ.copy_(occ_p, non_blocking=True)
      step()
      en_p.copy_(out[0], non_blocking=True)
      g_re_p.copy_(out[1], non_blocking=True)
      g_im_p.copy_(out[2], non_blocking=True)
      torch.cuda.current_stream().synchronize()
    e2e_path = 'pinned H2D + jrb_eval_begin/NCCL all-reduce/jrb_eval_finish + D2H per rank'
  e2e_steps = max(1, min(steps, 5))
  e2e_step()
  barrier()
  t0 = time.perf_counter()
  for _ in range(e2e_steps):
    e2e_step()
  barrier()
  e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device='cuda')
  if world > 1:
    dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
  e2e_value = e2e_steps / float(e2e_s.item())
  e2e_energy = float(en_p.sum())
  # what the host <-> device copies of one step cost with NO kernel in between (H2D then D2H of the
  # same pinned buffers, all ranks at once): the floor the platform's PCIe / host memory puts
  # under e2e, reported next to it
  def copy_only():
    w_re.copy_(w_re_p, non_blocking=True)
    w_im.copy_(w_im_p, non_blocking=True)
    g_re_p.copy_(out[1], non_blocking=True)
    g_im_p.copy_(out[2], non_blocking=True)
    torch.cuda.current_stream().synchronize()
  copy_only()
  barrier()
  t0 = time.perf_counter()
  for _ in range(3):
    copy_only()
  barrier()
  copy_s = torch.tensor([(time.perf_counter() - t0) / 3], dtype=torch.float64, device='cuda')
  if world > 1:
    dist.all_reduce(copy_s, op=dist.ReduceOp.MAX)
  copy_only_ms = float(copy_s.item()) * 1e3

  # ---- phase split (outside the timed region; explains the number) ----------------------
  # Timed INSIDE the evaluation: jrb_plan_phase_timing records CUDA events at the phase boundaries
  # of jrb_eval on the stream its kernels run on, so the H-apply measured here is the sweep the
  # timed steps ran (with the psi(r) cache: k_x_vmul_cached, not the stand-alone jrb_hpsi).
  phases = {}
  reps = 3
  plan.phase_timing(True)
  step()
  acc = dict.fromkeys(plan.PHASES, 0.0)
  for _ in range(reps):
    step()
    for k, v in plan.phase_times().items():
      acc[k] += v / reps
  plan.phase_timing(False)
  phases.update(acc)
  if world > 1 and ev.reduce_path == 'peer':
    rho2 = torch.empty((1,) + tuple(wl['grid']), dtype=torch.float64, device='cuda')
    ebuf = torch.zeros(1, dtype=torch.float64, device='cuda')
    phases['allreduce_rho'] = timed(lambda: plan.allreduce_rho(rho2, ebuf), 10) / 10
    del rho2, ebuf

  # ---- the caller of the path: one optimisation step of the energy-mode driver (evaluation +
  # device Adam), eager and replayed as a CUDA graph; informational
  driver = None
  if world == 1 and full:
    from jrystal_b200.optim import Adam
    pw_re, pw_im = w_re.clone(), w_im.clone()
    opt = Adam([pw_re, pw_im])
    rho_d = torch.empty((1,) + tuple(wl['grid']), dtype=torch.float64, device='cuda')

    def opt_step():
      plan.eval(pw_re, pw_im, occ, 'lda_x', out=out, rho=rho_d)
      opt.step([out[1], out[2]])

    opt_step()
    eager_ms = timed(opt_step, reps) / reps
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
      opt_step()
    graph.replay()
    graph_ms = timed(graph.replay, reps) / reps
    driver = {'eager_steps_per_s': 1e3 / eager_ms, 'graph_steps_per_s': 1e3 / graph_ms,
              'what': 'jrb_eval + jrb_adam_tick/apply per step'}
    del pw_re, pw_im, opt, graph, rho_d

  # ---- roofline (a dense 3-D transform is charged one read + one write of its box,
  # sphere data its true size), per GPU -------------------------------------------------------
  m_local = (k1 - k0) * (b1 - b0)
  peak, peak_src = measured_peak()
  # dominant kernel group: the H-apply sweep k_yx_vmul + k_z_fwd_gather on the z-transformed
  # columns kept from the density sweep (k_yx_vmul alone is ~37 % of the step, profiles/); it
  # performs the backward dense transform of every orbital and reads Q / writes HQ on the sphere.
  # The algorithmic bytes are the contract figure of SURVEY 8d on the reference's own grid N,
  # whatever box the orbitals are transformed on (config.orbital_grid).
  happly_bytes = m_local * (32.0 * ngrid + 32.0 * ng)
  happly_achieved = happly_bytes / (phases['hpsi'] * 1e-3) / 1e9
  bytes_alg = 64.0 * m_local * (ngrid + ng)
  whole_achieved = bytes_alg / (ms_per_step * 1e-3) / 1e9
  fft_ms = phases['density'] + phases['hpsi']
  fft_bytes = 64.0 * m_local * ngrid
  traffic, traffic_note = measured_traffic(name, world)
  # FP64 ceiling: what ncu says binds the path (the sweeps keep psi(r) in shared memory; DRAM runs
  # at ~10 % of peak, the FP64 pipe at 45-60 %, the shared-memory pipe at 50-65 %)
  psi_cache = plan.psi_cache_bytes > 0
  fft_fl, qr_fl = fp64_flops(wl['mask'], plan.orbital_grid, m_local, k1 - k0, ng, nb, psi_cache)
  fp64_tf = (fft_fl + qr_fl) / (ms_per_step * 1e-3) / 1e12
  # the design's own compulsory bytes of the sweep: with the psi(r) cache it streams psi(r) of the
  # orbital box once, writes and re-reads the z columns, reads Q and writes HQ
  ncol = int(np.asarray(wl['mask']).any(axis=2).sum())
  og = [int(v) for v in plan.orbital_grid]
  design_bytes = m_local * ((16.0 * og[0] * og[1] * og[2] if psi_cache else 16.0 * ncol * og[2]) +
                            32.0 * ncol * og[2] + 32.0 * ng)
  measured = None
  if traffic:
    measured = {'bytes_per_launch': traffic, 'achieved': traffic / (phases['hpsi'] * 1e-3) / 1e9,
                'frac': traffic / (phases['hpsi'] * 1e-3) / 1e9 / peak,
                'what': 'dram__bytes_read + dram__bytes_write of the sweep (roofline.traffic) over '
                        'its time in this run: the fraction of the measured HBM copy bandwidth the '
                        'sweep really sustains'}
  roofline = {
    'bound': 'hbm' if psi_cache else 'fp64', 'achieved': happly_achieved, 'peak': peak, 'unit': 'GB/s',
    'frac': happly_achieved / peak, 'traffic': traffic, 'traffic_source': traffic_note,
    'peak_source': peak_src,
    'bound_note': ('ncu (profiles/r02_kernel_tables.md): with the psi(r) cache the dominant sweep '
                   'streams psi(r) from HBM at ~0.74 of the measured copy bandwidth while its L1 data '
                   'pipe (shared-memory exchanges of the line FFTs + that global stream) is ~76 % busy: '
                   'HBM and the L1 pipe bind it together; the QR products are bound by the FP64 tensor '
                   'pipe.  achieved/peak/frac keep the HBM-convention contract figure of SURVEY 8d '
                   '(algorithmic bytes of the REFERENCE grid: one read + one write of the dense box per '
                   'transform), which exceeds 1 because the pruned passes on the smaller orbital box '
                   'never move those bytes; roofline.measured is the sustained fraction on the bytes '
                   'really moved, roofline.design_bytes what this design must move, roofline.fp64 the '
                   'fraction of the FP64 roof over the flops executed'
                   if psi_cache else
                   'ncu (profiles/): without the psi(r) cache the plane kernels are bound by the '
                   'L1/shared-memory data pipe and FP64 issue, the QR products by the FP64 tensor pipe; '
                   'none by HBM.  achieved/peak/frac keep the HBM-convention contract figure of SURVEY '
                   '8d, roofline.fp64 is the fraction of the FP64 roof over the flops executed'),
    'measured': measured,
    'design_bytes': {'bytes_per_launch': design_bytes,
                     'formula': 'M*(16*nxw*nyw*nzw [psi(r) cache read] + 32*ncol*nzw [z columns '
                                'written and re-read] + 32*ng [Q read, HQ write])' if psi_cache else
                                'M*(48*ncol*nzw [kept z columns read, written, re-read] + 32*ng)',
                     'achieved': design_bytes / (phases['hpsi'] * 1e-3) / 1e9,
                     'frac': design_bytes / (phases['hpsi'] * 1e-3) / 1e9 / peak},
    'kernel': ('H-apply sweep (k_x_vmul_cached on psi(r) of the density sweep + k_z_fwd_gather)'
               if psi_cache else
               'H-apply sweep (k_yx_vmul + k_z_fwd_gather on the kept z columns; k_yx_vmul is the '
               'dominant kernel)'),
    'psi_cache_bytes': plan.psi_cache_bytes,
    'bytes_per_launch': happly_bytes,
    'bytes_formula': 'M*(32*N + 32*ng): one dense transform (read+write of the box) per orbital '
                     '+ Q read + HQ write, SURVEY 8d',
    'ms_per_launch': phases['hpsi'],
    'fp64': {'flops_per_eval': fft_fl + qr_fl, 'fft_flops': fft_fl, 'qr_flops': qr_fl,
             'achieved': fp64_tf, 'peak': FP64_PEAK_TFLOPS, 'unit': 'TFLOP/s',
             'frac': fp64_tf / FP64_PEAK_TFLOPS, 'peak_source': FP64_PEAK_SOURCE,
             'flops_formula': 'pruned line FFTs at 5 n log2 n (2 z + %d y + %d x passes per orbital on '
                              'the orbital box) + 4 x 8 ng nb^2 per (spin, k) for the QR products '
                              % ((2, 2) if psi_cache else (3, 3)) +
                              '(Hermitian / triangular halves not charged); whole evaluation over '
                              'ms_per_step'},
    'whole_evaluation': {'achieved': whole_achieved, 'frac': whole_achieved / peak,
                         'bytes': '64*M*(N+ng)'},
    'fft_density_path': {'ms': fft_ms, 'achieved': fft_bytes / (fft_ms * 1e-3) / 1e9,
                         'frac': fft_bytes / (fft_ms * 1e-3) / 1e9 / peak,
                         'bytes': '64*M*N (two dense 3-D transforms per orbital)'},
  }

  line = None
  if rank == 0:
    line = {
      'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': steps,
      'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True,
      'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
      'config': {'workload': wl['text'], 'orbitals': nk * nb, 'ng': ng, 'grid': wl['grid'],
                 'sharding': sharding, 'reduce_path': ev.reduce_path, 'xc': 'lda_x',
                 'orbital_grid': list(plan.orbital_grid),
                 'orbital_grid_note': 'box of the per-orbital FFTs (alias-free, n >= 4 gmax + 1 = '
                                      f'{list(plan.min_orbital_grid)}); rho, potentials and all '
                                      'results live on `grid`',
                 'l2': f'inputs larger than L2 ({2 * nw * 8 / 2**20:.0f} MiB of parameters per '
                       'GPU); no flush' if not flush_l2 else
                       'working set fits L2: L2 flushed between timed steps (256 MiB rewrite), '
                       'per-step CUDA events',
                 'gradient_pin': GRADIENT_PIN,
                 'launch': ('CUDA graph replay of jrb_eval' if ms_graph is not None
                            and ms_graph < ms_eager else 'eager (one jrb_eval call per step)'),
                 'eager_ms': ms_eager, 'graph_ms': ms_graph,
                 'batch_groups': int(os.environ.get('JRB_BATCH_GROUPS', 0))},
      'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
              'd2h_bytes_per_step': d2h, 'path': e2e_path, 'ms_per_step': 1e3 / e2e_value,
              'copies_alone_ms': copy_only_ms, 'host_numa': numa,
              'copies_alone_note': 'the same H2D + D2H with no kernel in between, all ranks at once: '
                                   'the PCIe / host-memory floor of this box under e2e',
              'energy_rel_diff_vs_device_path': abs(e2e_energy - sum(energies)) / abs(sum(energies))},
      'gpu_launches': launches, 'roofline': roofline, 'clocks': clk.summary(),
      'phases_ms': phases, 'driver_step': driver, 'energies_ha': energies,
      'workspace_mib': plan.workspace_bytes / 2**20,
    }
    if world == 1 and full and not args.no_cpu and not emulated:
      # bounded sample: ~10-20 s of CPU work
      nks = {'C1': 8, 'C2': 8, 'C3a': 1, 'C3b': 1, 'C4': 8}[name]
      v, t, cores, nks = cpu_sample_eval(wl, nks, 2 if name != 'C1' else 20, 1)
      line['cpu_baseline'] = {
        'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
        'sample': f'{nks} of {nk} k-points x {nb} bands ({t:.2f} s), scaled by {nk / nks:g}; '
                  'oracle port (torch FP64 autograd); `--impl reference` runs all k-points'}
  del ev, plan, w_re, w_im, occ, out, w_re_p, w_im_p, g_re_p, g_im_p
  torch.cuda.empty_cache()
  return line


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=10)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--config', default=None, choices=list(WORKLOADS),
                  help='one workload; default: C2 + the diamond-64 configurations C3b, C3a')
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
  ap.add_argument('--no-graph', action='store_true', help='do not try the CUDA-graph replay')
  ap.add_argument('--orbital-grid', default=os.environ.get('JRB_ORBITAL_GRID', 'auto'),
                  help="box of the per-orbital FFTs: 'auto' (smallest efficient alias-free box; default), "
                       "'full' (the reference's own grid) or nx,ny,nz; results do not depend on it")
  ap.add_argument('--emulate-ranks', type=int, default=1,
                  help='tuning aid: run only the k-points rank 0 of an N-GPU run would own')
  args = ap.parse_args()
  _route_stdout_to_stderr()
  if args.impl == 'reference':
    args.config = args.config or 'C2'
    run_reference(args)
  else:
    run_b200(args)


if __name__ == '__main__':
  main()
