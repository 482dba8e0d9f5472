"""Host-side handle of a `jrb_plan`: owns the C plan, checks shapes, passes torch CUDA
tensors (device memory + stream plumbing only) through the C ABI.

One plan per (device, grid, mask, k-points, bands): it is what the reference's drivers
build before their loop (calc/calc_ground_state_energy_all_electrons.py:93-106).
"""
import ctypes
from typing import Optional

import numpy as np
import torch

from . import _lib


def _ptr(t: Optional[torch.Tensor]):
  return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(device=None):
  """torch's current stream ON `device` (the plan's device, which need not be torch's current
  one: the C side calls cudaSetDevice(plan.device) and the stream must belong to it)."""
  return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check_tensor(t, shape, dtype, name, device):
  if not isinstance(t, torch.Tensor) or not t.is_cuda:
    raise TypeError(f'{name} must be a CUDA tensor')
  if t.device.index != device:
    raise ValueError(f'{name} is on {t.device}, plan is on cuda:{device}')
  if tuple(t.shape) != tuple(shape):
    raise ValueError(f'{name} has shape {tuple(t.shape)}, expected {tuple(shape)}')
  if t.dtype != dtype:
    raise TypeError(f'{name} has dtype {t.dtype}, expected {dtype}')
  if not t.is_contiguous():
    raise ValueError(f'{name} must be contiguous')
  return t


class _CommMixin:
  """Plan-owned communicator over NVLink peer memory (jrb_comm_*; include/jrystal_b200.h)."""

  comm_world = 1

  def comm_init(self, group=None, capacity: int = 0) -> bool:
    """Create this rank's symmetric region, exchange the CUDA IPC handles over torch.distributed
    and map the peers.  Returns False (and leaves the plan without a communicator) when peer
    memory cannot be set up -- callers then keep the NCCL all-reduce.  JRB_NO_PEER=1 disables it."""
    import os
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
      return False
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1 or os.environ.get('JRB_NO_PEER', '0') not in ('', '0'):
      return False
    nbytes = int(self.lib.jrb_comm_handle_bytes())
    mine = (ctypes.c_ubyte * nbytes)()
    ok = self.lib.jrb_comm_create(self._h, rank, world, int(capacity), mine) == 0
    send = torch.zeros(nbytes + 1, dtype=torch.uint8, device=self.tdev)
    if ok:
      send[:nbytes] = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(self.tdev)
      send[nbytes] = 1
    recv = torch.empty(world * (nbytes + 1), dtype=torch.uint8, device=self.tdev)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.cpu().view(world, nbytes + 1)
    ok = bool(recv[:, nbytes].all())
    if ok:
      handles = recv[:, :nbytes].contiguous().numpy().tobytes()
      ok = self.lib.jrb_comm_connect(self._h, handles) == 0
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.tdev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)   # all ranks or none
    ok = bool(flag.item())
    self.comm_world = world if ok else 1
    self.comm_error = None if ok else _lib.load().jrb_last_error()
    return ok

  def allreduce(self, buf):
    """In-place SUM over the ranks of a contiguous float64 / complex128 CUDA tensor."""
    if not buf.is_cuda or not buf.is_contiguous() or buf.dtype not in (torch.float64, torch.complex128):
      raise TypeError('allreduce needs a contiguous float64 / complex128 CUDA tensor')
    if buf.device.index != self.device:
      raise ValueError(f'buffer is on {buf.device}, plan is on cuda:{self.device}')
    n = buf.numel() * (2 if buf.is_complex() else 1)
    _lib.check(self.lib.jrb_allreduce(self._h, _ptr(buf), n, _stream(self.device)))
    return buf


class Plan(_CommMixin):

  def __init__(self, cell_vectors, freq_mask, kpts, num_bands: int, num_spin: int = 1,
               device: Optional[int] = None, batch_groups: int = 0, orbital_grid=None):
    if not torch.cuda.is_available():
      raise RuntimeError('jrystal_b200 needs a CUDA device (there is no CPU fallback)')
    self.lib = _lib.load()
    self.device = torch.cuda.current_device() if device is None else int(device)
    self.cell = np.ascontiguousarray(np.asarray(cell_vectors, dtype=np.float64).reshape(3, 3))
    mask = np.asarray(freq_mask)
    if mask.ndim != 3:
      raise ValueError(f'freq_mask must be 3-D, got shape {mask.shape}')
    self.mask = np.ascontiguousarray(mask.astype(np.uint8))
    self.kpts = np.ascontiguousarray(np.asarray(kpts, dtype=np.float64).reshape(-1, 3))
    self.nx, self.ny, self.nz = (int(v) for v in mask.shape)
    self.ns, self.nk, self.nb = int(num_spin), int(self.kpts.shape[0]), int(num_bands)
    self.ngrid = self.nx * self.ny * self.nz
    self.vol = float(abs(np.linalg.det(self.cell)))
    desc = _lib.PlanDesc(
      self.nx, self.ny, self.nz, self.ns, self.nk, self.nb,
      self.mask.ctypes.data, self.kpts.ctypes.data, self.cell.ctypes.data,
      self.device, int(batch_groups)
    )
    handle = ctypes.c_void_p()
    _lib.check(self.lib.jrb_plan_create(ctypes.byref(desc), ctypes.byref(handle)))
    self._h = handle
    self.ng = int(self.lib.jrb_plan_num_g(self._h))
    self.tdev = torch.device('cuda', self.device)
    self._atoms = False
    self.orbital_grid = (self.nx, self.ny, self.nz)
    if orbital_grid is not None:
      self.set_orbital_grid(orbital_grid)

  # -- orbital grid -----------------------------------------------------------------
  @property
  def min_orbital_grid(self):
    """Smallest alias-free box per axis, 4 gmax + 1."""
    dims = (ctypes.c_int32 * 3)()
    _lib.check(self.lib.jrb_plan_min_orbital_grid(self._h, dims))
    return tuple(int(v) for v in dims)

  def auto_orbital_grid(self):
    """The box `orbital_grid='auto'` picks (measured on B200, profiles/r01_orbital_grid.md): z, the
    axis the fused y+x plane kernels loop over, shrinks to the smallest length >= 4 gmax + 1 with
    a light line plan; x and y keep the plan's lengths (the fused kernels are tuned for 64 and
    128) except 128 -> 81 when that fits and the fused 81 x 81 kernels have the shared memory."""
    from .grid import orbital_grid_candidates
    return orbital_grid_candidates((self.nx, self.ny, self.nz), self.min_orbital_grid)

  def set_orbital_grid(self, dims):
    """Run the per-orbital transforms on a smaller alias-free box (jrb_plan_set_orbital_grid);
    'auto' = auto_orbital_grid(), 'full' / None = the plan's grid."""
    if dims is None or dims == 'full':
      dims = (self.nx, self.ny, self.nz)
    elif dims == 'auto':
      cands = self.auto_orbital_grid()
      for cand in cands[:-1]:  # boxes that only pay with the fused plane kernels
        if self.lib.jrb_plan_set_orbital_grid(self._h, *cand) != 0:
          continue  # e.g. no memory for this box: the next candidate is smaller in work space
        if self.lib.jrb_plan_orbital_fused(self._h) > 0:
          self.orbital_grid = cand
          return
      dims = cands[-1]
    dims = tuple(int(v) for v in dims)
    if len(dims) != 3:
      raise ValueError('orbital_grid must be three axis lengths, "auto" or "full"')
    _lib.check(self.lib.jrb_plan_set_orbital_grid(self._h, *dims))
    self.orbital_grid = dims

  PHASES = ('qr_fwd', 'density', 'reduce_interp', 'grid_potential', 'hpsi', 'qr_bwd')

  def phase_timing(self, enable=True):
    """Record CUDA events at the phase boundaries of every following eval (jrb_plan_phase_timing).
    Not while capturing a CUDA graph."""
    _lib.check(self.lib.jrb_plan_phase_timing(self._h, 1 if enable else 0))

  def phase_times(self):
    """{phase: ms} of the last eval, timed inside it (jrb_plan_phase_times); waits for it."""
    import ctypes
    buf = (ctypes.c_double * 6)()
    _lib.check(self.lib.jrb_plan_phase_times(self._h, ctypes.cast(buf, ctypes.c_void_p)))
    return dict(zip(self.PHASES, [float(v) for v in buf]))

  @property
  def psi_cache_bytes(self):
    """Size of the psi(r) cache of the current orbital box (jrb_plan_psi_cache_bytes); 0 = the
    H-apply of eval_finish repeats the inverse transforms."""
    return int(self.lib.jrb_plan_psi_cache_bytes(self._h))

  def __del__(self):
    h = getattr(self, '_h', None)
    if h:
      self.lib.jrb_plan_destroy(h)
      self._h = None

  # -- shapes -----------------------------------------------------------------------
  @property
  def sphere_shape(self):
    return (self.ns, self.nk, self.ng, self.nb)

  @property
  def workspace_bytes(self):
    return int(self.lib.jrb_plan_workspace_bytes(self._h))

  def _chk(self, t, shape, dtype, name):
    return _check_tensor(t, shape, dtype, name, self.device)

  def _out(self, out, specs):
    """Caller-provided output tensors: every element is held to the same contract as the inputs
    (a wrong shape / device / stride would be an out-of-bounds device write, not a Python error).
    specs: (shape, dtype, name) per element; returns fresh tensors when out is None."""
    if out is None:
      return tuple(self._new(shape, dtype) for shape, dtype, _ in specs)
    if isinstance(out, torch.Tensor):
      out = (out,)
    if len(out) != len(specs):
      raise ValueError(f'out must hold {len(specs)} tensors ({[n for _, _, n in specs]})')
    return tuple(self._chk(t, shape, dtype, 'out.' + name) for t, (shape, dtype, name) in zip(out, specs))

  @property
  def grid_shape(self):
    return (self.ns, self.nx, self.ny, self.nz)

  @property
  def small_shape(self):
    return (self.ns, self.nk, self.nb, self.nb)

  @property
  def band_shape(self):
    return (self.ns, self.nk, self.nb)

  def _new(self, shape, dtype):
    return torch.empty(shape, dtype=dtype, device=self.tdev)

  # -- setup ------------------------------------------------------------------------
  def set_atoms(self, positions, charges):
    pos = np.ascontiguousarray(np.asarray(positions, dtype=np.float64).reshape(-1, 3))
    chg = np.ascontiguousarray(np.asarray(charges, dtype=np.float64).reshape(-1))
    if pos.shape[0] != chg.shape[0]:
      raise ValueError('positions and charges disagree on the number of atoms')
    _lib.check(self.lib.jrb_set_atoms(self._h, pos.ctypes.data, chg.ctypes.data,
                                      pos.shape[0], _stream(self.device)))
    self._atoms = True
    self._natoms = int(pos.shape[0])

  def set_nonlocal(self, phi):
    """Projectors of the non-local pseudopotential on the sphere: complex128 CUDA tensor
    (nk, nproj, ng) = potential_nl_psi_reciprocal[..., mask] (nloc.py:60-141); None removes them."""
    if phi is None:
      _lib.check(self.lib.jrb_set_nonlocal(self._h, None, 0, _stream(self.device)))
      self.nproj = 0
      return
    if phi.ndim != 3 or phi.shape[0] != self.nk or phi.shape[2] != self.ng:
      raise ValueError(f'phi must have shape (nk={self.nk}, nproj, ng={self.ng}), got {tuple(phi.shape)}')
    self._chk(phi, tuple(phi.shape), torch.complex128, 'phi')
    _lib.check(self.lib.jrb_set_nonlocal(self._h, _ptr(phi), int(phi.shape[1]), _stream(self.device)))
    self.nproj = int(phi.shape[1])

  def nonlocal_energy(self, q, occ):
    self._chk(q, self.sphere_shape, torch.complex128, 'q')
    self._chk(occ, (self.ns, self.nk, self.nb), torch.float64, 'occupation')
    e = self._new((1,), torch.float64)
    _lib.check(self.lib.jrb_nonlocal_energy(self._h, _ptr(q), _ptr(occ), _ptr(e), _stream(self.device)))
    return e

  def external_position_gradient(self, rho):
    """dE_ext / d position, (natoms, 3), for the atoms of set_atoms."""
    self._chk(rho, (self.ns, self.nx, self.ny, self.nz), torch.float64, 'density')
    if not getattr(self, '_natoms', 0):
      raise RuntimeError('call set_atoms(positions, charges) first')
    g = self._new((self._natoms, 3), torch.float64)
    _lib.check(self.lib.jrb_external_position_gradient(self._h, _ptr(rho), _ptr(g), _stream(self.device)))
    return g

  def set_external_potential(self, vhat):
    """V(G) of the external / local pseudopotential term, complex128 CUDA tensor (nx, ny, nz) in
    the convention of potential.external_reciprocal (potential.py:153-166)."""
    self._chk(vhat, (self.nx, self.ny, self.nz), torch.complex128, 'vhat')
    _lib.check(self.lib.jrb_set_external_potential(self._h, _ptr(vhat), _stream(self.device)))
    self._atoms = True
    self._natoms = 0

  def set_kpoints(self, kpts):
    """Same number of k-points, new vectors (band-structure walk along a k-path)."""
    k = np.ascontiguousarray(np.asarray(kpts, dtype=np.float64).reshape(-1, 3))
    if k.shape[0] != self.nk:
      raise ValueError(f'expected {self.nk} k-points, got {k.shape[0]}')
    _lib.check(self.lib.jrb_set_kpoints(self._h, k.ctypes.data, _stream(self.device)))
    self.kpts = k

  def check_status(self):
    """Synchronise and raise if an asynchronous call failed numerically (Cholesky breakdown)."""
    _lib.check(self.lib.jrb_check_status(self._h, _stream(self.device)))

  # -- orthonormalisation -------------------------------------------------------------
  def qr_fwd(self, w_re, w_im, out=None):
    self._chk(w_re, self.sphere_shape, torch.float64, 'w_re')
    self._chk(w_im, self.sphere_shape, torch.float64, 'w_im')
    q, r = self._out(out, [(self.sphere_shape, torch.complex128, 'q'),
                           (self.small_shape, torch.complex128, 'r')])
    _lib.check(self.lib.jrb_qr_fwd(self._h, _ptr(w_re), _ptr(w_im), _ptr(q), _ptr(r), _stream(self.device)))
    return q, r

  def qr_bwd(self, q, r, gq, out=None):
    self._chk(q, self.sphere_shape, torch.complex128, 'q')
    self._chk(r, (self.ns, self.nk, self.nb, self.nb), torch.complex128, 'r')
    self._chk(gq, self.sphere_shape, torch.complex128, 'gq')
    g_re, g_im = self._out(out, [(self.sphere_shape, torch.float64, 'g_re'),
                                 (self.sphere_shape, torch.float64, 'g_im')])
    _lib.check(self.lib.jrb_qr_bwd(self._h, _ptr(q), _ptr(r), _ptr(gq), _ptr(g_re), _ptr(g_im),
                                   _stream(self.device)))
    return g_re, g_im

  # -- sphere <-> box ---------------------------------------------------------------
  def expand(self, q):
    self._chk(q, self.sphere_shape, torch.complex128, 'q')
    out = self._new((self.ns, self.nk, self.nb, self.nx, self.ny, self.nz), torch.complex128)
    _lib.check(self.lib.jrb_expand(self._h, _ptr(q), _ptr(out), _stream(self.device)))
    return out

  def squeeze(self, coeff_dense):
    self._chk(coeff_dense, (self.ns, self.nk, self.nb, self.nx, self.ny, self.nz),
              torch.complex128, 'coeff')
    q = self._new(self.sphere_shape, torch.complex128)
    _lib.check(self.lib.jrb_squeeze(self._h, _ptr(coeff_dense), _ptr(q), _stream(self.device)))
    return q

  # -- hot path pieces --------------------------------------------------------------
  def density(self, q, occ, out=None):
    self._chk(q, self.sphere_shape, torch.complex128, 'q')
    self._chk(occ, (self.ns, self.nk, self.nb), torch.float64, 'occupation')
    rho, = self._out(out, [(self.grid_shape, torch.float64, 'rho')])
    _lib.check(self.lib.jrb_density(self._h, _ptr(q), _ptr(occ), _ptr(rho), _stream(self.device)))
    return rho

  def kinetic(self, q):
    self._chk(q, self.sphere_shape, torch.complex128, 'q')
    t = self._new((self.ns, self.nk, self.nb), torch.float64)
    _lib.check(self.lib.jrb_kinetic(self._h, _ptr(q), _ptr(t), _stream(self.device)))
    return t

  def grid_potential(self, rho, xc: str = 'lda_x', kohn_sham: bool = False, out=None):
    if not self._atoms:
      raise RuntimeError('call set_atoms(positions, charges) first')
    self._chk(rho, (self.ns, self.nx, self.ny, self.nz), torch.float64, 'density')
    if xc not in _lib.XC_IDS:
      raise NotImplementedError(f'xc "{xc}" is not implemented (implemented: {list(_lib.XC_IDS)})')
    en, veff = self._out(out, [((3,), torch.float64, 'energies'),
                               (self.grid_shape, torch.float64, 'veff')])
    _lib.check(self.lib.jrb_grid_potential(self._h, _ptr(rho), _lib.XC_IDS[xc],
                                           int(bool(kohn_sham)), _ptr(en), _ptr(veff), _stream(self.device)))
    return en, veff

  def potential(self, rho, xc: str = 'lda_x', kohn_sham: bool = False, parts: int = 7):
    """potential.effective with the reference's semantics (real part); parts: 1 Hartree,
    2 external, 4 xc (bitmask)."""
    if not self._atoms:
      raise RuntimeError('call set_atoms(positions, charges) first')
    self._chk(rho, (self.ns, self.nx, self.ny, self.nz), torch.float64, 'density')
    if xc not in _lib.XC_IDS:
      raise NotImplementedError(f'xc "{xc}" is not implemented (implemented: {list(_lib.XC_IDS)})')
    v = self._new((self.ns, self.nx, self.ny, self.nz), torch.float64)
    _lib.check(self.lib.jrb_potential(self._h, _ptr(rho), _lib.XC_IDS[xc], int(bool(kohn_sham)),
                                      int(parts), _ptr(v), _stream(self.device)))
    return v

  def density_reciprocal(self, rho):
    self._chk(rho, (self.ns, self.nx, self.ny, self.nz), torch.float64, 'density')
    out = self._new((self.ns, self.nx, self.ny, self.nz), torch.complex128)
    _lib.check(self.lib.jrb_density_reciprocal(self._h, _ptr(rho), _ptr(out), _stream(self.device)))
    return out

  def wave_grid(self, q):
    self._chk(q, self.sphere_shape, torch.complex128, 'q')
    out = self._new((self.ns, self.nk, self.nb, self.nx, self.ny, self.nz), torch.complex128)
    _lib.check(self.lib.jrb_wave_grid(self._h, _ptr(q), _ptr(out), _stream(self.device)))
    return out

  def prepare_potential(self, veff):
    """Fix the potential of the following hpsi(q, None) calls (jrb_hpsi_prepare): copied into plan
    work space and resampled onto the orbital grid once."""
    self._chk(veff, (self.ns, self.nx, self.ny, self.nz), torch.float64, 'veff')
    _lib.check(self.lib.jrb_hpsi_prepare(self._h, _ptr(veff), _stream(self.device)))

  def hpsi(self, q, veff, out=None):
    """veff=None applies the potential of the last prepare_potential()."""
    self._chk(q, self.sphere_shape, torch.complex128, 'q')
    if veff is not None:
      self._chk(veff, (self.ns, self.nx, self.ny, self.nz), torch.float64, 'veff')
    hq, = self._out(out, [(self.sphere_shape, torch.complex128, 'hq')])
    _lib.check(self.lib.jrb_hpsi(self._h, _ptr(q), _ptr(veff), _ptr(hq), _stream(self.device)))
    return hq

  def band_expect(self, q, hq, out=None):
    self._chk(q, self.sphere_shape, torch.complex128, 'q')
    self._chk(hq, self.sphere_shape, torch.complex128, 'hq')
    eps, = self._out(out, [(self.band_shape, torch.float64, 'eps')])
    _lib.check(self.lib.jrb_band_expect(self._h, _ptr(q), _ptr(hq), _ptr(eps), _stream(self.device)))
    return eps

  def overlap(self, q, hq):
    """H[s,k,i,j] = <q_i|hq_j> (Hermitian), jrb_hamiltonian_matrix."""
    self._chk(q, self.sphere_shape, torch.complex128, 'q')
    self._chk(hq, self.sphere_shape, torch.complex128, 'hq')
    h = self._new((self.ns, self.nk, self.nb, self.nb), torch.complex128)
    _lib.check(self.lib.jrb_hamiltonian_matrix(self._h, _ptr(q), _ptr(hq), _ptr(h), _stream(self.device)))
    return h

  def fft3d(self, x, inverse: bool, out=None):
    if x.dtype != torch.complex128 or not x.is_cuda or not x.is_contiguous():
      raise TypeError('fft3d needs a contiguous complex128 CUDA tensor')
    if x.ndim < 3:
      raise ValueError(f'Input must have at least 3 dimensions, got {x.ndim}')
    if tuple(x.shape[-3:]) != (self.nx, self.ny, self.nz):
      raise ValueError(f'last three axes {tuple(x.shape[-3:])} do not match the plan grid')
    out = torch.empty_like(x) if out is None else self._chk(out, tuple(x.shape), torch.complex128,
                                                            'out')
    if x.device.index != self.device:
      raise ValueError(f'x is on {x.device}, plan is on cuda:{self.device}')
    batch = int(np.prod(x.shape[:-3])) if x.ndim > 3 else 1
    _lib.check(self.lib.jrb_fft3d(self._h, _ptr(x), _ptr(out),
                                  _lib.FFT_INVERSE if inverse else _lib.FFT_FORWARD, batch,
                                  _stream(self.device)))
    return out

  # -- fused evaluation ---------------------------------------------------------------
  def eval_begin(self, w_re, w_im, occ, rho=None, e_kin=None):
    self._chk(w_re, self.sphere_shape, torch.float64, 'w_re')
    self._chk(w_im, self.sphere_shape, torch.float64, 'w_im')
    self._chk(occ, (self.ns, self.nk, self.nb), torch.float64, 'occupation')
    rho = self._new(self.grid_shape, torch.float64) if rho is None else self._chk(
      rho, self.grid_shape, torch.float64, 'rho')
    e_kin = self._new((1,), torch.float64) if e_kin is None else self._chk(
      e_kin, (1,), torch.float64, 'e_kin')
    _lib.check(self.lib.jrb_eval_begin(self._h, _ptr(w_re), _ptr(w_im), _ptr(occ), _ptr(rho),
                                       _ptr(e_kin), _stream(self.device)))
    return rho, e_kin

  def eval_finish(self, occ, rho, e_kin, xc: str = 'lda_x', want_occ_grad: bool = False,
                  out=None):
    if not self._atoms:
      raise RuntimeError('call set_atoms(positions, charges) first')
    if xc not in _lib.XC_IDS:
      raise NotImplementedError(f'xc "{xc}" is not implemented (implemented: {list(_lib.XC_IDS)})')
    # occ must be THIS plan's k-shard (ns, nk_local, nb), not the global occupation array
    self._chk(occ, self.band_shape, torch.float64, 'occupation')
    self._chk(rho, self.grid_shape, torch.float64, 'rho')
    self._chk(e_kin, (1,), torch.float64, 'e_kin')
    energies, g_re, g_im = self._out(out, [((4,), torch.float64, 'energies'),
                                           (self.sphere_shape, torch.float64, 'g_re'),
                                           (self.sphere_shape, torch.float64, 'g_im')])
    g_occ = self._new(self.band_shape, torch.float64) if want_occ_grad else None
    _lib.check(self.lib.jrb_eval_finish(self._h, _ptr(occ), _ptr(rho), _ptr(e_kin),
                                        _lib.XC_IDS[xc], _ptr(energies), _ptr(g_re), _ptr(g_im),
                                        _ptr(g_occ), _stream(self.device)))
    return energies, g_re, g_im, g_occ

  def allreduce_rho(self, rho, e_kin=None):
    """jrb_allreduce_rho: in-place SUM over the ranks of the partial density (+ E_kin)."""
    self._chk(rho, self.grid_shape, torch.float64, 'rho')
    if e_kin is not None:
      self._chk(e_kin, (1,), torch.float64, 'e_kin')
    _lib.check(self.lib.jrb_allreduce_rho(self._h, _ptr(rho), _ptr(e_kin), _stream(self.device)))

  def eval(self, w_re, w_im, occ, xc: str = 'lda_x', want_occ_grad: bool = False, out=None,
           rho=None):
    """jrb_eval: the whole evaluation in one call; with a connected communicator the partial
    density is all-reduced inside the library (on the orbital grid)."""
    if not self._atoms:
      raise RuntimeError('call set_atoms(positions, charges) first')
    if xc not in _lib.XC_IDS:
      raise NotImplementedError(f'xc "{xc}" is not implemented (implemented: {list(_lib.XC_IDS)})')
    self._chk(w_re, self.sphere_shape, torch.float64, 'w_re')
    self._chk(w_im, self.sphere_shape, torch.float64, 'w_im')
    self._chk(occ, self.band_shape, torch.float64, 'occupation')
    energies, g_re, g_im = self._out(out, [((4,), torch.float64, 'energies'),
                                           (self.sphere_shape, torch.float64, 'g_re'),
                                           (self.sphere_shape, torch.float64, 'g_im')])
    rho = self._new(self.grid_shape, torch.float64) if rho is None else self._chk(
      rho, self.grid_shape, torch.float64, 'rho')
    g_occ = self._new(self.band_shape, torch.float64) if want_occ_grad else None
    _lib.check(self.lib.jrb_eval(self._h, _ptr(w_re), _ptr(w_im), _ptr(occ), _lib.XC_IDS[xc],
                                 _ptr(energies), _ptr(g_re), _ptr(g_im), _ptr(g_occ), _ptr(rho),
                                 _stream(self.device)))
    return energies, g_re, g_im, g_occ, rho

  def energy_grad_host(self, w_re, w_im, occ, xc: str = 'lda_x', out=None, want_rho=False):
    """Host (numpy / pinned torch CPU) buffers in and out through jrb_energy_grad_host."""
    if not self._atoms:
      raise RuntimeError('call set_atoms(positions, charges) first')

    def host_ptr(a):
      if isinstance(a, torch.Tensor):
        assert not a.is_cuda and a.is_contiguous() and a.dtype == torch.float64
        return ctypes.c_void_p(a.data_ptr())
      assert a.flags['C_CONTIGUOUS'] and a.dtype == np.float64
      return ctypes.c_void_p(a.ctypes.data)

    if out is None:
      energies = np.empty(4)
      g_re = np.empty(self.sphere_shape)
      g_im = np.empty(self.sphere_shape)
    else:
      energies, g_re, g_im = out
    rho = np.empty((self.ns, self.nx, self.ny, self.nz)) if want_rho else None
    _lib.check(self.lib.jrb_energy_grad_host(
      self._h, host_ptr(w_re), host_ptr(w_im), host_ptr(occ), _lib.XC_IDS[xc],
      host_ptr(energies), host_ptr(g_re), host_ptr(g_im),
      None if rho is None else host_ptr(rho)))
    return energies, g_re, g_im, rho


class RowsPlan(_CommMixin):
  """QR-only plan over a block of `nrows` rows of the (ns, nk, ng, nb) coefficient matrices
  (jrb_plan_create_rows): the row-sharded half of the Gamma-only multi-GPU layout, SURVEY 8e.
  The split-phase methods also exist on nothing else; between `gram` and `apply` the caller
  all-reduces the small matrices over the ranks that hold the other row blocks."""

  def __init__(self, nrows: int, num_k: int, num_bands: int, num_spin: int = 1,
               device: Optional[int] = None):
    if not torch.cuda.is_available():
      raise RuntimeError('jrystal_b200 needs a CUDA device (there is no CPU fallback)')
    self.lib = _lib.load()
    self.device = torch.cuda.current_device() if device is None else int(device)
    self.ns, self.nk, self.nb, self.ng = int(num_spin), int(num_k), int(num_bands), int(nrows)
    handle = ctypes.c_void_p()
    _lib.check(self.lib.jrb_plan_create_rows(self.ng, self.ns, self.nk, self.nb, self.device,
                                             ctypes.byref(handle)))
    self._h = handle
    self.tdev = torch.device('cuda', self.device)

  def __del__(self):
    h = getattr(self, '_h', None)
    if h:
      self.lib.jrb_plan_destroy(h)
      self._h = None

  @property
  def rows_shape(self):
    return (self.ns, self.nk, self.ng, self.nb)

  @property
  def small_shape(self):
    return (self.ns, self.nk, self.nb, self.nb)

  def _new(self, shape, dtype):
    return torch.empty(shape, dtype=dtype, device=self.tdev)

  def _chk(self, t, shape, dtype, name):
    return _check_tensor(t, shape, dtype, name, self.device)

  def _rows_in(self, w_re, w_im, needed: bool):
    # pass 1 reads the Q1 of pass 0 from plan work space and ignores w_re / w_im
    if needed or w_re is not None:
      self._chk(w_re, self.rows_shape, torch.float64, 'w_re')
    if needed or w_im is not None:
      self._chk(w_im, self.rows_shape, torch.float64, 'w_im')

  def gram(self, w_re, w_im, pass_: int, out=None):
    if pass_ not in (0, 1):
      raise ValueError('pass_ must be 0 or 1')
    self._rows_in(w_re, w_im, pass_ == 0)
    s = self._new(self.small_shape, torch.complex128) if out is None else self._chk(
      out, self.small_shape, torch.complex128, 'out')
    _lib.check(self.lib.jrb_qr_rows_gram(self._h, _ptr(w_re), _ptr(w_im), int(pass_), _ptr(s),
                                         _stream(self.device)))
    return s

  def apply(self, w_re, w_im, pass_: int, s, q=None, r=None):
    if pass_ not in (0, 1):
      raise ValueError('pass_ must be 0 or 1')
    self._rows_in(w_re, w_im, pass_ == 0)
    self._chk(s, self.small_shape, torch.complex128, 's')
    if pass_ == 1:
      q = self._new(self.rows_shape, torch.complex128) if q is None else q
      r = self._new(self.small_shape, torch.complex128) if r is None else r
    if q is not None:
      self._chk(q, self.rows_shape, torch.complex128, 'q')
    if r is not None:
      self._chk(r, self.small_shape, torch.complex128, 'r')
    _lib.check(self.lib.jrb_qr_rows_apply(self._h, _ptr(w_re), _ptr(w_im), int(pass_), _ptr(s),
                                          _ptr(q), _ptr(r), _stream(self.device)))
    return q, r

  def bwd_gram(self, q, gq, out=None):
    self._chk(q, self.rows_shape, torch.complex128, 'q')
    self._chk(gq, self.rows_shape, torch.complex128, 'gq')
    m = self._new(self.small_shape, torch.complex128) if out is None else self._chk(
      out, self.small_shape, torch.complex128, 'out')
    _lib.check(self.lib.jrb_qr_rows_bwd_gram(self._h, _ptr(q), _ptr(gq), _ptr(m),
                                             _stream(self.device)))
    return m

  def bwd_apply(self, q, gq, occ, m, out=None):
    self._chk(q, self.rows_shape, torch.complex128, 'q')
    self._chk(gq, self.rows_shape, torch.complex128, 'gq')
    if occ is not None:
      self._chk(occ, (self.ns, self.nk, self.nb), torch.float64, 'occupation')
    self._chk(m, self.small_shape, torch.complex128, 'm')
    if out is None:
      g_re = self._new(self.rows_shape, torch.float64)
      g_im = self._new(self.rows_shape, torch.float64)
    else:
      g_re, g_im = out
      self._chk(g_re, self.rows_shape, torch.float64, 'out.g_re')
      self._chk(g_im, self.rows_shape, torch.float64, 'out.g_im')
    _lib.check(self.lib.jrb_qr_rows_bwd_apply(self._h, _ptr(q), _ptr(gq), _ptr(occ), _ptr(m),
                                              _ptr(g_re), _ptr(g_im), _stream(self.device)))
    return g_re, g_im

  def check_status(self):
    _lib.check(self.lib.jrb_check_status(self._h, _stream(self.device)))
