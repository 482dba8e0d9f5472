"""TEST INFRASTRUCTURE: a CPU stand-in for jrystal_b200.plan.Plan built on the oracle, so that the
HOST LOGIC of the drivers (jrystal_b200/calc/*: set-up, occupation handling, optimiser loop,
convergence, k-path walk, pseudopotential attachment, output containers) runs under
`-m "not gpu"`.  It implements the Plan methods the drivers call with the same argument order,
shapes, in-place `out=` semantics and conventions (kinetic slot = kinetic + non-local, external
slot = whatever V(G) was set), on CPU torch tensors.  It is NOT a fallback: only tests construct
it (by monkeypatching the `Plan` name inside the driver modules); the product's Plan raises without
a CUDA device.  `EmulatedAdam` is optax.adam on CPU tensors in place of the device optimiser."""
import numpy as np
import torch

from oracle import analytic
from oracle import reference_port as rp

C128 = torch.complex128


class EmulatedPlan:

  def __init__(self, cell_vectors, freq_mask, kpts, num_bands, num_spin=1, device=None,
               batch_groups=0, orbital_grid=None):
    self.cell = np.asarray(cell_vectors, dtype=np.float64).reshape(3, 3)
    self.mask = np.asarray(freq_mask).astype(np.uint8)
    self._m = self.mask.astype(bool)
    self.kpts = np.asarray(kpts, dtype=np.float64).reshape(-1, 3)
    self.nx, self.ny, self.nz = self.mask.shape
    self.ns, self.nk, self.nb = int(num_spin), self.kpts.shape[0], int(num_bands)
    self.ng = int(self._m.sum())
    self.vol = float(abs(np.linalg.det(self.cell)))
    self.tdev = torch.device('cpu')
    self.g_vec = rp.g_vectors(self.cell, [self.nx, self.ny, self.nz])
    self.orbital_grid = (self.nx, self.ny, self.nz)
    self.v_ext = None          # V_ext(G), complex (x, y, z)
    self.phi = None            # (nk, nproj, ng)
    self.nproj = 0
    self._atoms = False
    self._prepared = None
    self.calls = []

  sphere_shape = property(lambda self: (self.ns, self.nk, self.ng, self.nb))

  def comm_init(self, group=None, capacity=0):
    """No peer memory on the CPU: callers keep the torch.distributed all-reduce (gloo here)."""
    return False

  # -- set-up ---------------------------------------------------------------------------------
  def set_atoms(self, positions, charges):
    self.v_ext = torch.as_tensor(rp.external_reciprocal(
      np.asarray(positions, dtype=np.float64), np.asarray(charges, dtype=np.float64), self.g_vec,
      self.vol)).to(C128)
    self._atoms = True
    self.calls.append('set_atoms')

  def set_external_potential(self, vhat):
    assert tuple(vhat.shape) == (self.nx, self.ny, self.nz) and vhat.dtype == C128
    self.v_ext = vhat.clone()
    self._atoms = True
    self.calls.append('set_external_potential')

  def set_nonlocal(self, phi):
    self.calls.append('set_nonlocal')
    if phi is None:
      self.phi, self.nproj = None, 0
      return
    assert phi.dtype == C128 and phi.shape[0] == self.nk and phi.shape[2] == self.ng
    self.phi, self.nproj = phi.clone(), int(phi.shape[1])

  def set_kpoints(self, kpts):
    k = np.asarray(kpts, dtype=np.float64).reshape(-1, 3)
    assert k.shape[0] == self.nk
    self.kpts = k

  def check_status(self):
    pass

  # -- pieces ---------------------------------------------------------------------------------
  def _gk2(self):
    g = self.g_vec[self._m]
    return torch.from_numpy(np.stack([np.sum((g + k) ** 2, axis=-1) for k in self.kpts]))  # (k, g)

  def _box(self, q):
    """(s, k, g, b) -> (s, k, b, x, y, z)."""
    return rp.expand_coefficient(q, self._m)

  def qr_fwd(self, w_re, w_im, out=None):
    w = (w_re + 1j * w_im).numpy()
    q = np.empty_like(w)
    r = np.empty((self.ns, self.nk, self.nb, self.nb), dtype=np.complex128)
    for s in range(self.ns):
      for k in range(self.nk):
        q[s, k], r[s, k] = analytic.cholesky_qr2(w[s, k])
    q, r = torch.from_numpy(q), torch.from_numpy(r)
    if out is not None:
      out[0].copy_(q)
      out[1].copy_(r)
      return out
    return q, r

  def qr_bwd(self, q, r, gq, out=None):
    g_re = torch.empty(self.sphere_shape, dtype=torch.float64)
    g_im = torch.empty(self.sphere_shape, dtype=torch.float64)
    for s in range(self.ns):
      for k in range(self.nk):
        gw = analytic.qr_backward(q[s, k].numpy(), r[s, k].numpy(), gq[s, k].numpy())
        g_re[s, k] = torch.from_numpy(2 * gw.real)
        g_im[s, k] = torch.from_numpy(2 * gw.imag)
    if out is not None:
      out[0].copy_(g_re)
      out[1].copy_(g_im)
      return out
    return g_re, g_im

  def _nonlocal_f(self, q):
    return torch.einsum('skgb,kpg->skbp', q, self.phi)

  def nonlocal_energy(self, q, occ):
    f = self._nonlocal_f(q)
    return ((f.conj() * f).real.sum(-1) * occ).sum().reshape(1) / self.vol

  def kinetic(self, q):
    return 0.5 * torch.einsum('skgb,kg->skb', (q.conj() * q).real, self._gk2())

  def potential(self, rho, xc='lda_x', kohn_sham=False, parts=7):
    rho_g = torch.fft.fftn(rho, dim=(-3, -2, -1))
    v = torch.zeros_like(rho)
    if parts & 1:
      v = v + torch.fft.ifftn(rp.hartree_reciprocal(rho_g, self.g_vec, kohn_sham),
                              dim=(-3, -2, -1)).real[None]
    if parts & 2:
      v = v + torch.fft.ifftn(self.v_ext, dim=(-3, -2, -1)).real[None]
    if parts & 4:
      v = v + rp.xc_density(rho, kohn_sham, xc, self.g_vec)
    return v.contiguous()

  def expand(self, q):
    return self._box(q)

  def squeeze(self, c):
    return c[..., torch.from_numpy(self._m)].transpose(-1, -2).contiguous()

  def wave_grid(self, q):
    return rp.wave_grid(self._box(q), self.vol)

  def density(self, q, occ, out=None):
    rho = rp.density_grid(self._box(q), self.vol, occ)
    if out is not None:
      out.copy_(rho)
      return out
    return rho

  def density_reciprocal(self, rho):
    return torch.fft.fftn(rho, dim=(-3, -2, -1))

  def fft3d(self, x, inverse, out=None):
    assert x.dtype == C128 and tuple(x.shape[-3:]) == (self.nx, self.ny, self.nz)
    y = (torch.fft.ifftn if inverse else torch.fft.fftn)(x, dim=(-3, -2, -1))
    if out is not None:
      out.copy_(y)
      return out
    return y

  def grid_potential(self, rho, xc='lda_x', kohn_sham=False, out=None):
    """(E_H, E_ext, E_xc) of a density and v_eff = the potential the backward pass applies
    (dE/d rho: Hartree not halved, v_xc the functional derivative)."""
    assert self._atoms
    rho_g = torch.fft.fftn(rho, dim=(-3, -2, -1))
    en = torch.stack([rp.energy_hartree(rho_g, self.g_vec, self.vol, kohn_sham),
                      rp.reciprocal_braket(self.v_ext, rho_g, self.vol),
                      rp.energy_xc(rho, self.vol, xc, kohn_sham=kohn_sham, g_vector_grid=self.g_vec)])
    veff = self.potential(rho, xc, True, 7)
    if out is not None:
      out[0].copy_(en)
      out[1].copy_(veff)
      return out
    return en, veff

  def prepare_potential(self, veff):
    self._prepared = veff.clone()

  def hpsi(self, q, veff, out=None):
    veff = self._prepared if veff is None else veff
    psi = torch.fft.ifftn(self._box(q), dim=(-3, -2, -1))
    hq = torch.fft.fftn(psi * veff[:, None, None], dim=(-3, -2, -1))[..., torch.from_numpy(self._m)]  # (s,k,b,g)
    hq = hq.transpose(-1, -2) + 0.5 * self._gk2()[None, :, :, None] * q
    if self.nproj:
      f = self._nonlocal_f(q)
      hq = hq + torch.einsum('skbp,kpg->skgb', f, self.phi.conj()) / self.vol
    hq = hq.contiguous()
    if out is not None:
      out.copy_(hq)
      return out
    return hq

  def band_expect(self, q, hq, out=None):
    return (q.conj() * hq).real.sum(-2)

  def overlap(self, q, hq):
    return torch.einsum('skgi,skgj->skij', q.conj(), hq)

  # -- the evaluation -------------------------------------------------------------------------
  def eval_begin(self, w_re, w_im, occ, rho=None, e_kin=None):
    self._w = (w_re.clone(), w_im.clone())
    q = rp.unitary_matrix(w_re, w_im)
    dens = rp.density_grid(self._box(q), self.vol, occ)
    ek = (self.kinetic(q) * occ).sum().reshape(1)
    if self.nproj:
      ek = ek + self.nonlocal_energy(q, occ)
    rho = torch.empty_like(dens) if rho is None else rho
    e_kin = torch.empty(1, dtype=torch.float64) if e_kin is None else e_kin
    rho.copy_(dens)
    e_kin.copy_(ek)
    return rho, e_kin

  def eval_finish(self, occ, rho, e_kin, xc='lda_x', want_occ_grad=False, out=None):
    """`rho` / `e_kin` may have been all-reduced over a k mesh since eval_begin: the grid energies
    come from the TOTAL density, the kinetic slot reports the TOTAL `e_kin`, and the gradients
    are those of this plan's own orbitals in the total potential (what jrb_eval_finish does)."""
    assert self._atoms
    with torch.enable_grad():  # callable from inside a torch.autograd.Function forward
      wr = self._w[0].clone().requires_grad_(True)
      wi = self._w[1].clone().requires_grad_(True)
      oc = occ.clone().requires_grad_(True)
      q = rp.unitary_matrix(wr, wi)
      c = self._box(q)
      own = rp.density_grid(c, self.vol, oc)
      dens = own + (rho - own.detach())               # other ranks' share enters as a constant
      dens_g = torch.fft.fftn(dens, dim=(-3, -2, -1))
      e0 = (self.kinetic(q) * oc).sum()
      if self.nproj:
        e0 = e0 + self.nonlocal_energy(q, oc)[0]
      e1 = rp.reciprocal_braket(self.v_ext, dens_g, self.vol)
      e2 = rp.energy_hartree(dens_g, self.g_vec, self.vol)
      e3 = rp.energy_xc(dens, self.vol, xc, kohn_sham=False, g_vector_grid=self.g_vec)
      grads = torch.autograd.grad(e0 + e1 + e2 + e3, [wr, wi, oc])
    en = torch.stack([e_kin.reshape(()).to(e1.dtype), e1, e2, e3]).detach()
    if out is None:
      out = (torch.empty(4, dtype=torch.float64), torch.empty(self.sphere_shape, dtype=torch.float64),
             torch.empty(self.sphere_shape, dtype=torch.float64))
    out[0].copy_(en)
    out[1].copy_(grads[0])
    out[2].copy_(grads[1])
    return out[0], out[1], out[2], (grads[2].detach() if want_occ_grad else None)


class EmulatedRowsPlan:
  """CPU stand-in for jrystal_b200.plan.RowsPlan (jrb_qr_rows_*): the split-phase Cholesky-QR2 and
  its adjoint on a block of sphere rows; the caller all-reduces the small matrices between the
  phases (parallel.RowShardedEvaluator)."""

  def __init__(self, nrows, num_k, num_bands, num_spin=1, device=None):
    self.ns, self.nk, self.nb, self.ng = int(num_spin), int(num_k), int(num_bands), int(nrows)
    self.tdev = torch.device('cpu')
    self._r1 = self._q1 = self._r = None

  def comm_init(self, group=None, capacity=0):
    return False

  @staticmethod
  def _chol_upper(s):
    return torch.linalg.cholesky(s).conj().transpose(-1, -2)      # S = R^H R

  def gram(self, w_re, w_im, pass_, out=None):
    x = (w_re + 1j * w_im) if pass_ == 0 else self._q1
    return torch.einsum('skgi,skgj->skij', x.conj(), x)

  def apply(self, w_re, w_im, pass_, s, q=None, r=None):
    if pass_ == 0:
      self._r1 = self._chol_upper(s)
      self._q1 = (w_re + 1j * w_im) @ torch.linalg.inv(self._r1)
      return None, None
    r2 = self._chol_upper(s)
    self._r = r2 @ self._r1
    return self._q1 @ torch.linalg.inv(r2), self._r

  def bwd_gram(self, q, gq, out=None):
    return torch.einsum('skgi,skgj->skij', q.conj(), gq)

  def bwd_apply(self, q, gq, occ, m, out=None):
    f = occ[:, :, None, :].to(gq.dtype)
    g = gq * f                                                   # dE/dQ* = HQ diag(f)
    mf = m * f                                                   # Q^H G summed over ALL rows
    up = torch.triu(mf, 1)
    x = -(up + up.conj().transpose(-1, -2)
          + torch.diag_embed(torch.diagonal(mf, dim1=-2, dim2=-1).real.to(mf.dtype)))
    gw = (g + q @ x) @ torch.linalg.inv(self._r).conj().transpose(-1, -2)
    return (2 * gw.real).contiguous(), (2 * gw.imag).contiguous()

  def check_status(self):
    pass


class EmulatedAdam:
  """optax.adam (jrystal/calc/opt_utils.py:153-168) on CPU tensors; interface of jrystal_b200.optim.Adam."""

  def __init__(self, params, learning_rate=0.01, b1=0.9, b2=0.99, eps=1e-8):
    self.params = list(params)
    self.lr, self.b1, self.b2, self.eps = learning_rate, b1, b2, eps
    self.m = [torch.zeros_like(p) for p in self.params]
    self.v = [torch.zeros_like(p) for p in self.params]
    self.step_count = 0

  def step(self, grads):
    self.step_count += 1
    t = self.step_count
    for p, g, m, v in zip(self.params, grads, self.m, self.v):
      m.mul_(self.b1).add_((1 - self.b1) * g)
      v.mul_(self.b2).add_((1 - self.b2) * g * g)
      p.sub_(self.lr * (m / (1 - self.b1 ** t)) / (torch.sqrt(v / (1 - self.b2 ** t)) + self.eps))


# ---------------------------------------------------------------------------------------------
# The argument contract of the real Plan (jrystal_b200/plan.py:_chk): torch tensors of the exact
# shape and dtype, contiguous.  Enforced here too, so that a test body that would be refused on the
# GPU (numpy array, wrong dtype, a strided view) already fails on the CPU stand-in.
# ---------------------------------------------------------------------------------------------
_F64 = torch.float64
_SHAPES = {
  'sphere': lambda p: p.sphere_shape,
  'occ': lambda p: (p.ns, p.nk, p.nb),
  'grid': lambda p: (p.ns, p.nx, p.ny, p.nz),
  'grid3': lambda p: (p.nx, p.ny, p.nz),
  'small': lambda p: (p.ns, p.nk, p.nb, p.nb),
  'dense': lambda p: (p.ns, p.nk, p.nb, p.nx, p.ny, p.nz),
}
_CONTRACT = {  # method -> [(positional index, parameter name, shape key, dtype, may be None)]
  'nonlocal_energy': [(0, 'q', 'sphere', C128, False), (1, 'occ', 'occ', _F64, False)],
  'set_external_potential': [(0, 'vhat', 'grid3', C128, False)],
  'set_nonlocal': [(0, 'phi', None, C128, True)],
  'qr_fwd': [(0, 'w_re', 'sphere', _F64, False), (1, 'w_im', 'sphere', _F64, False)],
  'qr_bwd': [(0, 'q', 'sphere', C128, False), (1, 'r', 'small', C128, False),
             (2, 'gq', 'sphere', C128, False)],
  'expand': [(0, 'q', 'sphere', C128, False)],
  'squeeze': [(0, 'c', 'dense', C128, False)],
  'density': [(0, 'q', 'sphere', C128, False), (1, 'occ', 'occ', _F64, False)],
  'kinetic': [(0, 'q', 'sphere', C128, False)],
  'grid_potential': [(0, 'rho', 'grid', _F64, False)],
  'potential': [(0, 'rho', 'grid', _F64, False)],
  'density_reciprocal': [(0, 'rho', 'grid', _F64, False)],
  'wave_grid': [(0, 'q', 'sphere', C128, False)],
  'prepare_potential': [(0, 'veff', 'grid', _F64, False)],
  'hpsi': [(0, 'q', 'sphere', C128, False), (1, 'veff', 'grid', _F64, True)],
  'band_expect': [(0, 'q', 'sphere', C128, False), (1, 'hq', 'sphere', C128, False)],
  'overlap': [(0, 'q', 'sphere', C128, False), (1, 'hq', 'sphere', C128, False)],
  'eval_begin': [(0, 'w_re', 'sphere', _F64, False), (1, 'w_im', 'sphere', _F64, False),
                 (2, 'occ', 'occ', _F64, False)],
  'fft3d': [(0, 'x', None, C128, False)],
}


def _check_argument(plan, method, name, t, shape_key, dtype, optional):
  if t is None and optional:
    return
  if not isinstance(t, torch.Tensor):
    raise TypeError(f'{method}: {name} must be a tensor (the real Plan takes CUDA tensors), got {type(t)}')
  if shape_key is not None and tuple(t.shape) != tuple(_SHAPES[shape_key](plan)):
    raise ValueError(f'{method}: {name} has shape {tuple(t.shape)}, expected '
                     f'{tuple(_SHAPES[shape_key](plan))}')
  if t.dtype != dtype:
    raise TypeError(f'{method}: {name} has dtype {t.dtype}, expected {dtype}')
  if not t.is_contiguous():
    raise ValueError(f'{method}: {name} must be contiguous')


def _with_contract(method, fn, spec):
  import functools

  @functools.wraps(fn)
  def checked(self, *args, **kwargs):
    if getattr(self, '_inside', 0):          # the stand-in calling its own methods: not a caller
      return fn(self, *args, **kwargs)
    for index, name, shape_key, dtype, optional in spec:
      if index < len(args):
        _check_argument(self, method, name, args[index], shape_key, dtype, optional)
      elif name in kwargs:
        _check_argument(self, method, name, kwargs[name], shape_key, dtype, optional)
    self._inside = 1
    try:
      return fn(self, *args, **kwargs)
    finally:
      self._inside = 0
  return checked


for _method in [n for n, f in vars(EmulatedPlan).items() if callable(f) and not n.startswith('_')]:
  # every public method marks "inside", the ones of the contract also check their arguments
  setattr(EmulatedPlan, _method,
          _with_contract(_method, getattr(EmulatedPlan, _method), _CONTRACT.get(_method, [])))


def patch_drivers(monkeypatch):
  """Point the driver modules at the emulation (Plan, Adam) and make the CUDA-only calls no-ops."""
  from jrystal_b200.calc import (calc_band_structure_all_electrons as band,
                                 calc_ground_state_energy_all_electrons as energy, opt_utils)
  monkeypatch.setattr(energy, 'Plan', EmulatedPlan)
  monkeypatch.setattr(band, 'Plan', EmulatedPlan)
  monkeypatch.setattr(band, 'Adam', EmulatedAdam)
  monkeypatch.setattr(opt_utils, 'Adam', EmulatedAdam)
  monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a, **k: None)
