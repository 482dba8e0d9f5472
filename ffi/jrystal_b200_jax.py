"""jax.ffi + jax.custom_vjp binding of libjrystal_b200.so for jrystal (sail-sg/jrystal).

This is the file a jrystal maintainer drops next to `jrystal/_src/` (JAX >= 0.5.3, the reference's
own pin, requirements.txt:1-2).  It mirrors the registration idiom of the reference's one custom
primitive pair (`jrystal/_src/spmd/fft.py:79-134`: Primitive + impl + lowering + JVP + transpose +
batching) with XLA-FFI custom calls, and wraps the differentiable entry points in
`jax.custom_vjp` whose backward passes are the hand-written reverse kernels of the library:

  total_energy            calc/calc_ground_state_energy_all_electrons.py:119-137  ->  JrbEval
  hamiltonian_matrix_trace   _src/hamiltonian.py:147-168     ->  JrbQrFwd + JrbHpsi + JrbBandExpect
  coeff                   _src/pw.py:136-137 (unitary_matrix + expand)            ->  JrbQrFwd / JrbQrBwd
  density_grid            _src/pw.py:273-284                                      ->  JrbDensity / JrbHpsi
  ifftn3d / fftn3d        _src/spmd/fft.py:30-66                                  ->  JrbFft3d

JAX is NOT installable in this repo's build image (no network): the file is shipped as source and
held to the C ABI by tests/test_ffi_shim.py (every custom-call target it names is defined by
ffi/jrb_xla_ffi.cc with the same argument / result counts; that file compiles against
include/jrystal_b200.h).  The executable specification of the same wrappers over torch tensors is
jrystal_b200/autograd.py, which the GPU tests run.

Cotangent convention: for a real loss L and complex z = x + i y JAX's cotangent is
dL/dx - i dL/dy = 2 dL/dz; the library's reverse kernels take dL/dz* (jrb_qr_bwd: gq = dE/dQ*), so
complex cotangents are conjugated and halved on the way in, and dE/dQ* results are doubled and
conjugated on the way out.  Real leaves (w_re, w_im, occupation) need no conversion.
"""
import ctypes
import os
from functools import partial

import numpy as np

import jax
import jax.numpy as jnp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.environ.get('JRB_LIB', os.path.join(_HERE, '..', 'jrystal_b200', 'csrc', 'libjrystal_b200.so'))
_FFI = os.environ.get('JRB_FFI_LIB', os.path.join(_HERE, 'libjrb_xla_ffi.so'))

XC_IDS = {'lda_x': 1, 'lda_x+lda_c_pw': 2, 'gga_x_pbe': 3, 'gga_x_pbe+gga_c_pbe': 4}
FFT_FORWARD, FFT_INVERSE = -1, 1

# custom-call targets defined by ffi/jrb_xla_ffi.cc: name -> (number of operands, number of results)
TARGETS = {
  'JrbQrFwd': (2, 2), 'JrbQrBwd': (3, 2), 'JrbDensity': (2, 1), 'JrbWaveGrid': (1, 1),
  'JrbExpand': (1, 1), 'JrbSqueeze': (1, 1), 'JrbKinetic': (1, 1), 'JrbGridPotential': (1, 2),
  'JrbPotential': (1, 1), 'JrbDensityReciprocal': (1, 1), 'JrbHpsi': (2, 1),
  'JrbHpsiPrepare': (1, 1), 'JrbHpsiPrepared': (2, 1), 'JrbBandExpect': (2, 1),
  'JrbHamiltonianMatrix': (2, 1), 'JrbFft3d': (1, 1), 'JrbEvalBegin': (3, 2),
  'JrbEvalFinish': (3, 4), 'JrbEval': (3, 5), 'JrbAllreduceRho': (2, 2), 'JrbAdamTick': (1, 1),
}

_lib = None
_registered = False


def load():
  """Load libjrystal_b200.so (ctypes, for plan set-up) and register every XLA-FFI target."""
  global _lib, _registered
  if _lib is None:
    _lib = ctypes.CDLL(_LIB, mode=ctypes.RTLD_GLOBAL)
    _lib.jrb_last_error.restype = ctypes.c_char_p
    _lib.jrb_plan_num_g.restype = ctypes.c_int64
  if not _registered:
    ffi_lib = ctypes.CDLL(_FFI)
    for name in TARGETS:
      jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(ffi_lib, name)), platform='CUDA')
    _registered = True
  return _lib


def _check(rc):
  if rc != 0:
    msg = _lib.jrb_last_error()
    raise RuntimeError(f'jrystal_b200 error {rc}: {msg.decode() if msg else "?"}')


class _PlanDesc(ctypes.Structure):
  _fields_ = [('nx', ctypes.c_int32), ('ny', ctypes.c_int32), ('nz', ctypes.c_int32),
              ('ns', ctypes.c_int32), ('nk', ctypes.c_int32), ('nb', ctypes.c_int32),
              ('mask', ctypes.c_void_p), ('kpts', ctypes.c_void_p), ('cell', ctypes.c_void_p),
              ('device', ctypes.c_int32), ('batch_groups', ctypes.c_int32)]


class Plan:
  """Set-up object (not traced): what the reference's drivers build before their loop
  (calc_ground_state_energy_all_electrons.py:93-106).  `handle` is the int64 attribute every
  custom call takes."""

  def __init__(self, cell_vectors, freq_mask, kpts, num_bands, num_spin=1, device=0,
               orbital_grid=None):
    lib = load()
    self.cell = np.ascontiguousarray(np.asarray(cell_vectors, dtype=np.float64).reshape(3, 3))
    self.mask = np.ascontiguousarray(np.asarray(freq_mask).astype(np.uint8))
    self.kpts = np.ascontiguousarray(np.asarray(kpts, dtype=np.float64).reshape(-1, 3))
    self.nx, self.ny, self.nz = (int(v) for v in self.mask.shape)
    self.ns, self.nk, self.nb = int(num_spin), int(self.kpts.shape[0]), int(num_bands)
    desc = _PlanDesc(self.nx, self.ny, self.nz, self.ns, self.nk, self.nb, self.mask.ctypes.data,
                     self.kpts.ctypes.data, self.cell.ctypes.data, int(device), 0)
    h = ctypes.c_void_p()
    _check(lib.jrb_plan_create(ctypes.byref(desc), ctypes.byref(h)))
    self._h = h
    self.handle = np.int64(h.value)
    self.ng = int(lib.jrb_plan_num_g(h))
    self.vol = float(abs(np.linalg.det(self.cell)))
    if orbital_grid is not None:
      _check(lib.jrb_plan_set_orbital_grid(h, *(ctypes.c_int32(int(v)) for v in orbital_grid)))

  def set_atoms(self, positions, charges):
    pos = np.ascontiguousarray(np.asarray(positions, dtype=np.float64).reshape(-1, 3))
    chg = np.ascontiguousarray(np.asarray(charges, dtype=np.float64).reshape(-1))
    _check(_lib.jrb_set_atoms(self._h, ctypes.c_void_p(pos.ctypes.data),
                              ctypes.c_void_p(chg.ctypes.data), ctypes.c_int32(pos.shape[0]), None))

  def __del__(self):
    if getattr(self, '_h', None) and _lib is not None:
      _lib.jrb_plan_destroy(self._h)
      self._h = None

  # shapes (the reference's layouts)
  sphere = property(lambda s: (s.ns, s.nk, s.ng, s.nb))
  grid = property(lambda s: (s.ns, s.nx, s.ny, s.nz))
  bands = property(lambda s: (s.ns, s.nk, s.nb))
  small = property(lambda s: (s.ns, s.nk, s.nb, s.nb))


def _f64(shape):
  return jax.ShapeDtypeStruct(tuple(shape), jnp.float64)


def _c128(shape):
  return jax.ShapeDtypeStruct(tuple(shape), jnp.complex128)


def _call(name, out, *args, **attrs):
  assert TARGETS[name] == (len(args), len(out) if isinstance(out, (tuple, list)) else 1), name
  return jax.ffi.ffi_call(name, out)(*args, **attrs)


# ---------------------------------------------------------------------------------------------
# the primitive pair of _src/spmd/fft.py: linear, self-transposing custom calls
# ---------------------------------------------------------------------------------------------
def make_fft3d(plan: Plan):
  """(ifftn3d, fftn3d) with the reference's semantics (numpy normalisation, last three axes,
  ValueError below 3 dimensions, fft.py:30-66).  Linear maps: jax.custom_jvp with the transform
  itself as tangent rule, which also gives JAX the transpose (the DFT matrix is symmetric, so the
  transpose of fftn is fftn and of ifftn is ifftn, fft.py:88-101); vmap moves the batch axis to the
  front and re-binds (fft.py:104-115), which a leading-batch custom call does by itself."""

  def bind(direction):
    @jax.custom_jvp
    def f(x):
      if x.ndim < 3:
        raise ValueError(f'Input must have at least 3 dimensions, got {x.ndim}')
      return _call('JrbFft3d', _c128(x.shape), x.astype(jnp.complex128), plan=plan.handle,
                   direction=np.int32(direction))

    @f.defjvp
    def f_jvp(primals, tangents):
      return f(*primals), f(*tangents)

    return f

  return bind(FFT_INVERSE), bind(FFT_FORWARD)


# ---------------------------------------------------------------------------------------------
# fused losses
# ---------------------------------------------------------------------------------------------
def make_total_energy(plan: Plan, xc: str = 'lda_x'):
  """Drop-in for the `total_energy` closure of calc_ground_state_energy_all_electrons.py:119-137:
  (w_re, w_im, occupation) -> E_kin + E_ext + E_har + E_xc, differentiable in all three.  One
  custom call per jax.value_and_grad; on a k-sharded mesh with a connected communicator
  (jrb_comm_connect) the psum of rho runs inside it."""
  xc_id = np.int32(XC_IDS[xc])

  def run(w_re, w_im, occ):
    return _call('JrbEval', (_f64((4,)), _f64(plan.sphere), _f64(plan.sphere), _f64(plan.bands),
                             _f64(plan.grid)), w_re, w_im, occ, plan=plan.handle, xc_id=xc_id)

  @jax.custom_vjp
  def total_energy(w_re, w_im, occ):
    return run(w_re, w_im, occ)[0].sum()

  def fwd(w_re, w_im, occ):
    energies, g_re, g_im, g_occ, _ = run(w_re, w_im, occ)
    return energies.sum(), (g_re, g_im, g_occ)     # residuals = the gradients themselves

  def bwd(res, ct):
    g_re, g_im, g_occ = res
    return ct * g_re, ct * g_im, ct * g_occ

  total_energy.defvjp(fwd, bwd)

  def split(w_re, w_im, occ):
    """(kinetic, external, hartree, xc), density -- energy.total_energy(split=True) + rho."""
    energies, _, _, _, rho = run(w_re, w_im, occ)
    return energies, rho

  total_energy.split = split
  return total_energy


def make_total_energy_two_calls(plan: Plan, xc: str = 'lda_x', axis_name=None):
  """The same loss as two custom calls with the collective left to JAX: rho and E_kin are
  psum-ed over `axis_name` between JrbEvalBegin and JrbEvalFinish (shard_map / pmap over the k
  mesh, the reference's P('s', 'k') layout, calc_...all_electrons.py:83-91)."""
  xc_id = np.int32(XC_IDS[xc])

  def run(w_re, w_im, occ):
    rho, e_kin = _call('JrbEvalBegin', (_f64(plan.grid), _f64((1,))), w_re, w_im, occ,
                       plan=plan.handle)
    if axis_name is not None:
      rho, e_kin = jax.lax.psum(rho, axis_name), jax.lax.psum(e_kin, axis_name)
    return _call('JrbEvalFinish', (_f64((4,)), _f64(plan.sphere), _f64(plan.sphere),
                                   _f64(plan.bands)), occ, rho, e_kin, plan=plan.handle, xc_id=xc_id)

  @jax.custom_vjp
  def total_energy(w_re, w_im, occ):
    return run(w_re, w_im, occ)[0].sum()

  def fwd(w_re, w_im, occ):
    energies, g_re, g_im, g_occ = run(w_re, w_im, occ)
    return energies.sum(), (g_re, g_im, g_occ)

  def bwd(res, ct):
    return tuple(ct * g for g in res)

  total_energy.defvjp(fwd, bwd)
  return total_energy


def make_hamiltonian_matrix_trace(plan: Plan):
  """Band-mode loss (hamiltonian.hamiltonian_matrix_trace, hamiltonian.py:147-168, as
  calc_band_structure_all_electrons.py:75-108 uses it): (w_re, w_im, veff) -> sum_i <psi_i| T +
  v_eff |psi_i> for the FIXED potential veff (ns, x, y, z) = potential(rho_gs, kohn_sham=True);
  differentiable in w_re / w_im."""

  def run(w_re, w_im, veff):
    q, r = _call('JrbQrFwd', (_c128(plan.sphere), _c128(plan.small)), w_re, w_im, plan=plan.handle)
    hq = _call('JrbHpsi', _c128(plan.sphere), q, veff, plan=plan.handle)
    eps = _call('JrbBandExpect', _f64(plan.bands), q, hq, plan=plan.handle)
    return q, r, hq, eps

  @jax.custom_vjp
  def trace(w_re, w_im, veff):
    return run(w_re, w_im, veff)[3].sum()

  def fwd(w_re, w_im, veff):
    q, r, hq, eps = run(w_re, w_im, veff)
    # H is Hermitian: d(sum eps)/dQ* = H Q, pushed through the QR adjoint
    g_re, g_im = _call('JrbQrBwd', (_f64(plan.sphere), _f64(plan.sphere)), q, r, hq,
                       plan=plan.handle)
    return eps.sum(), (g_re, g_im, veff)

  def bwd(res, ct):
    g_re, g_im, veff = res
    return ct * g_re, ct * g_im, jnp.zeros_like(veff)   # the potential is held fixed (stop_gradient)

  trace.defvjp(fwd, bwd)
  trace.per_band = lambda w_re, w_im, veff: run(w_re, w_im, veff)[3]
  return trace


def make_effective_potential(plan: Plan, xc: str = 'lda_x'):
  """rho -> (energies[3] = E_har, E_ext, E_xc; v_eff) in one sweep (potential.effective +
  energy.hartree / external / xc_energy); kohn_sham as in hamiltonian.py:147-156."""
  xc_id = np.int32(XC_IDS[xc])

  def effective(rho, kohn_sham=False):
    return _call('JrbGridPotential', (_f64((3,)), _f64(plan.grid)), rho, plan=plan.handle,
                 xc_id=xc_id, kohn_sham=np.int32(bool(kohn_sham)))

  return effective


# ---------------------------------------------------------------------------------------------
# fine-grained pieces under the reference's names
# ---------------------------------------------------------------------------------------------
def make_coeff(plan: Plan):
  """pw.coeff without the dense expansion: (w_re, w_im) -> Q (ns, nk, ng, nb) complex, the sphere
  layout every other call takes.  Backward: the closed-form QR adjoint (JrbQrBwd)."""

  @jax.custom_vjp
  def coeff(w_re, w_im):
    return _call('JrbQrFwd', (_c128(plan.sphere), _c128(plan.small)), w_re, w_im,
                 plan=plan.handle)[0]

  def fwd(w_re, w_im):
    q, r = _call('JrbQrFwd', (_c128(plan.sphere), _c128(plan.small)), w_re, w_im, plan=plan.handle)
    return q, (q, r)

  def bwd(res, ct_q):
    q, r = res
    gq = 0.5 * jnp.conj(ct_q)          # JAX cotangent 2 dL/dQ  ->  dL/dQ*
    return _call('JrbQrBwd', (_f64(plan.sphere), _f64(plan.sphere)), q, r, gq, plan=plan.handle)

  coeff.defvjp(fwd, bwd)
  return coeff


def make_density_grid(plan: Plan):
  """pw.density_grid(coeff, vol, occupation) on the sphere layout: (Q, occ) -> rho (ns, x, y, z).
  Backward: with v = ct_rho, d<v, rho>/dQ* = occ * (sqrt(vol)/N fftn(v psi))|mask = occ * (H - T) Q
  -- one H-apply with v as the potential minus the kinetic term -- and d<v, rho>/d occ =
  <q| v |q> from the same sweep."""

  @jax.custom_vjp
  def density_grid(q, occ):
    return _call('JrbDensity', _f64(plan.grid), q, occ, plan=plan.handle)

  def fwd(q, occ):
    return density_grid(q, occ), (q, occ)

  def bwd(res, ct_rho):
    q, occ = res
    hq = _call('JrbHpsi', _c128(plan.sphere), q, ct_rho, plan=plan.handle)
    eps = _call('JrbBandExpect', _f64(plan.bands), q, hq, plan=plan.handle)
    t = _call('JrbKinetic', _f64(plan.bands), q, plan=plan.handle)
    # remove the kinetic part 1/2 |G+k|^2 q that jrb_hpsi adds: per band <q|T|q> = t, and
    # T q itself is diagonal on the sphere: T q = hq(v = 0); linearity gives it from one more call
    tq = _call('JrbHpsi', _c128(plan.sphere), q, jnp.zeros_like(ct_rho), plan=plan.handle)
    g_q = (hq - tq) * occ[:, :, None, :]
    return 2.0 * jnp.conj(g_q), eps - t          # dL/dQ*  ->  JAX cotangent 2 dL/dQ = conj(2 dL/dQ*)

  density_grid.defvjp(fwd, bwd)
  return density_grid


def make_wave_grid(plan: Plan):
  """pw.wave_grid o pw.coeff: Q -> psi (ns, nk, nb, x, y, z), dense (diagnostics; linear in Q)."""

  def wave_grid(q):
    return _call('JrbWaveGrid', _c128((plan.ns, plan.nk, plan.nb, plan.nx, plan.ny, plan.nz)), q,
                 plan=plan.handle)

  return wave_grid


def make_expand_squeeze(plan: Plan):
  """utils.expand_coefficient / squeeze_coefficient (utils.py:257-308): each the transpose of the
  other (a scatter and its gather)."""
  dense = (plan.ns, plan.nk, plan.nb, plan.nx, plan.ny, plan.nz)

  @jax.custom_vjp
  def expand(q):
    return _call('JrbExpand', _c128(dense), q, plan=plan.handle)

  @jax.custom_vjp
  def squeeze(c):
    return _call('JrbSqueeze', _c128(plan.sphere), c, plan=plan.handle)

  expand.defvjp(lambda q: (expand(q), None), lambda _, ct: (squeeze(ct),))
  squeeze.defvjp(lambda c: (squeeze(c), None), lambda _, ct: (expand(ct),))
  return expand, squeeze


def install(plan: Plan, xc: str = 'lda_x'):
  """Everything a driver needs, under the reference's names."""
  ifftn3d, fftn3d = make_fft3d(plan)
  expand, squeeze = make_expand_squeeze(plan)
  return dict(total_energy=make_total_energy(plan, xc),
              hamiltonian_matrix_trace=make_hamiltonian_matrix_trace(plan),
              effective=make_effective_potential(plan, xc), coeff=make_coeff(plan),
              density_grid=make_density_grid(plan), wave_grid=make_wave_grid(plan),
              ifftn3d=ifftn3d, fftn3d=fftn3d, expand_coefficient=expand,
              squeeze_coefficient=squeeze)
