"""TEST INFRASTRUCTURE (golden-vector generation only; never imported by the product).

A numpy stand-in for the handful of `jax` names the reference's forward path touches, so that the
reference's OWN source files under /root/reference/jrystal execute in this container (jax, jax_xc,
ase, chex, interpax are not installable here; numpy, scipy, einops, jaxtyping, absl are).  With it
`tests/golden/make_reference_golden.py` runs e.g. `jrystal._src.pw.density_grid`,
`jrystal._src.energy.hartree`, `jrystal.pseudopotential.nloc.potential_nonlocal_psi_reciprocal`
verbatim and stores their outputs as golden vectors, which pin the oracle restatement
(`oracle/`) and the host code of `jrystal_b200/` to the reference itself rather than to a
re-reading of it.

What the stand-in provides, and why each is faithful:
  * jax.numpy / jax.numpy.fft / jax.numpy.linalg -> the numpy functions of the same name in
    float64/complex128 (the reference runs with jax_enable_x64; jnp.linalg.qr on CPU and numpy
    both call LAPACK geqrf/orgqr; jnp.fft on CPU and numpy both use pocketfft).  Arrays are a
    numpy subclass that adds the functional update `x.at[idx].set(v)` / `.add(v)`.
  * jax.vmap -> a Python loop over the mapped axis + stack (in_axes / out_axes honoured).
  * jax.jit -> identity; jax.lax.stop_gradient -> identity (forward values only);
    jax.lax.select / dynamic_update_slice -> numpy equivalents.
  * the custom FFT primitives of _src/spmd/fft.py: `core.Primitive` with def_impl / bind calls the
    implementation the reference registers (jnp.fft.(i)fftn over the last three axes); the
    lowering / JVP / batching registrations are accepted and ignored; the SPMD partitioning
    decorator of _src/spmd/custom_sharding.py is the identity (it changes sharding, not values).
  * jax.scipy.special.erfc, sph_harm_y -> scipy.special.
  * ase.dft.kpoints.monkhorst_pack -> its published definition (i + 0.5)/n - 0.5 in C order;
    chex.dataclass -> dataclasses.dataclass; interpax.CubicSpline -> scipy's CubicSpline
    (interpax mirrors scipy's signature and not-a-knot default).
  * jax_xc is NOT stood in for: anything that evaluates an XC functional stays outside the
    reference goldens (XC arithmetic parity unpinned).
No automatic differentiation: jacfwd / value_and_grad / vjp raise; jax.grad is a complex-step
derivative for scalar analytic functions only (the per-point d(exc)/d(rho) of xc.py:118).  Gradients are pinned through finite
differences of the pinned energies (tests/test_oracle.py).
"""
import dataclasses
import importlib
import sys
import types

import numpy as np
import scipy.interpolate
import scipy.special

REFERENCE_ROOT = '/root/reference'


class Arr(np.ndarray):
  """numpy array with jax's functional-update syntax."""

  @property
  def at(self):
    return _At(self)

  def block_until_ready(self):
    return self

  # jax arrays are immutable: `x *= y` rebinds x to a NEW array (with type promotion)
  def __imul__(self, o):
    return self * o

  def __iadd__(self, o):
    return self + o

  def __isub__(self, o):
    return self - o

  def __itruediv__(self, o):
    return self / o

  def __rpow__(self, base):
    # jnp.power(int, negative int) does not raise (binary exponentiation on the bit pattern);
    # the only such use on this path is (-1) ** m in spherical.py:77-82, where it yields the
    # parity (-1)^|m| = the mathematical value.  numpy raises, so go through floats.
    if np.issubdtype(self.dtype, np.integer) and isinstance(base, (int, np.integer)):
      return np.rint(np.power(float(base), np.asarray(self, dtype=np.float64))).astype(np.int64).view(Arr)
    return np.power(base, np.asarray(self)).view(Arr)


class _At:
  def __init__(self, a):
    self.a = a

  def __getitem__(self, idx):
    return _AtIdx(self.a, idx)


class _AtIdx:
  def __init__(self, a, idx):
    self.a, self.idx = a, idx

  def set(self, v):
    out = np.array(self.a, copy=True).view(Arr)
    out[self.idx] = v
    return out

  def add(self, v):
    out = np.array(self.a, copy=True).view(Arr)
    np.add.at(out, self.idx, v)
    return out

  def get(self):
    return np.asarray(self.a)[self.idx].view(Arr) if np.ndim(np.asarray(self.a)[self.idx]) else np.asarray(self.a)[self.idx]

  def multiply(self, v):
    out = np.array(self.a, copy=True).view(Arr)
    out[self.idx] = out[self.idx] * v
    return out


def _wrap_out(x):
  if isinstance(x, np.ndarray):
    return x.view(Arr)
  if isinstance(x, tuple) and type(x) is not tuple:  # namedtuple results (qr, eigh, slogdet)
    return type(x)(*[_wrap_out(i) for i in x])
  if isinstance(x, tuple):
    return tuple(_wrap_out(i) for i in x)
  if isinstance(x, list):
    return [_wrap_out(i) for i in x]
  return x


def _wrap_fn(f):
  def g(*a, **k):
    # jnp takes any sequence of axes, numpy insists on tuples
    a = tuple(tuple(x) if isinstance(x, range) else x for x in a)
    k = {key: (tuple(v) if isinstance(v, range) else v) for key, v in k.items()}
    return _wrap_out(f(*a, **k))
  g.__name__ = getattr(f, '__name__', 'fn')
  g.__doc__ = getattr(f, '__doc__', None)
  return g


class _NumpyNamespace(types.ModuleType):
  """Module whose attributes are looked up in a numpy namespace; callables return Arr."""

  def __init__(self, name, source, extra=None):
    super().__init__(name)
    self.__dict__['_source'] = source
    self.__dict__['_extra'] = dict(extra or {})

  def __getattr__(self, key):
    if key.startswith('__'):
      raise AttributeError(key)
    extra = self.__dict__['_extra']
    if key in extra:
      return extra[key]
    v = getattr(self.__dict__['_source'], key)
    if isinstance(v, (type, types.ModuleType)) or not callable(v):
      return v
    w = _wrap_fn(v)
    setattr(self, key, w)
    return w


def _vmap(fun, in_axes=0, out_axes=0, **_unused):
  def mapped(*args):
    axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
    if len(axes) != len(args):
      raise ValueError('vmap stand-in: in_axes does not match the arguments')
    n = None
    for a, ax in zip(args, axes):
      if ax is not None:
        n = np.shape(a)[ax]
        break
    if n is None:
      raise ValueError('vmap stand-in: nothing to map over')
    outs = []
    for i in range(n):
      sl = [a if ax is None else np.take(np.asarray(a), i, axis=ax).view(Arr)
            for a, ax in zip(args, axes)]
      outs.append(fun(*sl))
    if isinstance(outs[0], (tuple, list)):
      return type(outs[0])(np.stack([o[j] for o in outs], axis=out_axes).view(Arr)
                           for j in range(len(outs[0])))
    return np.stack([np.asarray(o) for o in outs], axis=out_axes).view(Arr)
  return mapped


def _complex_step_grad(fun, argnums=0, **_unused):
  """jax.grad for a REAL ANALYTIC scalar function of a scalar: Im f(x + i h) / h, exact to rounding
  (no subtraction).  Enough for the one use on the LDA forward path (xc.py:118:
  jax.vmap(jax.grad(exc)) over the grid points); anything else must not rely on it."""
  h = 1e-40

  def g(*args):
    x = np.asarray(args[argnums])
    if x.ndim != 0:
      raise NotImplementedError('complex-step grad stand-in: scalar argument only')
    shifted = list(args)
    shifted[argnums] = (x.astype(np.complex128) + 1j * h).view(Arr)
    y = np.asarray(fun(*shifted))
    if y.size != 1:
      raise NotImplementedError('complex-step grad stand-in: scalar function only')
    return (np.imag(y).reshape(()) / h).view(Arr)
  return g


def _no_autodiff(*_a, **_k):
  raise NotImplementedError('the numpy stand-in has no automatic differentiation')


class _Primitive:
  """jax.extend.core.Primitive: bind() runs the registered implementation."""

  def __init__(self, name):
    self.name = name
    self.impl = None

  def def_impl(self, f):
    self.impl = f
    return f

  def def_abstract_eval(self, f):
    return f

  def bind(self, *args, **kw):
    return self.impl(*args, **kw)

  def __hash__(self):
    return hash(self.name)


def _module(name, **attrs):
  m = types.ModuleType(name)
  m.__dict__.update(attrs)
  sys.modules[name] = m
  return m


def _dynamic_update_slice(operand, update, start):
  out = np.array(operand, copy=True).view(Arr)
  idx = tuple(slice(int(s), int(s) + n) for s, n in zip(start, np.shape(update)))
  out[idx] = update
  return out


def _monkhorst_pack(size):
  """ase.dft.kpoints.monkhorst_pack: (i + 0.5) / n - 0.5 over np.indices in C order."""
  size = np.asarray(size)
  kpts = np.indices(size).transpose((1, 2, 3, 0)).reshape((-1, 3))
  return (kpts + 0.5) / size - 0.5


def install():
  """Put the stand-in modules into sys.modules (idempotent).  Refuses to shadow a real jax."""
  if 'jax' in sys.modules and not getattr(sys.modules['jax'], '_is_numpy_standin', False):
    raise RuntimeError('a real jax is imported; use it instead of the stand-in')
  if 'jax' in sys.modules:
    return
  fft = _NumpyNamespace('jax.numpy.fft', np.fft)
  linalg = _NumpyNamespace('jax.numpy.linalg', np.linalg)
  jnp = _NumpyNamespace('jax.numpy', np, extra=dict(
    fft=fft, linalg=linalg, ndarray=np.ndarray,
    array=_wrap_fn(lambda x, dtype=None, **k: np.array(x, dtype=dtype)),
    asarray=_wrap_fn(lambda x, dtype=None, **k: np.asarray(x, dtype=dtype)),
  ))
  sys.modules['jax.numpy'] = jnp
  sys.modules['jax.numpy.fft'] = fft
  sys.modules['jax.numpy.linalg'] = linalg

  lax = _module('jax.lax', stop_gradient=lambda x: x,
                select=_wrap_fn(lambda p, a, b: np.where(p, a, b)),
                dynamic_update_slice=_dynamic_update_slice)
  sharding = _module('jax.sharding', Sharding=type('Sharding', (), {}),
                     NamedSharding=type('NamedSharding', (), {}),
                     PartitionSpec=lambda *a: tuple(a), Mesh=type('Mesh', (), {}))
  special = _module('jax.scipy.special', erfc=_wrap_fn(scipy.special.erfc),
                    sph_harm=_wrap_fn(lambda m, n, theta, phi, n_max=None: scipy.special.sph_harm_y(n, m, phi, theta)),
                    sph_harm_y=_wrap_fn(scipy.special.sph_harm_y),
                    gammaln=_wrap_fn(scipy.special.gammaln))
  jscipy = _module('jax.scipy', special=special)
  random = _module('jax.random', PRNGKey=lambda s: np.random.default_rng(s),
                   split=lambda k, n=2: [np.random.default_rng(k.integers(1 << 31)) for _ in range(n)],
                   uniform=lambda k, shape, **kw: k.random(tuple(shape)).view(Arr))
  core = _module('jax.extend.core', Primitive=_Primitive,
                 ShapedArray=lambda shape, dtype=None: (shape, dtype))
  extend = _module('jax.extend', core=core)
  ad = _module('jax._src.interpreters.ad', primitive_jvps={}, deflinear2=lambda p, r: None)
  batching = _module('jax._src.interpreters.batching', primitive_batchers={},
                     moveaxis=_wrap_fn(np.moveaxis))
  src_interp = _module('jax._src.interpreters', ad=ad, batching=batching)
  src = _module('jax._src', interpreters=src_interp)
  mlir = _module('jax.interpreters.mlir', register_lowering=lambda *a, **k: None,
                 lower_fun=lambda f, *a, **k: f)
  interp = _module('jax.interpreters', mlir=mlir, ad=ad, batching=batching)
  nn = _module('jax.nn', sigmoid=_wrap_fn(scipy.special.expit))
  _module('jax', _is_numpy_standin=True, numpy=jnp, lax=lax, sharding=sharding, scipy=jscipy,
          random=random, extend=extend, _src=src, interpreters=interp, nn=nn,
          Array=np.ndarray, vmap=_vmap, jit=lambda f=None, **k: (f if f is not None else (lambda g: g)),
          grad=_complex_step_grad, value_and_grad=_no_autodiff, jacfwd=_no_autodiff,
          jacrev=_no_autodiff, hessian=_no_autodiff, jvp=_no_autodiff, vjp=_no_autodiff,
          tree_map=lambda f, t: f(t), device_count=lambda: 1, devices=lambda: [None],
          config=types.SimpleNamespace(update=lambda *a, **k: None))

  # third-party packages the reference imports next to jax
  kpoints = _module('ase.dft.kpoints', monkhorst_pack=_monkhorst_pack)
  dft = _module('ase.dft', kpoints=kpoints)
  _module('ase.io', read=_no_autodiff)
  _module('ase', dft=dft, io=sys.modules['ase.io'])
  _module('chex', dataclass=dataclasses.dataclass)
  _module('interpax', CubicSpline=scipy.interpolate.CubicSpline)


def load_reference(root=REFERENCE_ROOT):
  """Import the reference's modules by path WITHOUT running the package __init__ files (those pull
  in optax / ml_collections / jax_xc for the drivers).  Returns the bare `jrystal` package; use
  importlib.import_module('jrystal._src.pw') etc. afterwards."""
  install()
  if 'jrystal' in sys.modules:
    return sys.modules['jrystal']
  base = f'{root}/jrystal'

  def bare(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__package__ = name
    sys.modules[name] = m
    return m

  pkg = bare('jrystal', base)
  bare('jrystal._src', f'{base}/_src')
  spmd = bare('jrystal._src.spmd', f'{base}/_src/spmd')
  bare('jrystal.pseudopotential', f'{base}/pseudopotential')
  sbt_pkg = bare('jrystal.sbt', f'{base}/sbt')
  # sharding decorator of the FFT primitives: identity on values
  _module('jrystal._src.spmd.custom_sharding', custom_sharding_by_mesh=lambda f: f)
  # the functional library is not available: importing xc.py works, USING a functional fails
  # loudly unless the caller registers a formula with provide_functional()
  _module('jax_xc.utils', get_p=lambda name, polarized: None)
  _module('jax_xc.impl')
  _module('jax_xc', utils=sys.modules['jax_xc.utils'], impl=sys.modules['jax_xc.impl'])
  # a Crystal type for annotations only (the real one needs ase + chex)
  _module('jrystal._src.crystal', Crystal=type('Crystal', (), {}))
  # what jrystal/sbt/__init__.py exports
  num = importlib.import_module('jrystal.sbt.sbt_numerical')
  sbt_pkg.sbt_numerical = num.sbt
  try:
    s = importlib.import_module('jrystal.sbt.sbt')
    sbt_pkg.sbt, sbt_pkg.batch_sbt = s.sbt, s.batch_sbt
  except Exception:  # FFT-based SBT variant: not used by the default 'sbt' method
    sbt_pkg.sbt = sbt_pkg.batch_sbt = None
  del spmd
  return pkg


def provide_functional(name, unpol):
  """Register `jax_xc.impl.<name>.unpol(p, rho) -> eps_xc per particle` (jax_xc's calling
  convention, xc.py:25-27, 57-63).  The FORMULA is the caller's (jax_xc is absent): goldens made
  with it pin the reference's assembly AROUND the functional, not the functional."""
  load_reference()
  mod = _module(f'jax_xc.impl.{name}', unpol=unpol)
  setattr(sys.modules['jax_xc.impl'], name, mod)


def ref(module):
  """ref('_src.pw') -> the reference module jrystal/_src/pw.py executed over the stand-in."""
  load_reference()
  return importlib.import_module('jrystal.' + module)
