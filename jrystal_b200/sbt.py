"""jrystal.sbt: spherical Bessel transforms of radial pseudopotential data (host set-up, "not
differentiable", jrystal/sbt/__init__.py).  The reference's pseudopotential pipeline calls
`sbt_numerical` only (pseudopotential/beta.py:73, local.py:82; the call to the pySBT port is
commented out there), so that is the one carried here: direct quadrature of
int r^2 f(r) j_l(k r) dr on the UPF's radial grid, pinned to the reference's own output
(tests/test_pseudopotential.py).  `pyNumSBT` / `sbt` / `batch_sbt` (Talman's logarithmic-mesh FFT
algorithm, adapted by the reference from pySBT) are not on any path of the drivers and are absent."""
from .pseudopotential.beta import sbt_numerical  # noqa: F401

__all__ = ['sbt_numerical']
