"""Host logic of the drivers under `-m "not gpu"`: the GPU driver tests of tests/test_drivers_gpu.py
re-run with jrystal_b200.plan.Plan / optim.Adam replaced by the oracle-backed CPU stand-ins of
tests/emulated_plan.py.  What this covers is everything in jrystal_b200/calc/ that is not a kernel:
configuration, set-up helpers, occupation maps and their autograd chain, the Adam / temperature /
convergence loop, the k-path walk with warm start, the output containers.  The kernels themselves
are covered by the `-m gpu` runs of the same test bodies."""
import pytest

from tests import emulated_plan
from tests import test_drivers_gpu as gpu_tests


@pytest.fixture
def emulated(monkeypatch):
  emulated_plan.patch_drivers(monkeypatch)
  return 0  # stands in for the cuda_device fixture value


def test_energy_driver_follows_the_oracle_trajectory(emulated):
  gpu_tests.test_energy_driver_follows_the_oracle_trajectory(emulated, False)


@pytest.mark.parametrize('method', ['simplex-projector', 'idempotent'])
def test_energy_driver_with_trainable_occupations(emulated, method):
  gpu_tests.test_energy_driver_with_trainable_occupations(emulated, method)


def test_energy_driver_converges_and_stops(emulated):
  gpu_tests.test_energy_driver_converges_and_stops(emulated)


def test_band_driver_eigenvalues_are_variational_and_close(emulated):
  gpu_tests.test_band_driver_eigenvalues_are_variational_and_close(emulated)
