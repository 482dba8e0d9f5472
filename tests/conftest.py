import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def cuda_device():
  import torch
  if not torch.cuda.is_available():
    pytest.skip('no CUDA device')
  torch.cuda.set_device(0)
  return 0


# Dual-backend tests: the same body on the product ('cuda', marked gpu: jrystal_b200.Plan through
# the C ABI) and on the oracle-backed CPU stand-in ('emulated', tests/emulated_plan.py), which
# checks the host logic of the drivers and the test's own arithmetic without a GPU.
BACKENDS = [pytest.param('cuda', marks=pytest.mark.gpu), 'emulated']


@pytest.fixture
def backend(request, monkeypatch):
  """(kind, Plan class, to-device function); parametrize indirectly with BACKENDS."""
  import numpy as np
  import torch
  kind = request.param
  if kind == 'cuda':
    if not torch.cuda.is_available():
      pytest.skip('no CUDA device')
    torch.cuda.set_device(0)
    import jrystal_b200 as jb
    from tests.common import to_dev
    return kind, jb.Plan, to_dev
  from tests import emulated_plan
  emulated_plan.patch_drivers(monkeypatch)
  return kind, emulated_plan.EmulatedPlan, lambda a: torch.from_numpy(np.ascontiguousarray(a))
