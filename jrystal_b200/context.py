"""The plan the stateless reference-named functions (pw, energy, potential, hamiltonian) run on.

The reference's functions are pure functions of arrays; the CUDA kernels need the set-up the
reference's drivers build before their loop (index maps, |G+k|^2 tables, V_ext(G), work space),
which lives in a `Plan`.  `use_plan(plan)` makes a plan current (also usable as a context
manager); handles returned by `pw.coeff` carry their plan with them.
"""
import threading

_state = threading.local()


class use_plan:
  def __init__(self, plan):
    self._prev = getattr(_state, 'plan', None)
    _state.plan = plan

  def __enter__(self):
    return _state.plan

  def __exit__(self, *exc):
    _state.plan = self._prev
    return False


def current_plan():
  plan = getattr(_state, 'plan', None)
  if plan is None:
    raise RuntimeError('no current plan: call jrystal_b200.use_plan(Plan(...)) first')
  return plan
