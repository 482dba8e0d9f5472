"""Stopping rule of the drivers: the run has converged once the standard deviation of the most
recent `window_size` total energies drops below `threshold` (the rule of
jrystal/calc/convergence.py:13-35; config keys convergence_window_size / convergence_condition).
A bounded deque holds the window, so a check costs O(window) and nothing is ever popped by hand."""
import collections
import statistics


def create_convergence_checker(config):
  return ConvergenceChecker(config.convergence_window_size, config.convergence_condition)


class ConvergenceChecker:
  """check(energy) -> bool; `history` lists the energies currently inside the window."""

  def __init__(self, window_size: int = 20, threshold: float = 1e-5):
    if int(window_size) < 1:
      raise ValueError('window_size must be at least 1')
    self.window_size, self.threshold = int(window_size), float(threshold)
    self._recent = collections.deque(maxlen=self.window_size)

  @property
  def history(self):
    return list(self._recent)

  def check(self, value) -> bool:
    self._recent.append(float(value))
    window_full = len(self._recent) == self.window_size
    # population standard deviation, as numpy.std in the reference
    return window_full and statistics.pstdev(self._recent) < self.threshold

  def reset(self) -> None:
    self._recent.clear()
