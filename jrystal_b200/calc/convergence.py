"""Rolling-window convergence test of the drivers (jrystal/calc/convergence.py:13-35): converged
when the standard deviation of the last `window_size` energies falls below `threshold`."""
import numpy as np


def create_convergence_checker(config):
  return ConvergenceChecker(window_size=config.convergence_window_size,
                            threshold=config.convergence_condition)


class ConvergenceChecker:

  def __init__(self, window_size: int = 20, threshold: float = 1e-5):
    self.window_size = int(window_size)
    self.threshold = float(threshold)
    self.history = []

  def check(self, value: float) -> bool:
    self.history.append(float(value))
    if len(self.history) > self.window_size:
      self.history.pop(0)
    if len(self.history) < self.window_size:
      return False
    return bool(np.std(self.history) < self.threshold)

  def reset(self):
    self.history = []
