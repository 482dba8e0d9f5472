"""Device-side optimiser of the energy-mode driver: optax.adam as the reference configures it
(jrystal/calc/opt_utils.py:153-168, defaults of jrystal/config.py: lr 0.01, b1 0.9, b2 0.99),
on jrb_adam_tick / jrb_adam_apply.  All state (moments, step counter, bias corrections) lives in
device memory so a whole optimisation step can be replayed as a CUDA graph."""
import ctypes
from typing import Sequence

import torch

from . import _lib


class Adam:

  def __init__(self, params: Sequence[torch.Tensor], learning_rate: float = 0.01, b1: float = 0.9,
               b2: float = 0.99, eps: float = 1e-8):
    self.lib = _lib.load()
    self.params = list(params)
    for p in self.params:
      if not (p.is_cuda and p.dtype == torch.float64 and p.is_contiguous()):
        raise TypeError('Adam needs contiguous float64 CUDA tensors')
    self.lr, self.b1, self.b2, self.eps = float(learning_rate), float(b1), float(b2), float(eps)
    self.m = [torch.zeros_like(p) for p in self.params]
    self.v = [torch.zeros_like(p) for p in self.params]
    self.state = torch.zeros(4, dtype=torch.float64, device=self.params[0].device)

  @property
  def step_count(self) -> int:
    return int(self.state[0].item())

  def step(self, grads: Sequence[torch.Tensor]) -> None:
    """In-place update of every parameter with its gradient (asynchronous on the current stream)."""
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    sp = ctypes.c_void_p(self.state.data_ptr())
    _lib.check(self.lib.jrb_adam_tick(sp, self.b1, self.b2, st))
    for p, g, m, v in zip(self.params, grads, self.m, self.v):
      if g.shape != p.shape or g.dtype != p.dtype or not g.is_contiguous():
        raise ValueError('gradient does not match its parameter')
      _lib.check(self.lib.jrb_adam_apply(
        p.numel(), ctypes.c_void_p(p.data_ptr()), ctypes.c_void_p(g.data_ptr()),
        ctypes.c_void_p(m.data_ptr()), ctypes.c_void_p(v.data_ptr()), self.lr, self.b1, self.b2,
        self.eps, sp, st))
