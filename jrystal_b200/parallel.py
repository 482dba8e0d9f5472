"""Multi-GPU host logic: one process per GPU, k-points (then band blocks) sharded over ranks,
the only data-path collective is the all-reduce of the partial density (+ E_kin).

Mirrors the reference's `parallel_over_k_mesh` layout, NamedSharding(P('s','k')) on the
parameters (calc/calc_ground_state_energy_all_electrons.py:83-91,151-158), where XLA inserts the
all-reduce of rho where einsum('skb...,skb->s...') contracts the sharded k axis (pw.py:278).
"""
from typing import Tuple

import torch
import torch.distributed as dist


def shard_kpoints(num_k: int, world_size: int, rank: int) -> Tuple[int, int]:
  """Contiguous k-range [start, stop) of `rank`.  Like the reference (spmd/uniform.py:22-24)
  this needs num_k % world_size == 0."""
  if world_size < 1 or not (0 <= rank < world_size):
    raise ValueError(f'bad rank/world_size {rank}/{world_size}')
  if num_k % world_size != 0:
    raise ValueError(
      f'num_k ({num_k}) must be divisible by the number of devices ({world_size})')
  per = num_k // world_size
  return rank * per, (rank + 1) * per


def shard_bands(num_bands: int, world_size: int, rank: int) -> Tuple[int, int]:
  """Contiguous band block of `rank` (Gamma-only supercells: fewer k-points than ranks).
  Orthonormalisation then has to be done over the full band set (replicated QR); only the
  FFT/density/H-apply work is split."""
  base, extra = divmod(num_bands, world_size)
  start = rank * base + min(rank, extra)
  return start, start + base + (1 if rank < extra else 0)


def allreduce_density(rho: torch.Tensor, e_kin: torch.Tensor) -> None:
  """In-place SUM over ranks of the partial density and kinetic energy (no-op when
  torch.distributed is not initialised).  NCCL on GPUs, gloo in the CPU tests."""
  if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
    dist.all_reduce(rho, op=dist.ReduceOp.SUM)
    dist.all_reduce(e_kin, op=dist.ReduceOp.SUM)
