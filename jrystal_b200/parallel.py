"""Multi-GPU host logic: one process per GPU, k-points (then band blocks) sharded over ranks,
the only data-path collective is the all-reduce of the partial density (+ E_kin).

Mirrors the reference's `parallel_over_k_mesh` layout, NamedSharding(P('s','k')) on the
parameters (calc/calc_ground_state_energy_all_electrons.py:83-91,151-158), where XLA inserts the
all-reduce of rho where einsum('skb...,skb->s...') contracts the sharded k axis (pw.py:278).
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def _world(group=None) -> Tuple[int, int]:
  if dist.is_available() and dist.is_initialized():
    return dist.get_world_size(group), dist.get_rank(group)
  return 1, 0


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


def density_buffers(shape, device):
  """(buf, rho, e_kin): the partial density and the partial kinetic energy as views of ONE flat
  buffer, so that a single all-reduce carries both (one collective latency per evaluation)."""
  n = int(np.prod(shape))
  buf = torch.empty(n + 1, dtype=torch.float64, device=device)
  return buf, buf[:n].view(*shape), buf[n:]


def allreduce_density(rho: torch.Tensor, e_kin: torch.Tensor, buf: Optional[torch.Tensor] = None,
                      group=None) -> None:
  """In-place SUM over ranks of the partial density and kinetic energy (no-op when
  torch.distributed is not initialised).  NCCL on GPUs, gloo in the CPU tests.  With `buf` (from
  density_buffers) one collective reduces both."""
  if _world(group)[0] > 1:
    if buf is not None:
      dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    else:
      dist.all_reduce(rho, op=dist.ReduceOp.SUM, group=group)
      dist.all_reduce(e_kin, op=dist.ReduceOp.SUM, group=group)


def shard_rows(num_g: int, world_size: int, rank: int) -> Tuple[int, int]:
  """Contiguous block of sphere rows g of `rank` (row-sharded orthonormalisation)."""
  return shard_bands(num_g, world_size, rank)


def allreduce_sum(*tensors: torch.Tensor, group=None) -> None:
  if _world(group)[0] > 1:
    for t in tensors:
      dist.all_reduce(torch.view_as_real(t) if t.is_complex() else t, op=dist.ReduceOp.SUM,
                      group=group)


def rows_to_bands(x_rows: torch.Tensor, row_blocks: Sequence[Tuple[int, int]],
                  band_blocks: Sequence[Tuple[int, int]], group=None) -> torch.Tensor:
  """(ns, nk, rows of this rank, nb) -> (ns, nk, ng, bands of this rank): the all-to-all that
  turns the row-sharded Q of the QR into the band-sharded Q of the FFTs.  Rank j
  receives from every rank i that rank's row block of j's bands; row blocks are contiguous and
  in rank order, so the received pieces concatenate to the full sphere."""
  world, rank = _world(group)
  b0, b1 = band_blocks[rank]
  if world == 1:
    return x_rows[..., b0:b1].contiguous()
  lead = x_rows.shape[:2]
  nlead = int(np.prod(lead))
  send = torch.cat([x_rows[..., a:b].reshape(-1) for a, b in band_blocks])
  in_splits = [nlead * (row_blocks[rank][1] - row_blocks[rank][0]) * (b - a) for a, b in band_blocks]
  out_splits = [nlead * (g1 - g0) * (b1 - b0) for g0, g1 in row_blocks]
  recv = torch.empty(sum(out_splits), dtype=x_rows.dtype, device=x_rows.device)
  dist.all_to_all_single(torch.view_as_real(recv), torch.view_as_real(send),
                         out_splits, in_splits, group=group)
  pieces, off = [], 0
  for (g0, g1), n in zip(row_blocks, out_splits):
    pieces.append(recv[off:off + n].view(*lead, g1 - g0, b1 - b0))
    off += n
  return torch.cat(pieces, dim=2)


def bands_to_rows(x_bands: torch.Tensor, row_blocks: Sequence[Tuple[int, int]],
                  band_blocks: Sequence[Tuple[int, int]], group=None) -> torch.Tensor:
  """Inverse of rows_to_bands: (ns, nk, ng, bands of this rank) -> (ns, nk, rows of this rank, nb)."""
  world, rank = _world(group)
  g0, g1 = row_blocks[rank]
  if world == 1:
    return x_bands[:, :, g0:g1].contiguous()
  lead = x_bands.shape[:2]
  nlead = int(np.prod(lead))
  nbl = x_bands.shape[-1]
  send = torch.cat([x_bands[:, :, a:b].reshape(-1) for a, b in row_blocks])
  in_splits = [nlead * (b - a) * nbl for a, b in row_blocks]
  out_splits = [nlead * (g1 - g0) * (b - a) for a, b in band_blocks]
  recv = torch.empty(sum(out_splits), dtype=x_bands.dtype, device=x_bands.device)
  dist.all_to_all_single(torch.view_as_real(recv), torch.view_as_real(send), out_splits, in_splits,
                         group=group)
  pieces, off = [], 0
  for (a, b), n in zip(band_blocks, out_splits):
    pieces.append(recv[off:off + n].view(*lead, g1 - g0, b - a))
    off += n
  return torch.cat(pieces, dim=3)


class KShardedEvaluator:
  """Energy+gradient evaluation with whole k-points per rank -- the reference's
  `parallel_over_k_mesh` layout (calc/calc_ground_state_energy_all_electrons.py:83-91,151-158;
  nk % world == 0, spmd/uniform.py:22-24).  The one collective, the all-reduce of the partial
  density and E_kin, runs INSIDE the library over NVLink peer memory (jrb_eval with the plan's
  communicator: reduced on the orbital grid, bit-identical on every rank); when peer memory
  cannot be set up (or JRB_NO_PEER=1) it falls back to NCCL between jrb_eval_begin and
  jrb_eval_finish.  `reduce_path` says which."""

  def __init__(self, cell_vectors, freq_mask, kpts, num_bands, positions, charges, group=None,
               device: Optional[int] = None, batch_groups: int = 0, orbital_grid=None):
    from .plan import Plan
    self.group = group
    self.world, self.rank = _world(group)
    kpts = np.asarray(kpts, dtype=np.float64).reshape(-1, 3)
    self.k0, self.k1 = shard_kpoints(kpts.shape[0], self.world, self.rank)
    self.plan = Plan(cell_vectors, freq_mask, kpts[self.k0:self.k1], num_bands, device=device,
                     batch_groups=batch_groups, orbital_grid=orbital_grid)
    self.plan.set_atoms(positions, charges)
    self.reduce_path = 'none'
    if self.world > 1:
      self.reduce_path = 'peer' if self.plan.comm_init(group) else 'nccl'
    p = self.plan
    self._dbuf, self._rho, self._e_kin = density_buffers((p.ns, p.nx, p.ny, p.nz), p.tdev)

  def evaluate(self, w_re, w_im, occ, xc: str = 'lda_x', out=None, want_occ_grad: bool = False):
    """w_re, w_im: (ns, k-points of this rank, ng, nb); occ: (ns, k-points of this rank, nb).
    Returns energies[4] (whole system), dE/dw_re, dE/dw_im (this rank's k-points), rho (whole)."""
    p = self.plan
    if self.reduce_path != 'nccl':
      en, g_re, g_im, g_occ, rho = p.eval(w_re, w_im, occ, xc, want_occ_grad, out=out, rho=self._rho)
    else:
      p.eval_begin(w_re, w_im, occ, self._rho, self._e_kin)
      allreduce_density(self._rho, self._e_kin, self._dbuf, group=self.group)
      en, g_re, g_im, g_occ = p.eval_finish(occ, self._rho, self._e_kin, xc, want_occ_grad, out=out)
      rho = self._rho
    if want_occ_grad:
      return en, g_re, g_im, rho, g_occ
    return en, g_re, g_im, rho


class RowShardedEvaluator:
  """Energy+gradient evaluation when there are fewer k-points than GPUs (Gamma-only supercells,
  BASELINE config C3a): every rank owns a contiguous block of sphere ROWS of the parameters for
  the orthonormalisation (Gram matrices all-reduced, Cholesky replicated) and a contiguous block
  of BANDS for the FFT / density / H-apply work; two all-to-alls per evaluation move Q and HQ
  between the layouts, and the partial densities are all-reduced as in the k-sharded case.
  Parameters and gradients stay row-sharded: (ns, nk, rows of this rank, nb).

  The reference has no such mode (its pmap mesh needs nk % ndev == 0, spmd/uniform.py:22-24); the
  mathematics is that of the single-device evaluation, cut where sums cross ranks."""

  def __init__(self, cell_vectors, freq_mask, kpts, num_bands, positions, charges, group=None,
               device: Optional[int] = None, batch_groups: int = 0, orbital_grid=None):
    from .plan import Plan, RowsPlan
    self.group = group
    self.world, self.rank = _world(group)
    mask = np.asarray(freq_mask)
    self.ng = int(mask.sum())
    self.nb = int(num_bands)
    self.nk = int(np.asarray(kpts).reshape(-1, 3).shape[0])
    self.row_blocks: List[Tuple[int, int]] = [shard_rows(self.ng, self.world, r)
                                              for r in range(self.world)]
    self.band_blocks: List[Tuple[int, int]] = [shard_bands(self.nb, self.world, r)
                                               for r in range(self.world)]
    self.g0, self.g1 = self.row_blocks[self.rank]
    self.b0, self.b1 = self.band_blocks[self.rank]
    if self.b1 == self.b0 or self.g1 == self.g0:
      raise ValueError('more ranks than bands / rows')
    self.rows = RowsPlan(self.g1 - self.g0, self.nk, self.nb, 1, device=device)
    self.bands = Plan(cell_vectors, mask, kpts, self.b1 - self.b0, device=device,
                      batch_groups=batch_groups, orbital_grid=orbital_grid)
    self.bands.set_atoms(positions, charges)
    # peer-memory all-reduces of the library (bit-identical sums on every rank, which the
    # replicated Cholesky relies on); NCCL when they cannot be set up
    self.reduce_path = 'none'
    if self.world > 1:
      ok = self.rows.comm_init(group) and self.bands.comm_init(group)
      self.reduce_path = 'peer' if ok else 'nccl'

  def _sum_small(self, m):
    if self.reduce_path == 'peer':
      self.rows.allreduce(m)
    else:
      allreduce_sum(m, group=self.group)

  def evaluate(self, w_re_rows, w_im_rows, occ, xc: str = 'lda_x'):
    """w_re_rows, w_im_rows: (1, nk, rows of this rank, nb); occ: (1, nk, nb) full.
    Returns energies[4] = (E_kin, E_ext, E_har, E_xc), dE/dw_re, dE/dw_im (row blocks), rho."""
    rp, bp = self.rows, self.bands
    for pass_ in (0, 1):
      s = rp.gram(w_re_rows, w_im_rows, pass_)
      self._sum_small(s)
      q_rows, _ = rp.apply(w_re_rows, w_im_rows, pass_, s)
    q_band = rows_to_bands(q_rows, self.row_blocks, self.band_blocks, self.group)
    occ_band = occ[:, :, self.b0:self.b1].contiguous()
    buf, rho, e_kin = density_buffers((1, bp.nx, bp.ny, bp.nz), bp.tdev)
    bp.density(q_band, occ_band, out=rho)
    e_kin.copy_((bp.kinetic(q_band) * occ_band).sum().reshape(1))
    if self.reduce_path == 'peer':
      bp.allreduce_rho(rho, e_kin)
    else:
      allreduce_density(rho, e_kin, buf, group=self.group)
    en, veff = bp.grid_potential(rho, xc, False)
    hq_band = bp.hpsi(q_band, veff)
    hq_rows = bands_to_rows(hq_band, self.row_blocks, self.band_blocks, self.group)
    m = rp.bwd_gram(q_rows, hq_rows)
    self._sum_small(m)
    g_re, g_im = rp.bwd_apply(q_rows, hq_rows, occ, m)
    energies = torch.stack([e_kin[0], en[1], en[0], en[2]])
    return energies, g_re, g_im, rho
