"""Shared helpers of the parity tests: seeded systems + oracle evaluation."""
import numpy as np
import torch

from oracle import reference_port as rp


def make_system(name, grid, kgrid, cutoff=None, mask_method='spherical', reps=None):
  return rp.System.from_name(name, grid, kgrid, cutoff, reps=reps, mask_method=mask_method)


def make_inputs(system, nb, seed=123, occ='uniform', jitter=0.0):
  p = rp.param_init(seed, nb, system.num_k, system.mask)
  if occ == 'uniform':
    o = rp.occupation_uniform(system.num_k, system.num_electrons, num_bands=nb).numpy()
  elif occ == 'ones':
    o = np.ones((1, system.num_k, nb))
  else:
    raise ValueError(occ)
  if jitter:
    o = o * (1.0 + jitter * np.random.default_rng(seed + 1).random(o.shape))
  return p['w_re'], p['w_im'], o


def make_plan(system, nb, **kw):
  import jrystal_b200 as jb
  plan = jb.Plan(system.cell, system.mask, system.kpts, nb, **kw)
  plan.set_atoms(system.positions, system.charges)
  return plan


def to_dev(a, dtype=None):
  t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
  return t if dtype is None else t.to(dtype)


def relerr(a, b):
  a = np.asarray(a)
  b = np.asarray(b)
  return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
