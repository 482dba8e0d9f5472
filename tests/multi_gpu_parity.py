#!/usr/bin/env python
"""Worker of tests/test_full_size_oracle_gpu.py::test_two_gpus_equal_one_gpu (not collected):
  torchrun --nproc-per-node N --master-addr 127.0.0.1 tests/multi_gpu_parity.py k|rows [--tol T]
Every rank runs the WHOLE evaluation on its own GPU, then the sharded one over NCCL, and compares
energies, density and its own block of the gradient.  Prints 'PARITY OK' per rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jrystal_b200 as jb  # noqa: E402
from jrystal_b200 import parallel  # noqa: E402
from tests.common import make_inputs, make_system, relerr  # noqa: E402

layout = sys.argv[1] if len(sys.argv) > 1 else 'k'
tol = float(sys.argv[sys.argv.index('--tol') + 1]) if '--tol' in sys.argv else 1e-12
rank, lrank, world = int(os.environ['RANK']), int(os.environ['LOCAL_RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(lrank)
dist.init_process_group('nccl', device_id=torch.device('cuda', lrank))
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()

if layout == 'k':
  # whole k-points per rank (calc_ground_state_energy_all_electrons.py:83-91; nk % world == 0)
  s = make_system('si8', 32, [2, 2, 2], 8.0, 'spherical')
  nb = 40
else:
  # Gamma only: rows for the QR, bands for the FFTs
  s = make_system('si8', 32, [1, 1, 1], 8.0, 'spherical')
  nb = 130
w_re, w_im, occ = make_inputs(s, nb, jitter=0.1)
plan = jb.Plan(s.cell, s.mask, s.kpts, nb, orbital_grid='auto')
plan.set_atoms(s.positions, s.charges)
occ_d = dev(occ)
rho, e_kin = plan.eval_begin(dev(w_re), dev(w_im), occ_d)
en, g_re, g_im, _ = plan.eval_finish(occ_d, rho, e_kin, 'lda_x')
plan.check_status()

if layout == 'k':
  k0, k1 = parallel.shard_kpoints(s.num_k, world, rank)
  ev = parallel.KShardedEvaluator(s.cell, s.mask, s.kpts, nb, s.positions, s.charges,
                                  orbital_grid='auto')
  assert (ev.k0, ev.k1) == (k0, k1)
  en2, g_re2, g_im2, rho2 = ev.evaluate(dev(w_re[:, k0:k1]), dev(w_im[:, k0:k1]), dev(occ[:, k0:k1]))
  ev.plan.check_status()
  own = (slice(None), slice(k0, k1))
  what = f'k [{k0},{k1}) via {ev.reduce_path}'
else:
  ev = parallel.RowShardedEvaluator(s.cell, s.mask, s.kpts, nb, s.positions, s.charges,
                                    orbital_grid='auto')
  g0, g1 = ev.g0, ev.g1
  en2, g_re2, g_im2, rho2 = ev.evaluate(dev(w_re[:, :, g0:g1]), dev(w_im[:, :, g0:g1]), occ_d)
  ev.rows.check_status()
  own = (slice(None), slice(None), slice(g0, g1))
  what = f'rows [{g0},{g1}) bands [{ev.b0},{ev.b1})'
torch.cuda.synchronize()
de = abs(en2.sum().item() - en.sum().item()) / abs(en.sum().item())
ds = float((en2 - en).abs().max().item() / abs(en.sum().item()))
gmax = max(g_re.abs().max().item(), g_im.abs().max().item())
dg = max((g_re2 - g_re[own]).abs().max().item(), (g_im2 - g_im[own]).abs().max().item()) / gmax
dr = relerr(rho2.cpu().numpy(), rho.cpu().numpy())
print(f'rank {rank}/{world} layout {layout}: {what}  E rel {de:.2e} split {ds:.2e} '
      f'grad rel {dg:.2e} rho rel {dr:.2e}', flush=True)
assert de < tol and ds < tol and dg < 100 * tol and dr < 10 * tol, (de, ds, dg, dr)
# every rank must hold the SAME density bits (the replicated grid part relies on it)
chk = torch.stack([rho2.double().sum(), (rho2 * rho2).sum()])
lo, hi = chk.clone(), chk.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
assert torch.equal(lo, hi), 'ranks disagree on the reduced density'
print('PARITY OK', flush=True)
dist.barrier()
dist.destroy_process_group()
