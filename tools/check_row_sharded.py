#!/usr/bin/env python
"""N-rank check of the row-sharded (Gamma-only) evaluation against the single-GPU one:
  torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/check_row_sharded.py [case]
Every rank runs the full evaluation on its own GPU and compares its row block of the gradient,
the energies and the density with what RowShardedEvaluator returns over NCCL."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jrystal_b200 as jb  # noqa: E402
from jrystal_b200.parallel import RowShardedEvaluator  # noqa: E402
from tests.common import make_inputs, make_system, relerr  # noqa: E402

CASES = {
  'si8_32_nb130': ('si8', 32, [1, 1, 1], 8, 130),
  'diamond_16': ('diamond', 16, [1, 1, 1], 20, 9),
  'si8_64_k2': ('si8', 64, [1, 1, 2], 30, 70),
}
case = sys.argv[1] if len(sys.argv) > 1 else 'si8_32_nb130'
rank, lrank = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lrank)
dist.init_process_group('nccl', device_id=torch.device('cuda', lrank))
name, grid, kgrid, cutoff, nb = CASES[case]
s = make_system(name, grid, kgrid, cutoff, 'spherical')
w_re, w_im, occ = make_inputs(s, nb, jitter=0.1)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
plan = jb.Plan(s.cell, s.mask, s.kpts, nb)
plan.set_atoms(s.positions, s.charges)
occ_d = dev(occ)
rho, e_kin = plan.eval_begin(dev(w_re), dev(w_im), occ_d)
en, g_re, g_im, _ = plan.eval_finish(occ_d, rho, e_kin, 'lda_x')
ev = RowShardedEvaluator(s.cell, s.mask, s.kpts, nb, s.positions, s.charges, orbital_grid='auto')
g0, g1 = ev.g0, ev.g1
en2, g_re2, g_im2, rho2 = ev.evaluate(dev(w_re[:, :, g0:g1]), dev(w_im[:, :, g0:g1]), occ_d)
ev.rows.check_status()
torch.cuda.synchronize()
de = abs(en2.sum().item() - en.sum().item()) / abs(en.sum().item())
dg = max(relerr(g_re2.cpu().numpy(), g_re[:, :, g0:g1].cpu().numpy()),
         relerr(g_im2.cpu().numpy(), g_im[:, :, g0:g1].cpu().numpy()))
dr = relerr(rho2.cpu().numpy(), rho.cpu().numpy())
print(f'rank {rank}/{ev.world} case {case}: rows [{g0},{g1}) bands [{ev.b0},{ev.b1}) '
      f'E rel {de:.2e} grad rel {dg:.2e} rho rel {dr:.2e}', flush=True)
assert de < 1e-11 and dg < 1e-9 and dr < 1e-10
dist.barrier()
dist.destroy_process_group()
