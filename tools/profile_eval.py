#!/usr/bin/env python
"""Minimal driver for ncu captures: builds a workload plan and runs N evaluations."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import jrystal_b200 as jb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--config', default='C2')
ap.add_argument('--evals', type=int, default=2)
ap.add_argument('--orbital-grid', default='auto')
args = ap.parse_args()
wl = bench.build_workload(args.config)
c = wl['crystal']
nk = wl['kpts'].shape[0]
plan = jb.Plan(c.cell_vectors, wl['mask'], wl['kpts'], wl['nb'],
               orbital_grid=bench.parse_orbital_grid(args.orbital_grid))
plan.set_atoms(c.positions, c.charges)
w_re_h, w_im_h = bench.synthetic_params(wl['ng'], nk, wl['nb'], 0, nk)
w_re, w_im = torch.from_numpy(w_re_h).cuda(), torch.from_numpy(w_im_h).cuda()
occ = torch.from_numpy(wl['occ']).cuda()
for _ in range(args.evals):
  en, g_re, g_im, _, rho = plan.eval(w_re, w_im, occ, 'lda_x')
torch.cuda.synchronize()
print('energies', en.cpu().numpy())
