"""Tuning aid: phase clocks of the factor + inverse kernel (JRB_FS_TIMING=1) for the C2 (nb = 66,
8 k-points) and C3a (nb = 208, Gamma) shapes, and the time of jrb_qr_fwd."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench, jrystal_b200 as jb
for name, nk in (('C2', 8), ('C3a', 1)):
  wl = bench.build_workload(name)
  c = wl['crystal']
  plan = jb.Plan(c.cell_vectors, wl['mask'], wl['kpts'][:nk], wl['nb'], orbital_grid='auto')
  w_re, w_im = bench.synthetic_params(wl['ng'], wl['kpts'].shape[0], wl['nb'], 0, nk)
  w_re, w_im = torch.from_numpy(w_re).cuda(), torch.from_numpy(w_im).cuda()
  q = torch.empty(w_re.shape, dtype=torch.complex128, device='cuda')
  r = torch.empty((1, nk, wl['nb'], wl['nb']), dtype=torch.complex128, device='cuda')
  for i in range(2):
    plan.qr_fwd(w_re, w_im, out=(q, r))
  torch.cuda.synchronize()
  if not os.environ.get('JRB_FS_TIMING'):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
      plan.qr_fwd(w_re, w_im, out=(q, r))
    e1.record()
    torch.cuda.synchronize()
    qk = q[0, 0]
    err = (qk.conj().T @ qk - torch.eye(wl['nb'], dtype=torch.complex128, device='cuda')).abs().max().item()
    print(name, 'nk', nk, 'nb', wl['nb'], 'qr_fwd ms', e0.elapsed_time(e1) / 10, 'orth err', err, flush=True)
  del plan
