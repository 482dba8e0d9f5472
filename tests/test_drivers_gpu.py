"""Drivers above the hot path: device Adam, the energy-mode loop (eager and
CUDA-graph replay) and the band-mode walk, against the oracle driven by the same optimiser
arithmetic on the CPU."""
import numpy as np
import pytest
import torch

from oracle import reference_port as rp
from tests.common import relerr

pytestmark = pytest.mark.gpu


def _adam_numpy(p, g, m, v, t, lr=0.01, b1=0.9, b2=0.99, eps=1e-8):
  m[:] = b1 * m + (1 - b1) * g
  v[:] = b2 * v + (1 - b2) * g * g
  p -= lr * (m / (1 - b1 ** t)) / (np.sqrt(v / (1 - b2 ** t)) + eps)


def test_adam_matches_optax_formula(cuda_device):
  from jrystal_b200.optim import Adam
  rng = np.random.default_rng(0)
  n = 100003  # odd: exercises the scalar tail
  p0 = rng.standard_normal(n)
  p = torch.from_numpy(p0.copy()).cuda()
  opt = Adam([p], learning_rate=0.02, b1=0.8, b2=0.95, eps=1e-7)
  pr, m, v = p0.copy(), np.zeros(n), np.zeros(n)
  for t in range(1, 5):
    g = rng.standard_normal(n)
    opt.step([torch.from_numpy(g).cuda()])
    _adam_numpy(pr, g, m, v, t, 0.02, 0.8, 0.95, 1e-7)
  assert opt.step_count == 4
  assert relerr(p.cpu().numpy(), pr) < 1e-14


def _config(**kw):
  from jrystal_b200.config import get_config
  base = dict(crystal='diamond', grid_sizes=12, k_grid_sizes=[1, 1, 2], cutoff_energy=10,
              empty_bands=2, epoch=12, convergence_window_size=5, convergence_condition=1e-12,
              verbose=False, seed=7)
  base.update(kw)
  return get_config(**base)


@pytest.mark.parametrize('graph', [False, True])
def test_energy_driver_follows_the_oracle_trajectory(cuda_device, graph):
  """12 Adam steps of the energy-mode loop == 12 steps of oracle value_and_grad + the same Adam
  on the CPU (energies to 1e-10 relative at every step)."""
  from jrystal_b200 import calc
  cfg = _config()
  out = calc.energy(cfg, use_cuda_graph=graph)
  assert out.steps == cfg.epoch and not out.converged
  c = out.crystal
  s = rp.System(c.cell_vectors, c.positions, c.charges, [12, 12, 12], k_grid_sizes=[1, 1, 2],
                cutoff_energy=10, mask_method='spherical')
  nb = int(np.ceil(c.num_electron / 2)) + 2
  rng = np.random.default_rng(cfg.seed)
  shape = (1, s.num_k, s.num_g, nb)
  w_re, w_im = rng.random(shape), rng.random(shape)
  occ = rp.occupation_uniform(s.num_k, s.num_electrons, num_bands=nb).numpy()
  st = [np.zeros(shape) for _ in range(4)]
  for t in range(1, cfg.epoch + 1):
    ref = rp.energy_and_grad(s, w_re, w_im, occ)
    assert abs(out.total_energy_history[t - 1] - ref['e_tot']) < 1e-10 * abs(ref['e_tot']), t
    _adam_numpy(w_re, ref['g_re'], st[0], st[1], t)
    _adam_numpy(w_im, ref['g_im'], st[2], st[3], t)
  ref = rp.energy_and_grad(s, w_re, w_im, occ)
  assert abs(out.total_energy - out.energies['ewald'] - ref['e_tot']) < 1e-10 * abs(ref['e_tot'])
  assert relerr(out.density.cpu().numpy(), ref['density']) < 1e-8
  assert out.total_energy_history[-1] < out.total_energy_history[0]


@pytest.mark.parametrize('method', ['simplex-projector', 'idempotent'])
def test_energy_driver_with_trainable_occupations(cuda_device, method):
  """Free-energy minimisation over plane-wave AND occupation parameters (the reference's default
  occupation scheme): 8 steps against the oracle (evaluation + literal occupation map + entropy,
  torch autograd) with the same Adam and temperature schedule."""
  from jrystal_b200 import calc
  from jrystal_b200.calc.calc_ground_state_energy_all_electrons import temperature_scheduler
  cfg = _config(epoch=8, occupation=method, smearing=0.01)
  out = calc.energy(cfg)
  c = out.crystal
  s = rp.System(c.cell_vectors, c.positions, c.charges, [12, 12, 12], k_grid_sizes=[1, 1, 2],
                cutoff_energy=10, mask_method='spherical')
  nb = int(np.ceil(c.num_electron / 2)) + 2
  ne = int(c.num_electron)
  rng = np.random.default_rng(cfg.seed)
  if method == 'simplex-projector':
    n = nb * s.num_k
    v = ((np.arange(n) - n // 2) * 0.1).reshape(s.num_k, nb)
    leaves = [torch.tensor(v.copy(), requires_grad=True), torch.tensor(v.copy(), requires_grad=True)]
    occ_fn = lambda: rp.occupation_simplex_projector(leaves[0], leaves[1], ne)
  else:
    w = rng.random((nb * s.num_k, (ne // 2) * s.num_k))
    leaves = [torch.tensor(w, requires_grad=True)]
    occ_fn = lambda: rp.occupation_idempotent(leaves[0], leaves[0], s.num_k)
  shape = (1, s.num_k, s.num_g, nb)
  w_re, w_im = rng.random(shape), rng.random(shape)
  st = [np.zeros(shape) for _ in range(4)]
  st_occ = [(np.zeros(l.shape), np.zeros(l.shape)) for l in leaves]
  sched = temperature_scheduler(cfg)
  for t in range(1, cfg.epoch + 1):
    o = occ_fn()
    ref = rp.energy_and_grad(s, w_re, w_im, o.detach().numpy(), occ_grad=True)
    assert abs(out.total_energy_history[t - 1] - ref['e_tot']) < 1e-10 * abs(ref['e_tot']), t
    sur = (torch.from_numpy(ref['g_occ']) * o).sum() - sched(t - 1) * rp.entropy_fermi_dirac(o, cfg.eps)
    grads = torch.autograd.grad(sur, leaves)
    _adam_numpy(w_re, ref['g_re'], st[0], st[1], t)
    _adam_numpy(w_im, ref['g_im'], st[2], st[3], t)
    for l, g, (m, v) in zip(leaves, grads, st_occ):
      p = l.detach().numpy().copy()
      _adam_numpy(p, g.numpy(), m, v, t)
      with torch.no_grad():
        l.copy_(torch.from_numpy(p))
  o = occ_fn().detach()
  assert relerr(out.occupation.cpu().numpy(), o.numpy()) < 1e-9
  assert abs(float(out.occupation.sum()) - ne) < 1e-9


def test_energy_driver_converges_and_stops(cuda_device):
  from jrystal_b200 import calc
  cfg = _config(epoch=400, convergence_window_size=5, convergence_condition=5e-2)
  out = calc.energy(cfg)
  assert out.converged and out.steps < 400


def test_band_driver_eigenvalues_are_variational_and_close(cuda_device):
  """Band mode on a tiny cell: subspace eigenvalues from the driver lie above the exact lowest
  eigenvalues of the dense plane-wave Hamiltonian (variational) and approach them."""
  from jrystal_b200 import calc
  cfg = _config(epoch=3, band_structure_empty_bands=2, band_structure_epoch=1500,
                k_path_fine_tuning_epoch=600, num_kpoints=2, k_path_special_points='GX',
                optimizer_args={'learning_rate': 0.02, 'b1': 0.9, 'b2': 0.99})
  out = calc.band(cfg)
  c = out.ground_state.crystal
  assert out.eigenvalues.shape == (1, 2, int(np.ceil(c.num_electron / 2)) + 2)
  rho = out.ground_state.density.cpu()
  for ik in range(2):
    s = rp.System(c.cell_vectors, c.positions, c.charges, [12, 12, 12],
                  kpts=out.k_path[ik:ik + 1], cutoff_energy=10, mask_method='spherical')
    v = rp.effective(rho, s.positions, s.charges, s.g_vec, s.vol, False, 'lda_x', True)
    v_hat = torch.fft.fftn(v[0].real.to(torch.complex128)) / s.mask.size
    idx = np.argwhere(s.mask)
    d = (idx[:, None, :] - idx[None, :, :]) % np.array(s.mask.shape)
    h = v_hat[d[..., 0], d[..., 1], d[..., 2]].numpy()
    gk = s.g_vec[s.mask] + out.k_path[ik]
    h = h + np.diag(0.5 * np.sum(gk * gk, axis=1))
    exact = np.linalg.eigvalsh(h)[:out.eigenvalues.shape[-1]]
    got = out.eigenvalues[0, ik]
    assert (got >= exact - 1e-9).all()
    assert np.abs(got - exact).max() < 1e-3
