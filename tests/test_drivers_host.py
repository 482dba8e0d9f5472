"""Host logic of the drivers under `-m "not gpu"`: the GPU driver tests of tests/test_drivers_gpu.py
re-run with jrystal_b200.plan.Plan / optim.Adam replaced by the oracle-backed CPU stand-ins of
tests/emulated_plan.py.  What this covers is everything in jrystal_b200/calc/ that is not a kernel:
configuration, set-up helpers, occupation maps and their autograd chain, the Adam / temperature /
convergence loop, the k-path walk with warm start, the output containers.  The kernels themselves
are covered by the `-m gpu` runs of the same test bodies."""
import pytest

from tests import emulated_plan
from tests import test_drivers_gpu as gpu_tests


@pytest.fixture
def emulated(monkeypatch):
  emulated_plan.patch_drivers(monkeypatch)
  return 0  # stands in for the cuda_device fixture value


def test_energy_driver_follows_the_oracle_trajectory(emulated):
  gpu_tests.test_energy_driver_follows_the_oracle_trajectory(emulated, False)


@pytest.mark.parametrize('method', ['simplex-projector', 'idempotent'])
def test_energy_driver_with_trainable_occupations(emulated, method):
  gpu_tests.test_energy_driver_with_trainable_occupations(emulated, method)


def test_energy_driver_converges_and_stops(emulated):
  gpu_tests.test_energy_driver_converges_and_stops(emulated)


def test_band_driver_eigenvalues_are_variational_and_close(emulated):
  gpu_tests.test_band_driver_eigenvalues_are_variational_and_close(emulated)


# ---------------------------------------------------------------------------------------------
# N > 1: the k-mesh layout of the energy drivers on two gloo ranks (one k-point each)
# ---------------------------------------------------------------------------------------------

def _kmesh_config(upf_dir, normcons):
  from jrystal_b200.config import get_config
  kw = dict(crystal='si', grid_sizes=12, k_grid_sizes=[1, 1, 2], cutoff_energy=6.0, empty_bands=2,
            epoch=4, verbose=False, orbital_grid='full', convergence_condition=1e-12, seed=3,
            smearing=0.01)
  if normcons:
    kw.update(use_pseudopotential=True, pseudopotential_file_dir=upf_dir,
              occupation='simplex-projector')
  else:
    kw.update(crystal='diamond', cutoff_energy=10, occupation='uniform')
  return get_config(**kw)


def _patch_plain():
  import torch
  torch.set_num_threads(2)  # two ranks share the box with the test runner
  from jrystal_b200.calc import (calc_band_structure_all_electrons as band,
                                 calc_ground_state_energy_all_electrons as energy, opt_utils)
  energy.Plan = band.Plan = emulated_plan.EmulatedPlan
  band.Adam = opt_utils.Adam = emulated_plan.EmulatedAdam
  torch.cuda.synchronize = lambda *a, **k: None


def _rank_kmesh(rank, world, port, upf_dir, out_dir):
  import os
  import numpy as np
  import torch.distributed as dist
  from jrystal_b200 import calc
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    _patch_plain()
    for normcons in (False, True):
      cfg = _kmesh_config(upf_dir, normcons)
      cfg.parallel_over_k_mesh = True
      out = calc.energy(cfg)
      assert tuple(out.params_pw['w_re'].shape)[1] == 1       # this rank's k-point only
      np.save(os.path.join(out_dir, f'hist{int(normcons)}_{rank}.npy'),
              np.array(out.total_energy_history + [out.total_energy]
                       + [out.energies[k] for k in sorted(out.energies)]))
      np.save(os.path.join(out_dir, f'rho{int(normcons)}_{rank}.npy'), out.density.numpy())
  finally:
    dist.destroy_process_group()


def test_gloo_world2_energy_drivers_on_a_k_mesh(emulated, tmp_path):
  """parallel_over_k_mesh on 2 ranks == 1 rank: parameters / occupations sharded by k-point, rho and
  E_kin all-reduced between eval_begin and eval_finish, dE/d occ gathered for the occupation
  update, E_nl all-reduced for the final split (all-electron + uniform occupations; norm-conserving
  + simplex-projector occupations)."""
  import os
  import numpy as np
  import torch.multiprocessing as mp
  from jrystal_b200 import calc
  from tests.test_pseudopotential import golden, write_upf
  upf_dir = str(tmp_path / 'upf')
  os.makedirs(upf_dir)
  write_upf(os.path.join(upf_dir, 'Si.pz-vbc.UPF'), golden())
  port = 33500 + (os.getpid() % 2000)
  mp.spawn(_rank_kmesh, args=(2, port, upf_dir, str(tmp_path)), nprocs=2, join=True)
  for normcons in (False, True):
    one = calc.energy(_kmesh_config(upf_dir, normcons))
    want = np.array(one.total_energy_history + [one.total_energy]
                    + [one.energies[k] for k in sorted(one.energies)])
    for rank in range(2):
      got = np.load(tmp_path / f'hist{int(normcons)}_{rank}.npy')
      np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-12)
      np.testing.assert_allclose(np.load(tmp_path / f'rho{int(normcons)}_{rank}.npy'),
                                 one.density.numpy(), rtol=1e-9, atol=1e-12)


def _band_config():
  return gpu_tests._config(epoch=3, band_structure_empty_bands=2, band_structure_epoch=1500,
                           k_path_fine_tuning_epoch=600, num_kpoints=2, k_path_special_points='GX',
                           optimizer_args={'learning_rate': 0.02, 'b1': 0.9, 'b2': 0.99})


def _rank_kpath(rank, world, port, out_dir):
  import os
  import numpy as np
  import torch.distributed as dist
  from jrystal_b200 import calc
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    _patch_plain()
    out = calc.band(_band_config())
    np.save(os.path.join(out_dir, f'eig{rank}.npy'), out.eigenvalues)
  finally:
    dist.destroy_process_group()


def test_gloo_world2_band_driver_splits_the_k_path(emulated, tmp_path):
  """parallel_over_k_path: each rank walks its contiguous chunk of the path without communication
  (the reference's pmap), the eigenvalues are gathered at the end; every rank holds the whole band
  structure and it agrees with the one-rank walk at the convergence level of the minimisation."""
  import os
  import numpy as np
  import torch.multiprocessing as mp
  from jrystal_b200 import calc
  port = 35500 + (os.getpid() % 2000)
  mp.spawn(_rank_kpath, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  e0, e1 = np.load(tmp_path / 'eig0.npy'), np.load(tmp_path / 'eig1.npy')
  np.testing.assert_array_equal(e0, e1)
  one = calc.band(_band_config()).eigenvalues
  assert e0.shape == one.shape
  np.testing.assert_allclose(e0[0, 0], one[0, 0], rtol=0, atol=1e-10)   # same first k-point, same walk
  assert np.abs(e0 - one).max() < 2e-3                                   # second: cold vs warm start


# ---------------------------------------------------------------------------------------------
# spin-unrestricted energy mode (spin_restricted: false), both backends
# ---------------------------------------------------------------------------------------------

from tests.conftest import BACKENDS  # noqa: E402


@pytest.mark.parametrize('backend', BACKENDS, indirect=True)
def test_spin_unrestricted_energy_driver(backend):
  """Two spin channels through the energy-mode loop (pw.py:88-91 ns = 2, spin-scaled lda_x,
  xc.py:54-64): 5 Adam steps == the oracle's value_and_grad + the same Adam; per-spin density."""
  import numpy as np
  from oracle import reference_port as rp
  from jrystal_b200 import calc
  from tests.common import relerr
  kind = backend[0]
  cfg = gpu_tests._config(epoch=5, spin_restricted=False, empty_bands=3, orbital_grid='full')
  out = calc.energy(cfg, use_cuda_graph=(kind == 'cuda'))
  c = out.crystal
  s = rp.System(c.cell_vectors, c.positions, c.charges, [12, 12, 12], k_grid_sizes=[1, 1, 2],
                cutoff_energy=10, mask_method='spherical')
  nb = int(np.ceil(c.num_electron / 2)) + 3
  shape = (2, s.num_k, s.num_g, nb)
  assert tuple(out.params_pw['w_re'].shape) == shape and out.density.shape[0] == 2
  rng = np.random.default_rng(cfg.seed)
  w_re, w_im = rng.random(shape), rng.random(shape)
  occ = rp.occupation_uniform(s.num_k, s.num_electrons, spin=c.spin, num_bands=nb,
                              spin_restricted=False).numpy()
  assert relerr(out.occupation.cpu().numpy(), occ) < 1e-14
  st = [np.zeros(shape) for _ in range(4)]
  for t in range(1, cfg.epoch + 1):
    ref = rp.energy_and_grad(s, w_re, w_im, occ)
    assert abs(out.total_energy_history[t - 1] - ref['e_tot']) < 1e-10 * abs(ref['e_tot']), t
    gpu_tests._adam_numpy(w_re, ref['g_re'], st[0], st[1], t)
    gpu_tests._adam_numpy(w_im, ref['g_im'], st[2], st[3], t)
  ref = rp.energy_and_grad(s, w_re, w_im, occ)
  assert abs(out.total_energy - out.energies['ewald'] - ref['e_tot']) < 1e-10 * abs(ref['e_tot'])
  assert relerr(out.density.cpu().numpy(), ref['density']) < 1e-8


def test_command_line_with_the_shipped_config_keys(emulated, tmp_path, capsys):
  """`python -m jrystal_b200 -m energy -c config.yaml` with the keys of the reference's shipped
  config.yaml (GGA-PBE, simplex-projector occupations, norm-conserving pseudopotentials), scaled
  down to a 12^3 grid; the pseudopotential directory is the one thing a user must point somewhere."""
  import os
  import yaml
  from jrystal_b200.__main__ import main
  from tests.test_pseudopotential import golden, write_upf
  upf_dir = str(tmp_path / 'upf')
  os.makedirs(upf_dir)
  write_upf(os.path.join(upf_dir, 'Si.pz-vbc.UPF'), golden())
  cfg = {
    'crystal': 'si', 'crystal_file_path_path': None, 'spin': 0, 'save_dir': str(tmp_path),
    'xc': 'gga_x_pbe+gga_c_pbe', 'use_pseudopotential': True, 'pseudopotential_type': 'nc',
    'pseudopotential_file_dir': upf_dir, 'freq_mask_method': 'spherical', 'cutoff_energy': 6,
    'grid_sizes': 12, 'k_grid_sizes': [1, 1, 1], 'occupation': 'simplex-projector',
    'smearing': 0.001, 'empty_bands': 2, 'spin_restricted': True,
    'ewald_args': {'ewald_eta': 0.1, 'ewald_cutoff': 2.0e4}, 'epoch': 6, 'optimizer': 'adam',
    'optimizer_args': {'learning_rate': 0.01, 'b1': 0.9, 'b2': 0.99}, 'scheduler': None,
    'convergence_window_size': 20, 'convergence_condition': 1e-4, 'seed': 123, 'verbose': True,
    'orbital_grid': 'full',
  }
  path = tmp_path / 'config.yaml'
  path.write_text(yaml.safe_dump(cfg))
  assert main(['-m', 'energy', '-c', str(path)]) == 0
  out = capsys.readouterr().out
  for line in ('Hartree Energy:', 'External (local) Energy:', 'External (nonlocal) Energy:',
               'XC Energy:', 'Kinetic Energy:', 'Nuclear repulsion Energy:', 'Total Energy:',
               'Did not converge after 6 steps'):
    assert line in out, line
  assert (tmp_path / 'density.npy').exists()


def test_command_line_band_mode_loads_the_saved_ground_state(emulated, tmp_path, capsys, monkeypatch):
  """`-m energy` with save_dir writes ground_state.npz; `-m band -l` takes the density from it
  instead of minimising again (the reference declares -l / save_dir, main.py:31-38, config.py:64,
  without wiring them) and gives the eigenvalues of the in-process flow; a file made for another
  grid is refused."""
  import numpy as np
  import yaml
  from jrystal_b200 import calc
  from jrystal_b200.__main__ import main
  from jrystal_b200.calc import calc_band_structure_all_electrons as band_mod
  from jrystal_b200.calc import ground_state_io
  from jrystal_b200.config import get_config
  cfg = dict(crystal='diamond', grid_sizes=12, k_grid_sizes=[1, 1, 2], cutoff_energy=10, empty_bands=2,
             epoch=6, occupation='uniform', xc='lda_x', verbose=False, orbital_grid='full',
             convergence_condition=1e-12, save_dir=str(tmp_path), band_structure_empty_bands=2,
             num_kpoints=2, k_path_special_points='GX', band_structure_epoch=30,
             k_path_fine_tuning_epoch=10)
  path = tmp_path / 'config.yaml'
  path.write_text(yaml.safe_dump(cfg))
  monkeypatch.chdir(tmp_path)                       # the band mode writes its .npy into the cwd
  assert main(['-m', 'energy', '-c', str(path)]) == 0
  assert (tmp_path / ground_state_io.FILE_NAME).exists()
  config = get_config(str(path))
  gs = ground_state_io.load(str(tmp_path), config)
  assert gs.steps == 6 and len(gs.total_energy_history) == 6
  assert tuple(gs.density.shape) == (1, 12, 12, 12)
  np.testing.assert_array_equal(gs.density.numpy(), np.load(tmp_path / 'density.npy'))
  assert abs(sum(gs.energies[k] for k in ('kinetic', 'external', 'hartree', 'xc', 'ewald'))
             - gs.total_energy) < 1e-12 * abs(gs.total_energy)
  expected = calc.band(config, ground_state=gs).eigenvalues

  def no_energy_run(*a, **k):
    raise AssertionError('the band mode minimised the energy again although -l was given')
  monkeypatch.setattr(band_mod, 'energy_calc', no_energy_run)
  # no -c: ./config.yaml of the working directory, the reference's default (main.py:24-29)
  assert main(['-m', 'band', '-l', str(tmp_path / ground_state_io.FILE_NAME)]) == 0
  assert 'CC_band_structure.npy' in capsys.readouterr().out
  np.testing.assert_allclose(np.load(tmp_path / 'CC_band_structure.npy'), expected, rtol=0, atol=1e-12)
  with pytest.raises(ValueError):
    ground_state_io.load(str(tmp_path), get_config(str(path), grid_sizes=16))
  with pytest.raises(ValueError):
    ground_state_io.load(str(tmp_path), get_config(str(path), xc='lda_x+lda_c_pw'))


# ---------------------------------------------------------------------------------------------
# N > 1, fewer k-points than ranks: the row/band-sharded evaluation (Gamma-only supercells, C3a)
# ---------------------------------------------------------------------------------------------

def _rank_row_sharded(rank, world, port, out_dir):
  import os
  import numpy as np
  import torch
  import torch.distributed as dist
  import jrystal_b200.plan as plan_mod
  from jrystal_b200 import parallel
  from oracle import reference_port as rp
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    torch.set_num_threads(2)
    plan_mod.Plan, plan_mod.RowsPlan = emulated_plan.EmulatedPlan, emulated_plan.EmulatedRowsPlan
    s = rp.System.from_name('diamond', [12, 12, 12], [1, 1, 1], 10.0)     # Gamma only
    nb = 7                                                                 # odd: uneven band blocks
    p = rp.param_init(3, nb, s.num_k, s.mask)
    occ = rp.occupation_uniform(s.num_k, s.num_electrons, num_bands=nb).numpy()
    occ = occ * (1.0 + 0.1 * np.random.default_rng(4).random(occ.shape))
    ev = parallel.RowShardedEvaluator(s.cell, s.mask, s.kpts, nb, s.positions, s.charges)
    w_re = torch.from_numpy(np.ascontiguousarray(p['w_re'][:, :, ev.g0:ev.g1]))
    w_im = torch.from_numpy(np.ascontiguousarray(p['w_im'][:, :, ev.g0:ev.g1]))
    en, g_re, g_im, rho = ev.evaluate(w_re, w_im, torch.from_numpy(occ), 'lda_x')
    np.savez(os.path.join(out_dir, f'rows{rank}.npz'), en=en.numpy(), g_re=g_re.numpy(),
             g_im=g_im.numpy(), rho=rho.numpy(), g0=ev.g0, g1=ev.g1, b0=ev.b0, b1=ev.b1)
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 4])
def test_gloo_world2_row_band_sharded_evaluation(tmp_path, world):
  """parallel.RowShardedEvaluator on 2 and 4 ranks (rows of the parameters for the QR, bands for
  the FFT work, two all-to-alls, Gram / rho / E_kin all-reduces) == the oracle's single evaluation:
  energies, density and the row blocks of both gradients.  7 bands: blocks of 4 + 3 on two ranks,
  2 + 2 + 2 + 1 on four (the driver's scaling run stops at 1, 2, 4 and 8 GPUs)."""
  import os
  import numpy as np
  import torch.multiprocessing as mp
  from oracle import reference_port as rp
  port = 37500 + (os.getpid() % 2000) + world
  mp.spawn(_rank_row_sharded, args=(world, port, str(tmp_path)), nprocs=world, join=True)
  s = rp.System.from_name('diamond', [12, 12, 12], [1, 1, 1], 10.0)
  nb = 7
  p = rp.param_init(3, nb, s.num_k, s.mask)
  occ = rp.occupation_uniform(s.num_k, s.num_electrons, num_bands=nb).numpy()
  occ = occ * (1.0 + 0.1 * np.random.default_rng(4).random(occ.shape))
  ref = rp.energy_and_grad(s, p['w_re'], p['w_im'], occ)
  rows = 0
  for rank in range(world):
    d = np.load(tmp_path / f'rows{rank}.npz')
    g0, g1 = int(d['g0']), int(d['g1'])
    rows += g1 - g0
    np.testing.assert_allclose(d['en'], [ref['e_kin'], ref['e_ext'], ref['e_har'], ref['e_xc']],
                               rtol=1e-10)
    assert np.abs(d['rho'] - ref['density']).max() < 1e-10 * np.abs(ref['density']).max()
    scale = np.abs(ref['g_re']).max()
    assert np.abs(d['g_re'] - ref['g_re'][:, :, g0:g1]).max() < 1e-9 * scale
    assert np.abs(d['g_im'] - ref['g_im'][:, :, g0:g1]).max() < 1e-9 * scale
    assert int(d['b1']) - int(d['b0']) in ((3, 4) if world == 2 else (1, 2))
  assert rows == s.num_g


def _rank_k_sharded(rank, world, port, out_dir):
  import os
  import numpy as np
  import torch
  import torch.distributed as dist
  import jrystal_b200.plan as plan_mod
  from jrystal_b200 import parallel
  from oracle import reference_port as rp
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    torch.set_num_threads(2)
    plan_mod.Plan = emulated_plan.EmulatedPlan
    s = rp.System.from_name('diamond', [12, 12, 12], [1, 2, 2], 10.0)     # 4 k-points, 2 per rank
    nb = 6
    p = rp.param_init(3, nb, s.num_k, s.mask)
    occ = rp.occupation_uniform(s.num_k, s.num_electrons, num_bands=nb).numpy()
    occ = occ * (1.0 + 0.1 * np.random.default_rng(4).random(occ.shape))
    ev = parallel.KShardedEvaluator(s.cell, s.mask, s.kpts, nb, s.positions, s.charges)
    assert ev.reduce_path == 'nccl'     # = torch.distributed (gloo here): no peer memory on the CPU
    sl = slice(ev.k0, ev.k1)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    en, g_re, g_im, rho = ev.evaluate(t(p['w_re'][:, sl]), t(p['w_im'][:, sl]), t(occ[:, sl]))
    np.savez(os.path.join(out_dir, f'k{rank}.npz'), en=en.numpy(), g_re=g_re.numpy(),
             g_im=g_im.numpy(), rho=rho.numpy(), k0=ev.k0, k1=ev.k1)
  finally:
    dist.destroy_process_group()


def test_gloo_world2_k_sharded_evaluator(tmp_path):
  """parallel.KShardedEvaluator on 2 ranks (whole k-points per rank, the reference's k-mesh layout,
  calc_ground_state_energy_all_electrons.py:83-91; rho + E_kin all-reduced) == the oracle's single
  evaluation: energies, density and each rank's k-block of both gradients."""
  import os
  import numpy as np
  import torch.multiprocessing as mp
  from oracle import reference_port as rp
  port = 39500 + (os.getpid() % 2000)
  mp.spawn(_rank_k_sharded, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  s = rp.System.from_name('diamond', [12, 12, 12], [1, 2, 2], 10.0)
  nb = 6
  p = rp.param_init(3, nb, s.num_k, s.mask)
  occ = rp.occupation_uniform(s.num_k, s.num_electrons, num_bands=nb).numpy()
  occ = occ * (1.0 + 0.1 * np.random.default_rng(4).random(occ.shape))
  ref = rp.energy_and_grad(s, p['w_re'], p['w_im'], occ)
  seen = 0
  for rank in range(2):
    d = np.load(tmp_path / f'k{rank}.npz')
    k0, k1 = int(d['k0']), int(d['k1'])
    seen += k1 - k0
    np.testing.assert_allclose(d['en'], [ref['e_kin'], ref['e_ext'], ref['e_har'], ref['e_xc']],
                               rtol=1e-10)
    assert np.abs(d['rho'] - ref['density']).max() < 1e-10 * np.abs(ref['density']).max()
    scale = np.abs(ref['g_re']).max()
    assert np.abs(d['g_re'] - ref['g_re'][:, k0:k1]).max() < 1e-9 * scale
    assert np.abs(d['g_im'] - ref['g_im'][:, k0:k1]).max() < 1e-9 * scale
  assert seen == s.num_k


def test_create_crystal_checks_the_spin_number_and_prefers_the_builtin_name(tmp_path):
  """opt_utils.py:105-114 of the reference: check_spin_number on the created crystal (the default
  `spin: 0` with the 13 electrons of al_primitive must raise, not drop an electron) and a built-in
  `crystal` name wins over `crystal_file_path_path`."""
  from jrystal_b200.calc.opt_utils import create_crystal, create_optimizer
  from jrystal_b200.config import get_config
  with pytest.raises(ValueError):
    create_crystal(get_config(crystal='al_primitive', spin=0))
  assert create_crystal(get_config(crystal='al_primitive', spin=1)).num_electron == 13
  c = create_crystal(get_config(crystal='si', crystal_file_path_path=str(tmp_path / 'missing.xyz'), spin=0))
  assert c.num_electron == 28


def test_ground_state_file_rejects_other_settings(tmp_path):
  """ground_state_io.load compares every setting the density depends on (cut-off, mask method,
  k-mesh, occupation scheme, smearing, spin) and refuses the k-block of a sharded run."""
  import torch
  from jrystal_b200.calc import ground_state_io
  from jrystal_b200.calc.calc_ground_state_energy_all_electrons import GroundStateEnergyOutput
  from jrystal_b200.calc.opt_utils import create_crystal
  from jrystal_b200.config import get_config
  cfg = get_config(crystal='diamond', grid_sizes=8, k_grid_sizes=[1, 1, 2], cutoff_energy=10.0,
                   save_dir=str(tmp_path))
  out = GroundStateEnergyOutput(
    config=cfg, crystal=create_crystal(cfg),
    params_pw={'w_re': torch.zeros(1, 2, 5, 3), 'w_im': torch.zeros(1, 2, 5, 3)},
    occupation=torch.zeros(1, 2, 3), density=torch.zeros(1, 8, 8, 8), total_energy=-1.0,
    energies={'kinetic': 1.0}, total_energy_history=[-1.0], converged=True, steps=1,
    seconds_per_step=0.0, k_range=(0, 2))
  path = ground_state_io.save(out, str(tmp_path))
  assert ground_state_io.load(path, cfg).total_energy == -1.0
  for key, val in [('cutoff_energy', 12.0), ('k_grid_sizes', [1, 1, 1]), ('occupation', 'gamma'),
                   ('freq_mask_method', 'cubic'), ('smearing', 0.01), ('spin_restricted', False)]:
    with pytest.raises(ValueError):
      ground_state_io.load(path, get_config(**{**dict(cfg), key: val}))
  out.k_range = (0, 1)
  path = ground_state_io.save(out, str(tmp_path / 'shard'))
  with pytest.raises(ValueError):
    ground_state_io.load(path, cfg)
