"""Host-side logic (no GPU): grid / occupation / crystal mirrors against the oracle, k-point
and band sharding, and the N>1 density all-reduce over gloo with world_size 2."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jrystal_b200 import grid, occupation, parallel
from jrystal_b200.crystal import Crystal
from oracle import reference_port as rp
from oracle import structures


@pytest.mark.parametrize('name', ['diamond', 'si', 'si8', 'diamond8'])
def test_builtin_crystals_match_shipped_geometry(name):
  cell, pos, chg = structures.load(name)
  c = Crystal.create_builtin(name)
  np.testing.assert_allclose(c.cell_vectors, cell, atol=1e-7)
  np.testing.assert_allclose(c.positions, pos, atol=1e-7)
  np.testing.assert_array_equal(c.charges, chg)


def test_grid_mirrors_oracle():
  cell, _, _ = structures.load('diamond')
  for gs in ([7, 8, 9], [16, 12, 24]):
    np.testing.assert_allclose(grid.g_vectors(cell, gs), rp.g_vectors(cell, gs), atol=1e-14)
    np.testing.assert_allclose(grid.r_vectors(cell, gs), rp.r_vectors(cell, gs), atol=1e-14)
    np.testing.assert_array_equal(grid.cubic_mask(gs), rp.cubic_mask(gs))
    np.testing.assert_array_equal(grid.spherical_mask(cell, gs, 15.0),
                                  rp.spherical_mask(cell, gs, 15.0))
  np.testing.assert_allclose(grid.k_vectors(cell, [2, 3, 1]), rp.k_vectors(cell, [2, 3, 1]),
                             atol=1e-15)
  assert list(grid.proper_grid_size([11, 13, 17])) == [12, 14, 18]
  with pytest.raises(ValueError):
    grid.fft_factor(4096)


def test_occupation_mirrors_oracle():
  np.testing.assert_array_equal(occupation.uniform(4, 12, 0, 10),
                                rp.occupation_uniform(4, 12, num_bands=10).numpy())
  np.testing.assert_array_equal(occupation.gamma(4, 12, 0, 10),
                                rp.occupation_gamma(4, 12, num_bands=10).numpy())
  f = occupation.uniform(2, 8, 0, 6)
  assert abs(occupation.fermi_dirac_entropy(f) -
             rp.entropy_fermi_dirac(torch.from_numpy(f)).item()) < 1e-12
  with pytest.raises(ValueError):
    occupation.param_init(None, 4, 8, 2, method='nope')


def test_sharding():
  assert [parallel.shard_kpoints(64, 8, r) for r in (0, 7)] == [(0, 8), (56, 64)]
  with pytest.raises(ValueError):
    parallel.shard_kpoints(6, 4, 0)   # reference: nk % ndev == 0 (spmd/uniform.py:22-24)
  blocks = [parallel.shard_bands(208, 8, r) for r in range(8)]
  assert blocks[0] == (0, 26) and blocks[-1] == (182, 208)
  blocks = [parallel.shard_bands(10, 4, r) for r in range(4)]
  assert blocks == [(0, 3), (3, 6), (6, 8), (8, 10)]


def _rank_density(rank, world, port, out_dir):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    s = rp.System.from_name('diamond', [7, 8, 9], [2, 1, 1], mask_method='cubic')
    nb = 6
    p = rp.param_init(5, nb, s.num_k, s.mask)
    occ = rp.occupation_uniform(s.num_k, s.num_electrons, num_bands=nb)
    k0, k1 = parallel.shard_kpoints(s.num_k, world, rank)
    c = rp.coeff(torch.from_numpy(p['w_re'][:, k0:k1]), torch.from_numpy(p['w_im'][:, k0:k1]),
                 s.mask)
    rho = rp.density_grid(c, s.vol, occ[:, k0:k1]).contiguous()
    e_kin = rp.energy_kinetic(s.g_vec, s.kpts[k0:k1], c, occ[:, k0:k1]).reshape(1).contiguous()
    parallel.allreduce_density(rho, e_kin)
    np.save(os.path.join(out_dir, f'rho{rank}.npy'), rho.numpy())
    np.save(os.path.join(out_dir, f'ekin{rank}.npy'), e_kin.numpy())
  finally:
    dist.destroy_process_group()


def test_gloo_world2_density_allreduce(tmp_path):
  """k-sharded partial densities summed over 2 ranks equal the unsharded density (the only
  data-path collective of the evaluation; pw.py:278 under the reference's k mesh)."""
  port = 29500 + (os.getpid() % 2000)
  mp.spawn(_rank_density, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  s = rp.System.from_name('diamond', [7, 8, 9], [2, 1, 1], mask_method='cubic')
  nb = 6
  p = rp.param_init(5, nb, s.num_k, s.mask)
  occ = rp.occupation_uniform(s.num_k, s.num_electrons, num_bands=nb)
  c = rp.coeff(torch.from_numpy(p['w_re']), torch.from_numpy(p['w_im']), s.mask)
  rho = rp.density_grid(c, s.vol, occ).numpy()
  e_kin = rp.energy_kinetic(s.g_vec, s.kpts, c, occ).item()
  for r in range(2):
    np.testing.assert_allclose(np.load(tmp_path / f'rho{r}.npy'), rho, rtol=1e-13, atol=1e-15)
    assert abs(np.load(tmp_path / f'ekin{r}.npy')[0] - e_kin) < 1e-12 * abs(e_kin)


def _rank_transposes(rank, world, port, out_dir):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    ng, nb, nk = 23, 7, 2
    rng = np.random.default_rng(0)
    full = torch.from_numpy(rng.standard_normal((1, nk, ng, nb)) +
                            1j * rng.standard_normal((1, nk, ng, nb)))
    rows = [parallel.shard_rows(ng, world, r) for r in range(world)]
    bands = [parallel.shard_bands(nb, world, r) for r in range(world)]
    g0, g1 = rows[rank]
    b0, b1 = bands[rank]
    x_rows = full[:, :, g0:g1].contiguous()
    x_bands = parallel.rows_to_bands(x_rows, rows, bands)
    assert torch.equal(x_bands, full[..., b0:b1])
    back = parallel.bands_to_rows(x_bands, rows, bands)
    assert torch.equal(back, x_rows)
    s = torch.full((2, 2), float(rank + 1), dtype=torch.complex128)
    parallel.allreduce_sum(s)
    assert torch.equal(s, torch.full((2, 2), 3.0, dtype=torch.complex128))
    open(os.path.join(out_dir, f'ok{rank}'), 'w').write('ok')
  finally:
    dist.destroy_process_group()


def test_gloo_world2_row_band_transposes(tmp_path):
  """The two all-to-alls of the row-sharded (Gamma-only) layout: rows -> bands -> rows."""
  port = 31500 + (os.getpid() % 2000)
  mp.spawn(_rank_transposes, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  assert (tmp_path / 'ok0').exists() and (tmp_path / 'ok1').exists()


def test_kinetic_operator_mirrors_oracle():
  from jrystal_b200 import kinetic
  cell, _, _ = structures.load('diamond')
  g = grid.g_vectors(cell, [7, 8, 9])
  k = grid.k_vectors(cell, [2, 1, 1])
  np.testing.assert_allclose(kinetic.kinetic_operator(g, k), rp.kinetic_operator(g, k).numpy(),
                             rtol=1e-14)
  np.testing.assert_allclose(kinetic.kinetic_operator(g), rp.kinetic_operator(g).numpy(),
                             rtol=1e-14)


def test_use_plan_context():
  from jrystal_b200 import context
  with pytest.raises(RuntimeError):
    context.current_plan()
  with context.use_plan('sentinel') as p:
    assert p == 'sentinel' and context.current_plan() == 'sentinel'
  with pytest.raises(RuntimeError):
    context.current_plan()


def test_convergence_checker_is_the_reference_rule():
  """std of the last `window` values below the threshold (jrystal/calc/convergence.py:13-35)."""
  from jrystal_b200.calc.convergence import ConvergenceChecker
  c = ConvergenceChecker(window_size=3, threshold=1e-2)
  assert [c.check(v) for v in (1.0, 0.5, 0.4, 0.399, 0.3985, 0.398)] == \
      [False, False, False, False, True, True]
  c.reset()
  assert c.history == []


def test_ewald_madelung_constant_and_eta_independence():
  """Simple-cubic lattice of unit charges in a neutralising background: E = xi / (2 L),
  xi = -2.837297479 (known answer); the converged sums do not depend on eta."""
  from jrystal_b200.ewald import ewald_coulomb_repulsion
  ref = -2.837297479 / 2
  for eta in (None, 0.6, 2.5):
    assert abs(ewald_coulomb_repulsion([[0, 0, 0]], [1.0], np.eye(3), eta) - ref) < 1e-9
  cell, pos, chg = structures.load('diamond')
  e1 = ewald_coulomb_repulsion(pos, chg, cell, 0.1)
  e2 = ewald_coulomb_repulsion(pos, chg, cell, 0.35)
  assert abs(e1 - e2) < 1e-10 * abs(e1)
  # translation invariance
  e3 = ewald_coulomb_repulsion(pos + 0.37, chg, cell)
  assert abs(e1 - e3) < 1e-10 * abs(e1)


def test_k_path_and_config_defaults():
  from jrystal_b200 import k_path
  from jrystal_b200.config import get_config
  cell, _, _ = structures.load('diamond')
  assert k_path.lattice_type(cell) == 'fcc'
  frac = k_path.get_k_path(cell, 'GXL', 7, fractional=True)
  np.testing.assert_allclose(frac[0], [0, 0, 0])
  assert any(np.allclose(f, [0.5, 0, 0.5]) for f in frac)       # X is on a sample
  np.testing.assert_allclose(frac[-1], [0.5, 0.5, 0.5])
  cart = k_path.get_k_path(cell, 'GXL', 7)
  np.testing.assert_allclose(cart, frac @ (2 * np.pi * np.linalg.inv(cell).T), atol=1e-14)
  with pytest.raises(ValueError):
    k_path.get_k_path(cell, 'GQ', 5)
  cfg = get_config(epoch=10)
  assert cfg.optimizer_args['b2'] == 0.99 and cfg.epoch == 10 and cfg.convergence_window_size == 20
  assert cfg.band_structure_empty_bands == 8


def test_k_path_tables_of_the_other_shipped_lattices():
  """BCC primitive (li), hexagonal (mg, graphene, zns_w), tetragonal (li_slab), orthorhombic cells:
  lattice detection and the special points at their textbook Cartesian places."""
  from jrystal_b200 import k_path
  from jrystal_b200.crystal import Crystal
  tp = 2 * np.pi
  a = 3.51
  bcc = a / 2 * np.array([[-1.0, 1, 1], [1, -1, 1], [1, 1, -1]])
  assert k_path.lattice_type(bcc) == 'bcc'
  pts = k_path.get_k_path(bcc, 'GHNGPH', 41)
  for want in ([0, 1, 0], [0.5, 0.5, 0], [0.5, 0.5, 0.5]):            # H, N, P in units of 2 pi / a
    assert any(np.allclose(p, np.array(want) * tp / a, atol=1e-12) for p in pts), want
  a, c = 3.20302773, 5.126691
  hexc = np.array([[a, 0, 0], [-a / 2, a * np.sqrt(3) / 2, 0], [0, 0, c]])
  assert k_path.lattice_type(hexc) == 'hex'
  pts = k_path.get_k_path(hexc, 'GMKGALHA', 61)
  norms = np.linalg.norm(pts, axis=1)
  assert any(abs(n - tp / a / np.sqrt(3)) < 1e-12 for n in norms)      # |M| = 2 pi / (sqrt(3) a)
  assert any(abs(n - 2 * tp / (3 * a)) < 1e-12 for n in norms)         # |K| = 4 pi / (3 a)
  assert any(np.allclose(p, [0, 0, tp / (2 * c)], atol=1e-12) for p in pts)   # A
  # K is a zone corner: equidistant from Gamma and the two nearest reciprocal lattice points
  b = tp * np.linalg.inv(hexc).T
  k = (b[0] + b[1]) / 3
  assert abs(np.linalg.norm(k) - np.linalg.norm(k - b[0])) < 1e-12
  assert abs(np.linalg.norm(k) - np.linalg.norm(k - b[1])) < 1e-12
  tet = np.diag([3.285, 3.285, 21.14])
  assert k_path.lattice_type(tet) == 'tet'
  f = k_path.get_k_path(tet, None, 50, fractional=True)                # default path GXMGZRAZ
  np.testing.assert_allclose(f[0], [0, 0, 0])
  np.testing.assert_allclose(f[-1], [0, 0, 0.5])
  assert any(np.allclose(x, [0.5, 0.5, 0.5]) for x in f)
  assert k_path.lattice_type(np.diag([3.0, 4.0, 5.0])) == 'orc'
  assert k_path.lattice_type(np.diag([4.2906] * 3)) == 'cubic'
  for bad in (np.array([[a, 0, 0], [a / 2, a * np.sqrt(3) / 2, 0], [0, 0, c]]),   # 60 degree setting
              np.array([[3.0, 0, 0], [0.4, 4.0, 0], [0, 0, 5.0]])):               # monoclinic
    with pytest.raises(NotImplementedError):
      k_path.lattice_type(bad)
  with pytest.raises(ValueError, match='crystal_file_path_path'):
    Crystal.create_builtin('graphene')


def test_temperature_schedule_matches_optax_exponential_decay():
  from jrystal_b200.calc.calc_ground_state_energy_all_electrons import temperature_scheduler
  from jrystal_b200.config import get_config
  s = temperature_scheduler(get_config(epoch=100, smearing=0.001))
  assert s(0) == 100.0
  assert abs(s(50) - 0.001) < 1e-12 and s(99) == 0.001      # init * rate^(i / (epoch // 2)), floored
  assert abs(s(25) - 100.0 * (0.001 / 100.0) ** 0.5) < 1e-12
  assert temperature_scheduler(get_config(smearing=0.0))(10) == 0.0


def test_trainable_occupations_match_the_literal_restatement():
  """jrystal_b200.occupation (sort/cumsum projection, autograd) against the step-by-step
  restatement of the reference's `proj` / simplex_projector / idempotent in the oracle, and the
  constraints they must satisfy (occupation.py:299-366, 55-80)."""
  from jrystal_b200 import occupation as occ
  g = torch.Generator().manual_seed(3)
  for n, m in [(12, 5), (40, 17), (7, 6), (9, 1), (16, 8)]:
    x = torch.rand(n, dtype=torch.float64, generator=g)
    a = occ._capped_simplex(x, float(m))
    b = rp.occupation_proj(x, float(m))
    assert float((a - b).abs().max()) < 1e-14
    assert abs(float(a.sum()) - m) < 1e-12 and float(a.min()) >= 0 and float(a.max()) <= 1
    # projecting a feasible point changes nothing
    assert float((occ._capped_simplex(a, float(m)) - a).abs().max()) < 1e-14
  # edges without an index in the sort/cumsum rule: an empty channel and a completely filled one
  # (simplex-projector with empty_bands = 0; num_electrons == spin)
  x = torch.rand(6, dtype=torch.float64, generator=g)
  assert float(occ._capped_simplex(x, 0.0).abs().max()) == 0.0
  assert float((occ._capped_simplex(x, 6.0) - 1.0).abs().max()) == 0.0
  full = occ.simplex_projector({k: v.detach().cpu() for k, v in occ.simplex_projector_init(4, 2).items()}, 8)
  assert float((full - 2.0 / 2).abs().max()) == 0.0          # every band full, 2 / nk each
  one = occ.simplex_projector({k: v.detach().cpu() for k, v in occ.simplex_projector_init(3, 1).items()},
                              1, spin=1, spin_restricted=False)
  assert float(one[1].abs().max()) == 0.0 and abs(float(one[0].sum()) - 1.0) < 1e-14
  nk, nb, ne = 3, 6, 8
  p = {k: v.detach().cpu().requires_grad_(True) for k, v in occ.simplex_projector_init(nb, nk).items()}
  o = occ.simplex_projector(p, ne)
  ref = rp.occupation_simplex_projector(p['param_up'].detach(), p['param_down'].detach(), ne)
  assert o.shape == (1, nk, nb) and float((o.detach() - ref).abs().max()) < 1e-14
  assert abs(float(o.sum()) - ne) < 1e-12
  # autograd through the projection against central differences
  w = torch.rand(1, nk, nb, dtype=torch.float64, generator=g)
  gu, = torch.autograd.grad((o * w).sum(), [p['param_up']])
  i, j, h = 1, 2, 1e-6
  vals = []
  for sgn in (+1, -1):
    q = {k: v.detach().clone() for k, v in p.items()}
    q['param_up'][i, j] += sgn * h
    vals.append(float((occ.simplex_projector(q, ne) * w).sum()))
  assert abs(gu[i, j].item() - (vals[0] - vals[1]) / (2 * h)) < 1e-6
  # idempotent: diag of a rank-(ne/2 nk) projector, trace fixed
  ip = occ.idempotent_param_init(5, nb, ne, nk)
  ip = {k: {'w_re': v['w_re'].detach().cpu()} for k, v in ip.items()}
  oi = occ.idempotent(ip, nk)
  refi = rp.occupation_idempotent(ip['param_up']['w_re'], ip['param_down']['w_re'], nk)
  assert float((oi - refi).abs().max()) < 1e-13 and abs(float(oi.sum()) - ne) < 1e-10
  assert float(oi.min()) >= 0 and float(oi.max()) <= 2.0 / nk + 1e-12


def test_orbital_grid_selection():
  """Host logic of orbital_grid='auto' (no GPU): alias-free minimum from the mask and the boxes
  tried, for the benchmark geometries."""
  from jrystal_b200 import grid
  from jrystal_b200.crystal import Crystal
  si8 = Crystal.create_builtin('si8')
  mask = grid.spherical_mask(si8.cell_vectors, [64, 64, 64], 30.0)
  need = grid.min_orbital_grid(mask)
  assert need == (49, 49, 49)
  assert grid.orbital_grid_candidates((64, 64, 64), need) == [(64, 64, 49)]
  # diamond-64 at 40 Ha on 128^3: 4 * 19 + 1 = 77 -> fused 81^3 first, z-only fallback
  assert grid.orbital_grid_candidates((128, 128, 128), (77, 77, 77)) == [(81, 81, 81), (128, 128, 81)]
  assert grid.orbital_grid_candidates((128, 128, 128), (85, 85, 85)) == [(128, 128, 100)]
  assert grid.orbital_grid_candidates((128, 128, 128), (105, 105, 105)) == [(128, 128, 128)]
  # small grids are left alone; an axis the caller's grid under-resolves stays as it is
  assert grid.orbital_grid_candidates((32, 32, 32), (21, 21, 21)) == [(32, 32, 32)]
  assert grid.orbital_grid_candidates((12, 12, 12), (13, 13, 13)) == [(12, 12, 12)]
  assert grid.orbital_grid_candidates((48, 48, 48), (49, 49, 49)) == [(48, 48, 48)]
  assert grid.orbital_grid_candidates((48, 48, 48), (33, 33, 33)) == [(36, 36, 36), (48, 48, 36)]
  assert grid.orbital_grid_candidates((48, 48, 48), (37, 33, 33)) == [(48, 48, 36)]
  assert grid.orbital_grid_candidates((48, 48, 64), (49, 49, 49)) == [(48, 48, 49)]
  # anisotropic mask: per-axis minimum
  m = np.zeros((16, 24, 32), dtype=bool)
  m[0, 0, 0] = m[2, 0, 0] = m[-1, 3, 0] = m[0, -5, 7] = True
  assert grid.min_orbital_grid(m) == (9, 21, 29)
