#!/usr/bin/env python
"""Golden vectors produced by the REFERENCE'S OWN SOURCE (not by the oracle restatement).

jax is not installable in this container, so the reference cannot run as shipped; but its forward
path is plain array code, and `tests/golden/numpy_jax_standin.py` lets the files under
/root/reference/jrystal execute over numpy (see that file for what is stood in and why it is
faithful).  This script calls the reference functions VERBATIM on the seeded inputs of the
existing golden recipes (tests/golden/make_golden.py: same structure / grid / k-grid / mask /
seed) and stores their outputs in tests/golden/reference_<case>.npz:

  jrystal/_src/grid.py        g_vectors, r_vectors, k_vectors, spherical_mask / cubic_mask
  jrystal/_src/utils.py       volume
  jrystal/_src/pw.py          coeff (QR + expand), wave_grid, density_grid, density_grid_reciprocal
  jrystal/_src/energy.py      kinetic, hartree, external, nuclear_repulsion
  jrystal/_src/potential.py   hartree_reciprocal, external_reciprocal
  jrystal/_src/kinetic.py     kinetic_operator
  jrystal/_src/braket.py      expectation (kinetic mode diagonal; real mode diagonal and full)
  jrystal/_src/occupation.py  uniform, gamma, simplex_projector(_init), proj, idempotent
  jrystal/_src/entropy.py     fermi_dirac
  jrystal/_src/xc.py          sigma_r_fn (the functional-free part of the GGA branch)

and, for the norm-conserving rows (tests/golden/reference_si_normcons.npz), from the shipped
pseudopotential/normconserving/Si.pz-vbc.UPF:

  jrystal/pseudopotential/load.py       parse_upf
  jrystal/pseudopotential/dataclass.py  NormConservingPseudopotential.create
  jrystal/pseudopotential/beta.py       beta_sbt_grid
  jrystal/pseudopotential/local.py      potential_local_reciprocal, energy_local
  jrystal/pseudopotential/nloc.py       potential_nonlocal_psi_reciprocal, hamiltonian_nonlocal,
                                        energy_nonlocal
  jrystal/pseudopotential/spherical.py  batch_sph_harm_real, cartesian_to_spherical

NOT covered (needs jax_xc / automatic differentiation): every XC value and every gradient.

  python tests/golden/make_reference_golden.py      # needs /root/reference; deterministic

The fixtures travel to the GPU box; /root/reference does not (and is never read by a test).
"""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
warnings.filterwarnings('ignore', category=SyntaxWarning)

import numpy_jax_standin as standin  # noqa: E402
from make_golden import CASES  # noqa: E402  (recipes only; the oracle is not called here)
from oracle import structures  # noqa: E402  (geometry literals, each citing its .xyz file)

ref = standin.ref
A = lambda x: np.ascontiguousarray(np.asarray(x))
SAMPLE_STRIDE = (2, 3, 5)  # grids above 8192 points are stored as strided samples + sums


def G(x):
  """A field on the FFT grid: whole for small grids, a strided sample otherwise (the test takes
  the same sample; the *_sum entries cover the rest)."""
  x = A(x)
  if x.shape[-3] * x.shape[-2] * x.shape[-1] <= 8192:
    return x
  return A(x[..., ::SAMPLE_STRIDE[0], ::SAMPLE_STRIDE[1], ::SAMPLE_STRIDE[2]])


def seeded_params(seed, nb, nk, ng):
  """The parameter recipe of the golden cases (oracle.reference_port.param_init: numpy
  default_rng(seed), U[0,1), shape (1, nk, ng, nb)); restated so this script does not call the
  oracle."""
  rng = np.random.default_rng(seed)
  shape = (1, nk, ng, nb)
  return rng.random(shape), rng.random(shape)


def core_case(key, c):
  grid_m, utils, pw, energy = ref('_src.grid'), ref('_src.utils'), ref('_src.pw'), ref('_src.energy')
  potential, kinetic, braket = ref('_src.potential'), ref('_src.kinetic'), ref('_src.braket')
  occupation, entropy = ref('_src.occupation'), ref('_src.entropy')
  xc_m = ref('_src.xc')  # imports with jax_xc stubbed; only the functional-free sigma_r_fn is called

  cell, pos, chg = structures.load(c['name'], None)
  gs = [int(g) for g in grid_m.proper_grid_size(c['grid'])]
  ks = [int(g) for g in grid_m.proper_grid_size(c['kgrid'])]
  vol = float(utils.volume(cell))
  g_vec = grid_m.g_vectors(cell, gs)
  r_vec = grid_m.r_vectors(cell, gs)
  kpts = grid_m.k_vectors(cell, ks)
  if c['mask'] == 'spherical':
    mask = grid_m.spherical_mask(cell, gs, c['cutoff'])
  else:
    mask = grid_m.cubic_mask(gs)
  mask = np.asarray(mask)
  nk, ng, nb = kpts.shape[0], int(mask.sum()), c['nb']
  w_re, w_im = seeded_params(c['seed'], nb, nk, ng)
  ne = int(round(float(np.sum(chg))))
  occ = A(occupation.uniform(nk, ne, num_bands=nb))
  occ = occ * (1.0 + 0.1 * np.random.default_rng(c['seed'] + 1).random(occ.shape))

  coeff = pw.coeff({'w_re': w_re, 'w_im': w_im}, mask)
  psi = pw.wave_grid(coeff, vol)
  rho = pw.density_grid(coeff, vol, occ)
  rho_g = pw.density_grid_reciprocal(coeff, vol, occ)
  e_kin = energy.kinetic(g_vec, kpts, coeff, occ)
  e_har = energy.hartree(rho_g, g_vec, vol)
  e_har_ks = energy.hartree(rho_g, g_vec, vol, kohn_sham=True)
  e_ext = energy.external(rho_g, pos, chg, g_vec, vol)
  e_nuc = energy.nuclear_repulsion(pos, chg, cell, g_vec, vol, 0.1, 2e4)
  v_har = potential.hartree_reciprocal(rho_g, g_vec)
  v_ext = potential.external_reciprocal(pos, chg, g_vec, vol)
  t_k = kinetic.kinetic_operator(g_vec, kpts)
  kin_band = braket.expectation(coeff, t_k, vol, diagonal=True, mode='kinetic')
  ent_in = np.random.default_rng(c['seed'] + 2).random(occ.shape) * 2.0 / nk  # occupations in (0, 2/nk)
  # <psi_i| v |psi_j> in real space for a seeded real potential (braket.py:167-207), the
  # contraction band mode is built on; v is regenerated by the test from the same seed
  v_r = np.random.default_rng(c['seed'] + 3).standard_normal(tuple(gs))
  v_diag = braket.expectation(psi, v_r, vol, diagonal=True, mode='real')
  v_full = braket.expectation(psi, v_r, vol, diagonal=False, mode='real')
  out = dict(
    expect_v_diag=A(v_diag), expect_v_full=A(v_full),
    sigma_r=G(xc_m.sigma_r_fn(rho, g_vec)),  # |grad rho|^2 of the GGA branch (xc.py:94-112)
    vol=np.array(vol), grid=np.array(gs), kpts=A(kpts), mask=mask,
    g_vec_sum=np.array(np.abs(A(g_vec)).sum()), g_vec_corner=A(g_vec)[1, 2, 3],
    r_vec_corner=A(r_vec)[1, 2, 3], occ=occ,
    w_re_sum=np.array(w_re.sum()), w_im_sum=np.array(w_im.sum()),
    # Q on the sphere, (ns, nk, ng, nb): the reference's compact layout
    q=A(np.swapaxes(A(coeff)[..., mask], -1, -2)),
    psi_band0=G(A(psi)[0, 0, 0]), density=G(rho), density_reciprocal=G(rho_g),
    density_sum=np.array(A(rho).sum()), density_abs2_sum=np.array((A(rho) ** 2).sum()),
    e_kin=np.array(float(e_kin)), e_har=np.array(float(e_har)),
    e_har_kohn_sham=np.array(float(e_har_ks)), e_ext=np.array(float(e_ext)),
    e_nuc=np.array(float(e_nuc)), v_har_reciprocal=G(v_har), v_ext_reciprocal=G(v_ext),
    v_har_abs_sum=np.array(np.abs(A(v_har)).sum()), v_ext_abs_sum=np.array(np.abs(A(v_ext)).sum()),
    kinetic_per_band=A(kin_band),
    occ_uniform=A(occupation.uniform(nk, ne, num_bands=nb)),
    occ_gamma=A(occupation.gamma(nk, ne, num_bands=nb)),
    entropy_input=ent_in, entropy_fermi_dirac=np.array(float(entropy.fermi_dirac(ent_in))),
  )
  path = os.path.join(HERE, f'reference_{key}.npz')
  np.savez_compressed(path, **out)
  print(f'{key}: E_kin {float(e_kin):.12f} E_H {float(e_har):.12f} E_ext {float(e_ext):.12f} '
        f'E_nuc {float(e_nuc):.12f} -> {os.path.basename(path)}')


NORMCONS = dict(name='si', grid=[12, 12, 12], kgrid=[1, 1, 2], cutoff=6.0, nb=6, seed=21)


def normcons_case():
  """Si2 with the shipped Si.pz-vbc.UPF: the reference's pseudopotential set-up and its
  local / non-local energy terms, executed verbatim."""
  c = NORMCONS
  grid_m, utils, pw = ref('_src.grid'), ref('_src.utils'), ref('_src.pw')
  occupation = ref('_src.occupation')
  load, dc = ref('pseudopotential.load'), ref('pseudopotential.dataclass')
  beta, local, nloc = ref('pseudopotential.beta'), ref('pseudopotential.local'), ref('pseudopotential.nloc')
  sph = ref('pseudopotential.spherical')

  cell, pos, _ = structures.load(c['name'], None)
  upf_dir = os.path.join(standin.REFERENCE_ROOT, 'pseudopotential', 'normconserving') + '/'
  upf = load.parse_upf(load.find_upf(upf_dir, 'Si'))
  crystal = types.SimpleNamespace(positions=pos, charges=np.array([14, 14]), symbols=['Si', 'Si'])
  pp = dc.NormConservingPseudopotential.create(crystal, upf_dir)

  gs = [int(g) for g in grid_m.proper_grid_size(c['grid'])]
  vol = float(utils.volume(cell))
  g_vec = grid_m.g_vectors(cell, gs)
  kpts = grid_m.k_vectors(cell, c['kgrid'])
  mask = np.asarray(grid_m.spherical_mask(cell, gs, c['cutoff']))
  nk, ng, nb = kpts.shape[0], int(mask.sum()), c['nb']
  w_re, w_im = seeded_params(c['seed'], nb, nk, ng)
  ne = int(sum(pp.valence_charges))
  occ = A(occupation.uniform(nk, ne, num_bands=nb))
  coeff = pw.coeff({'w_re': w_re, 'w_im': w_im}, mask)
  rho_g = pw.density_grid_reciprocal(coeff, vol, occ)

  v_loc = local.potential_local_reciprocal(
    pos, g_vec, pp.r_grid, pp.local_potential_grid, pp.local_potential_charge, vol)
  e_loc = local.energy_local(rho_g, v_loc, vol)
  beta_gk = beta.beta_sbt_grid(pp.r_grid, pp.nonlocal_beta_grid, pp.nonlocal_angular_momentum,
                               g_vec, kpts)
  phi = nloc.potential_nonlocal_psi_reciprocal(
    pos, g_vec, kpts, pp.r_grid, pp.nonlocal_beta_grid, pp.nonlocal_angular_momentum,
    pp.nonlocal_d_matrix, beta_gk)          # (kpt, beta, m, x, y, z)
  h_nl = nloc.hamiltonian_nonlocal(coeff, phi, vol)
  e_nl = nloc.energy_nonlocal(coeff, phi, vol, occ)

  # spherical harmonics on a fixed set of directions
  dirs = np.random.default_rng(3).normal(size=(7, 3))
  sp = A(sph.cartesian_to_spherical(dirs))
  ylm = {f'ylm_real_l{l}': A(sph.batch_sph_harm_real(l, sp[:, 1], sp[:, 2])) for l in range(3)}

  # sbt_numerical with the UPF's own quadrature weights (the call of jrystal/sbt/sbt_test.py:57-63)
  sbt_num = ref('sbt.sbt_numerical')
  kk, bk = sbt_num.sbt(pp.r_grid[0], pp.nonlocal_beta_grid[0], l=list(pp.nonlocal_angular_momentum[0]),
                       kmax=100, delta_r=pp.r_ab[0])
  phi = A(phi)
  out = dict(
    pp_r_ab=A(pp.r_ab[0]), sbt_rab_k=A(kk)[::37], sbt_rab_beta=A(bk)[:, ::37],
    vol=np.array(vol), grid=np.array(gs), kpts=A(kpts), mask=mask, occ=occ,
    w_re_sum=np.array(w_re.sum()), w_im_sum=np.array(w_im.sum()),
    # parse_upf / NormConservingPseudopotential.create
    upf_header_keys=np.array(sorted(upf['PP_HEADER'].keys())),
    upf_z_valence=np.array(float(upf['PP_HEADER']['z_valence'])),
    upf_mesh_size=np.array(int(upf['PP_HEADER']['mesh_size'])),
    upf_r=np.array(upf['PP_MESH']['PP_R']), upf_rab=np.array(upf['PP_MESH']['PP_RAB']),
    upf_local=np.array(upf['PP_LOCAL']),
    upf_beta=np.array([b['values'] for b in upf['PP_NONLOCAL']['PP_BETA']]),
    upf_beta_l=np.array([int(b['angular_momentum']) for b in upf['PP_NONLOCAL']['PP_BETA']]),
    upf_dij=np.array(upf['PP_NONLOCAL']['PP_DIJ']),
    pp_valence_charges=np.array(pp.valence_charges), pp_r_grid=A(pp.r_grid[0]),
    pp_local_potential_grid=A(pp.local_potential_grid[0]),
    pp_nonlocal_beta_grid=A(pp.nonlocal_beta_grid[0]),
    pp_nonlocal_d_matrix=A(pp.nonlocal_d_matrix[0]),
    pp_nonlocal_angular_momentum=A(pp.nonlocal_angular_momentum[0]),
    # potentials and energies
    v_loc_reciprocal=A(v_loc), e_loc=np.array(float(np.real(e_loc))),
    beta_gk_atom0=A(beta_gk[0]),
    phi_shape=np.array(phi.shape), phi_on_sphere=A(phi[..., mask]),
    phi_abs_sum_off_sphere=np.array(np.abs(phi[..., ~mask]).sum()),
    phi_abs_sum=np.array(np.abs(phi).sum()),
    h_nonlocal=A(h_nl), e_nonlocal=np.array(float(np.real(e_nl))),
    directions=dirs, directions_spherical=sp, **ylm,
  )
  path = os.path.join(HERE, 'reference_si_normcons.npz')
  np.savez_compressed(path, **out)
  print(f'si_normcons: E_loc {float(np.real(e_loc)):.12f} E_nl {float(np.real(e_nl)):.12f} '
        f'phi {phi.shape} -> {os.path.basename(path)}')


def occupation_case():
  """occupation.py: simplex_projector (+ its init), idempotent, on seeded parameters."""
  occupation = ref('_src.occupation')
  nk, nb, ne = 3, 7, 8
  rng = np.random.default_rng(17)
  logits_up, logits_dn = rng.normal(size=(nk, nb)) * 2, rng.normal(size=(nk, nb)) * 2
  params = {'param_up': logits_up, 'param_down': logits_dn}
  w_up = rng.random((nb * nk, (ne + 2) // 2 * nk))
  w_dn = rng.random((nb * nk, (ne - 2) // 2 * nk))
  idem = {'param_up': {'w_re': w_up}, 'param_down': {'w_re': w_dn}}
  init = occupation.simplex_projector_init(nb, nk)
  proj_x = np.random.default_rng(8).random(11)
  out = dict(
    nk=np.array(nk), nb=np.array(nb), ne=np.array(ne), logits_up=logits_up, logits_down=logits_dn,
    w_up=w_up, w_down=w_dn,
    simplex_restricted=A(occupation.simplex_projector(params, ne)),
    simplex_spin2=A(occupation.simplex_projector(params, ne, spin=2, spin_restricted=False)),
    simplex_init_up=A(init['param_up']), simplex_init_down=A(init['param_down']),
    simplex_from_init=A(occupation.simplex_projector(init, ne)),
    proj_input=proj_x, proj_down=A(occupation.proj(proj_x, 3.0)),   # sum(x) > 3: push down
    proj_up=A(occupation.proj(proj_x * 0.3, 6.0)),                   # sum(0.3 x) < 6: push up
    idempotent_unrestricted=A(occupation.idempotent(idem, nk, spin_restricted=False)),
    idempotent_restricted=A(occupation.idempotent({'param_up': {'w_re': w_up},
                                                   'param_down': {'w_re': w_up}}, nk)),
  )
  np.savez_compressed(os.path.join(HERE, 'reference_occupation.npz'), **out)
  print('occupation: sums', out['simplex_restricted'].sum(), out['idempotent_unrestricted'].sum(),
        '-> reference_occupation.npz')


# ---- LDA formulas handed to the reference in place of jax_xc (OUR restatement of LibXC's lda_x
# and lda_c_pw, the same closed forms as oracle/reference_port.py:_eps_lda_x/_eps_lda_c_pw, written
# for numpy and analytic in rho so that the complex-step derivative of the stand-in applies).
_LDA_X = -0.75 * (3.0 / np.pi) ** (1.0 / 3.0)


def _stub_lda_x(p, rho):
  return np.where(np.real(rho) > 1e-15, _LDA_X * np.asarray(rho) ** (1.0 / 3.0), 0.0)


def _stub_lda_c_pw(p, rho):
  a, a1, b1, b2, b3, b4 = 0.031091, 0.21370, 7.5957, 3.5876, 1.6382, 0.49294
  rho = np.asarray(rho)
  safe = np.where(np.real(rho) > 1e-15, rho, 1e-15)
  rs = (3.0 / (4.0 * np.pi * safe)) ** (1.0 / 3.0)
  den = 2 * a * (b1 * rs ** 0.5 + b2 * rs + b3 * rs ** 1.5 + b4 * rs ** 2)
  return np.where(np.real(rho) > 1e-15, -2 * a * (1 + a1 * rs) * np.log(1 + 1 / den), 0.0)


def _stub_gga_x_pbe(p, rho, sigma):
  kappa, mu = 0.8040, 0.2195149727645171
  safe = np.where(rho > 1e-15, rho, 1e-15)
  s2 = sigma / (4.0 * (3.0 * np.pi ** 2) ** (2.0 / 3.0) * safe ** (8.0 / 3.0))
  fx = 1.0 + kappa - kappa ** 2 / (kappa + mu * s2)
  return np.where(rho > 1e-15, _LDA_X * safe ** (1.0 / 3.0) * fx, 0.0)


def _stub_gga_c_pbe(p, rho, sigma):
  a, a1, b1, b2, b3, b4 = 0.0310907, 0.21370, 7.5957, 3.5876, 1.6382, 0.49294
  beta, gamma = 0.06672455060314922, (1.0 - np.log(2.0)) / np.pi ** 2
  safe = np.where(rho > 1e-12, rho, 1e-12)
  rs = (3.0 / (4.0 * np.pi * safe)) ** (1.0 / 3.0)
  den = 2 * a * (b1 * rs ** 0.5 + b2 * rs + b3 * rs ** 1.5 + b4 * rs ** 2)
  ec = -2 * a * (1 + a1 * rs) * np.log1p(1 / den)
  kf = (3.0 * np.pi ** 2 * safe) ** (1.0 / 3.0)
  t2 = sigma * np.pi / (16.0 * kf * safe ** 2)
  aa = (beta / gamma) / np.expm1(-ec / gamma)
  f1 = t2 + aa * t2 ** 2
  return np.where(rho > 1e-12, ec + gamma * np.log1p((beta / gamma) * f1 / (1 + aa * f1)), 0.0)


def gga_assembly_case():
  """GGA energy branch (kohn_sham=False): xc.sigma_r_fn + the per-point call of the functional +
  energy.xc_energy / total_energy, verbatim, with our PBE closed forms standing in for jax_xc."""
  standin.provide_functional('gga_x_pbe', _stub_gga_x_pbe)
  standin.provide_functional('gga_c_pbe', _stub_gga_c_pbe)
  key = 'diamond_12_pbe'
  c = CASES[key]
  grid_m, utils, pw, energy = ref('_src.grid'), ref('_src.utils'), ref('_src.pw'), ref('_src.energy')
  xc_m, occupation = ref('_src.xc'), ref('_src.occupation')
  cell, pos, chg = structures.load(c['name'], None)
  gs = [int(g) for g in grid_m.proper_grid_size(c['grid'])]
  vol = float(utils.volume(cell))
  g_vec = grid_m.g_vectors(cell, gs)
  kpts = grid_m.k_vectors(cell, [int(g) for g in grid_m.proper_grid_size(c['kgrid'])])
  mask = np.asarray(grid_m.spherical_mask(cell, gs, c['cutoff']))
  nk, ng, nb = kpts.shape[0], int(mask.sum()), c['nb']
  w_re, w_im = seeded_params(c['seed'], nb, nk, ng)
  occ = A(occupation.uniform(nk, int(round(float(np.sum(chg)))), num_bands=nb))
  occ = occ * (1.0 + 0.1 * np.random.default_rng(c['seed'] + 1).random(occ.shape))
  coeff = pw.coeff({'w_re': w_re, 'w_im': w_im}, mask)
  rho = pw.density_grid(coeff, vol, occ)
  out = dict(w_re_sum=np.array(w_re.sum()), occ=occ)
  for xc in ('gga_x_pbe', 'gga_x_pbe+gga_c_pbe'):
    tag = xc.replace('+', '_')
    out[f'{tag}_eps'] = G(xc_m.xc_density(rho, g_vec, False, xc))
    out[f'{tag}_e_xc'] = np.array(float(energy.xc_energy(rho, g_vec, vol, xc, False)))
    out[f'{tag}_total_energy'] = np.array(
      [float(np.real(e)) for e in energy.total_energy(coeff, pos, chg, g_vec, kpts, vol, occ,
                                                       kohn_sham=False, xc=xc, split=True)])
  np.savez_compressed(os.path.join(HERE, f'reference_xc_assembly_{key}.npz'), **out)
  print(f'gga assembly {key}: E_xc(pbe) {float(out["gga_x_pbe_gga_c_pbe_e_xc"]):.12f}')


def xc_assembly_case(key, c):
  """Everything the reference builds AROUND the functional, executed verbatim with the two LDA
  formulas above registered as jax_xc.impl.lda_x / lda_c_pw: xc.xc_density (both kohn_sham
  flags; vxc_lda = eps + rho d eps / d rho with jax.grad per grid point), energy.xc_energy,
  potential.effective (split and summed), energy.total_energy,
  hamiltonian.hamiltonian_matrix_trace.  The functional FORMULA stays unpinned (jax_xc absent);
  these vectors pin its calling convention, the kohn_sham semantics, the normalisations and the
  assembly of the band-mode loss."""
  standin.provide_functional('lda_x', _stub_lda_x)
  standin.provide_functional('lda_c_pw', _stub_lda_c_pw)
  grid_m, utils, pw, energy = ref('_src.grid'), ref('_src.utils'), ref('_src.pw'), ref('_src.energy')
  potential, hamiltonian, xc_m = ref('_src.potential'), ref('_src.hamiltonian'), ref('_src.xc')
  occupation = ref('_src.occupation')
  cell, pos, chg = structures.load(c['name'], None)
  gs = [int(g) for g in grid_m.proper_grid_size(c['grid'])]
  vol = float(utils.volume(cell))
  g_vec = grid_m.g_vectors(cell, gs)
  kpts = grid_m.k_vectors(cell, [int(g) for g in grid_m.proper_grid_size(c['kgrid'])])
  mask = np.asarray(grid_m.spherical_mask(cell, gs, c['cutoff']) if c['mask'] == 'spherical'
                    else grid_m.cubic_mask(gs))
  nk, ng, nb = kpts.shape[0], int(mask.sum()), c['nb']
  w_re, w_im = seeded_params(c['seed'], nb, nk, ng)
  ne = int(round(float(np.sum(chg))))
  occ = A(occupation.uniform(nk, ne, num_bands=nb))
  occ = occ * (1.0 + 0.1 * np.random.default_rng(c['seed'] + 1).random(occ.shape))
  coeff = pw.coeff({'w_re': w_re, 'w_im': w_im}, mask)
  rho = pw.density_grid(coeff, vol, occ)
  out = dict(w_re_sum=np.array(w_re.sum()), occ=occ)
  for xc in ('lda_x', 'lda_x+lda_c_pw'):
    tag = xc.replace('+', '_')
    out[f'{tag}_eps'] = G(xc_m.xc_density(rho, g_vec, False, xc))
    out[f'{tag}_vxc'] = G(xc_m.xc_density(rho, g_vec, True, xc))
    out[f'{tag}_e_xc'] = np.array(float(energy.xc_energy(rho, g_vec, vol, xc, False)))
    out[f'{tag}_e_xc_kohn_sham'] = np.array(float(energy.xc_energy(rho, g_vec, vol, xc, True)))
    for ks in (False, True):
      v_h, v_e, v_x = potential.effective(rho, pos, chg, g_vec, vol, split=True, xc_type=xc,
                                          kohn_sham=ks)
      v_sum = potential.effective(rho, pos, chg, g_vec, vol, split=False, xc_type=xc, kohn_sham=ks)
      # real parts (what every consumer keeps, hamiltonian.py:164); the imaginary part is the
      # Nyquist residue of an even grid
      out[f'{tag}_veff_ks{int(ks)}'] = G(np.real(A(v_sum)))
      out[f'{tag}_veff_imag_max_ks{int(ks)}'] = np.array(np.abs(np.imag(A(v_sum))).max())
      out[f'{tag}_veff_parts_ks{int(ks)}'] = np.stack(
        [G(np.real(np.broadcast_to(A(v), A(v_sum).shape))) for v in (v_h, v_e, v_x)])
      out[f'{tag}_total_energy_ks{int(ks)}'] = np.array(
        [float(np.real(e)) for e in energy.total_energy(coeff, pos, chg, g_vec, kpts, vol, occ,
                                                         kohn_sham=ks, xc=xc, split=True)])
      # energy.band_energy is not callable as shipped (energy.py:361: real_braket rejects the
      # (s, k, b, x, y, z) density against the (s, x, y, z) potential), so it has no golden
    out[f'{tag}_hamiltonian_trace'] = np.array(float(np.real(hamiltonian.hamiltonian_matrix_trace(
      coeff, pos, chg, rho, g_vec, kpts, vol, xc=xc, kohn_sham=True))))
    out[f'{tag}_hamiltonian_trace_per_k'] = A(hamiltonian.hamiltonian_matrix_trace(
      coeff, pos, chg, rho, g_vec, kpts, vol, xc=xc, kohn_sham=True, keep_kpts_axis=True))
  path = os.path.join(HERE, f'reference_xc_assembly_{key}.npz')
  np.savez_compressed(path, **out)
  print(f'xc assembly {key}: E_xc(lda_x) {float(out["lda_x_e_xc"]):.12f} trace '
        f'{float(out["lda_x_hamiltonian_trace"]):.12f} -> {os.path.basename(path)}')


API_MODULES = {  # jrystal_b200 module -> reference module it mirrors
  'pw': '_src.pw', 'grid': '_src.grid', 'energy': '_src.energy', 'potential': '_src.potential',
  'kinetic': '_src.kinetic', 'hamiltonian': '_src.hamiltonian', 'occupation': '_src.occupation',
  'ewald': '_src.ewald', 'pseudopotential.load': 'pseudopotential.load',
  'pseudopotential.local': 'pseudopotential.local', 'pseudopotential.beta': 'pseudopotential.beta',
  'pseudopotential.nloc': 'pseudopotential.nloc',
  'pseudopotential.spherical': 'pseudopotential.spherical',
  'utils': '_src.utils', 'entropy': '_src.entropy',
}


def api_signatures():
  """Parameter names, order and defaults of every public function of the reference modules the
  host package mirrors (inspect.signature over the modules as imported here) ->
  tests/golden/reference_api_signatures.json; tests/test_reference_golden.py holds
  jrystal_b200's functions of the same name to them (the drop-in claim of the Python API)."""
  import inspect
  import json
  out = {}
  for ours, theirs in API_MODULES.items():
    mod = ref(theirs)
    out[ours] = {
      name: [[p.name, None if p.default is inspect.Parameter.empty else repr(p.default)]
             for p in inspect.signature(fn).parameters.values()]
      for name, fn in inspect.getmembers(mod, inspect.isfunction)
      if not name.startswith('_') and fn.__module__ == mod.__name__}
  with open(os.path.join(HERE, 'reference_api_signatures.json'), 'w') as f:
    json.dump(out, f, indent=1, sort_keys=True)
  print('api signatures:', sum(len(v) for v in out.values()), 'functions ->',
        'reference_api_signatures.json')


def grid_helpers_case():
  """grid.py set-up helpers and the Ewald sum on given grids, on the Si2 cell."""
  grid_m, ewald = ref('_src.grid'), ref('_src.ewald')
  cell, pos, chg = structures.load('si', None)
  gs = [8, 9, 10]
  g, r = grid_m.g_vectors(cell, gs), grid_m.r_vectors(cell, gs)
  tv = grid_m.translation_vectors(cell, 3e3)
  vol = float(abs(np.linalg.det(cell)))
  out = dict(
    grid=np.array(gs), translation_cutoff=np.array(3e3), translation_vectors=A(tv),
    g2cell=A(grid_m.g2cell_vectors(g)), r2cell=A(grid_m.r2cell_vectors(r)),
    g2r_corner=A(grid_m.g2r_vector_grid(g))[1, 2, 3], r2g_corner=A(grid_m.r2g_vector_grid(r))[1, 2, 3],
    radius_corner=A(grid_m.grid_vector_radius(g))[1, 2, 3],
    half_frequency_shapes=np.array([grid_m.half_frequency_shape(n) for n in
                                    ([7, 8, 9], [12, 12, 12], [16, 5, 6], [1, 2, 3])]),
    ewald_eta=np.array(0.25),
    ewald=np.array(float(ewald.ewald_coulomb_repulsion(pos, chg, g, vol, 0.25, tv))),
  )
  np.savez_compressed(os.path.join(HERE, 'reference_grid_helpers.npz'), **out)
  print('grid helpers: ewald', float(out['ewald']), '-> reference_grid_helpers.npz')


def main():
  only = sys.argv[1:]
  for key, c in CASES.items():
    if only and key not in only:
      continue
    core_case(key, c)
  if not only or 'si_normcons' in only:
    normcons_case()
  for key in ('diamond_789_cubic', 'diamond_16_sph'):
    if not only or 'xc_assembly' in only or f'xc_assembly_{key}' in only:
      xc_assembly_case(key, CASES[key])
  if not only or 'gga_assembly' in only:
    gga_assembly_case()
  if not only or 'api' in only:
    api_signatures()
  if not only or 'grid_helpers' in only:
    grid_helpers_case()
  if not only or 'occupation' in only:
    with np.errstate(divide='ignore', invalid='ignore'):  # proj() divides by n - arange(n) - 1
      occupation_case()


if __name__ == '__main__':
  main()
