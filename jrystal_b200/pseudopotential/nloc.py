"""Projectors of the non-local (Kleinman-Bylander) part on the cut-off SPHERE, ready for
`Plan.set_nonlocal` (jrb_set_nonlocal), and host-side twins of the reference's contractions.

Reference: jrystal/pseudopotential/nloc.py:43-141 builds, for the whole box,
  Phi[k, (a, b1), m, G] = 4 pi i^{l_b1} e^{-i (G+k).R_a} sum_b2 S_a[b1, b2] Y_{l_b2, m}(G+k) beta_b2(|G+k|)
with S_a = eigvec * sqrt(eigval) of the atom's D matrix, Y the REAL harmonics of spherical.py and
beta the transforms of beta.py; energy: F = sum_G c_G Phi_G, E_nl = sum f |F|^2 / Omega
(nloc.py:143-158, 217-236).  The (kpt, beta, m, x, y, z) array is 68 GB at BASELINE config 4; the
coefficients vanish outside the sphere, so only Phi[..., mask] ever contributes: (kpt, proj, g),
1.5 GB there.  Rows that are identically zero (the m-padding of l < l_max) are dropped.

Reference behaviour kept on purpose: |F|^2 is weighted by |eigval| (the sign of a negative D_ii is
lost through conj(sqrt(eigval)) * sqrt(eigval)), and i^{l} is indexed by the OUTPUT projector b1;
both are exact for the shipped pseudopotentials (diagonal positive D)."""
from typing import List, Optional

import numpy as np

from .beta import beta_sbt_sphere, max_radius
from .spherical import batch_sph_harm_real, cartesian_to_spherical


def sphere_vectors(g_vector_grid, kpts, freq_mask) -> np.ndarray:
  """G + k on the sphere, (kpt, g, 3), g in the C-order enumeration of the mask (the compact index
  of the coefficients, utils.py:279-281)."""
  g = np.asarray(g_vector_grid, dtype=np.float64)[np.asarray(freq_mask, dtype=bool)]
  return g[None] + np.asarray(kpts, dtype=np.float64).reshape(-1, 1, 3)


def projector_rows(nonlocal_angular_momentum, nonlocal_d_matrix):
  """(keep, atom): for the (atom, beta, m) rows in the reference's order, which rows are not
  structurally zero (m < 2 l + 1 for a projector that feeds the row) and the atom each row
  belongs to.  Independent of the k-points, so the row count of a plan never changes."""
  nm = 2 * int(max(int(np.max(l)) for l in nonlocal_angular_momentum)) + 1
  keep, atom = [], []
  for a, (d, l) in enumerate(zip(nonlocal_d_matrix, nonlocal_angular_momentum)):
    mixes = np.abs(np.linalg.eigh(np.asarray(d, dtype=np.float64))[1]) > 0       # (b1, b2)
    has_m = np.arange(nm)[None, :] < 2 * np.asarray(l).astype(int)[:, None] + 1  # (b2, m)
    rows = (mixes.astype(int) @ has_m.astype(int)).reshape(-1) > 0
    keep.append(rows)
    atom.append(np.full(rows.shape, a))
  return np.concatenate(keep), np.concatenate(atom)


def potential_nonlocal_psi_sphere(position, g_vector_grid, kpts, freq_mask, r_grid,
                                  nonlocal_beta_grid, nonlocal_angular_momentum, nonlocal_d_matrix,
                                  drop_zero_rows: bool = True,
                                  kmax: Optional[float] = None) -> np.ndarray:
  """<beta_i | G + k> on the sphere: complex128 (kpt, proj, g), proj = (atom, beta, m) flattened in
  the reference's order (atoms concatenated along beta, m fastest).
  `kmax`: upper end of the radial-transform grid; default max |G + k| over the box and `kpts`, what
  the reference uses when it transforms these k-points in one call.  The band driver passes the
  maximum over its whole k-path (the reference transforms the path at once,
  calc_band_structure_normcons.py:118-123) while building the projectors one k-point at a time."""
  pos = np.asarray(position, dtype=np.float64).reshape(-1, 3)
  gk = sphere_vectors(g_vector_grid, kpts, freq_mask)                 # (k, g, 3)
  kmax = max_radius(g_vector_grid, kpts) if kmax is None else float(kmax)
  sph = cartesian_to_spherical(gk)
  radius = np.sqrt((gk * gk).sum(-1))
  l_max = int(max(int(np.max(l)) for l in nonlocal_angular_momentum))
  nm = 2 * l_max + 1
  # real harmonics for every l, m padded to 2 l_max + 1: (l, k, g, m)
  y_lm = np.zeros((l_max + 1,) + radius.shape + (nm,))
  for l in range(l_max + 1):
    y_lm[l, ..., :2 * l + 1] = batch_sph_harm_real(l, sph[..., 1], sph[..., 2])
  radial = {}
  blocks: List[np.ndarray] = []
  for a in range(pos.shape[0]):
    ls = np.asarray(nonlocal_angular_momentum[a]).astype(int)
    d = np.asarray(nonlocal_d_matrix[a], dtype=np.float64)
    key = (id(r_grid[a]), id(nonlocal_beta_grid[a]), tuple(ls))
    if key not in radial:  # atoms of one species share the radial transform
      radial[key] = beta_sbt_sphere(r_grid[a], nonlocal_beta_grid[a], ls, radius, kmax)  # (k, b, g)
    beta = radial[key]
    eigval, eigvec = np.linalg.eigh(d)
    s = eigvec * np.sqrt(eigval + 0j)                                 # (b1, b2)
    yb = y_lm[ls] * np.swapaxes(beta, 0, 1)[..., None]               # (b2, k, g, m)
    out = np.einsum('ab,bkgm->kamg', s, yb)                          # (k, b1, m, g)
    phase = np.exp(-1j * (gk @ pos[a]))                               # (k, g)
    out = out * phase[:, None, None, :] * ((1j) ** ls)[None, :, None, None] * (4 * np.pi)
    blocks.append(out.reshape(out.shape[0], -1, out.shape[-1]))
  phi = np.concatenate(blocks, axis=1)
  if drop_zero_rows:
    phi = phi[:, projector_rows(nonlocal_angular_momentum, nonlocal_d_matrix)[0]]
  return np.ascontiguousarray(phi)


def potential_nonlocal_psi_reciprocal(position, g_vector_grid, kpts, r_grid, nonlocal_beta_grid,
                                      nonlocal_angular_momentum, nonlocal_d_matrix, beta_gk=None,
                                      fourier_transform_method: str = 'sbt', concat: bool = True):
  """Dense drop-in for the reference's function of this name (nloc.py:43-141): the whole
  (kpt, beta, m, x, y, z) array, for callers that want it; the drivers never build it.
  `beta_gk` (a precomputed transform) is accepted and ignored: the transform is recomputed on the
  same radial grid.  concat=False returns the per-atom list."""
  del beta_gk
  if fourier_transform_method != 'sbt':
    raise ValueError("only the reference's default method 'sbt' is implemented")
  g = np.asarray(g_vector_grid)
  full = np.ones(g.shape[:-1], dtype=bool)
  nm = 2 * int(max(int(np.max(l)) for l in nonlocal_angular_momentum)) + 1
  rows = potential_nonlocal_psi_sphere(position, g, kpts, full, r_grid, nonlocal_beta_grid,
                                       nonlocal_angular_momentum, nonlocal_d_matrix,
                                       drop_zero_rows=False)
  dense = rows.reshape(rows.shape[0], -1, nm, *g.shape[:-1])
  if concat:
    return dense
  edges = np.cumsum([0] + [len(l) for l in nonlocal_angular_momentum])
  return [dense[:, a:b] for a, b in zip(edges[:-1], edges[1:])]


def hamiltonian_nonlocal(coeff_sphere, phi_sphere, vol: float) -> np.ndarray:
  """Host twin of nloc.py:143-158 on the sphere: coeff (spin, kpt, g, band) (the compact layout of
  the parameters), phi (kpt, proj, g) -> (spin, kpt, band, band)."""
  f = np.einsum('skgb,kpg->skbp', np.asarray(coeff_sphere), np.asarray(phi_sphere))
  return np.einsum('skap,skbp->skab', np.conj(f), f) / vol


def energy_nonlocal(coeff_sphere, phi_sphere, vol: float, occupation) -> float:
  """Host twin of nloc.py:217-236."""
  h = hamiltonian_nonlocal(coeff_sphere, phi_sphere, vol)
  return float(np.real(np.sum(np.diagonal(h, axis1=-2, axis2=-1) * np.asarray(occupation))))


def energy_nonlocal_position_gradient(coeff_sphere, phi_sphere, row_atom, gk_sphere, vol: float,
                                      occupation, num_atom: int) -> np.ndarray:
  """dE_nl / dR_a, (atom, 3): the position cotangent jax.grad gives the reference through the
  structure factor e^{-i (G+k).R_a} of nloc.py:122-129.  With F_p = sum_g c_g Phi_pg:
  dE_nl/dR_a = (2 / Omega) sum_skb f sum_{p in a} Re( conj(F_p) sum_g c_g Phi_pg (-i (G+k)_g) ).
  coeff_sphere (s, k, g, b); phi_sphere (k, proj, g); row_atom (proj,) from projector_rows (kept
  rows); gk_sphere (k, g, 3) from sphere_vectors."""
  c = np.asarray(coeff_sphere)
  phi = np.asarray(phi_sphere)
  f = np.einsum('skgb,kpg->skbp', c, phi)
  df = np.einsum('skgb,kpg,kgd->skbpd', c, phi, -1j * np.asarray(gk_sphere))
  per_row = 2.0 / vol * np.real(np.einsum('skb,skbp,skbpd->pd', np.asarray(occupation), np.conj(f), df))
  out = np.zeros((num_atom, 3))
  np.add.at(out, np.asarray(row_atom), per_row)
  return out


# -- the band-mode loss and subspace matrix of the norm-conserving drivers, on the current plan ----

def _plan_with_pseudopotential(coefficient, potential_local_grid_reciprocal,
                               potential_nl_psi_reciprocal):
  """Hand V_loc(G) and the projectors to the plan the coefficients live on (once per distinct pair
  of arrays; jrb_set_external_potential / jrb_set_nonlocal).  Projectors may come dense,
  (kpt, proj, x, y, z) as the reference builds them, or on the sphere, (kpt, proj, g)."""
  import torch
  from .. import pw as _pw
  c = _pw._as_coeff(coefficient)
  plan = c.plan
  host = lambda a: a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
  v_loc = np.ascontiguousarray(host(potential_local_grid_reciprocal), dtype=np.complex128)
  phi = host(potential_nl_psi_reciprocal)
  if phi.ndim == 5:
    phi = phi[:, :, plan.mask.astype(bool)]
  phi = np.ascontiguousarray(phi, dtype=np.complex128)
  key = (v_loc.tobytes(), phi.tobytes())
  if getattr(plan, '_pseudopotential_key', None) != key:
    plan.set_external_potential(torch.from_numpy(v_loc).to(c.q.device))
    plan.set_nonlocal(torch.from_numpy(phi).to(c.q.device) if phi.shape[1] else None)
    plan._pseudopotential_key = key
    plan._atom_key = None      # energy._plan_with_atoms must re-send its all-electron table
  return c, plan


def _h_apply(coefficient, hamiltonian_density_grid, potential_local_grid_reciprocal,
             potential_nl_psi_reciprocal, xc, kohn_sham):
  c, plan = _plan_with_pseudopotential(coefficient, potential_local_grid_reciprocal,
                                       potential_nl_psi_reciprocal)
  rho = hamiltonian_density_grid
  if rho.ndim == 3:
    rho = rho[None]
  v = plan.potential(rho.contiguous(), xc, kohn_sham, 7)    # v_H + V_loc + v_xc of the given density
  return c, plan, plan.hpsi(c.q, v)                          # + T + V_nl


def hamiltonian_matrix(coefficient, hamiltonian_density_grid, potential_local_grid_reciprocal,
                       potential_nl_psi_reciprocal, g_vector_grid, kpts, vol, xc: str = 'lda_x',
                       kohn_sham: bool = True):
  """jrystal/pseudopotential/nloc.py:161-214: <psi_i| T + v_H[rho] + v_xc[rho] + V_loc + V_nl |psi_j>,
  (spin, kpt, band, band), from ONE H-apply (jrb_potential, jrb_hpsi with the projectors on the
  sphere) and one FP64 tensor-core Gram (jrb_hamiltonian_matrix) instead of four dense einsums."""
  del g_vector_grid, kpts, vol
  c, plan, hq = _h_apply(coefficient, hamiltonian_density_grid, potential_local_grid_reciprocal,
                         potential_nl_psi_reciprocal, xc, kohn_sham)
  return plan.overlap(c.q, hq)


def hamiltonian_trace(coefficient, hamiltonian_density_grid, potential_local_grid_reciprocal,
                      potential_nl_psi_reciprocal, g_vector_grid, kpts, vol, xc: str = 'lda_x',
                      kohn_sham: bool = True):
  """jrystal/pseudopotential/nloc.py:230-279: the band-mode loss of the norm-conserving driver,
  sum over (spin, kpt, band) of the diagonal above with unit occupations (a device scalar)."""
  del g_vector_grid, kpts, vol
  c, plan, hq = _h_apply(coefficient, hamiltonian_density_grid, potential_local_grid_reciprocal,
                         potential_nl_psi_reciprocal, xc, kohn_sham)
  return plan.band_expect(c.q, hq).sum()
