"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's norm-conserving pseudopotential
set-up and energy terms on the DENSE (kpt, beta, m, x, y, z) box, as the reference computes them.

Checker for jrystal_b200/pseudopotential/ (which works on the cut-off sphere) and for the CUDA
path with a local + non-local pseudopotential attached.  PIN STATUS: pinned to the reference's
own source: tests/golden/reference_si_normcons.npz holds the outputs of
/root/reference/jrystal/pseudopotential/{load,dataclass,beta,local,nloc,spherical}.py executed
verbatim (tests/golden/make_reference_golden.py) on the shipped Si.pz-vbc.UPF, and
tests/test_pseudopotential.py holds this module to them at 1e-12.  Gradients: torch autograd in
place of jax.value_and_grad."""
import numpy as np
import torch
from scipy.interpolate import CubicSpline
from scipy.special import sph_harm_y, spherical_jn

from . import reference_port as rp


def sbt_numerical(r, f, l, kmax):
  """jrystal/sbt/sbt_numerical.py:40-77: k = linspace(1e-4, kmax, 2 nr);
  F_l(k) = sum_r f(r) r^2 j_l(k r) dr, dr[:-1] = diff(r), dr[-1] = 0."""
  r = np.asarray(r)
  f = np.atleast_2d(np.asarray(f))
  dr = np.zeros_like(r)
  dr[:-1] = r[1:] - r[:-1]
  k = np.linspace(1e-4, kmax, 2 * len(r))
  kr = np.einsum('g,r->gr', k, r)
  ls = [l] * f.shape[0] if np.ndim(l) == 0 else list(l)
  jn = np.stack([spherical_jn(int(li), kr) for li in ls])
  return k, np.einsum('lr,r,lgr,r->lg', f, r ** 2, jn, dr)


def beta_sbt_grid(r_grid, beta_grid, angular_momentum, g_vec, kpts):
  """jrystal/pseudopotential/beta.py:27-120: list over atoms of (kpt, beta, x, y, z)."""
  gk = np.expand_dims(np.asarray(kpts), (1, 2, 3)) + np.asarray(g_vec)[None]
  radius = np.sqrt(np.sum(gk ** 2, axis=-1))
  out = []
  for r, b, l in zip(r_grid, beta_grid, angular_momentum):
    k, bk = sbt_numerical(r, b, list(l), np.max(radius))
    out.append(np.swapaxes(CubicSpline(k, bk, axis=1)(radius), 0, 1))
  return out


def potential_local_reciprocal(positions, g_vec, r_grid, v_grid, z_list, vol):
  """jrystal/pseudopotential/local.py:32-136, method 'sbt'."""
  g_vec = np.asarray(g_vec)
  g_radius = np.sqrt(np.sum(g_vec ** 2, axis=-1))
  n = g_radius.size
  v1 = []
  for r, v, z in zip(r_grid, v_grid, z_list):
    kk, fk = sbt_numerical(r, (v + z / r)[None], 0, np.max(g_radius))
    v1.append(4 * np.pi * CubicSpline(kk, fk[0])(g_radius))
  v1 = np.stack(v1)
  safe = g_radius.copy()
  safe[0, 0, 0] = 1e10
  v2 = 4 * np.pi * np.asarray(z_list, dtype=np.float64)[:, None, None, None] / safe[None] ** 2
  vg = v1 - v2
  vg[:, 0, 0, 0] = 0
  sf = np.exp(-1j * (g_vec @ np.asarray(positions).T))          # (x, y, z, atom)
  vg = np.sum(vg * np.transpose(sf, (3, 0, 1, 2)), axis=0)
  return vg * n / vol


def real_sph_harm(l, theta, phi):
  """jrystal/pseudopotential/spherical.py:66-89 (theta azimuth, phi polar), m = -l..l last."""
  m = np.arange(-l, l + 1)
  y = sph_harm_y(l, m, phi[..., None], theta[..., None])
  sgn = (-1.0) ** np.abs(m)
  y2 = np.conj(y) * sgn
  out = np.where(m >= 0, y.real * np.sqrt(2) * sgn, y2.imag * np.sqrt(2) * sgn)
  out[..., l] = y[..., l].real
  return out


def potential_nonlocal_psi_reciprocal(positions, g_vec, kpts, r_grid, beta_grid, angular_momentum,
                                      d_matrix, beta_gk=None):
  """jrystal/pseudopotential/nloc.py:43-141: (kpt, beta_total, m, x, y, z), atoms concatenated
  along beta."""
  g_vec = np.asarray(g_vec)
  gk = np.expand_dims(np.asarray(kpts), (1, 2, 3)) + g_vec[None]
  if beta_gk is None:
    beta_gk = beta_sbt_grid(r_grid, beta_grid, angular_momentum, g_vec, kpts)
  r = np.linalg.norm(gk, axis=-1)
  r = np.where(r == 0., 1e-10, r)
  phi = np.arccos(np.clip(gk[..., 2] / r, -1.0, 1.0))
  theta = np.mod(np.arctan2(gk[..., 1], gk[..., 0]) + 2 * np.pi, 2 * np.pi)
  l_max = int(np.max(np.hstack(angular_momentum)))
  y_lm = np.zeros((l_max + 1,) + r.shape + (2 * l_max + 1,))
  for l in range(l_max + 1):
    y_lm[l, ..., :2 * l + 1] = real_sph_harm(l, theta, phi)
  out = []
  for pos, ls, d, b in zip(positions, angular_momentum, d_matrix, beta_gk):
    ls = np.asarray(ls).astype(int)
    w, v = np.linalg.eigh(np.asarray(d))
    dsqrt = v * np.sqrt(w + 0j)
    o = np.einsum('ab,bkxyzm,kbxyz->kamxyz', dsqrt, y_lm[ls], b)
    o = o * np.exp(-1j * (gk @ np.asarray(pos)))[:, None, None]
    o = o * ((1j) ** ls)[None, :, None, None, None, None]
    out.append(o * 4 * np.pi)
  return np.concatenate(out, axis=1)


def energy_local(rho_g, v_loc, vol):
  """jrystal/pseudopotential/local.py:166-187."""
  return rp.reciprocal_braket(torch.as_tensor(v_loc), rho_g, vol)


def energy_terms(system, w_re, w_im, occupation, v_loc, phi_dense, xc='lda_x'):
  """total_energy of jrystal/calc/calc_ground_state_energy_normcons.py:175-192:
  (kinetic, hartree, external_local, external_nonlocal, xc, density).
  phi_dense: (kpt, proj, x, y, z) with (beta, m) flattened."""
  c = rp.coeff(w_re, w_im, system.mask)
  dens = rp.density_grid(c, system.vol, occupation)
  dens_g = torch.fft.fftn(dens, dim=(-3, -2, -1))
  e_kin = rp.energy_kinetic(system.g_vec, system.kpts, c, occupation)
  e_har = rp.energy_hartree(dens_g, system.g_vec, system.vol)
  e_loc = energy_local(dens_g, v_loc, system.vol)
  e_nl = rp.energy_nonlocal(c, torch.as_tensor(phi_dense), system.vol, occupation)
  e_xc = rp.energy_xc(dens, system.vol, xc, kohn_sham=False, g_vector_grid=system.g_vec)
  return e_kin, e_har, e_loc, e_nl, e_xc, dens


def energy_and_grad(system, w_re, w_im, occupation, v_loc, phi_dense, xc='lda_x', occ_grad=False):
  """value_and_grad of that loss w.r.t. {'w_re', 'w_im'} (+ occupation)
  (calc_ground_state_energy_normcons.py:226-233, optimiser excluded)."""
  wr = torch.from_numpy(np.asarray(w_re)).clone().requires_grad_(True)
  wi = torch.from_numpy(np.asarray(w_im)).clone().requires_grad_(True)
  occ = torch.from_numpy(np.asarray(occupation)).clone().requires_grad_(occ_grad)
  e_kin, e_har, e_loc, e_nl, e_xc, dens = energy_terms(system, wr, wi, occ, v_loc, phi_dense, xc)
  e_tot = e_kin + e_har + e_loc + e_nl + e_xc
  grads = torch.autograd.grad(e_tot, [wr, wi] + ([occ] if occ_grad else []))
  out = dict(e_kin=e_kin.item(), e_har=e_har.item(), e_loc=e_loc.item(), e_nl=e_nl.item(),
             e_xc=e_xc.item(), e_tot=e_tot.item(), density=dens.detach().numpy(),
             g_re=grads[0].numpy(), g_im=grads[1].numpy())
  if occ_grad:
    out['g_occ'] = grads[2].numpy()
  return out
