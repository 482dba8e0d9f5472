"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Second, independent CPU route to the same numbers as ``reference_port.energy_and_grad``:
the hand-derived forward/backward that the CUDA kernels implement (DESIGN.md "Math"),
written with numpy only (no autograd).  tests/test_oracle.py checks it against the
autograd route, which pins the algebra (H-apply backward, Cholesky-QR gauge, thin-QR
adjoint) before any CUDA is involved.
"""

import numpy as np

from . import reference_port as rp


def cholesky_qr2(w: np.ndarray):
  """W = Q R with R upper triangular, real POSITIVE diagonal (the product's gauge).
  LAPACK/Householder (reference, unitary_module.py:74-75) returns Q diag(+-1)."""
  s1 = w.conj().T @ w
  r1 = np.linalg.cholesky(s1).conj().T
  q1 = w @ np.linalg.inv(r1)
  s2 = q1.conj().T @ q1
  r2 = np.linalg.cholesky(s2).conj().T
  q = q1 @ np.linalg.inv(r2)
  return q, r2 @ r1


def qr_backward(q, r, g):
  """dE/dW* from g = dE/dQ* for W = Q R (thin QR, real-diagonal R):
  M = Q^H g ; X = -(up(M) + up(M)^H + diag(Re M)) ; gW = (g + Q X) R^{-H}."""
  m = q.conj().T @ g
  up = np.triu(m, 1)
  x = -(up + up.conj().T + np.diag(np.real(np.diag(m))))
  return (g + q @ x) @ np.linalg.inv(r).conj().T


def lda_x_eps_v(rho):
  pos = rho > rp.DENS_THRESHOLD
  eps = np.where(pos, rp.LDA_X_FACTOR * np.cbrt(np.where(pos, rho, 1.0)), 0.0)
  return eps, eps * (4.0 / 3.0)


def energy_and_grad(system: rp.System, w_re, w_im, occupation) -> dict:
  s = system
  ns, nk, ng, nb = w_re.shape
  n = int(np.prod(s.grid_sizes))
  mask = s.mask
  gk2 = np.stack([np.sum((s.g_vec[mask] + k)**2, axis=-1) for k in s.kpts])  # [k,g]
  q = np.zeros((ns, nk, ng, nb), dtype=np.complex128)
  r = np.zeros((ns, nk, nb, nb), dtype=np.complex128)
  psi = np.zeros((ns, nk, nb) + mask.shape, dtype=np.complex128)
  for i in range(ns):
    for k in range(nk):
      q[i, k], r[i, k] = cholesky_qr2(w_re[i, k] + 1j * w_im[i, k])
      box = np.zeros((nb,) + mask.shape, dtype=np.complex128)
      box[:, mask] = q[i, k].T
      psi[i, k] = np.fft.ifftn(box, axes=(-3, -2, -1)) * (n / np.sqrt(s.vol))
  rho = np.einsum('skbxyz,skb->sxyz', np.abs(psi)**2, occupation)
  t_kb = 0.5 * np.einsum('kg,skgb->skb', gk2, np.abs(q)**2)
  e_kin = float(np.sum(occupation * t_kb))
  # grid terms
  g2 = np.sum(s.g_vec**2, axis=-1)
  g2s = g2.copy()
  g2s[0, 0, 0] = 1.0
  rho_g = np.fft.fftn(rho, axes=(-3, -2, -1))
  n_g = rho_g.sum(0)
  vh_full = 4 * np.pi * n_g / g2s          # un-halved Hartree potential (= dE_H/drho)
  vh_full[0, 0, 0] = 0
  vext_g = rp.external_reciprocal(s.positions, s.charges, s.g_vec, s.vol).numpy()
  e_har = float(np.real(np.sum(np.conj(0.5 * vh_full) * n_g)) * s.vol / n / n)
  e_ext = float(np.real(np.sum(np.conj(vext_g) * n_g)) * s.vol / n / n)
  eps, vxc = lda_x_eps_v(rho[0])
  e_xc = float(np.sum(eps * rho[0]) * s.vol / n)
  v_eff = np.real(np.fft.ifftn(vh_full + vext_g)) + vxc
  # H-apply and QR adjoint
  g_re = np.zeros_like(w_re)
  g_im = np.zeros_like(w_im)
  eps_kb = np.zeros((ns, nk, nb))
  for i in range(ns):
    for k in range(nk):
      vpsi = np.fft.fftn(v_eff[None] * psi[i, k], axes=(-3, -2, -1))
      hq = 0.5 * gk2[k][:, None] * q[i, k] + (np.sqrt(s.vol) / n) * vpsi[:, mask].T
      eps_kb[i, k] = np.real(np.sum(np.conj(q[i, k]) * hq, axis=0))
      gq = hq * occupation[i, k][None, :]           # dE/dQ*
      gw = qr_backward(q[i, k], r[i, k], gq)        # dE/dW*
      g_re[i, k] = 2 * np.real(gw)
      g_im[i, k] = 2 * np.imag(gw)
  return dict(e_kin=e_kin, e_ext=e_ext, e_har=e_har, e_xc=e_xc,
              e_tot=e_kin + e_ext + e_har + e_xc, density=rho, g_re=g_re, g_im=g_im,
              g_occ=eps_kb, q=q, r=r, v_eff=v_eff)
