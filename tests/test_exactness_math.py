"""Mathematics behind two exact shortcuts of the product, checked on the CPU with numpy only (no CUDA,
no product kernels).

ORBITAL GRID (jrb_plan_set_orbital_grid): transforming the orbitals on any box with
n >= 4 gmax + 1 and moving rho / v_eff between the boxes by Fourier interpolation / truncation
reproduces what the reference computes on its own grid (pw.wave_grid + density_grid,
jrystal/_src/pw.py:208-284; the sphere part of fftn(v_eff psi), the backward pass) to rounding --
including the case where the caller's grid itself is too coarse on some axes and only the other
axis shrinks; one point below the bound aliases.

SECOND CHOLESKY-QR PASS (k_near_identity): the closed-form factor of I + E."""
import numpy as np
import pytest

from jrystal_b200 import grid as jgrid
from oracle import reference_port as rp


def _freqs(mask):
  idx = np.argwhere(mask)
  dims = np.array(mask.shape)
  return np.where(idx < (dims + 1) // 2, idx, idx - dims)


def _scatter(coeff, freqs, box):
  """coeff (nb, ng) on the sphere -> dense (nb, *box) with the frequencies folded into the box."""
  out = np.zeros((coeff.shape[0],) + tuple(box), dtype=np.complex128)
  i = freqs % np.array(box)
  out[:, i[:, 0], i[:, 1], i[:, 2]] = coeff
  return out


def _resample(a_hat, box):
  """Fourier coefficients on one box -> another: bins both hold symmetrically (2|f| < min n) on
  resized axes, bin by bin on unchanged axes (what k_resample does)."""
  src = a_hat.shape
  out = np.zeros(box, dtype=np.complex128)
  sel_src, sel_dst = [], []
  for ns_, nd in zip(src, box):
    if ns_ == nd:
      f = np.arange(nd)
      sel_src.append(f)
      sel_dst.append(f)
    else:
      m = min(ns_, nd)
      f = np.array([k for k in range(-(m // 2), m // 2 + 1) if 2 * abs(k) < m])
      sel_src.append(f % ns_)
      sel_dst.append(f % nd)
  out[np.ix_(*sel_dst)] = a_hat[np.ix_(*sel_src)]
  return out


CASES = [
  # (grid, cutoff, orbital box)
  ([24, 24, 24], 8.0, (16, 16, 16)),
  ([24, 24, 24], 8.0, (24, 24, 15)),      # odd length
  ([20, 24, 32], 8.0, (20, 16, 18)),
  ([12, 12, 32], 10.0, (12, 12, 16)),     # x, y under-resolved by the caller's grid: z only
]


@pytest.mark.parametrize('gridsz,cutoff,box', CASES)
def test_orbital_box_reproduces_density_and_happly(gridsz, cutoff, box):
  s = rp.System.from_name('diamond', gridsz, [1, 1, 1], cutoff)
  mask = s.mask
  full = mask.shape
  need = jgrid.min_orbital_grid(mask)
  assert all(b >= m or b == f for b, m, f in zip(box, need, full)), (need, box)
  rng = np.random.default_rng(4)
  nb = 5
  fr = _freqs(mask)
  c = rng.standard_normal((nb, fr.shape[0])) + 1j * rng.standard_normal((nb, fr.shape[0]))
  occ = rng.random(nb)
  n_full, n_box = int(np.prod(full)), int(np.prod(box))

  # reference: everything on the full grid
  psi_f = np.fft.ifftn(_scatter(c, fr, full), axes=(1, 2, 3)) * n_full
  rho_f = np.einsum('b,bxyz->xyz', occ, np.abs(psi_f)**2)
  v = rng.standard_normal(full)                       # any real potential on the full grid
  hv_f = np.fft.fftn(v * psi_f, axes=(1, 2, 3)) / n_full
  i = fr % np.array(full)
  hv_f = hv_f[:, i[:, 0], i[:, 1], i[:, 2]]

  # orbital box
  psi_b = np.fft.ifftn(_scatter(c, fr, box), axes=(1, 2, 3)) * n_box
  rho_b = np.einsum('b,bxyz->xyz', occ, np.abs(psi_b)**2)
  rho_up = np.fft.ifftn(_resample(np.fft.fftn(rho_b) / n_box, full)) * n_full
  assert np.abs(rho_up.imag).max() < 1e-12 * np.abs(rho_f).max()
  assert np.abs(rho_up.real - rho_f).max() < 1e-12 * np.abs(rho_f).max()

  v_b = np.fft.ifftn(_resample(np.fft.fftn(v) / n_full, box)) * n_box
  assert np.abs(v_b.imag).max() < 1e-12 * np.abs(v).max()
  hv_b = np.fft.fftn(v_b.real * psi_b, axes=(1, 2, 3)) / n_box
  j = fr % np.array(box)
  hv_b = hv_b[:, j[:, 0], j[:, 1], j[:, 2]]
  assert np.abs(hv_b - hv_f).max() < 1e-12 * np.abs(hv_f).max()


def test_a_box_below_the_bound_aliases():
  """4 gmax + 1 is sharp: one point less on an axis and the density differs."""
  s = rp.System.from_name('diamond', [24, 24, 24], [1, 1, 1], 8.0)
  need = jgrid.min_orbital_grid(s.mask)
  box = (24, 24, need[2] - 1)
  fr = _freqs(s.mask)
  rng = np.random.default_rng(4)
  c = rng.standard_normal((3, fr.shape[0])) + 1j * rng.standard_normal((3, fr.shape[0]))
  n_full, n_box = 24**3, int(np.prod(box))
  rho_f = (np.abs(np.fft.ifftn(_scatter(c, fr, (24, 24, 24)), axes=(1, 2, 3)) * n_full)**2).sum(0)
  rho_b = (np.abs(np.fft.ifftn(_scatter(c, fr, box), axes=(1, 2, 3)) * n_box)**2).sum(0)
  rho_up = np.fft.ifftn(_resample(np.fft.fftn(rho_b) / n_box, (24, 24, 24))).real * n_full
  assert np.abs(rho_up - rho_f).max() > 1e-6 * np.abs(rho_f).max()


@pytest.mark.parametrize('nb,scale', [(66, 1e-12), (208, 9e-11)])
def test_closed_form_second_pass_cholesky(nb, scale):
  """k_near_identity (qr.cu): for S = I + E, max|E| < 1e-10, R = I + up(E) + diag(E)/2 and
  R^-1 = I - U (+ u^2 on the diagonal) differ from the true Cholesky factor and its inverse by
  O(nb |E|^2), far below the rounding of the factorisation they replace."""
  rng = np.random.default_rng(nb)
  a = rng.standard_normal((nb, nb)) + 1j * rng.standard_normal((nb, nb))
  e = (a + a.conj().T) * (scale / np.abs(a + a.conj().T).max())
  s = np.eye(nb) + e
  u = np.triu(e, 1) + np.diag(np.diag(e).real / 2)
  r = np.eye(nb) + u
  d = np.diag(u).real
  rinv = np.eye(nb) - u + np.diag(d * d)
  r_true = np.linalg.cholesky(s).conj().T
  assert np.abs(r - r_true).max() < 4 * nb * scale**2 + 4e-16
  assert np.abs(r.conj().T @ r - s).max() < 4 * nb * scale**2 + 4e-16
  assert np.abs(rinv @ r_true - np.eye(nb)).max() < 4 * nb * scale**2 + 4e-16
