#!/usr/bin/env python
"""Design study (CPU, numpy) for DESIGN.md section 9 item 1b: the Gram product S = W^H W of the
Cholesky-QR on INT8 tensor cores by Ozaki splitting.  Emulates exactly what the device kernel
would do -- per-ROW power-of-two scaling of the (ng x nb) panel, `slices` signed slices of `bits`
bits each, every slice-pair product accumulated exactly in int32 (as tcgen05.mma kind::i8 does),
recombination in FP64 -- and reports, for the benchmark shapes, the error of S against a long
double Gram and the orthogonality of the Q that Cholesky-QR2 then produces.

  python tools/ozaki_gram_study.py            # C2-like (ng 8409, nb 66) and C3a-like (29423, 208)

Scaling: the exponent must not depend on the summation index g, so the split is per COLUMN (per
band) of W: a_gb = 2^e_b * sum_s q_s 2^(-bits (s+1)), q_s integers in [-2^(bits-1), 2^(bits-1)].
That loses nothing here: the columns of W ~ U[0, 1) and of Q1 (orthonormal) are uniformly scaled.
The integer GEMMs are evaluated in float64 BLAS, which is exact for these sizes
(|sum| < ng 2^(2 bits - 2) << 2^53) and stands for the int32 accumulators of the device."""
import sys
import time

import numpy as np


def split_columns(a, slices, bits):
  """a (ng, nb) real -> (exponents (nb,), q (slices, ng, nb) int8/int16 slices)."""
  e = np.ceil(np.log2(np.abs(a).max(axis=0) + 1e-300)).astype(np.int64) + 1
  r = a / np.exp2(e)[None, :]                    # |r| < 1/2
  q = np.empty((slices,) + a.shape, dtype=np.int16)
  for s in range(slices):
    r = r * (1 << bits)
    q[s] = np.rint(r)                             # |q| <= 2^(bits-1)
    r = r - q[s]
  return e, q


def ozaki_matmul_t(qa, ea, qb, eb, bits, keep):
  """sum_g A_gi B_gj from slices; slice pairs (s, t) with s + t < keep are formed (the rest is
  below the target accuracy): the number of integer GEMMs is what the device pays."""
  slices = qa.shape[0]
  nb_a, nb_b = qa.shape[2], qb.shape[2]
  out = np.zeros((nb_a, nb_b), dtype=np.float64)
  gemms = 0
  for level in range(keep - 1, -1, -1):           # small terms first
    acc = np.zeros((nb_a, nb_b), dtype=np.float64)
    for s in range(min(level, slices - 1) + 1):
      t = level - s
      if t >= slices:
        continue
      acc += qa[s].T.astype(np.float64) @ qb[t].astype(np.float64)   # exact, see the header
      gemms += 1
    out += acc * np.exp2(-bits * (level + 2))
  return out * np.exp2(ea)[:, None] * np.exp2(eb)[None, :], gemms


def gram(w, slices, bits, keep):
  """Complex Hermitian Gram by four real slice products (re re + im im, re im - im re)."""
  er, qr = split_columns(w.real, slices, bits)
  ei, qi = split_columns(w.imag, slices, bits)
  rr, g1 = ozaki_matmul_t(qr, er, qr, er, bits, keep)
  ii, g2 = ozaki_matmul_t(qi, ei, qi, ei, bits, keep)
  ri, g3 = ozaki_matmul_t(qr, er, qi, ei, bits, keep)
  return (rr + ii) + 1j * (ri - ri.T), g1 + g2 + g3


def cholqr2(w, gram_fn):
  s1 = gram_fn(w)
  r1 = np.linalg.cholesky(s1).conj().T
  q1 = w @ np.linalg.inv(r1)
  s2 = gram_fn(q1)
  r2 = np.linalg.cholesky(s2).conj().T
  return q1 @ np.linalg.inv(r2)


def study(ng, nb, bits=7):
  rng = np.random.default_rng(0)
  w = rng.random((ng, nb)) + 1j * rng.random((ng, nb))
  wl = w.astype(np.clongdouble)
  ref = (wl.conj().T @ wl)
  s64 = w.conj().T @ w
  print(f'ng {ng} nb {nb}: FP64 Gram rel err {np.abs(s64 - ref).max() / np.abs(ref).max():.1e}')
  for slices, keep in ((6, 6), (7, 7), (8, 8)):
    t0 = time.time()
    s, gemms = gram(w, slices, bits, keep)
    err = float(np.abs(s - ref).max() / np.abs(ref).max())
    q = cholqr2(w, lambda m: gram(m, slices, bits, keep)[0])
    orth = float(np.abs(q.conj().T @ q - np.eye(nb)).max())
    print(f'  {slices} slices x {bits} bits, levels < {keep}: {gemms:3d} int8 GEMMs per complex Gram, '
          f'Gram rel err {err:.1e}, |Q^H Q - I| after Cholesky-QR2 {orth:.1e}  ({time.time() - t0:.1f} s)')


if __name__ == '__main__':
  shapes = [(8409, 66), (29423, 208)] if len(sys.argv) == 1 else [tuple(map(int, sys.argv[1:3]))]
  for ng, nb in shapes:
    study(ng, nb)
