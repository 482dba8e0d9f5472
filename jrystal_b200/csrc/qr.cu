// Orthonormalisation of the plane-wave coefficient matrix and its adjoint (K1).
//
// Replaces jnp.linalg.qr(w_re + i w_im, mode='reduced')[0] batched over (spin, k)
// (jrystal/_src/unitary_module.py:66-81; XLA: cuSOLVER geqrf+orgqr) and the QR AD rule that
// jax.value_and_grad applies to it (calc/calc_ground_state_energy_all_electrons.py:178).
//
// Algorithm: Cholesky-QR2.  S = W^H W (tall-skinny Gram on the FP64 tensor cores, DMMA
// mma.sync.m8n8k4.f64 -- tcgen05/TMEM have no FP64 type), R = chol(S)^H, Q1 = W R1^-1, then the
// same once more on Q1 for orthogonality at round-off level; R = R2 R1 has a real positive
// diagonal (gauge documented in include/jrystal_b200.h).
// Adjoint (DESIGN.md "Math"):  M = Q^H G,  X = -(up(M) + up(M)^H + diag Re M),
//   dE/dW* = (G + Q X) R^-H  with G = HQ diag(occ)  ->  dE/dW* = HQ (F R^-H) + Q (X R^-H),
// i.e. one more Gram-shaped product and one two-term tall-skinny product, both on DMMA.
//
// Complex products are 4 real DMMA products on separate re/im planes in shared memory
// (leading dimensions == 4 mod 16 doubles make every fragment load bank-conflict free).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include <cooperative_groups.h>

#include "plan.h"
#include "qr_gemm.cuh"

namespace jrb {

// ---------------------------------------------------------------------------------------
// Small (nb x nb) per-(spin,k) algebra.  Everything that is element-parallel runs on a
// (tiles, nsk) grid; only the Cholesky recurrence is one CTA per (spin,k), blocked so that the
// active 32-column panel lives in shared memory.
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) {  // a * conj(b)
  return cmake(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

constexpr int SMALL_T = 256;  // threads of the element-parallel kernels
constexpr int CHOL_PB = 32;   // Cholesky panel width
constexpr int CHOL_KC = 16;   // depth of one staged slice of the left-looking update
constexpr int CHOL_T = 512;

// S[sk] = sum over chunks of the Gram partials (upper blocks), mirrored as Hermitian.
// grid: (ceil(nb^2 / 256), nsk)
__global__ void __launch_bounds__(SMALL_T)
k_gram_reduce(const cplx* __restrict__ partial, int nchunks, int nb, cplx* __restrict__ S) {
  const int sk = blockIdx.y, nsk = gridDim.y;
  const long long nn = (long long)nb * nb;
  const long long e = (long long)blockIdx.x * SMALL_T + threadIdx.x;
  if (e >= nn) return;
  const int i = (int)(e / nb), j = (int)(e % nb);
  const long long src = j >= i ? e : (long long)j * nb + i;
  cplx s = cmake(0.0, 0.0);
  for (int c = 0; c < nchunks; ++c) {
    const cplx v = partial[((long long)c * nsk + sk) * nn + src];
    s.x += v.x; s.y += v.y;
  }
  S[sk * nn + e] = j >= i ? s : cconj(s);
}

// In-place blocked left-looking Cholesky S = L L^H of one Hermitian nb x nb matrix per CTA.
// On exit the lower triangle of S holds L and Rt = L^H (upper triangular, zeros below).
// dynamic smem: panel [nb][PB + 1] + stage [nb][KC + 1] complex
__global__ void __launch_bounds__(CHOL_T)
k_chol_blocked(cplx* __restrict__ S, cplx* __restrict__ Rt, int nb, int* __restrict__ fail_flag,
               const int* __restrict__ skip) {
  if (skip && skip[blockIdx.x]) return;  // factor already written by k_near_identity
  extern __shared__ __align__(16) unsigned char smem_raw_[];
  cplx* P = reinterpret_cast<cplx*>(smem_raw_);     // [rows][PB + 1]
  cplx* T = P + (size_t)nb * (CHOL_PB + 1);         // [rows][KC + 1]
  constexpr int LP = CHOL_PB + 1, LT = CHOL_KC + 1;
  const long long nn = (long long)nb * nb;
  S += blockIdx.x * nn;
  Rt += blockIdx.x * nn;
  const int tid = threadIdx.x;
  for (int p0 = 0; p0 < nb; p0 += CHOL_PB) {
    const int pw = min(CHOL_PB, nb - p0), rows = nb - p0;
    // panel <- S[p0.., p0..p0+pw)
    for (int e = tid; e < rows * pw; e += CHOL_T) {
      const int r = e / pw, c = e % pw;
      P[r * LP + c] = S[(long long)(p0 + r) * nb + p0 + c];
    }
    // left-looking update with the finished columns k < p0, KC at a time:
    //   P[r][c] -= sum_k L[p0 + r][k] conj(L[p0 + c][k])
    for (int k0 = 0; k0 < p0; k0 += CHOL_KC) {
      __syncthreads();
      for (int e = tid; e < rows * CHOL_KC; e += CHOL_T) {
        const int r = e / CHOL_KC, k = e % CHOL_KC;
        T[r * LT + k] = S[(long long)(p0 + r) * nb + k0 + k];  // k0 + k < p0 (p0 multiple of KC)
      }
      __syncthreads();
      for (int e = tid; e < rows * pw; e += CHOL_T) {
        const int r = e / pw, c = e % pw;
        if (r >= c) {
          cplx acc = P[r * LP + c];
#pragma unroll
          for (int k = 0; k < CHOL_KC; ++k) {
            const cplx v = cmulc(T[r * LT + k], T[c * LT + k]);
            acc.x -= v.x; acc.y -= v.y;
          }
          P[r * LP + c] = acc;
        }
      }
    }
    __syncthreads();
    // factor the panel in shared memory
    for (int j = 0; j < pw; ++j) {
      const double djj = P[j * LP + j].x;
      if (!(djj > 0.0) && tid == 0) atomicExch(fail_flag, 1);
      const double d = sqrt(djj > 0.0 ? djj : 1.0);
      const double inv = 1.0 / d;
      __syncthreads();
      for (int r = j + tid; r < rows; r += CHOL_T) {
        cplx v = P[r * LP + j];
        P[r * LP + j] = r == j ? cmake(d, 0.0) : cmake(v.x * inv, v.y * inv);
      }
      __syncthreads();
      const int wrem = pw - j - 1;
      for (int e = tid; e < (rows - j - 1) * wrem; e += CHOL_T) {
        const int r = j + 1 + e / wrem, c = j + 1 + e % wrem;
        if (r >= c) {
          const cplx v = cmulc(P[r * LP + j], P[c * LP + j]);
          cplx u = P[r * LP + c];
          P[r * LP + c] = cmake(u.x - v.x, u.y - v.y);
        }
      }
      __syncthreads();
    }
    // write L (lower) back and R = L^H
    for (int e = tid; e < rows * pw; e += CHOL_T) {
      const int r = e / pw, c = e % pw;
      const cplx v = r >= c ? P[r * LP + c] : cmake(0.0, 0.0);
      S[(long long)(p0 + r) * nb + p0 + c] = v;
      Rt[(long long)(p0 + c) * nb + p0 + r] = cconj(v);
      if (r < c) Rt[(long long)(p0 + c) * nb + p0 + r] = cmake(0.0, 0.0);
    }
    __syncthreads();
  }
  // strictly lower part of Rt that no panel touched (columns left of each panel)
  for (long long e = tid; e < nn; e += CHOL_T) {
    const int i = (int)(e / nb), j = (int)(e % nb);
    if (j < i && (i / CHOL_PB) != (j / CHOL_PB)) Rt[e] = cmake(0.0, 0.0);
  }
}

// Rinv = R^-1 for upper-triangular R (complex diagonal allowed): one warp per column of the
// inverse (columns are independent), back substitution with the column kept in shared memory.
// grid: (ceil(nb / 8), nsk), block 256; dynamic smem: 8 * nb complex
__global__ void __launch_bounds__(256)
k_tri_inv_cols(const cplx* __restrict__ R, int nb, cplx* __restrict__ Rinv,
               const int* __restrict__ skip) {
  if (skip && skip[blockIdx.y]) return;
  extern __shared__ __align__(16) unsigned char smem_raw_[];
  const long long nn = (long long)nb * nb;
  R += blockIdx.y * nn;
  Rinv += blockIdx.y * nn;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + warp;
  if (j >= nb) return;
  cplx* x = reinterpret_cast<cplx*>(smem_raw_) + (size_t)warp * nb;
  {
    const cplx d = R[(long long)j * nb + j];
    const double n2 = d.x * d.x + d.y * d.y;
    if (lane == 0) x[j] = cmake(d.x / n2, -d.y / n2);
  }
  __syncwarp();
  for (int i = j - 1; i >= 0; --i) {
    double sx = 0.0, sy = 0.0;
    const cplx* row = R + (long long)i * nb;
    for (int k = i + 1 + lane; k <= j; k += 32) {
      const cplx v = cmul(row[k], x[k]);
      sx += v.x; sy += v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sy += __shfl_xor_sync(0xffffffffu, sy, o);
    }
    if (lane == 0) {
      const cplx d = row[i];
      const double n2 = d.x * d.x + d.y * d.y;
      x[i] = cmake(-(sx * d.x + sy * d.y) / n2, -(sy * d.x - sx * d.y) / n2);  // -(s / d)
    }
    __syncwarp();
  }
  for (int i = lane; i < nb; i += 32) Rinv[(long long)i * nb + j] = i <= j ? x[i] : cmake(0.0, 0.0);
}

// ---------------------------------------------------------------------------------------
// Large matrices (Gamma-only supercells: one or a few 200 x 200 matrices): the one-CTA
// recurrences above are pure latency (0.45 ms + 0.27 ms per call at nb = 208,
// profiles/r01_ncu_c3a_baseline.md), so the work of one panel is spread over a CTA per
// 32-row tile and the inverse over a CTA per 8 columns.
//
// Left-looking panel step of the Cholesky factorisation, one launch per 32-column panel p0.
// Tile 0 owns the diagonal block, tile t > 0 the rows [p0 + 32 t, p0 + 32 t + 32).  Every CTA
// updates and factors the diagonal block itself (redundant, but it removes a grid-wide
// dependency), then solves its own rows against it.
// grid: (1 + ceil((nb - p0 - 32) / 32), nsk), block 256
constexpr int CP = 32;        // panel width == tile height
constexpr int CP_LD = CP + 1;
constexpr int CP_SMEM = 4 * CP * CP_LD * (int)sizeof(cplx);  // D, T, Ld, Lr
__global__ void __launch_bounds__(256)
k_chol_panel(cplx* __restrict__ S, cplx* __restrict__ Rt, int nb, int p0, int* __restrict__ fail_flag,
             const int* __restrict__ skip, int right_looking) {
  if (skip && skip[blockIdx.y]) return;
  extern __shared__ __align__(16) unsigned char smem_raw_[];
  cplx* D = reinterpret_cast<cplx*>(smem_raw_);
  cplx* T = D + CP * CP_LD;
  cplx* Ld = T + CP * CP_LD;
  cplx* Lr = Ld + CP * CP_LD;
  const long long nn = (long long)nb * nb;
  S += blockIdx.y * nn;
  Rt += blockIdx.y * nn;
  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int pw = min(CP, nb - p0);
  const int r0 = p0 + CP * tile;
  const int nr = tile > 0 ? min(CP, nb - r0) : 0;
  // each thread owns the entries (r, c), r = tid / 32 + 8 j, c = tid % 32
  const int c = tid & 31, rb = tid >> 5;
  cplx accD[4], accT[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = rb + 8 * j;
    accD[j] = (r < pw && c < pw) ? S[(long long)(p0 + r) * nb + p0 + c] : cmake(0.0, 0.0);
    accT[j] = (r < nr && c < pw) ? S[(long long)(r0 + r) * nb + p0 + c] : cmake(0.0, 0.0);
  }
  // left-looking update with the finished columns k < p0, 32 at a time; the next slice is
  // fetched into registers while the current one is multiplied
  cplx pd[4], pr[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = rb + 8 * j;
      pd[j] = r < pw ? S[(long long)(p0 + r) * nb + k0 + c] : cmake(0.0, 0.0);
      pr[j] = r < nr ? S[(long long)(r0 + r) * nb + k0 + c] : cmake(0.0, 0.0);
    }
  };
  // right-looking mode: the finished panels were already subtracted from the trailing matrix by
  // k_trailing_update (many CTAs per panel instead of this CTA's serial slices)
  const int kend = right_looking ? 0 : p0;
  if (kend > 0) fetch(0);
  for (int k0 = 0; k0 < kend; k0 += CP) {  // p0 is a multiple of CP
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      Ld[(rb + 8 * j) * CP_LD + c] = pd[j];
      Lr[(rb + 8 * j) * CP_LD + c] = pr[j];
    }
    __syncthreads();
    if (k0 + CP < kend) fetch(k0 + CP);
    if (tile == 0) {  // the diagonal tile has no rows below (whole-CTA branch)
#pragma unroll 8
      for (int k = 0; k < CP; ++k) {
        const cplx lc = Ld[c * CP_LD + k];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const cplx u = cmulc(Ld[(rb + 8 * j) * CP_LD + k], lc);
          accD[j].x -= u.x; accD[j].y -= u.y;
        }
      }
    } else {
#pragma unroll 8
      for (int k = 0; k < CP; ++k) {
        const cplx lc = Ld[c * CP_LD + k];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = rb + 8 * j;
          const cplx u = cmulc(Ld[r * CP_LD + k], lc);
          accD[j].x -= u.x; accD[j].y -= u.y;
          const cplx v = cmulc(Lr[r * CP_LD + k], lc);
          accT[j].x -= v.x; accT[j].y -= v.y;
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    D[(rb + 8 * j) * CP_LD + c] = accD[j];
    T[(rb + 8 * j) * CP_LD + c] = accT[j];
  }
  __syncthreads();
  // Unscaled right-looking elimination of the diagonal block and, in the same sweep, of the rows
  // below it (one barrier per column): after step j column j of D and T is final up to the
  // factor 1 / sqrt(d_j), d_j = D[j][j].
  for (int j = 0; j < pw; ++j) {
    const double djj = D[j * CP_LD + j].x;
    const double rinv = 1.0 / (djj > 0.0 ? djj : 1.0);
    if (c > j && c < pw) {
      const cplx lc = D[c * CP_LD + j];
      const cplx lcs = cmake(lc.x * rinv, lc.y * rinv);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int r = rb + 8 * jj;
        if (r >= c && r < pw) {
          const cplx v = cmulc(D[r * CP_LD + j], lcs);
          D[r * CP_LD + c].x -= v.x;
          D[r * CP_LD + c].y -= v.y;
        }
        if (r < nr) {
          const cplx v = cmulc(T[r * CP_LD + j], lcs);
          T[r * CP_LD + c].x -= v.x;
          T[r * CP_LD + c].y -= v.y;
        }
      }
    }
    __syncthreads();
  }
  const double dcc = c < pw ? D[c * CP_LD + c].x : 1.0;
  if (!(dcc > 0.0) && tile == 0 && rb == 0) atomicExch(fail_flag, 1);
  const double sc = 1.0 / sqrt(dcc > 0.0 ? dcc : 1.0);
  if (tile == 0) {
    // L (lower) back into S, R = L^H into Rt (zeros below its diagonal)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = rb + 8 * j;
      if (r < pw && c < pw) {
        cplx v = cmake(0.0, 0.0);
        if (r > c) v = cmake(D[r * CP_LD + c].x * sc, D[r * CP_LD + c].y * sc);
        if (r == c) v = cmake(dcc > 0.0 ? sqrt(dcc) : 1.0, 0.0);
        S[(long long)(p0 + r) * nb + p0 + c] = v;
        Rt[(long long)(p0 + c) * nb + p0 + r] = cconj(v);
      }
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = rb + 8 * j;
    if (r < nr && c < pw) {
      const cplx v = cmake(T[r * CP_LD + c].x * sc, T[r * CP_LD + c].y * sc);
      S[(long long)(r0 + r) * nb + p0 + c] = v;
      Rt[(long long)(p0 + c) * nb + r0 + r] = cconj(v);
      Rt[(long long)(r0 + r) * nb + p0 + c] = cmake(0.0, 0.0);
    }
  }
}

// Right-looking companion of k_chol_panel: after panel p0 is factored (its L sits in the lower
// part of S), the trailing matrix gets  S[i][j] -= sum_{k in panel} L[i][k] conj(L[j][k])  for
// i >= j >= p0 + 32, one CTA per 32 x 32 lower tile -- the rank-32 update spread over up to 21 CTAs
// instead of the serial left-looking slices inside each panel CTA (the 61 us per panel at nb = 208
// were half this update).   grid: (nt (nt + 1) / 2, nsk), nt = ceil((nb - p0 - 32) / 32); block 256
__global__ void __launch_bounds__(256)
k_trailing_update(cplx* __restrict__ S, int nb, int p0, const int* __restrict__ skip) {
  if (skip && skip[blockIdx.y]) return;
  __shared__ cplx Li[CP * CP_LD], Lj[CP * CP_LD];
  S += (long long)blockIdx.y * nb * nb;
  int ti = 0, rem = blockIdx.x;  // lower tiles enumerated row by row: ti >= tj
  while (rem > ti) {
    rem -= ti + 1;
    ++ti;
  }
  const int tj = rem;
  const int base = p0 + CP, i0 = base + CP * ti, j0 = base + CP * tj;
  const int pw = min(CP, nb - p0);
  const int tid = threadIdx.x, c = tid & 31, rb = tid >> 5;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int r = rb + 8 * q;
    Li[r * CP_LD + c] = (i0 + r < nb && c < pw) ? S[(long long)(i0 + r) * nb + p0 + c] : cmake(0.0, 0.0);
    Lj[r * CP_LD + c] = (j0 + r < nb && c < pw) ? S[(long long)(j0 + r) * nb + p0 + c] : cmake(0.0, 0.0);
  }
  __syncthreads();
  cplx acc[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) acc[q] = cmake(0.0, 0.0);
#pragma unroll 8
  for (int k = 0; k < CP; ++k) {
    const cplx lc = Lj[c * CP_LD + k];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const cplx u = cmulc(Li[(rb + 8 * q) * CP_LD + k], lc);
      acc[q].x += u.x;
      acc[q].y += u.y;
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = i0 + rb + 8 * q, j = j0 + c;
    if (i < nb && j < nb && i >= j) {
      cplx v = S[(long long)i * nb + j];
      S[(long long)i * nb + j] = cmake(v.x - acc[q].x, v.y - acc[q].y);
    }
  }
}

// Blocked inverse of an upper-triangular R, step 1: the 32 x 32 diagonal blocks (one thread per
// column, the block staged in shared memory).   grid: (ceil(nb / 32), nsk), block 32
__global__ void __launch_bounds__(32)
k_tri_inv_diag(const cplx* __restrict__ R, int nb, cplx* __restrict__ Rinv,
               const int* __restrict__ skip) {
  if (skip && skip[blockIdx.y]) return;
  __shared__ cplx U[CP * CP_LD], X[CP * CP_LD];
  const long long nn = (long long)nb * nb;
  R += blockIdx.y * nn;
  Rinv += blockIdx.y * nn;
  const int b0 = blockIdx.x * CP, bw = min(CP, nb - b0);
  const int j = threadIdx.x;
  for (int i = 0; i < bw; ++i)
    if (j < bw) U[i * CP_LD + j] = R[(long long)(b0 + i) * nb + b0 + j];
  __syncwarp();
  if (j < bw) {
    for (int i = bw - 1; i > j; --i) X[i * CP_LD + j] = cmake(0.0, 0.0);
    {
      const cplx d = U[j * CP_LD + j];
      const double n2 = d.x * d.x + d.y * d.y;
      X[j * CP_LD + j] = cmake(d.x / n2, -d.y / n2);
    }
    for (int i = j - 1; i >= 0; --i) {
      double sx = 0.0, sy = 0.0;
      for (int k = i + 1; k <= j; ++k) {
        const cplx v = cmul(U[i * CP_LD + k], X[k * CP_LD + j]);
        sx += v.x; sy += v.y;
      }
      const cplx d = U[i * CP_LD + i];
      const double n2 = d.x * d.x + d.y * d.y;
      X[i * CP_LD + j] = cmake(-(sx * d.x + sy * d.y) / n2, -(sy * d.x - sx * d.y) / n2);
    }
  }
  __syncwarp();
  for (int i = 0; i < bw; ++i)
    if (j < bw) Rinv[(long long)(b0 + i) * nb + b0 + j] = X[i * CP_LD + j];
}

// Step 2: the blocks above the diagonal.  A CTA owns 8 columns of block column J and walks the
// block rows I = J-1 .. 0:  X[I] = -Xd[I] (sum_{K = I+1..J} R[I][K] X[K]),  Xd = diagonal-block
// inverses of step 1 (already in Rinv).  Also zeroes its columns below the diagonal block.
// grid: (4 * ceil(nb / 32), nsk), block 256 = 32 rows x 8 columns;
// dynamic smem: (ceil(nb / 32) * 32) x 8 complex (the CTA's columns of X)
__global__ void __launch_bounds__(256)
k_tri_inv_offdiag(const cplx* __restrict__ R, int nb, cplx* __restrict__ Rinv,
                  const int* __restrict__ skip) {
  if (skip && skip[blockIdx.y]) return;
  extern __shared__ __align__(16) unsigned char smem_raw_[];
  cplx* Xc = reinterpret_cast<cplx*>(smem_raw_);  // [rows][8]
  __shared__ cplx A[CP * CP_LD], Y[CP * 8];
  const long long nn = (long long)nb * nb;
  R += blockIdx.y * nn;
  Rinv += blockIdx.y * nn;
  const int J = blockIdx.x >> 2, c0 = J * CP + (blockIdx.x & 3) * 8;
  const int tid = threadIdx.x, r = tid >> 3, c = tid & 7;
  const int col = c0 + c;
  const bool col_ok = col < nb;
  // diagonal block rows of these columns (from step 1); rows below the block are zero
  {
    const int row = J * CP + r;
    Xc[row * 8 + c] = (row < nb && col_ok) ? Rinv[(long long)row * nb + col] : cmake(0.0, 0.0);
    for (int rr = (J + 1) * CP + r; rr < nb; rr += CP)
      if (col_ok) Rinv[(long long)rr * nb + col] = cmake(0.0, 0.0);
  }
  for (int I = J - 1; I >= 0; --I) {
    cplx acc = cmake(0.0, 0.0);
    for (int K = I + 1; K <= J; ++K) {
      __syncthreads();
      for (int e = tid; e < CP * CP; e += 256) {
        const int i = e >> 5, k = e & 31;
        const int gc = K * CP + k;
        A[i * CP_LD + k] = gc < nb ? R[(long long)(I * CP + i) * nb + gc] : cmake(0.0, 0.0);
      }
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < CP; ++k) {
        const cplx v = cmul(A[r * CP_LD + k], Xc[(K * CP + k) * 8 + c]);
        acc.x += v.x; acc.y += v.y;
      }
    }
    __syncthreads();
    Y[r * 8 + c] = acc;
    for (int e = tid; e < CP * CP; e += 256) {
      const int i = e >> 5, k = e & 31;
      A[i * CP_LD + k] = Rinv[(long long)(I * CP + i) * nb + I * CP + k];  // I < J: full block
    }
    __syncthreads();
    cplx out = cmake(0.0, 0.0);
#pragma unroll 8
    for (int k = 0; k < CP; ++k) {
      const cplx v = cmul(A[r * CP_LD + k], Y[k * 8 + c]);
      out.x -= v.x; out.y -= v.y;
    }
    Xc[(I * CP + r) * 8 + c] = out;
    if (col_ok) Rinv[(long long)(I * CP + r) * nb + col] = out;
  }
}


// ---------------------------------------------------------------------------------------
// Blocked inverse of an upper-triangular R by recursive doubling (large matrices, few of them):
//   [[A, B], [0, C]]^-1 = [[A^-1, -A^-1 B C^-1], [0, C^-1]]
// level 0 inverts the 32 x 32 diagonal blocks (k_tri_inv_diag256); level l joins neighbouring
// groups of g = 2^(l-1) blocks: T = B C^-1, then X_B = -A^-1 T, every 32 x 32 tile of every group a
// CTA of its own -- 1 + 2 ceil(log2 nblk) launches with no serial block walk inside a CTA
// (k_tri_inv_offdiag: one CTA per 8 columns walking up to 6 block rows, 120 us at nb = 208).
// T lives in the strictly upper part of the scratch matrix `Tm` (the Cholesky work matrix keeps L
// in its lower part; its upper part is free).
// grid: (ceil(nb / 32), nsk), block 256: thread (column j, part) of one diagonal block
__global__ void __launch_bounds__(256)
k_tri_inv_diag256(const cplx* __restrict__ R, int nb, cplx* __restrict__ Rinv,
                  const int* __restrict__ skip) {
  if (skip && skip[blockIdx.y]) return;
  __shared__ cplx U[CP * CP_LD], X[CP * CP_LD], P[8 * CP];
  const long long nn = (long long)nb * nb;
  R += blockIdx.y * nn;
  Rinv += blockIdx.y * nn;
  const int b0 = blockIdx.x * CP, bw = min(CP, nb - b0);
  const int tid = threadIdx.x, j = tid & 31, part = tid >> 5;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = part + 8 * q;
    U[i * CP_LD + j] = (i < bw && j < bw) ? R[(long long)(b0 + i) * nb + b0 + j] : cmake(i == j ? 1.0 : 0.0, 0.0);
    X[i * CP_LD + j] = cmake(0.0, 0.0);
  }
  __syncthreads();
  // back substitution, rows bottom up: X[i][j] = (delta_ij - sum_{k = i+1}^{j} U[i][k] X[k][j]) / U[i][i]
  for (int i = CP - 1; i >= 0; --i) {
    double sx = 0.0, sy = 0.0;
    if (j > i)
      for (int k = i + 1 + part; k <= j; k += 8) {
        const cplx v = cmul(U[i * CP_LD + k], X[k * CP_LD + j]);
        sx += v.x;
        sy += v.y;
      }
    P[part * CP + j] = cmake(sx, sy);
    __syncthreads();
    if (part == 0 && j >= i) {
      double tx = j == i ? 1.0 : 0.0, ty = 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        tx -= P[q * CP + j].x;
        ty -= P[q * CP + j].y;
      }
      const cplx d = U[i * CP_LD + i];
      const double n2 = d.x * d.x + d.y * d.y;
      X[i * CP_LD + j] = cmake((tx * d.x + ty * d.y) / n2, (ty * d.x - tx * d.y) / n2);  // t / d
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = part + 8 * q;
    if (i < bw && j < bw) Rinv[(long long)(b0 + i) * nb + b0 + j] = X[i * CP_LD + j];
  }
}

// One product of a doubling level, a 32 x 32 output tile per CTA.  Group G joins the block ranges
// L = [2 g G, 2 g G + g) and Rr = [2 g G + g, min(2 g G + 2 g, nblk)).
//   phase 0:  T[L][Rr]   =  R[L][Rr] X[Rr][Rr]        (k over Rr)
//   phase 1:  X[L][Rr]   = -X[L][L]  T[L][Rr]         (k over L)
// grid: (g * g, ngroups, nsk): tile (bi, bj) = (blockIdx.x / g, blockIdx.x % g) of the group
__global__ void __launch_bounds__(256)
k_tri_inv_level(const cplx* __restrict__ R, cplx* __restrict__ X, cplx* __restrict__ Tm, int nb,
                int nblk, int g, int phase, const int* __restrict__ skip) {
  if (skip && skip[blockIdx.z]) return;
  __shared__ cplx As[CP * CP_LD], Bs[CP * CP_LD];
  const long long off = (long long)blockIdx.z * nb * nb;
  const int lo = 2 * g * blockIdx.y, mid = lo + g, hi = min(lo + 2 * g, nblk);
  const int bi = lo + blockIdx.x / g, bj = mid + blockIdx.x % g;
  if (mid >= nblk || bj >= hi) return;
  const cplx* A = (phase == 0 ? R : X) + off;
  const cplx* B = (phase == 0 ? X : Tm) + off;
  cplx* C = (phase == 0 ? Tm : X) + off;
  const int k_lo = phase == 0 ? mid : lo, k_hi = phase == 0 ? hi : mid;
  const int tid = threadIdx.x, c = tid & 31, rb = tid >> 5;
  const int i0 = bi * CP, j0 = bj * CP;
  cplx acc[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) acc[q] = cmake(0.0, 0.0);
  for (int K = k_lo; K < k_hi; ++K) {
    const int k0 = K * CP;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = rb + 8 * q;
      As[r * CP_LD + c] = (i0 + r < nb && k0 + c < nb) ? A[(long long)(i0 + r) * nb + k0 + c] : cmake(0.0, 0.0);
      Bs[r * CP_LD + c] = (k0 + r < nb && j0 + c < nb) ? B[(long long)(k0 + r) * nb + j0 + c] : cmake(0.0, 0.0);
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < CP; ++k) {
      const cplx b = Bs[k * CP_LD + c];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const cplx u = cmul(As[(rb + 8 * q) * CP_LD + k], b);
        acc[q].x += u.x;
        acc[q].y += u.y;
      }
    }
  }
  const double sgn = phase == 0 ? 1.0 : -1.0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = i0 + rb + 8 * q, j = j0 + c;
    if (i < nb && j < nb) C[(long long)i * nb + j] = cmake(sgn * acc[q].x, sgn * acc[q].y);
  }
}

// Second Cholesky-QR pass: S = Q1^H Q1 = I + E with |E| ~ kappa(W)^2 eps.  For |E|_max < tol the
// factor is written in closed form, R = I + U, R^-1 = I - U + U^2 - ... with U = up(E) + diag(E)/2:
// R^H R = I + E + U^H U, so the neglected terms are O(nb |E|^2) < 1e-17 for tol = 1e-10, below
// the rounding of the factorisation it replaces (one launch instead of the 9 of the panel
// Cholesky + blocked inverse at nb = 208).  skip[sk] tells the regular kernels to stand down.
// grid: (nsk), block 256
__global__ void __launch_bounds__(256)
k_near_identity(const cplx* __restrict__ S, int nb, double tol, cplx* __restrict__ Rt,
                cplx* __restrict__ Rit, int* __restrict__ skip) {
  __shared__ double red[8];
  __shared__ int ok;
  const long long nn = (long long)nb * nb;
  S += blockIdx.x * nn;
  Rt += blockIdx.x * nn;
  Rit += blockIdx.x * nn;
  double m = 0.0;
  for (long long e = threadIdx.x; e < nn; e += 256) {
    const int i = (int)(e / nb), j = (int)(e % nb);
    const cplx v = S[e];
    m = fmax(m, fmax(fabs(v.x - (i == j ? 1.0 : 0.0)), fabs(v.y)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t = fmax(t, red[w]);
    ok = (t < tol) ? 1 : 0;   // NaN compares false: the regular path reports it
    skip[blockIdx.x] = ok;
  }
  __syncthreads();
  if (!ok) return;
  for (long long e = threadIdx.x; e < nn; e += 256) {
    const int i = (int)(e / nb), j = (int)(e % nb);
    cplx r = cmake(0.0, 0.0), ri = cmake(0.0, 0.0);
    if (j > i) {
      const cplx u = S[e];
      r = u;
      ri = cmake(-u.x, -u.y);
    } else if (j == i) {
      const double u = 0.5 * (S[e].x - 1.0);
      r = cmake(1.0 + u, 0.0);
      ri = cmake(1.0 - u + u * u, 0.0);
    }
    Rt[e] = r;
    Rit[e] = ri;
  }
}

// Second Cholesky-QR pass: r = R2 r_prev, rinv = rinv_prev R2inv (all upper triangular).
// Outputs must not alias the inputs.   grid: (ceil(nb^2 / 256), nsk)
__global__ void __launch_bounds__(SMALL_T)
k_tri_compose(const cplx* __restrict__ R2, const cplx* __restrict__ R2inv,
              const cplx* __restrict__ r_prev, const cplx* __restrict__ rinv_prev, int nb,
              cplx* __restrict__ r_out, cplx* __restrict__ rinv_out) {
  const long long nn = (long long)nb * nb;
  const long long off = blockIdx.y * nn;
  const long long e = (long long)blockIdx.x * SMALL_T + threadIdx.x;
  if (e >= nn) return;
  const int i = (int)(e / nb), j = (int)(e % nb);
  cplx a = cmake(0.0, 0.0), b = cmake(0.0, 0.0);
  if (j >= i) {
    for (int k = i; k <= j; ++k) {
      const cplx u = cmul(R2[off + (long long)i * nb + k], r_prev[off + (long long)k * nb + j]);
      a.x += u.x; a.y += u.y;
      const cplx v = cmul(rinv_prev[off + (long long)i * nb + k], R2inv[off + (long long)k * nb + j]);
      b.x += v.x; b.y += v.y;
    }
  }
  r_out[off + e] = a;
  rinv_out[off + e] = b;
}

// ---------------------------------------------------------------------------------------
// Cholesky factor AND triangular inverse of one Hermitian nb x nb matrix per CTA (or per cluster of
// 8 CTAs) in ONE launch and ONE sweep of nb - 1 steps.  Replaces the chains k_chol_blocked +
// k_tri_inv_cols (one CTA per matrix) and k_chol_panel x7 + k_tri_inv_diag + k_tri_inv_offdiag
// (nb = 208: 16 dependent launches, 0.6 ms of the 5.1 ms diamond-64 evaluation), whose latency
// does not shrink with the number of k-points a rank holds.
//
// Measured on B200 (tools/lat_probe.cu: DFMA 10, LDS ~30, 512-thread barrier 44 cycles; a lone warp
// issues one dependent instruction per ~4.5 cycles): a factorisation is bound by the chain of nb
// pivots and by the instructions PER THREAD between two barriers, not by flops.  So:
//   * S = L~ D L~^H by unscaled right-looking elimination with the lower-triangle slots (r, c)
//     DISTRIBUTED OVER THE REGISTERS of all threads (slot e -> thread e mod T: the slots a step
//     finishes are spread one per thread); L = L~ D^1/2.
//   * the inverse rides on the same steps: W = L~^-1 is the same row operations applied to the
//     identity, and slot (r, c) is needed for the factor only while j < c and for W only while
//     c <= j < r -- so ONE register per slot serves both, one update per slot and step, no second
//     phase:  L^-1 = D^-1/2 W.
//   * one published vector per step makes the update uniform: U_j[i] = a_ij (i > j),
//     U_j[j] = 1, U_j[i] = conj(W[j][i]) (i < j):   v(r, c) -= U_j[r] conj(U_j[c]) / d_j  for every
//     slot with r > j, whatever it currently holds.  Finished slots keep being updated (their
//     results were saved when they were published): no liveness branches, the slots of a thread
//     interleave.  U lives in shared memory, two interleaved buffers (the parity is an immediate
//     offset of the step unrolled by two); with a cluster every CTA holds a copy, written through
//     distributed shared memory by the slot's owner, and the barrier is the cluster barrier.
//   * 1 / d_j = rcp.approx + two Newton steps per thread, under the latency of its loads.
// PASS2 (single CTA): Q1^H Q1 = I + E, closed form for max|E| < tol (see k_near_identity).
// Outputs (row major, zeros below the diagonal): Rt = L^H, Rit = Rt^-1.
// grid: (nsk * NCTA) in clusters of NCTA, block CF_T; dynamic smem: cf_smem_bytes(nb)
namespace cg = cooperative_groups;
constexpr int CF_T = 512;
__device__ long long g_fs_clocks[8];      // JRB_FS_TIMING: phase boundaries of CTA 0 (tuning aid)
static int cf_smem_bytes(int nb) {
  return 2 * nb * (int)sizeof(cplx) + (nb + 2) * (int)sizeof(double);  // U[nb][2], pivots, piv[2]
}

__device__ __forceinline__ double fast_rcp(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double t = fma(-d, x, 1.0);
  x = fma(x, t, x);
  t = fma(-d, x, 1.0);
  return fma(x, t, x);
}

// store to the same shared-memory location of every CTA of the cluster (distributed shared
// memory: mapa + st.shared::cluster on 32-bit shared addresses, emitted where the value is ready)
template <int NCTA>
__device__ __forceinline__ void cf_publish(cplx* local, cplx val) {
  if constexpr (NCTA > 1) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(local);
#pragma unroll
    for (int k = 0; k < NCTA; ++k)
      asm volatile(
        "{ .reg .u32 ra; mapa.shared::cluster.u32 ra, %0, %1;\n"
        "  st.shared::cluster.v2.f64 [ra], {%2, %3}; }\n" ::"r"(a), "r"(k), "d"(val.x), "d"(val.y)
        : "memory");
  } else {
    *local = val;
  }
}
template <int NCTA>
__device__ __forceinline__ void cf_publish(double* local, double val) {
  if constexpr (NCTA > 1) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(local);
#pragma unroll
    for (int k = 0; k < NCTA; ++k)
      asm volatile(
        "{ .reg .u32 ra; mapa.shared::cluster.u32 ra, %0, %1;\n"
        "  st.shared::cluster.f64 [ra], %2; }\n" ::"r"(a), "r"(k), "d"(val)
        : "memory");
  } else {
    *local = val;
  }
}

// one step: U_j at parity PAR is complete; updates every slot and publishes U_{j+1} at PAR ^ 1.
// ua / ub: byte offsets of U[er] / U[ec] (16 bytes per index) within one parity array of pitch
// `pitch` bytes; a slot's finished factor and inverse entries stay in registers (af, wf).
template <int EPT, int NCTA, int PAR>
__device__ __forceinline__ void cf_step(int j, int nb, int pitch, cplx (&v)[EPT], cplx (&af)[EPT],
                                        cplx (&wf)[EPT], const int (&ua)[EPT], const int (&ub)[EPT],
                                        unsigned char* Ub, double* dvec) {
  const double djj = dvec[nb + PAR];
  const double rinv = fast_rcp(djj > 0.0 ? djj : 1.0);
  const int jn = 16 * (j + 1);
  const unsigned char* Uc = Ub + PAR * pitch;
  unsigned char* Un = Ub + (PAR ^ 1) * pitch;
#pragma unroll
  for (int q = 0; q < EPT; ++q) {
    const cplx a = *reinterpret_cast<const cplx*>(Uc + ua[q]);
    const cplx b = *reinterpret_cast<const cplx*>(Uc + ub[q]);
    const double ax = a.x * rinv, ay = a.y * rinv;
    v[q].x -= ax * b.x + ay * b.y;     // (a rinv) conj(b)
    v[q].y -= ay * b.x - ax * b.y;
  }
#pragma unroll
  for (int q = 0; q < EPT; ++q) {
    if (ub[q] == jn) {               // column j + 1 of the factor is final (r >= j + 1)
      if (ua[q] == jn) {             // its pivot
        cf_publish<NCTA>(dvec + j + 1, v[q].x);
        cf_publish<NCTA>(dvec + nb + (PAR ^ 1), v[q].x);
        cf_publish<NCTA>(reinterpret_cast<cplx*>(Un + jn), cmake(1.0, 0.0));
      } else {
        cf_publish<NCTA>(reinterpret_cast<cplx*>(Un + ua[q]), v[q]);
        af[q] = v[q];                // a_rc, c = j + 1
        v[q] = cmake(0.0, 0.0);      // the slot now accumulates W[r][c]
      }
    } else if (ua[q] == jn) {        // row j + 1 of W is final (c <= j)
      wf[q] = v[q];                  // W[r][c], r = j + 1
      cf_publish<NCTA>(reinterpret_cast<cplx*>(Un + ub[q]), cconj(v[q]));
    }
  }
}

template <int EPT, int NCTA, bool PASS2>
__global__ void __launch_bounds__(CF_T, 1)
k_factor_stream(const cplx* __restrict__ S, int nb, double tol, cplx* __restrict__ Rt,
                cplx* __restrict__ Rit, int* __restrict__ fail_flag, const int* __restrict__ skip,
                int timing) {
  extern __shared__ __align__(16) unsigned char smem_raw_[];
  cplx* U = reinterpret_cast<cplx*>(smem_raw_);               // [2][nb]: U_j by parity of j
  double* dvec = reinterpret_cast<double*>(U + 2 * nb);       // [nb] pivots d_i, then [2] d_j by parity
  const int pitch = nb * (int)sizeof(cplx);
  const int mat = blockIdx.x / NCTA;
  if (skip && skip[mat]) return;  // factor already written by k_near_identity (cluster-uniform)
  const long long nn = (long long)nb * nb;
  const long long off = mat * nn;
  S += off;
  Rt += off;
  Rit += off;
  const int tid = threadIdx.x;
  const int ntri = nb * (nb + 1) / 2;
  auto tick = [&](int slot) {
    if (timing && blockIdx.x == 0 && threadIdx.x == 0) g_fs_clocks[slot] = clock64();
  };
  tick(0);

  if (PASS2) {
    static_assert(!PASS2 || NCTA == 1, "the fused second pass is a single-CTA variant");
    __shared__ double red[CF_T / 32];
    __shared__ int near_id;
    double m = 0.0;
    for (int e = tid; e < nb * nb; e += CF_T) {
      const int i = e / nb, j = e - i * nb;
      const cplx x = S[e];
      m = fmax(m, fmax(fabs(x.x - (i == j ? 1.0 : 0.0)), fabs(x.y)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0) red[tid >> 5] = m;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < CF_T / 32; ++w) t = fmax(t, red[w]);
      near_id = (t < tol) ? 1 : 0;   // NaN compares false: the factorisation reports it
    }
    __syncthreads();
    if (near_id) {
      // R2 = I + U, R2^-1 = I - U (+ U^2 on the diagonal), U = up(E) + diag(E) / 2
      for (int e = tid; e < nb * nb; e += CF_T) {
        const int i = e / nb, j = e - i * nb;
        cplx r = cmake(0.0, 0.0), ri = cmake(0.0, 0.0);
        if (j > i) {
          const cplx u = S[e];
          r = u;
          ri = cmake(-u.x, -u.y);
        } else if (j == i) {
          const double u = 0.5 * (S[e].x - 1.0);
          r = cmake(1.0 + u, 0.0);
          ri = cmake(1.0 - u + u * u, 0.0);
        }
        Rt[e] = r;
        Rit[e] = ri;
      }
      tick(1);
      tick(2);
      tick(3);
      return;
    }
  }
  tick(1);

  int rank = 0;
  if constexpr (NCTA > 1) {
    rank = (int)cg::this_cluster().block_rank();
    cg::this_cluster().sync();  // every CTA of the cluster is running before remote stores
  }
  auto sync = [&]() {
    if constexpr (NCTA > 1) cg::this_cluster().sync();
    else __syncthreads();
  };

  // slots, row-major lower triangle e = r (r + 1) / 2 + c: a WARP owns 32 EPT consecutive slots (a
  // few neighbouring rows, so it retires as a whole once its last row is done: rows finish top
  // down), lane l its slots l, l + 32, ... (consecutive lanes read consecutive U entries)
  const int gw = (rank * CF_T + tid) >> 5, lane = tid & 31;
  unsigned char* Ub = smem_raw_;
  cplx v[EPT], af[EPT], wf[EPT];
  int ua[EPT], ub[EPT];
  int rmax = 0;  // last row of this warp
#pragma unroll
  for (int q = 0; q < EPT; ++q) {
    const int e = (gw * EPT + q) * 32 + lane;
    // absent slots alias (0, 0), which publishes at set-up only (never equals j + 1)
    ua[q] = 0;
    ub[q] = 0;
    v[q] = af[q] = wf[q] = cmake(0.0, 0.0);
    if (e < ntri) {
      int r = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
      while (r * (r + 1) / 2 > e) --r;
      while ((r + 1) * (r + 2) / 2 <= e) ++r;
      const int c = e - r * (r + 1) / 2;
      ua[q] = 16 * r;
      ub[q] = 16 * c;
      rmax = r;
      v[q] = S[(long long)r * nb + c];
      if (r == c) v[q].y = 0.0;
      if (c == 0) {  // column 0 is final as it stands: U_0
        if (r == 0) {
          cf_publish<NCTA>(dvec, v[q].x);
          cf_publish<NCTA>(dvec + nb, v[q].x);
          cf_publish<NCTA>(U, cmake(1.0, 0.0));
        } else {
          cf_publish<NCTA>(U + r, v[q]);
          af[q] = v[q];
          v[q] = cmake(0.0, 0.0);
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rmax = max(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
  sync();
  for (int j = 0; j < nb - 1; j += 2) {
    if (rmax > j) cf_step<EPT, NCTA, 0>(j, nb, pitch, v, af, wf, ua, ub, Ub, dvec);
    sync();
    if (j + 1 < nb - 1) {
      if (rmax > j + 1) cf_step<EPT, NCTA, 1>(j + 1, nb, pitch, v, af, wf, ua, ub, Ub, dvec);
      sync();
    }
  }
  tick(2);
  // L = L~ D^1/2, L^-1 = D^-1/2 W:  Rt[c][r] = conj(L[r][c]) = conj(a_rc) / sqrt(d_c),
  // Rit[c][r] = conj(X[r][c]) = conj(W[r][c]) / sqrt(d_r)
#pragma unroll
  for (int q = 0; q < EPT; ++q) {
    const int e = (gw * EPT + q) * 32 + lane;
    if (e < ntri) {
      const int r = ua[q] >> 4, c = ub[q] >> 4;
      const double dc = dvec[c];
      if (r == c) {
        if (!(dc > 0.0)) atomicExch(fail_flag, 1);
        const double l = sqrt(dc > 0.0 ? dc : 1.0);
        Rt[(long long)r * nb + r] = cmake(l, 0.0);
        Rit[(long long)r * nb + r] = cmake(1.0 / l, 0.0);
      } else {
        const double dr = dvec[r];
        const double sc = rsqrt(dc > 0.0 ? dc : 1.0), sr = rsqrt(dr > 0.0 ? dr : 1.0);
        Rt[(long long)c * nb + r] = cmake(af[q].x * sc, -af[q].y * sc);
        Rit[(long long)c * nb + r] = cmake(wf[q].x * sr, -wf[q].y * sr);
        Rt[(long long)r * nb + c] = cmake(0.0, 0.0);
        Rit[(long long)r * nb + c] = cmake(0.0, 0.0);
      }
    }
  }
  tick(3);
}

// Adjoint, step 1: M = (sum partial) diag(f); X = -(up(M) + up(M)^H + diag Re M);
// T1 = diag(f) Rinv^H.   grid: (ceil(nb^2 / 256), nsk)
__global__ void __launch_bounds__(SMALL_T)
k_bwd_x(const cplx* __restrict__ partial, int nchunks, int nb, const double* __restrict__ occ,
        const cplx* __restrict__ rinv, cplx* __restrict__ X, cplx* __restrict__ T1) {
  const int sk = blockIdx.y, nsk = gridDim.y;
  const long long nn = (long long)nb * nb;
  const long long e = (long long)blockIdx.x * SMALL_T + threadIdx.x;
  if (e >= nn) return;
  const int i = (int)(e / nb), j = (int)(e % nb);
  const double* f = occ ? occ + (long long)sk * nb : nullptr;
  // only the upper triangle of M (written by k_gram) is needed: m = M[min][max]
  const int a = min(i, j), b = max(i, j);
  cplx m = cmake(0.0, 0.0);
  for (int c = 0; c < nchunks; ++c) {
    const cplx v = partial[((long long)c * nsk + sk) * nn + (long long)a * nb + b];
    m.x += v.x; m.y += v.y;
  }
  const double fb = f ? f[b] : 1.0;  // M = Q^H (HQ diag f): column scaling
  m = cmake(m.x * fb, m.y * fb);
  cplx x;
  if (j > i) x = cmake(-m.x, -m.y);
  else if (j < i) x = cmake(-m.x, m.y);
  else x = cmake(-m.x, 0.0);
  X[sk * nn + e] = x;
  const double fi = f ? f[i] : 1.0;
  const cplx r = rinv[sk * nn + (long long)j * nb + i];
  T1[sk * nn + e] = cmake(fi * r.x, -fi * r.y);  // f_i conj(rinv[j][i])
}

// Adjoint, step 2: T2 = X Rinv^H  (rinv upper triangular -> k >= j).  grid as above.
__global__ void __launch_bounds__(SMALL_T)
k_bwd_t2(const cplx* __restrict__ X, const cplx* __restrict__ rinv, int nb, cplx* __restrict__ T2) {
  const long long nn = (long long)nb * nb;
  const long long off = blockIdx.y * nn;
  const long long e = (long long)blockIdx.x * SMALL_T + threadIdx.x;
  if (e >= nn) return;
  const int i = (int)(e / nb), j = (int)(e % nb);
  cplx s = cmake(0.0, 0.0);
  for (int k = j; k < nb; ++k) {
    const cplx v = cmulc(X[off + (long long)i * nb + k], rinv[off + (long long)j * nb + k]);
    s.x += v.x; s.y += v.y;
  }
  T2[off + e] = s;
}

// ---------------------------------------------------------------------------------------
// nb x nb products of the small algebra as 32 x 32 tiles, one CTA each (large nb, few matrices: the
// one-thread-per-element kernels k_tri_compose / k_bwd_t2 walk 200 strided L2 reads per thread).
//   MODE 0: C = A B, A and B upper triangular (blocks K in [bi, bj]; C's lower blocks are zeroed)
//   MODE 1: C = A B^H, B upper triangular (B^H[k][j] = conj(B[j][k]): blocks K >= bj)
// blockIdx.z = which * nsk + sk: `which` selects the operand set (two products per launch).
struct TileGemmOps {
  const cplx* A[2];
  const cplx* B[2];
  cplx* C[2];
};
template <int MODE>
__global__ void __launch_bounds__(256)
k_tile_gemm(TileGemmOps ops, int nb, int nsk) {
  __shared__ cplx As[CP * CP_LD], Bs[CP * CP_LD];
  const int which = blockIdx.z / nsk, sk = blockIdx.z % nsk;
  const long long off = (long long)sk * nb * nb;
  const cplx* A = ops.A[which] + off;
  const cplx* B = ops.B[which] + off;
  cplx* C = ops.C[which] + off;
  const int nblk = (nb + CP - 1) / CP;
  const int bi = blockIdx.y, bj = blockIdx.x;
  const int tid = threadIdx.x, c = tid & 31, rb = tid >> 5;
  const int i0 = bi * CP, j0 = bj * CP;
  cplx acc[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) acc[q] = cmake(0.0, 0.0);
  const int k_lo = MODE == 0 ? bi : bj, k_hi = MODE == 0 ? bj + 1 : nblk;
  for (int K = k_lo; K < k_hi; ++K) {
    const int k0 = K * CP;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = rb + 8 * q;
      As[r * CP_LD + c] = (i0 + r < nb && k0 + c < nb) ? A[(long long)(i0 + r) * nb + k0 + c] : cmake(0.0, 0.0);
      if (MODE == 0) {
        Bs[r * CP_LD + c] = (k0 + r < nb && j0 + c < nb) ? B[(long long)(k0 + r) * nb + j0 + c] : cmake(0.0, 0.0);
      } else {  // Bs[k][j] = conj(B[j0 + j][k0 + k]): read row-wise, store transposed
        const cplx v = (j0 + r < nb && k0 + c < nb) ? B[(long long)(j0 + r) * nb + k0 + c] : cmake(0.0, 0.0);
        Bs[c * CP_LD + r] = cconj(v);
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < CP; ++k) {
      const cplx b = Bs[k * CP_LD + c];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const cplx u = cmul(As[(rb + 8 * q) * CP_LD + k], b);
        acc[q].x += u.x;
        acc[q].y += u.y;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = i0 + rb + 8 * q, j = j0 + c;
    if (i < nb && j < nb) C[(long long)i * nb + j] = acc[q];
  }
}

// k_near_identity in two grid-wide kernels (one CTA scanning a 208 x 208 matrix took 143 us):
// max|S - I| by blocks into emax[sk] (non-negative doubles order like their bit patterns), then
// the closed form and the skip flag.   grid: (ceil(nb^2 / 256), nsk), block 256
__global__ void __launch_bounds__(256)
k_near_identity_max(const cplx* __restrict__ S, int nb, unsigned long long* __restrict__ emax) {
  __shared__ double red[8];
  const long long nn = (long long)nb * nb;
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  double m = 0.0;
  if (e < nn) {
    const int i = (int)(e / nb), j = (int)(e % nb);
    const cplx v = S[blockIdx.y * nn + e];
    m = fmax(fabs(v.x - (i == j ? 1.0 : 0.0)), fabs(v.y));
    if (!(m == m)) m = 1e300;  // NaN: not near the identity
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t = fmax(t, red[w]);
    atomicMax(emax + blockIdx.y, (unsigned long long)__double_as_longlong(t));
  }
}
__global__ void __launch_bounds__(256)
k_near_identity_apply(const cplx* __restrict__ S, int nb, double tol,
                      const unsigned long long* __restrict__ emax, cplx* __restrict__ Rt,
                      cplx* __restrict__ Rit, int* __restrict__ skip) {
  const bool ok = __longlong_as_double((long long)emax[blockIdx.y]) < tol;
  if (blockIdx.x == 0 && threadIdx.x == 0) skip[blockIdx.y] = ok ? 1 : 0;
  if (!ok) return;
  const long long nn = (long long)nb * nb;
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  if (e >= nn) return;
  const int i = (int)(e / nb), j = (int)(e % nb);
  cplx r = cmake(0.0, 0.0), ri = cmake(0.0, 0.0);
  const cplx u0 = S[blockIdx.y * nn + e];
  if (j > i) {
    r = u0;
    ri = cmake(-u0.x, -u0.y);
  } else if (j == i) {
    const double u = 0.5 * (u0.x - 1.0);
    r = cmake(1.0 + u, 0.0);
    ri = cmake(1.0 - u + u * u, 0.0);
  }
  Rt[blockIdx.y * nn + e] = r;
  Rit[blockIdx.y * nn + e] = ri;
}

// ---------------------------------------------------------------------------------------
static int gram_chunks(const jrb_plan* p, int tiles, int nsk) {
  // CTAs = upper super-tile pairs x chunks x (spin,k); two CTAs are resident per SM.  Pick the
  // chunk count (>= 256 rows each) whose last wave is fullest, preferring >= 2 waves.
  const int per_chunk = nsk * (tiles * (tiles + 1) / 2);
  const int slots = (tiles == 1 && p->nb <= 32 ? 3 : 2) * 148;  // resident CTAs of the variant used
  const int max_chunks = (int)std::max<int64_t>(1, std::min<int64_t>(128, p->ng / 256));
  int best = 1;
  double best_score = -1.0;
  for (int c = 1; c <= max_chunks; ++c) {
    const double waves = (double)per_chunk * c / slots;
    double eff = waves / std::ceil(waves);
    if (waves < 2.0) eff *= 0.5 + 0.25 * waves;      // too few CTAs to hide the prologues
    if (waves > 6.0) eff *= 6.0 / waves;             // shorter k-loops cost more per row
    if (eff > best_score + 1e-9) {
      best_score = eff;
      best = c;
    }
  }
  return best;
}

// upper bound of chunks x nsk for sizing the partial buffer (any sub-range of (spin,k))
int qr_gram_partial_mats(const jrb_plan* p) {
  const int tiles = (p->nb + QT - 1) / QT;
  int worst = 1;
  for (int n = 1; n <= p->ns * p->nk; ++n) worst = std::max(worst, n * gram_chunks(p, tiles, n));
  return worst;
}

template <class K>
static int opt_in_smem(K kernel, int bytes) {
  if (bytes > 48 * 1024)
    JRB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return 0;
}

// partial = upper blocks of A^H B, split over row chunks
static int run_gram(jrb_plan* p, int nsk, TallMat A, TallMat B, bool same, cplx* partial,
                    int* nchunks_out, cudaStream_t st) {
  constexpr int ST = 3;
  const int tiles = (p->nb + QT - 1) / QT;
  const int nchunks = gram_chunks(p, tiles, nsk);
  long long rows = (p->ng + nchunks - 1) / nchunks;
  rows = (rows + QK - 1) / QK * QK;
  dim3 grid(tiles * (tiles + 1) / 2, nchunks, nsk);
  const long long sks = p->ng * p->nb;
  const int panels = (same && tiles == 1) ? 1 : 2;
  const int smem = panels * ST * QK * QLDB * (int)sizeof(cplx);
  if (tiles == 1 && p->nb <= 32) {
    // few bands: 10 blocks at most (2 per warp), narrow panels, six-stage ring
    constexpr int STS = 6;
    const int ldb = 8 * ((p->nb + 7) / 8) + 2;
    const int smem_s = panels * STS * QK * ldb * (int)sizeof(cplx);
    static int once = opt_in_smem(k_gram<STS, 2, 0>, 2 * STS * QK * 34 * (int)sizeof(cplx));
    if (once) return once;
    k_gram<STS, 2, 0><<<grid, QTHREADS, smem_s, st>>>(A, B, same ? 1 : 0, p->ng, p->nb, sks, tiles,
                                                     rows, partial);
  } else if (tiles == 1) {
    static int once = opt_in_smem(k_gram<ST, QSLOT_DIAG>, 2 * ST * QK * QLDB * (int)sizeof(cplx));
    if (once) return once;
    k_gram<ST, QSLOT_DIAG><<<grid, QTHREADS, smem, st>>>(A, B, same ? 1 : 0, p->ng, p->nb, sks,
                                                        tiles, rows, partial);
  } else {
    static int once = opt_in_smem(k_gram<ST, QMAXSLOT>, 2 * ST * QK * QLDB * (int)sizeof(cplx));
    if (once) return once;
    k_gram<ST, QMAXSLOT><<<grid, QTHREADS, smem, st>>>(A, B, same ? 1 : 0, p->ng, p->nb, sks,
                                                      tiles, rows, partial);
  }
  JRB_CHECK_LAUNCH("k_gram");
  *nchunks_out = nchunks;
  return 0;
}

template <int MODE, int NCB, bool SPLIT>
static int run_apply_ncb(jrb_plan* p, int nsk, TallMat in1, const cplx* t1, int tri1, TallMat in2,
                         const cplx* t2, int tri2, int nterms, double* out_a, double* out_b,
                         cudaStream_t st) {
  constexpr int ST = 2;
  const int smem = ST * (QROWS * QLDA + QK * (8 * NCB + 2)) * (int)sizeof(cplx);
  static int once = opt_in_smem(k_apply<MODE, NCB, ST, SPLIT>, smem);
  if (once) return once;
  dim3 grid((unsigned)((p->ng + QROWS - 1) / QROWS), (p->nb + 8 * NCB - 1) / (8 * NCB), nsk);
  k_apply<MODE, NCB, ST, SPLIT><<<grid, QTHREADS, smem, st>>>(
    in1.re, in1.im, t1, tri1, reinterpret_cast<const cplx*>(in2.re), t2, tri2, nterms, p->ng,
    p->nb, p->ng * p->nb, out_a, out_b);
  JRB_CHECK_LAUNCH("k_apply");
  return 0;
}

// column-tile width: 32 columns for few bands, else 72 (blocks past nb cost no tensor work)
template <int MODE>
static int run_apply(jrb_plan* p, int nsk, TallMat in1, const cplx* t1, int tri1, TallMat in2,
                     const cplx* t2, int tri2, int nterms, double* out_a, double* out_b,
                     cudaStream_t st) {
  const int nb = p->nb;
  const bool split = in1.im != nullptr;
  if (MODE == 0 && split) {
    if (nb <= 32)
      return run_apply_ncb<0, 4, true>(p, nsk, in1, t1, tri1, in2, t2, tri2, nterms, out_a, out_b, st);
    return run_apply_ncb<0, 9, true>(p, nsk, in1, t1, tri1, in2, t2, tri2, nterms, out_a, out_b, st);
  }
  if (nb <= 32)
    return run_apply_ncb<MODE, 4, false>(p, nsk, in1, t1, tri1, in2, t2, tri2, nterms, out_a, out_b, st);
  return run_apply_ncb<MODE, 9, false>(p, nsk, in1, t1, tri1, in2, t2, tri2, nterms, out_a, out_b, st);
}

// Rectangular products on the same DMMA kernels (the projector products of the non-local
// pseudopotential, nonlocal.cu):
//   gram : partial[chunk][sk][i][j] = sum_{g in chunk} conj(At[sk % a_mod][g][i]) B[sk][g][j]
//   apply: out[sk][g][j] += sum_i In[sk % in_mod][g][i] T[sk][i][j]
int launch_gram_rect(jrb_plan* p, int nsk, const cplx* At, int nbA, int a_mod, const cplx* B,
                     int nchunks, cplx* partial, cudaStream_t st) {
  constexpr int ST = 3;
  const int ti = (nbA + QT - 1) / QT, tj = (p->nb + QT - 1) / QT;
  long long rows = (p->ng + nchunks - 1) / nchunks;
  rows = (rows + QK - 1) / QK * QK;
  dim3 grid(ti * tj, nchunks, nsk);
  const int smem = 2 * ST * QK * QLDB * (int)sizeof(cplx);
  TallMat A{reinterpret_cast<const double*>(At), nullptr, nbA};
  TallMat Bm{reinterpret_cast<const double*>(B), nullptr, p->nb};
  const GramRect rect{p->nb, (long long)p->ng * nbA, a_mod};
  const int blocks = ((std::min(nbA, QT) + 7) / 8) * ((std::min(p->nb, QT) + 7) / 8);
  if (blocks <= 8 * QSLOT_DIAG) {
    static int once = opt_in_smem(k_gram<ST, QSLOT_DIAG>, 2 * ST * QK * QLDB * (int)sizeof(cplx));
    if (once) return once;
    k_gram<ST, QSLOT_DIAG><<<grid, QTHREADS, smem, st>>>(A, Bm, 0, p->ng, nbA, p->ng * p->nb, 0, rows,
                                                        partial, rect);
  } else {
    static int once = opt_in_smem(k_gram<ST, QMAXSLOT>, 2 * ST * QK * QLDB * (int)sizeof(cplx));
    if (once) return once;
    k_gram<ST, QMAXSLOT><<<grid, QTHREADS, smem, st>>>(A, Bm, 0, p->ng, nbA, p->ng * p->nb, 0, rows,
                                                      partial, rect);
  }
  JRB_CHECK_LAUNCH("k_gram (rectangular)");
  return 0;
}

template <int NCB>
static int run_apply_rect(jrb_plan* p, int nsk, const cplx* In, int kdim, int in_mod, const cplx* T,
                          cplx* out, cudaStream_t st) {
  constexpr int ST = 2;
  const int smem = ST * (QROWS * QLDA + QK * (8 * NCB + 2)) * (int)sizeof(cplx);
  static int once = opt_in_smem(k_apply<0, NCB, ST, false>, smem);
  if (once) return once;
  dim3 grid((unsigned)((p->ng + QROWS - 1) / QROWS), (p->nb + 8 * NCB - 1) / (8 * NCB), nsk);
  const ApplyRect rect{kdim, (long long)p->ng * kdim, in_mod};
  k_apply<0, NCB, ST, false><<<grid, QTHREADS, smem, st>>>(
    reinterpret_cast<const double*>(In), nullptr, T, TRI_FULL, nullptr, nullptr, TRI_FULL, 1, p->ng,
    p->nb, p->ng * p->nb, reinterpret_cast<double*>(out), nullptr, rect);
  JRB_CHECK_LAUNCH("k_apply (rectangular)");
  return 0;
}

int launch_apply_rect(jrb_plan* p, int nsk, const cplx* In, int kdim, int in_mod, const cplx* T,
                      cplx* out, cudaStream_t st) {
  if (p->nb <= 32) return run_apply_rect<4>(p, nsk, In, kdim, in_mod, T, out, st);
  return run_apply_rect<9>(p, nsk, In, kdim, in_mod, T, out, st);
}

// H_ij = <q_i | hq_j> for Hermitian H (hamiltonian.hamiltonian_matrix): one Gram on DMMA
int launch_hamiltonian_matrix(jrb_plan* p, const cplx* q, const cplx* hq, cplx* h, cudaStream_t st) {
  int nchunks = 0, rc = 0;
  const int nsk = p->ns * p->nk;
  TallMat Q{reinterpret_cast<const double*>(q), nullptr, p->nb};
  TallMat G{reinterpret_cast<const double*>(hq), nullptr, p->nb};
  if ((rc = run_gram(p, nsk, Q, G, false, p->d_gpart, &nchunks, st))) return rc;
  dim3 grid((unsigned)(((long long)p->nb * p->nb + SMALL_T - 1) / SMALL_T), nsk);
  k_gram_reduce<<<grid, SMALL_T, 0, st>>>(p->d_gpart, nchunks, p->nb, h);
  JRB_CHECK_LAUNCH("k_gram_reduce");
  return 0;
}

// Configuration of k_factor_stream for nb bands: slots per thread and CTAs per matrix (0 = use
// the multi-launch kernels).  One CTA while the lower triangle fits 6 registers-slots per thread
// (nb <= 77), else a cluster of 8 (nb <= 221).
struct CfConfig {
  int ept, ncta;
};
static CfConfig cf_config(int nb) {
  static int off = [] {
    const char* env = std::getenv("JRB_NO_FUSED_SMALL");
    return env ? std::atoi(env) : 0;
  }();
  const int ntri = nb * (nb + 1) / 2;
  if (off) return {0, 0};
  static int cluster = [] {
    const char* env = std::getenv("JRB_FACTOR_CLUSTER");  // tuning aid: 8-CTA clusters for nb > 77
    return env ? std::atoi(env) : 0;
  }();
  if (ntri <= 2 * CF_T) return {2, 1};
  if (ntri <= 5 * CF_T) return {5, 1};
  if (ntri <= 6 * CF_T) return {6, 1};
  if (cluster && ntri <= 6 * CF_T * 8) return {6, 8};
  return {0, 0};
}

template <int EPT, int NCTA, bool PASS2>
static int run_factor_stream(jrb_plan* p, int nsk, const cplx* S, double tol, cplx* Rt, cplx* Rit,
                             const int* skip, cudaStream_t st) {
  int* fail = reinterpret_cast<int*>(p->d_scal + 32);
  static int timing = std::getenv("JRB_FS_TIMING") ? std::atoi(std::getenv("JRB_FS_TIMING")) : 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nsk * NCTA);
  cfg.blockDim = dim3(CF_T);
  cfg.dynamicSmemBytes = cf_smem_bytes(p->nb);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NCTA;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const int nb = p->nb;
  JRB_CUDA(cudaLaunchKernelEx(&cfg, k_factor_stream<EPT, NCTA, PASS2>, S, nb, tol, Rt, Rit, fail,
                              skip, timing));
  JRB_CHECK_LAUNCH("k_factor_stream");
  if (timing) {
    long long h[8];
    cudaStreamSynchronize(st);
    cudaMemcpyFromSymbol(h, g_fs_clocks, sizeof(h));
    fprintf(stderr, "k_factor_stream<%d,%d,%d> nb %d cycles: scan %lld sweep %lld out %lld total %lld\n",
            EPT, NCTA, (int)PASS2, nb, h[1] - h[0], h[2] - h[1], h[3] - h[2], h[3] - h[0]);
  }
  return 0;
}

// pass2: the single-CTA variants also test for the near-identity closed form (tol > 0)
static int factor_stream(jrb_plan* p, CfConfig c, bool pass2, int nsk, const cplx* S, double tol,
                         cplx* Rt, cplx* Rit, const int* skip, cudaStream_t st) {
  if (c.ncta == 8) return run_factor_stream<6, 8, false>(p, nsk, S, 0.0, Rt, Rit, skip, st);
  if (c.ept == 2)
    return pass2 ? run_factor_stream<2, 1, true>(p, nsk, S, tol, Rt, Rit, skip, st)
                 : run_factor_stream<2, 1, false>(p, nsk, S, 0.0, Rt, Rit, skip, st);
  if (c.ept == 5)
    return pass2 ? run_factor_stream<5, 1, true>(p, nsk, S, tol, Rt, Rit, skip, st)
                 : run_factor_stream<5, 1, false>(p, nsk, S, 0.0, Rt, Rit, skip, st);
  return pass2 ? run_factor_stream<6, 1, true>(p, nsk, S, tol, Rt, Rit, skip, st)
               : run_factor_stream<6, 1, false>(p, nsk, S, 0.0, Rt, Rit, skip, st);
}

// r = R2 R1, rinv = R1^-1 R2^-1 (all upper triangular): both products as 32 x 32 tiles, one launch
static int compose_factors(jrb_plan* p, int nsk, const cplx* R2, const cplx* R2inv, const cplx* R1,
                           const cplx* R1inv, cplx* r_out, cplx* rinv_out, cudaStream_t st) {
  TileGemmOps ops;
  ops.A[0] = R2;    ops.B[0] = R1;    ops.C[0] = r_out;
  ops.A[1] = R1inv; ops.B[1] = R2inv; ops.C[1] = rinv_out;
  const int nblk = (p->nb + CP - 1) / CP;
  k_tile_gemm<0><<<dim3(nblk, nblk, 2 * nsk), 256, 0, st>>>(ops, p->nb, nsk);
  JRB_CHECK_LAUNCH("k_tile_gemm");
  return 0;
}

__global__ void k_zero_unless_skipped(cplx* __restrict__ a, long long n, const int* __restrict__ skip) {
  if (skip && skip[blockIdx.y]) return;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[(long long)blockIdx.y * n + i] = cmake(0.0, 0.0);
}

// Above this size the per-matrix recurrences are spread over several CTAs (more launches, far
// less latency); below it one CTA per (spin,k) with batch parallelism is the better shape.
static int large_nb_threshold() {
  static int v = [] {
    if (const char* env = std::getenv("JRB_LARGE_NB")) return std::atoi(env);
    return 96;
  }();
  return v;
}
#define LARGE_NB large_nb_threshold()
// One CTA per matrix only pays when there are enough matrices to fill the GPU: with few (spin,k)
// items (a k-sharded rank of an 8-GPU run holds 8) the multi-CTA kernels win from 33 bands on.
static bool multi_cta_small(int nb, int nsk) {
  static int few = [] {
    if (const char* env = std::getenv("JRB_MULTI_CTA_NSK")) return std::atoi(env);
    return 16;
  }();
  return nb > LARGE_NB || (nb > 32 && nsk <= few);
}

// Rinv = R^-1 (upper triangular)
static int tri_inverse(const cplx* R, int nb, int nsk, cplx* Rinv, cudaStream_t st,
                       const int* skip = nullptr, cplx* scratch = nullptr) {
  static int no_doubling = std::getenv("JRB_NO_TRI_DOUBLING") ? std::atoi(std::getenv("JRB_NO_TRI_DOUBLING")) : 0;
  if (multi_cta_small(nb, nsk) && scratch && !no_doubling) {
    // recursive doubling over 32 x 32 blocks; `scratch`: nb x nb per matrix whose strictly upper
    // part is free (the Cholesky work matrix).  Rinv is zeroed first unless the matrix is skipped
    // (its factor and inverse were written by k_near_identity).
    const int nblk = (nb + CP - 1) / CP;
    k_zero_unless_skipped<<<dim3((unsigned)(((long long)nb * nb + 255) / 256), nsk), 256, 0, st>>>(
      Rinv, (long long)nb * nb, skip);
    JRB_CHECK_LAUNCH("k_zero_unless_skipped");
    k_tri_inv_diag256<<<dim3(nblk, nsk), 256, 0, st>>>(R, nb, Rinv, skip);
    JRB_CHECK_LAUNCH("k_tri_inv_diag256");
    for (int g = 1; g < nblk; g *= 2) {
      const int ngroups = (nblk + 2 * g - 1) / (2 * g);
      for (int phase = 0; phase < 2; ++phase) {
        k_tri_inv_level<<<dim3(g * g, ngroups, nsk), 256, 0, st>>>(R, Rinv, scratch, nb, nblk, g, phase,
                                                                  skip);
        JRB_CHECK_LAUNCH("k_tri_inv_level");
      }
    }
    return 0;
  }
  if (multi_cta_small(nb, nsk)) {
    const int nblk = (nb + CP - 1) / CP;
    k_tri_inv_diag<<<dim3(nblk, nsk), 32, 0, st>>>(R, nb, Rinv, skip);
    JRB_CHECK_LAUNCH("k_tri_inv_diag");
    static int once = opt_in_smem(k_tri_inv_offdiag, 160 * 1024);
    if (once) return once;
    k_tri_inv_offdiag<<<dim3(4 * nblk, nsk), 256, nblk * CP * 8 * (int)sizeof(cplx), st>>>(R, nb, Rinv,
                                                                                    skip);
    JRB_CHECK_LAUNCH("k_tri_inv_offdiag");
    return 0;
  }
  dim3 igrid((nb + 7) / 8, nsk);
  k_tri_inv_cols<<<igrid, 256, 8 * nb * (int)sizeof(cplx), st>>>(R, nb, Rinv, skip);
  JRB_CHECK_LAUNCH("k_tri_inv_cols");
  return 0;
}

// S[sk] = Hermitian sum of the Gram partials in d_gpart
static int reduce_gram(jrb_plan* p, int nsk, int nchunks, cplx* S, cudaStream_t st) {
  dim3 egrid((unsigned)(((long long)p->nb * p->nb + SMALL_T - 1) / SMALL_T), nsk);
  k_gram_reduce<<<egrid, SMALL_T, 0, st>>>(p->d_gpart, nchunks, p->nb, S);
  JRB_CHECK_LAUNCH("k_gram_reduce");
  return 0;
}

// S (Hermitian, destroyed) -> Rt = chol(S)^H, Rit = Rt^-1
static int chol_and_inverse(jrb_plan* p, int nsk, cplx* S, cplx* Rt, cplx* Rit, cudaStream_t st,
                            const int* skip = nullptr, bool prefer_single = false) {
  const int nb = p->nb;
  int* fail = reinterpret_cast<int*>(p->d_scal + 32);
  const CfConfig cf = cf_config(nb);
  if (cf.ept) return factor_stream(p, cf, false, nsk, S, 0.0, Rt, Rit, skip, st);
  const int smem1 = nb * (CHOL_PB + 1 + CHOL_KC + 1) * (int)sizeof(cplx);
  // prefer_single: the caller expects every matrix to be skipped (second pass behind the closed
  // form): two launches of the one-CTA kernels instead of the 20+ of the multi-CTA ones
  if (multi_cta_small(nb, nsk) && !(prefer_single && smem1 <= 200 * 1024)) {
    for (int p0 = 0; p0 < nb; p0 += CP) {
      const int below = std::max(0, nb - p0 - CP);
      dim3 grid(1 + (below + CP - 1) / CP, nsk);
      static int once = opt_in_smem(k_chol_panel, CP_SMEM);
      if (once) return once;
      static int left = std::getenv("JRB_CHOL_LEFT") ? std::atoi(std::getenv("JRB_CHOL_LEFT")) : 0;
      k_chol_panel<<<grid, 256, CP_SMEM, st>>>(S, Rt, nb, p0, fail, skip, left ? 0 : 1);
      JRB_CHECK_LAUNCH("k_chol_panel");
      if (!left && below > 0) {
        const int nt = (below + CP - 1) / CP;
        k_trailing_update<<<dim3(nt * (nt + 1) / 2, nsk), 256, 0, st>>>(S, nb, p0, skip);
        JRB_CHECK_LAUNCH("k_trailing_update");
      }
    }
    return tri_inverse(Rt, nb, nsk, Rit, st, skip, S);  // S: L below, free above the diagonal
  }
  const int smem = nb * (CHOL_PB + 1 + CHOL_KC + 1) * (int)sizeof(cplx);
  static int once = opt_in_smem(k_chol_blocked, 200 * 1024);
  if (once) return once;
  if (smem > 200 * 1024) {
    set_error("Cholesky panel does not fit shared memory (too many bands)");
    return JRB_EUNSUPPORTED;
  }
  k_chol_blocked<<<nsk, CHOL_T, smem, st>>>(S, Rt, nb, fail, skip);
  JRB_CHECK_LAUNCH("k_chol_blocked");
  dim3 igrid((nb + 7) / 8, nsk);
  k_tri_inv_cols<<<igrid, 256, 8 * nb * (int)sizeof(cplx), st>>>(Rt, nb, Rit, skip);
  JRB_CHECK_LAUNCH("k_tri_inv_cols");
  return 0;
}

// Cholesky-QR2 of the (spin,k) range [sk0, sk0 + nsk), in phases so that a row-sharded caller can
// all-reduce the Gram matrices in between (jrb_qr_rows_*): pass 0 works on W, pass 1 on Q1.
struct QrSlots {
  cplx *S, *Rt, *Rit, *R1, *R1inv, *tmp, *rinv;
  long long soff, moff;
};
static QrSlots qr_slots(jrb_plan* p, int sk0) {
  const long long nn = (long long)p->nb * p->nb;
  const long long nall = (long long)p->ns * p->nk * nn;
  QrSlots q;
  q.soff = (long long)sk0 * p->ng * p->nb;
  q.moff = (long long)sk0 * nn;
  q.S = p->d_small + q.moff;                 // slot 0: Gram / Cholesky work
  q.Rt = p->d_small + nall + q.moff;         // slot 1: R of the current pass
  q.Rit = p->d_small + 2 * nall + q.moff;    // slot 2: its inverse
  q.R1 = p->d_small + 3 * nall + q.moff;     // slot 3: R1 (first pass), kept for the compose
  q.R1inv = p->d_small + 4 * nall + q.moff;  // slot 4
  q.tmp = p->d_tmp + q.soff;
  q.rinv = p->d_rinv + q.moff;
  return q;
}

// Gram matrix of pass `pass` into S (Hermitian, full storage), S indexed from sk0
int launch_qr_gram_phase(jrb_plan* p, int sk0, int nsk, const double* w_re, const double* w_im,
                         int pass, cplx* S, cudaStream_t st) {
  const QrSlots q = qr_slots(p, sk0);
  int nchunks = 0, rc = 0;
  TallMat A = pass == 0 ? TallMat{w_re + q.soff, w_im + q.soff, p->nb}
                        : TallMat{reinterpret_cast<const double*>(q.tmp), nullptr, p->nb};
  if ((rc = run_gram(p, nsk, A, A, true, p->d_gpart, &nchunks, st))) return rc;
  return reduce_gram(p, nsk, nchunks, S, st);
}

// Factor S (destroyed) and apply: pass 0: Q1 = W R1^-1 (plan work space); pass 1: R2, Q = Q1 R2^-1,
// R = R2 R1, R^-1 = R1^-1 R2^-1.
int launch_qr_apply_phase(jrb_plan* p, int sk0, int nsk, const double* w_re, const double* w_im,
                          int pass, cplx* S, cplx* qout, cplx* r, cudaStream_t st) {
  const QrSlots q = qr_slots(p, sk0);
  const int nb = p->nb;
  int rc = 0;
  TallMat none{nullptr, nullptr, 0};
  const CfConfig cf = cf_config(nb);
  if (pass == 0) {
    TallMat W{w_re + q.soff, w_im + q.soff, nb};
    if ((rc = chol_and_inverse(p, nsk, S, q.R1, q.R1inv, st))) return rc;
    return run_apply<0>(p, nsk, W, q.R1inv, TRI_UPPER, none, nullptr, TRI_FULL, 1,
                        reinterpret_cast<double*>(q.tmp), nullptr, st);
  }
  TallMat Q1{reinterpret_cast<const double*>(q.tmp), nullptr, nb};
  const char* no_shortcut = std::getenv("JRB_NO_QR_SHORTCUT");
  const bool shortcut = !(no_shortcut && std::atoi(no_shortcut) != 0);
  if (cf.ncta == 1) {
    // one launch: closed form or factorisation + inverse; then the composition with the first pass
    if ((rc = factor_stream(p, cf, true, nsk, S, shortcut ? 1e-10 : 0.0, q.Rt, q.Rit, nullptr, st)))
      return rc;
    if ((rc = compose_factors(p, nsk, q.Rt, q.Rit, q.R1, q.R1inv, r + q.moff, q.rinv, st))) return rc;
    return run_apply<0>(p, nsk, Q1, q.Rit, TRI_UPPER, none, nullptr, TRI_FULL, 1,
                        reinterpret_cast<double*>(qout + q.soff), nullptr, st);
  }
  const int* skip = nullptr;
  const long long nn = (long long)nb * nb;
  dim3 egrid((unsigned)((nn + SMALL_T - 1) / SMALL_T), nsk);
  if (shortcut) {
    unsigned long long* emax = reinterpret_cast<unsigned long long*>(p->d_emax) + sk0;
    JRB_CUDA(cudaMemsetAsync(emax, 0, sizeof(unsigned long long) * nsk, st));
    k_near_identity_max<<<egrid, 256, 0, st>>>(S, nb, emax);
    JRB_CHECK_LAUNCH("k_near_identity_max");
    k_near_identity_apply<<<egrid, 256, 0, st>>>(S, nb, 1e-10, emax, q.Rt, q.Rit, p->d_skip + sk0);
    JRB_CHECK_LAUNCH("k_near_identity_apply");
    skip = p->d_skip + sk0;
  }
  // the regular factorisation stands down where the closed form applied; with a skip list the
  // one-CTA kernels serve (2 launches that return at once instead of 22)
  if ((rc = chol_and_inverse(p, nsk, S, q.Rt, q.Rit, st, skip, skip != nullptr))) return rc;
  if ((rc = compose_factors(p, nsk, q.Rt, q.Rit, q.R1, q.R1inv, r + q.moff, q.rinv, st))) return rc;
  return run_apply<0>(p, nsk, Q1, q.Rit, TRI_UPPER, none, nullptr, TRI_FULL, 1,
                      reinterpret_cast<double*>(qout + q.soff), nullptr, st);
}

int launch_qr_fwd_range(jrb_plan* p, int sk0, int nsk, const double* w_re, const double* w_im,
                        cplx* q, cplx* r, cudaStream_t st) {
  const QrSlots sl = qr_slots(p, sk0);
  int rc = 0;
  for (int pass = 0; pass < 2; ++pass) {
    if ((rc = launch_qr_gram_phase(p, sk0, nsk, w_re, w_im, pass, sl.S, st))) return rc;
    if ((rc = launch_qr_apply_phase(p, sk0, nsk, w_re, w_im, pass, sl.S, q, r, st))) return rc;
  }
  return 0;
}

int launch_qr_fwd(jrb_plan* p, const double* w_re, const double* w_im, cplx* q, cplx* r,
                  cudaStream_t st) {
  return launch_qr_fwd_range(p, 0, p->ns * p->nk, w_re, w_im, q, r, st);
}

// Adjoint for the (spin,k) range [sk0, sk0 + nsk); all arrays indexed from (spin,k) 0.  Two phases
// (M = Q^H G, then the small algebra + the tall product) for the same reason as the forward.
int launch_qr_bwd_gram_phase(jrb_plan* p, int sk0, int nsk, const cplx* q, const cplx* gq, cplx* M,
                             cudaStream_t st) {
  const QrSlots sl = qr_slots(p, sk0);
  int nchunks = 0, rc = 0;
  TallMat Q{reinterpret_cast<const double*>(q + sl.soff), nullptr, p->nb};
  TallMat G{reinterpret_cast<const double*>(gq + sl.soff), nullptr, p->nb};
  if ((rc = run_gram(p, nsk, Q, G, false, p->d_gpart, &nchunks, st))) return rc;
  return reduce_gram(p, nsk, nchunks, M, st);  // only the upper triangle of M is consumed
}

// M: (nsk, nb, nb) reduced Q^H G of this range (upper triangle), must not alias the plan's slots 0-2
int launch_qr_bwd_apply_phase(jrb_plan* p, int sk0, int nsk, const cplx* q, const cplx* r,
                              const cplx* gq, const double* occ, const cplx* M, double* g_re,
                              double* g_im, cudaStream_t st) {
  const int nb = p->nb;
  const long long nn = (long long)nb * nb;
  const QrSlots sl = qr_slots(p, sk0);
  const cplx* rinv = sl.rinv;  // R^-1 of the plan's own forward call
  int rc = 0;
  if (r != p->d_r) {
    if ((rc = tri_inverse(r + sl.moff, nb, nsk, sl.R1, st))) return rc;
    rinv = sl.R1;
  }
  cplx* X = sl.S;
  cplx* T1 = sl.Rt;
  cplx* T2 = sl.Rit;
  TallMat Q{reinterpret_cast<const double*>(q + sl.soff), nullptr, nb};
  TallMat G{reinterpret_cast<const double*>(gq + sl.soff), nullptr, nb};
  dim3 egrid((unsigned)((nn + SMALL_T - 1) / SMALL_T), nsk);
  k_bwd_x<<<egrid, SMALL_T, 0, st>>>(M, 1, nb, occ ? occ + (long long)sk0 * nb : nullptr, rinv, X, T1);
  JRB_CHECK_LAUNCH("k_bwd_x");
  if (nb > 96) {  // T2 = X Rinv^H as 32 x 32 tiles
    TileGemmOps ops;
    ops.A[0] = X; ops.B[0] = rinv; ops.C[0] = T2;
    ops.A[1] = X; ops.B[1] = rinv; ops.C[1] = T2;
    const int nblk = (nb + CP - 1) / CP;
    k_tile_gemm<1><<<dim3(nblk, nblk, nsk), 256, 0, st>>>(ops, nb, nsk);
    JRB_CHECK_LAUNCH("k_tile_gemm");
  } else {
    k_bwd_t2<<<egrid, SMALL_T, 0, st>>>(X, rinv, nb, T2);
    JRB_CHECK_LAUNCH("k_bwd_t2");
  }
  return run_apply<1>(p, nsk, G, T1, TRI_LOWER, Q, T2, TRI_FULL, 2, g_re + sl.soff, g_im + sl.soff, st);
}

int launch_qr_bwd_range(jrb_plan* p, int sk0, int nsk, const cplx* q, const cplx* r,
                        const cplx* gq, const double* occ, double* g_re, double* g_im,
                        cudaStream_t st) {
  const QrSlots sl = qr_slots(p, sk0);
  int rc = 0;
  if ((rc = launch_qr_bwd_gram_phase(p, sk0, nsk, q, gq, sl.R1inv, st))) return rc;
  return launch_qr_bwd_apply_phase(p, sk0, nsk, q, r, gq, occ, sl.R1inv, g_re, g_im, st);
}

int launch_qr_bwd(jrb_plan* p, const cplx* q, const cplx* r, const cplx* gq, const double* occ,
                  double* g_re, double* g_im, cudaStream_t st) {
  return launch_qr_bwd_range(p, 0, p->ns * p->nk, q, r, gq, occ, g_re, g_im, st);
}

}  // namespace jrb
