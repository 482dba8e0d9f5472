// Tall-skinny complex-FP64 products on the FP64 tensor cores (DMMA mma.sync.m8n8k4.f64;
// tcgen05/TMEM have no FP64 kind):
//   k_gram : partial[chunk][sk] = upper blocks of A^H B over a chunk of rows  (split-K Gram)
//   k_apply: Out = In1 T1 (+ In2 T2), T1/T2 small, optionally triangular       (tall x small)
// Measured on B200 (tools/fp64_peak.cu): DMMA peaks at 36.9 TFLOP/s = the DFMA rate, so the
// only way to go faster is to issue fewer of them.  Both kernels therefore skip the 8x8 blocks
// that are structurally zero or not needed:
//   * Gram matrices are Hermitian (W^H W) or only their upper triangle is consumed (the QR
//     adjoint uses up(Q^H G) and Re diag) -> only blocks bi <= bj are computed (45 of 81 per
//     72 x 72 super-tile on the diagonal);
//   * R^-1 is upper triangular and diag(f) R^-H lower triangular -> the k-range of every
//     8-column block of the apply is clipped to its non-zero part.
// CTA = 256 threads = 8 warps (2 per SM sub-partition).  Operands are staged as interleaved
// complex in shared memory by cp.async (LDGSTS) in a multi-stage ring; fragment loads are
// LDS.128 (re, im together) and bank-conflict free by construction: leading dimensions are
// == 2 (mod 8) complex for [k][cols] panels and == 4 (mod 8) for [rows][k] panels.  A complex
// product is 4 real DMMAs.
#pragma once
#include <cuda_runtime.h>

#include "plan.h"

namespace jrb {

constexpr int QT = 72;        // Gram super-tile edge = 9 blocks of 8
constexpr int QK = 16;        // reduction depth per pipeline stage
constexpr int QLDB = QT + 2;  // [k][72] panel leading dimension (complex), 74 % 8 == 2
constexpr int QLDA = QK + 4;  // [rows][k] panel leading dimension (complex), 20 % 8 == 4
constexpr int QTHREADS = 256;
constexpr int QWARPS = 8;
constexpr int QROWS = 8 * QWARPS;  // rows of the apply CTA tile (one 8-row block per warp)
constexpr int QMAXSLOT = 11;       // ceil(81 / 8) Gram blocks per warp
constexpr int QGRP = 3;            // Gram blocks whose DMMAs are interleaved
constexpr int QSLOT_DIAG = 6;      // ceil(45 / 8): one diagonal super-tile only (nb <= 72)

enum TriKind { TRI_FULL = 0, TRI_UPPER = 1, TRI_LOWER = 2 };

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  // volatile: keeps the hand-interleaved issue order (and the register pressure) as written
  asm volatile(
    "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
    : "+d"(c[0]), "+d"(c[1])
    : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(n));
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(n));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// A tall matrix [rows][ld] given either as interleaved complex or as split re / im arrays.
struct TallMat {
  const double* re;
  const double* im;  // nullptr => interleaved complex at `re`
  long long ld;
  // async copy of element (r, c) into an interleaved complex slot
  __device__ __forceinline__ void fetch(cplx* dst, long long r, int c, bool valid) const {
    const long long o = valid ? r * ld + c : 0;
    if (im == nullptr) {
      cp_async16(dst, reinterpret_cast<const cplx*>(re) + o, valid);
    } else {
      cp_async8(&dst->x, re + o, valid);
      cp_async8(&dst->y, im + o, valid);
    }
  }
  __device__ __forceinline__ TallMat offset(long long elems) const {
    TallMat t = *this;
    t.re += elems * (im ? 1 : 2);
    if (im) t.im += elems;
    return t;
  }
};

// ---------------------------------------------------------------------------------------
// partial[chunk][sk][i][j] = sum_{g in chunk} conj(A[g][i]) B[g][j]   for 8x8 blocks bi <= bj
// (entries below the block diagonal are NOT written; consumers mirror / ignore them).
// grid: (ntiles (ntiles + 1) / 2 upper super-tiles, nchunks, nsk)
// dynamic smem: STAGES * 2 panels [QK][ldb]
// LDB_T: panel leading dimension; 0 = narrow panels sized at run time (ldb = pw + 2, still == 2 mod
// 8), used with a deeper ring for nb <= 32 where a 16-row stage carries too few DMMAs to hide the
// load latency behind a 3-stage ring.
// Rectangular form (GramRect with nbB > 0): A has nb columns, B has nbB, every super-tile pair
// (ti, tj) is computed (grid.x = tiles_A * tiles_B), A is indexed by sk % a_mod with its own
// stride -- the projector product P = Phi Q of the non-local pseudopotential, where Phi belongs
// to the k-point, not to (spin, k).
struct GramRect {
  int nbB;              // 0: square Gram (the QR products)
  long long a_stride;   // elements between consecutive A matrices
  int a_mod;            // A matrix of item sk is sk % a_mod
};
template <int STAGES, int MAXSLOT, int LDB_T = QLDB>
__global__ void __launch_bounds__(QTHREADS, (MAXSLOT <= 2 ? 3 : (MAXSLOT <= 6 ? 2 : 1)))
k_gram(TallMat A, TallMat B, int same, long long ng, int nb, long long sk_stride, int ntiles,
       long long rows_per_chunk, cplx* __restrict__ partial, GramRect rect = GramRect{0, 0, 1}) {
  extern __shared__ __align__(16) unsigned char smem_raw_[];
  cplx* sA = reinterpret_cast<cplx*>(smem_raw_);
  const int nbB = rect.nbB > 0 ? rect.nbB : nb;
  int ti = 0, tj;
  if (rect.nbB > 0) {
    const int ntj = (nbB + QT - 1) / QT;
    ti = blockIdx.x / ntj;
    tj = blockIdx.x % ntj;
  } else {
    // upper super-tile (ti <= tj) of this CTA
    int rem = blockIdx.x;
    while (rem >= ntiles - ti) {
      rem -= ntiles - ti;
      ++ti;
    }
    tj = ti + rem;
  }
  const bool diag = rect.nbB == 0 && ti == tj;
  const bool one_panel = same && diag;

  const int chunk = blockIdx.y, sk = blockIdx.z, nsk = gridDim.z;
  const long long g_begin = (long long)chunk * rows_per_chunk;
  const long long g_end = min(ng, g_begin + rows_per_chunk);
  const int i0 = ti * QT, j0 = tj * QT;
  const int wi_cols = min(QT, nb - i0), wj_cols = min(QT, nbB - j0);
  const int nbi = (wi_cols + 7) >> 3, nbj = (wj_cols + 7) >> 3;
  const TallMat a = rect.nbB > 0 ? A.offset((sk % rect.a_mod) * rect.a_stride) : A.offset(sk * sk_stride);
  const TallMat bm = B.offset(sk * sk_stride);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lr = lane >> 2, lc = lane & 3;
  const int nsteps = g_end > g_begin ? (int)((g_end - g_begin + QK - 1) / QK) : 0;

  // blocks of this warp: ids warp, warp + 8, ... in the row-major list of needed blocks
  // packed per slot: 8 bi in the low half, 8 bj in the high half, -1 = no block
  int blk[MAXSLOT];
#pragma unroll
  for (int s = 0; s < MAXSLOT; ++s) {
    int id = warp + QWARPS * s, bi, bj;
    if (diag) {
      bi = 0;
      while (bi < nbi && id >= nbj - bi) {
        id -= nbj - bi;
        ++bi;
      }
      bj = bi + id;
    } else {
      bi = id / nbj;
      bj = id % nbj;
    }
    blk[s] = bi < nbi ? (8 * bi) | ((8 * bj) << 16) : -1;
  }
  const int ldb = LDB_T ? LDB_T : 8 * max(nbi, nbj) + 2;
  cplx* sB = one_panel ? sA : sA + STAGES * QK * ldb;
  const int lofs = lc * ldb + lr;

  // panel elements of this thread: e = tid + 256 q -> (row e / pw, col e % pw), fixed per thread.
  // pw = the panel columns any block of this CTA reads (whole 8-column blocks): with few bands
  // (nb = 15 ... 32) a 72-wide panel would be 55-80 % zero-fill copies, and the copy instructions,
  // not the DMMAs, bound the stage (measured: the same 4.4 us per 16-row stage at nb = 30 and 66).
  constexpr int PWMAX = LDB_T ? QT : 32;  // the run-time-narrow variant serves nb <= 32 only
  constexpr int GLOADS = (QK * PWMAX + QTHREADS - 1) / QTHREADS;
  const int pw = 8 * max(nbi, nbj);
  int lrc[GLOADS];
#pragma unroll
  for (int q = 0; q < GLOADS; ++q) {
    const int e = threadIdx.x + QTHREADS * q;
    lrc[q] = e < QK * pw ? (e / pw) | ((e % pw) << 8) : -1;
  }
  auto load_stage = [&](int step, int stage) {
    const long long g0 = g_begin + (long long)step * QK;
#pragma unroll
    for (int q = 0; q < GLOADS; ++q) {
      if (lrc[q] >= 0) {
        const int r = lrc[q] & 0xff, c = lrc[q] >> 8;
        const long long g = g0 + r;
        a.fetch(&sA[(stage * QK + r) * ldb + c], g, i0 + c, g < g_end && c < wi_cols);
        if (!one_panel)
          bm.fetch(&sB[(stage * QK + r) * ldb + c], g, j0 + c, g < g_end && c < wj_cols);
      }
    }
  };

  double cre[MAXSLOT][2], cim[MAXSLOT][2];
#pragma unroll
  for (int s = 0; s < MAXSLOT; ++s) cre[s][0] = cre[s][1] = cim[s][0] = cim[s][1] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nsteps) load_stage(s, s);
    cp_async_commit();
  }
  for (int it = 0; it < nsteps; ++it) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (it + STAGES - 1 < nsteps) load_stage(it + STAGES - 1, (it + STAGES - 1) % STAGES);
    cp_async_commit();
    const cplx* pa = sA + (it % STAGES) * QK * ldb;
    const cplx* pb = sB + (it % STAGES) * QK * ldb;
#pragma unroll
    for (int k4 = 0; k4 < QK / 4; ++k4) {
      // slots in groups of QGRP: all first products, then all second products, so that the two
      // DMMAs accumulating into the same registers are QGRP * 2 - 1 issues apart
#pragma unroll
      for (int s0 = 0; s0 < MAXSLOT; s0 += QGRP) {
        cplx fa[QGRP], fb[QGRP];
#pragma unroll
        for (int q = 0; q < QGRP; ++q) {
          const int s = s0 + q;
          if (s < MAXSLOT && blk[s] >= 0) {
            fa[q] = pa[k4 * 4 * ldb + lofs + (blk[s] & 0xffff)];  // A frag (row i, col k)
            fb[q] = pb[k4 * 4 * ldb + lofs + (blk[s] >> 16)];     // B frag (row k, col j)
          }
        }
        // conj(a) b = (ar br + ai bi) + i (ar bi - ai br)
#pragma unroll
        for (int q = 0; q < QGRP; ++q) {
          const int s = s0 + q;
          if (s < MAXSLOT && blk[s] >= 0) {
            dmma(cre[s], fa[q].x, fb[q].x);
            dmma(cim[s], fa[q].x, fb[q].y);
          }
        }
#pragma unroll
        for (int q = 0; q < QGRP; ++q) {
          const int s = s0 + q;
          if (s < MAXSLOT && blk[s] >= 0) {
            dmma(cre[s], fa[q].y, fb[q].y);
            dmma(cim[s], -fa[q].y, fb[q].x);
          }
        }
      }
    }
  }
  cp_async_wait<0>();
  cplx* out = partial + ((long long)chunk * nsk + sk) * nb * nbB;
#pragma unroll
  for (int s = 0; s < MAXSLOT; ++s) {
    if (blk[s] >= 0) {
      const int i = i0 + (blk[s] & 0xffff) + lr;
      const int j = j0 + (blk[s] >> 16) + 2 * lc;
      if (i < nb) {
        if (j < nbB) out[(long long)i * nbB + j] = cmake(cre[s][0], cim[s][0]);
        if (j + 1 < nbB) out[(long long)i * nbB + j + 1] = cmake(cre[s][1], cim[s][1]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Out[g][j] = sum_i In1[g][i] T1[i][j] (+ sum_i In2[g][i] T2[i][j])
// tri1 / tri2 say which part of T1 / T2 is structurally zero (TriKind).  Per 16-deep k-step
// the needed 8-column blocks are a contiguous range [UB, UE): the step body is instantiated
// for every range start (upper T) / end (lower T), so the hot code is straight-line DMMAs.
// MODE 0: store interleaved complex;  MODE 1: store 2 Re / 2 Im into split real arrays.
// SPLIT: In1 is given as separate re / im arrays (the parameters w_re, w_im).
// CTA tile: 64 rows (one 8-row block per warp) x 8 NCB columns.
// grid: (ceil(ng / 64), ceil(nb / (8 NCB)), nsk)

// one k-step on blocks [UB, UE): nk4 sub-steps of depth 4
template <int NCB, int UB, int UE>
__device__ __forceinline__ void apply_step(double (&cre)[NCB][2], double (&cim)[NCB][2],
                                           const cplx* __restrict__ pa,
                                           const cplx* __restrict__ pb, int nk4) {
  constexpr int LDB = 8 * NCB + 2;
#pragma unroll
  for (int k4 = 0; k4 < QK / 4; ++k4) {
    if (k4 < nk4) {
      const cplx fa = pa[k4 * 4];  // A frag (row g, col k)
      const double nai = -fa.y;
      cplx fb[UE - UB];
#pragma unroll
      for (int u = UB; u < UE; ++u) fb[u - UB] = pb[k4 * 4 * LDB + 8 * u];  // B frag (row k, col j)
      // a b = (ar br - ai bi) + i (ar bi + ai br): first products of every block, then the
      // second ones, so the two DMMAs into one accumulator are far apart
#pragma unroll
      for (int u = UB; u < UE; ++u) {
        dmma(cre[u], fa.x, fb[u - UB].x);
        dmma(cim[u], fa.x, fb[u - UB].y);
      }
#pragma unroll
      for (int u = UB; u < UE; ++u) {
        dmma(cre[u], nai, fb[u - UB].y);
        dmma(cim[u], fa.y, fb[u - UB].x);
      }
    }
  }
}

template <int NCB, int U>
struct ApplyDispatch {
  // blocks [ub, NCB)
  static __device__ __forceinline__ void from(int ub, double (&cre)[NCB][2], double (&cim)[NCB][2],
                                              const cplx* pa, const cplx* pb, int nk4) {
    if (ub <= U) {
      apply_step<NCB, U, NCB>(cre, cim, pa, pb, nk4);
    } else if constexpr (U + 1 < NCB) {
      ApplyDispatch<NCB, U + 1>::from(ub, cre, cim, pa, pb, nk4);
    }
  }
  // blocks [0, ue), ue >= 1
  static __device__ __forceinline__ void upto(int ue, double (&cre)[NCB][2], double (&cim)[NCB][2],
                                              const cplx* pa, const cplx* pb, int nk4) {
    if (ue >= NCB - U) {
      apply_step<NCB, 0, NCB - U>(cre, cim, pa, pb, nk4);
    } else if constexpr (U + 1 < NCB) {
      ApplyDispatch<NCB, U + 1>::upto(ue, cre, cim, pa, pb, nk4);
    }
  }
};

// few bands (NCB <= 4): the k-loop is 2-4 steps, the kernel streams rows; three resident CTAs
// (84 registers) overlap more row tiles
// Rectangular form (ApplyRect with kdim > 0): In1 is [ng][kdim] (its own stride, indexed by
// sk % in_mod), T1 is [kdim][nb], and the product is ADDED to the interleaved complex output --
// hq += Phi^H P of the non-local pseudopotential.  MODE 0, one term, TRI_FULL only.
struct ApplyRect {
  int kdim;             // 0: square (the QR products: In1 is [ng][nb], T is [nb][nb])
  long long in_stride;
  int in_mod;
};
template <int MODE, int NCB, int STAGES, bool SPLIT>
__global__ void __launch_bounds__(QTHREADS, (NCB <= 4 ? 3 : 2))
k_apply(const double* __restrict__ in1_re, const double* __restrict__ in1_im,
        const cplx* __restrict__ T1, int tri1, const cplx* __restrict__ in2,
        const cplx* __restrict__ T2, int tri2, int nterms, long long ng, int nb,
        long long sk_stride, double* __restrict__ out_a, double* __restrict__ out_b,
        ApplyRect rect = ApplyRect{0, 0, 1}) {
  constexpr int LDB = 8 * NCB + 2;
  constexpr int ALOADS = QROWS * QK / QTHREADS;                    // 4
  constexpr int BLOADS = (QK * 8 * NCB + QTHREADS - 1) / QTHREADS;
  extern __shared__ __align__(16) unsigned char smem_raw_[];
  cplx* sA = reinterpret_cast<cplx*>(smem_raw_);  // [STAGES][64][QLDA]
  cplx* sB = sA + STAGES * QROWS * QLDA;          // [STAGES][QK][LDB]

  const long long g0 = (long long)blockIdx.x * QROWS;
  const int j0 = blockIdx.y * 8 * NCB;
  const int jend = min(nb, j0 + 8 * NCB);
  const int nblk = (jend - j0 + 7) >> 3;  // column blocks that exist in this tile
  const int sk = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lr = lane >> 2, lc = lane & 3;
  const long long skoff = (long long)sk * sk_stride;
  const int kdim = rect.kdim > 0 ? rect.kdim : nb;   // rows of T1 = columns of In1
  const cplx* t1 = T1 + (long long)sk * kdim * nb;
  const cplx* t2 = nterms > 1 ? T2 + (long long)sk * nb * nb : t1;

  // k-range of each term for this column tile (multiples of 4 at the lower end)
  const int kb1 = tri1 == TRI_LOWER ? (j0 & ~3) : 0;
  const int ke1 = tri1 == TRI_UPPER ? jend : kdim;
  const int kb2 = tri2 == TRI_LOWER ? (j0 & ~3) : 0;
  const int ke2 = nterms > 1 ? (tri2 == TRI_UPPER ? jend : nb) : kb2;
  const int ns1 = (ke1 - kb1 + QK - 1) / QK;
  const int ns2 = nterms > 1 ? (ke2 - kb2 + QK - 1) / QK : 0;
  const int nsteps = ns1 + ns2;

  // fixed per-thread load coordinates
  const int ar = threadIdx.x / QK, ac = threadIdx.x % QK;  // A rows ar + 16 q, column ac
  bool arow_ok[ALOADS];
#pragma unroll
  for (int q = 0; q < ALOADS; ++q) arow_ok[q] = g0 + ar + (QTHREADS / QK) * q < ng;
  // + 16 q ld + k0; the rectangular In1 has its own leading dimension, stride and index
  const long long a_off = rect.kdim > 0
                            ? (long long)(sk % rect.in_mod) * rect.in_stride + (g0 + ar) * (long long)kdim + ac
                            : skoff + (g0 + ar) * (long long)nb + ac;
  const int ld_in = kdim;
  int b_r[BLOADS], b_c[BLOADS];
#pragma unroll
  for (int q = 0; q < BLOADS; ++q) {
    const int e = threadIdx.x + QTHREADS * q;
    b_r[q] = e / (8 * NCB);
    b_c[q] = e % (8 * NCB);
  }

  auto load_stage = [&](int step, int stage) {
    const bool second = step >= ns1;
    const int k0 = second ? kb2 + (step - ns1) * QK : kb1 + step * QK;
    const int kend = second ? ke2 : ke1;
    const cplx* T = second ? t2 : t1;
    cplx* da = sA + (stage * QROWS + ar) * QLDA + ac;
    const bool kok = k0 + ac < kend;
#pragma unroll
    for (int q = 0; q < ALOADS; ++q) {
      const bool v = arow_ok[q] && kok;
      const long long o = v ? a_off + (long long)(QTHREADS / QK) * q * (second ? nb : ld_in) + k0 : 0;
      cplx* d = da + (QTHREADS / QK) * q * QLDA;
      if (SPLIT && !second) {
        cp_async8(&d->x, in1_re + o, v);
        cp_async8(&d->y, in1_im + o, v);
      } else {
        const cplx* src = second ? in2 : reinterpret_cast<const cplx*>(in1_re);
        cp_async16(d, src + o, v);
      }
    }
#pragma unroll
    for (int q = 0; q < BLOADS; ++q) {
      if (BLOADS * QTHREADS == QK * 8 * NCB || b_r[q] < QK) {
        const bool v = k0 + b_r[q] < kend && j0 + b_c[q] < nb;
        cp_async16(&sB[(stage * QK + b_r[q]) * LDB + b_c[q]],
                   T + (v ? (long long)(k0 + b_r[q]) * nb + j0 + b_c[q] : 0), v);
      }
    }
  };

  double cre[NCB][2], cim[NCB][2];
#pragma unroll
  for (int u = 0; u < NCB; ++u) cre[u][0] = cre[u][1] = cim[u][0] = cim[u][1] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nsteps) load_stage(s, s);
    cp_async_commit();
  }
  for (int it = 0; it < nsteps; ++it) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (it + STAGES - 1 < nsteps) load_stage(it + STAGES - 1, (it + STAGES - 1) % STAGES);
    cp_async_commit();
    const bool second = it >= ns1;
    const int k0 = second ? kb2 + (it - ns1) * QK : kb1 + it * QK;
    const int kend = second ? ke2 : ke1;
    const int tri = second ? tri2 : tri1;
    const int nk4 = min(QK / 4, (kend - k0 + 3) >> 2);
    const cplx* pa = sA + ((it % STAGES) * QROWS + warp * 8 + lr) * QLDA + lc;
    const cplx* pb = sB + (it % STAGES) * QK * LDB + lc * LDB + lr;
    if (tri == TRI_UPPER) {
      // T[i][j] = 0 for i > j: block u is needed while k0 <= j0 + 8 u + 7
      const int ub = max(0, (k0 - j0 - 7 + 7) >> 3);
      ApplyDispatch<NCB, 0>::from(ub, cre, cim, pa, pb, nk4);
    } else if (tri == TRI_LOWER) {
      // T[i][j] = 0 for i < j: block u is needed once k0 + 15 >= j0 + 8 u
      const int ue = min(nblk, ((k0 + QK - 1 - j0) >> 3) + 1);
      ApplyDispatch<NCB, 0>::upto(ue, cre, cim, pa, pb, nk4);
    } else {
      ApplyDispatch<NCB, 0>::upto(nblk, cre, cim, pa, pb, nk4);
    }
  }
  cp_async_wait<0>();
  const long long g = g0 + warp * 8 + lr;
  if (g < ng) {
#pragma unroll
    for (int u = 0; u < NCB; ++u) {
      const int j = j0 + 8 * u + 2 * lc;
      if (MODE == 0) {
        cplx* o = reinterpret_cast<cplx*>(out_a) + skoff + g * nb + j;
        if (rect.kdim > 0) {  // accumulate
          if (j < nb) o[0] = cmake(o[0].x + cre[u][0], o[0].y + cim[u][0]);
          if (j + 1 < nb) o[1] = cmake(o[1].x + cre[u][1], o[1].y + cim[u][1]);
        } else {
          if (j < nb) o[0] = cmake(cre[u][0], cim[u][0]);
          if (j + 1 < nb) o[1] = cmake(cre[u][1], cim[u][1]);
        }
      } else {
        const long long o = skoff + g * nb + j;
        if (j < nb) {
          out_a[o] = 2.0 * cre[u][0];
          out_b[o] = 2.0 * cim[u][0];
        }
        if (j + 1 < nb) {
          out_a[o + 1] = 2.0 * cre[u][1];
          out_b[o + 1] = 2.0 * cim[u][1];
        }
      }
    }
  }
}

}  // namespace jrb
