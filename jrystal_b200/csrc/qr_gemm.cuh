// Tall-skinny complex-FP64 products on the FP64 tensor cores (DMMA mma.sync.m8n8k4.f64):
//   k_gram : partial[chunk][sk] = A^H B over a chunk of rows           (split-K Gram)
//   k_apply: Out = In1 T1 (+ In2 T2)                                   (tall times small)
// CTA = 288 threads = 3 x 3 warps, CTA tile 72 x 72, warp tile 24 x 24 = 3 x 3 DMMA tiles;
// a complex product is 4 real DMMAs.  Operands are staged as interleaved complex in shared
// memory by cp.async (LDGSTS) in a multi-stage ring so global loads overlap the tensor pipe;
// fragment loads are LDS.128 (re, im together) and conflict free by construction:
// leading dimensions are == 2 (mod 8) complex for [k][72] panels and == 4 (mod 8) for
// [72][k] panels.
#pragma once
#include <cuda_runtime.h>

#include "plan.h"

namespace jrb {

constexpr int QT = 72;     // CTA tile edge
constexpr int QK = 16;     // reduction depth per stage
constexpr int QLDB = 74;   // [k][72] panel leading dimension (complex), 74 % 8 == 2
constexpr int QLDA = 20;   // [72][k] panel leading dimension (complex), 20 % 8 == 4
constexpr int QTHREADS = 288;

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
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

struct Acc {
  double re[3][3][2], im[3][3][2];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int u = 0; u < 3; ++u) re[s][u][0] = re[s][u][1] = im[s][u][0] = im[s][u][1] = 0.0;
  }
};

// ---------------------------------------------------------------------------------------
// partial[chunk][sk][i][j] = sum_{g in chunk} conj(A[g][i]) B[g][j]
// grid: (tiles * tiles, nchunks, nsk); dynamic smem: STAGES * (SAME ? 1 : 2) panels
template <int STAGES, bool SAME>
__global__ void __launch_bounds__(QTHREADS, 1)
k_gram(TallMat A, TallMat B, long long ng, int nb, long long sk_stride, int tiles,
       long long rows_per_chunk, cplx* __restrict__ partial) {
  extern __shared__ __align__(16) unsigned char smem_raw_[];
  cplx* sA = reinterpret_cast<cplx*>(smem_raw_);
  cplx* sB = SAME ? sA : sA + STAGES * QK * QLDB;

  const int ti = blockIdx.x / tiles, tj = blockIdx.x % tiles;
  const int chunk = blockIdx.y, sk = blockIdx.z, nsk = gridDim.z;
  const long long g_begin = (long long)chunk * rows_per_chunk;
  const long long g_end = min(ng, g_begin + rows_per_chunk);
  const int i0 = ti * QT, j0 = tj * QT;
  const TallMat a = A.offset(sk * sk_stride), bm = B.offset(sk * sk_stride);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wi = warp / 3, wj = warp % 3;
  const int lr = lane >> 2, lc = lane & 3;
  const int nsteps = g_end > g_begin ? (int)((g_end - g_begin + QK - 1) / QK) : 0;

  auto load_stage = [&](int step, int stage) {
    const long long g0 = g_begin + (long long)step * QK;
#pragma unroll
    for (int e0 = 0; e0 < QK * QT; e0 += QTHREADS) {
      const int e = e0 + threadIdx.x;
      const int r = e / QT, c = e % QT;
      const long long g = g0 + r;
      a.fetch(&sA[(stage * QK + r) * QLDB + c], g, i0 + c, g < g_end && i0 + c < nb);
      if (!SAME) bm.fetch(&sB[(stage * QK + r) * QLDB + c], g, j0 + c, g < g_end && j0 + c < nb);
    }
  };

  Acc acc;
  acc.zero();
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
    const cplx* pa = sA + (it % STAGES) * QK * QLDB;
    const cplx* pb = sB + (it % STAGES) * QK * QLDB;
#pragma unroll
    for (int k4 = 0; k4 < QK / 4; ++k4) {
      cplx fa[3], fb[3];
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        fa[s] = pa[(k4 * 4 + lc) * QLDB + wi * 24 + s * 8 + lr];  // A frag (row i, col k)
        fb[s] = pb[(k4 * 4 + lc) * QLDB + wj * 24 + s * 8 + lr];  // B frag (row k, col j)
      }
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const double nai = -fa[s].y;
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          // conj(a) b = (ar br + ai bi) + i (ar bi - ai br)
          dmma(acc.re[s][u], fa[s].x, fb[u].x);
          dmma(acc.im[s][u], fa[s].x, fb[u].y);
          dmma(acc.re[s][u], fa[s].y, fb[u].y);
          dmma(acc.im[s][u], nai, fb[u].x);
        }
      }
    }
  }
  cp_async_wait<0>();
  cplx* out = partial + ((long long)chunk * nsk + sk) * nb * nb;
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int i = i0 + wi * 24 + s * 8 + lr;
      const int j = j0 + wj * 24 + u * 8 + 2 * lc;
      if (i < nb) {
        if (j < nb) out[(long long)i * nb + j] = cmake(acc.re[s][u][0], acc.im[s][u][0]);
        if (j + 1 < nb) out[(long long)i * nb + j + 1] = cmake(acc.re[s][u][1], acc.im[s][u][1]);
      }
    }
}

// ---------------------------------------------------------------------------------------
// Out[g][j] = sum_i In1[g][i] T1[i][j] (+ sum_i In2[g][i] T2[i][j])
// MODE 0: store interleaved complex;  MODE 1: store 2 Re / 2 Im into split real arrays.
// grid: (row tiles, col tiles, nsk)
template <int MODE, int STAGES>
__global__ void __launch_bounds__(QTHREADS, 1)
k_apply(TallMat In1, const cplx* __restrict__ T1, TallMat In2, const cplx* __restrict__ T2,
        int nterms, long long ng, int nb, long long sk_stride, double* __restrict__ out_a,
        double* __restrict__ out_b) {
  extern __shared__ __align__(16) unsigned char smem_raw_[];
  cplx* sA = reinterpret_cast<cplx*>(smem_raw_);           // [STAGES][72][QLDA]
  cplx* sB = sA + STAGES * QT * QLDA;                       // [STAGES][QK][QLDB]

  const long long g0 = (long long)blockIdx.x * QT;
  const int j0 = blockIdx.y * QT;
  const int sk = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wi = warp / 3, wj = warp % 3;
  const int lr = lane >> 2, lc = lane & 3;
  const int ksteps = (nb + QK - 1) / QK;
  const int nsteps = ksteps * nterms;
  const TallMat in1 = In1.offset(sk * sk_stride);
  const TallMat in2 = nterms > 1 ? In2.offset(sk * sk_stride) : in1;
  const cplx* t1 = T1 + (long long)sk * nb * nb;
  const cplx* t2 = nterms > 1 ? T2 + (long long)sk * nb * nb : t1;

  auto load_stage = [&](int step, int stage) {
    const bool second = step >= ksteps;
    const int k0 = (second ? step - ksteps : step) * QK;
    const TallMat& in = second ? in2 : in1;
    const cplx* T = second ? t2 : t1;
#pragma unroll
    for (int e0 = 0; e0 < QT * QK; e0 += QTHREADS) {
      const int e = e0 + threadIdx.x;
      {
        const int r = e / QK, c = e % QK;
        in.fetch(&sA[(stage * QT + r) * QLDA + c], g0 + r, k0 + c, g0 + r < ng && k0 + c < nb);
      }
      {
        const int r = e / QT, c = e % QT;
        const bool v = k0 + r < nb && j0 + c < nb;
        cp_async16(&sB[(stage * QK + r) * QLDB + c], T + (v ? (long long)(k0 + r) * nb + j0 + c : 0),
                   v);
      }
    }
  };

  Acc acc;
  acc.zero();
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
    const cplx* pa = sA + (it % STAGES) * QT * QLDA;
    const cplx* pb = sB + (it % STAGES) * QK * QLDB;
#pragma unroll
    for (int k4 = 0; k4 < QK / 4; ++k4) {
      cplx fa[3], fb[3];
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        fa[s] = pa[(wi * 24 + s * 8 + lr) * QLDA + k4 * 4 + lc];  // A frag (row g, col k)
        fb[s] = pb[(k4 * 4 + lc) * QLDB + wj * 24 + s * 8 + lr];  // B frag (row k, col j)
      }
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const double nai = -fa[s].y;
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          // a b = (ar br - ai bi) + i (ar bi + ai br)
          dmma(acc.re[s][u], fa[s].x, fb[u].x);
          dmma(acc.im[s][u], fa[s].x, fb[u].y);
          dmma(acc.re[s][u], nai, fb[u].y);
          dmma(acc.im[s][u], fa[s].y, fb[u].x);
        }
      }
    }
  }
  cp_async_wait<0>();
  const long long base = (long long)sk * sk_stride;
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const long long g = g0 + wi * 24 + s * 8 + lr;
      const int j = j0 + wj * 24 + u * 8 + 2 * lc;
      if (g < ng) {
        if (MODE == 0) {
          cplx* o = reinterpret_cast<cplx*>(out_a) + base + g * nb + j;
          if (j < nb) o[0] = cmake(acc.re[s][u][0], acc.im[s][u][0]);
          if (j + 1 < nb) o[1] = cmake(acc.re[s][u][1], acc.im[s][u][1]);
        } else {
          const long long o = base + g * nb + j;
          if (j < nb) {
            out_a[o] = 2.0 * acc.re[s][u][0];
            out_b[o] = 2.0 * acc.im[s][u][0];
          }
          if (j + 1 < nb) {
            out_a[o + 1] = 2.0 * acc.re[s][u][1];
            out_b[o + 1] = 2.0 * acc.im[s][u][1];
          }
        }
      }
    }
}

}  // namespace jrb
