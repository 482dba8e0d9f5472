// Pencil passes of the pruned 3-D transforms (hand-written sm_100a kernels).
//
// Forward (density) sweep, one batch of band groups at a time, everything FP64:
//   k_z_inv_scatter : sphere coefficients Q[s,k,g,b]  --z IFFT on occupied columns-->  A
//   k_y_inv         : A --y IFFT on occupied x planes--> B
//   k_x_inv_density : B --x IFFT--> |psi|^2, occupation weighted, accumulated into rho
// Reverse (H-apply) sweep:
//   k_z_inv_scatter, k_y_inv (recompute), then
//   k_x_vmul        : B --x IFFT--> * v_eff(r)/N --x FFT--> B (occupied x planes only)
//   k_y_fwd         : B --y FFT--> A (occupied columns only)
//   k_z_fwd_gather  : A --z FFT--> gather on the sphere + 1/2|G+k|^2 q  -> HQ[s,k,g,b]
// The dense box of an orbital therefore never exists in global memory: the scatter
// (utils.expand_coefficient, jrystal/_src/utils.py:277-281) is fused into the first pass,
// |psi|^2 + einsum (pw.py:273-278) into the last inverse pass, and the gather of the AD
// transpose into the last forward pass.
//
// Work-space layouts (NB = 8 band lanes innermost -> every access is a 128-byte line):
//   A[group][z][col][NB]          col = occupied (x, y) column (z-major: a z-plane of one
//                                 group is contiguous, which is what the fused y+x kernels of
//                                 fft_fused.cuh read)
//   B[group][xo][y][z][NB]        xo  = occupied x plane
#pragma once
#include <cuda_runtime.h>

#include "fft_lines.cuh"
#include "plan.h"

namespace jrb {

enum PassKind {
  PASS_Z_INV_SCATTER = 0,
  PASS_Y_INV = 1,
  PASS_X_DENSITY = 2,
  PASS_X_VMUL = 3,
  PASS_Y_FWD = 4,
  PASS_Z_FWD_GATHER = 5,
};

struct PassArgs {
  SphereMaps m;
  const cplx* q;    // [ns*nk][ng][nb]
  cplx* hq;         // [ns*nk][ng][nb]
  cplx* wa;
  cplx* wb;
  const cplx* wa_add;  // k_z_fwd_gather: optional second share of A, added on load (fused 128)
  const cplx* tw;   // exp(-2 pi i t / n) of the pass axis
  const double* focc;  // [groups][NB] occupation / Omega
  const double* gk2;   // [nk][ng]
  const double* veff;  // [nx*ny*nz] of the current spin
  double* rho;         // [nx*ny*nz] of the current spin
  int nb, nk, ngpk;    // bands, k-points, band groups per (spin, k)
  int g0, ngroups;     // first global group id and number of groups of this batch
  double vscale;       // 1 / N applied with v_eff
};

struct DenseArgs {
  const cplx* in;
  cplx* out;
  const cplx* tw;
  int n0, lane_total;        // line groups: n0 x ceil(lane_total / NB)
  long long stride0, lane_stride, elem_stride, batch_stride;
  double scale;
};

constexpr int pick_lpc(int n, int tpl, int nbuf) {
  int by_threads = 256 / (NB * tpl);
  int by_smem = (64 * 1024) / (n * NB * 16 * nbuf) * nbuf;  // <= 64 KB per buffer set
  if (by_smem < 1) by_smem = 1;
  int l = by_threads < by_smem ? by_threads : by_smem;
  if (l < 1) l = 1;
  while (l > 1 && (NB * tpl * l) % 32 != 0) --l;
  while ((NB * tpl * l) % 32 != 0) ++l;
  return l;
}

template <int N>
struct Cfg {
  static constexpr int TPL = LinePlan<N>::tpl;
  static constexpr int LPC = pick_lpc(N, TPL, 1);
  static constexpr int NT = NB * TPL * LPC;
  static constexpr int SMEM = LPC * N * NB * (int)sizeof(cplx);
};

#define JRB_THREAD_COORDS(N_)                          \
  const int t_ = threadIdx.x;                           \
  const int b = t_ % NB;                                \
  const int tj = (t_ / NB) % Cfg<N_>::TPL;              \
  const int ls = t_ / (NB * Cfg<N_>::TPL);              \
  extern __shared__ __align__(16) unsigned char smem_raw_[]; \
  cplx* smem = reinterpret_cast<cplx*>(smem_raw_);

__device__ __forceinline__ cplx czero() { return cmake(0.0, 0.0); }

// The z passes are latency bound gathers/scatters: 4 resident CTAs (64 registers) instead of 3
// where that costs no real spilling (checked with -Xptxas -v).
constexpr bool vmul_two_buffers(int n) { return n <= 96; }
constexpr int z_min_blocks(int n) {
  return (n == 64 || n == 49 || n == 36) ? 4 : (n == 81 || n == 100) ? 3 : (n > 96 ? 2 : 1);
}
// Long lines (16 elements per thread): cap the registers at 128 so that two CTAs are resident
// (one CTA of 8 warps per SM left every 128-point pass latency bound at 12 % occupancy).
constexpr int long_min_blocks(int n) { return n > 96 ? 2 : 1; }

// ---------------------------------------------------------------------------------------
// z pass, inverse, with the sphere scatter fused into the loads.
// grid: (ceil(ncol / LPC), ngroups)
template <int NZ>
__global__ void __launch_bounds__(Cfg<NZ>::NT, z_min_blocks(NZ)) k_z_inv_scatter(PassArgs a) {
  using F = LineFFT<NZ, +1>;
  using C = Cfg<NZ>;
  JRB_THREAD_COORDS(NZ)
  const int col = blockIdx.x * C::LPC + ls;
  const int gl = blockIdx.y;
  const int gid = a.g0 + gl;
  const int sk = gid / a.ngpk;
  const int b0 = (gid % a.ngpk) * NB;
  const bool lane_ok = (b0 + b) < a.nb;
  const bool line_ok = col < a.m.ncol;
  cplx* sm = smem + (size_t)ls * NZ * NB + b;
  cplx tw[F::CB][F::NTW];
  F::load_twiddles(tw, a.tw, tj);

  const cplx* qbase = a.q + ((long long)sk * a.m.ng) * a.nb + b0 + b;
  const int32_t* zm = a.m.zmap + (long long)(line_ok ? col : 0) * NZ;
  cplx va[F::CA][F::RA];
#pragma unroll
  for (int i = 0; i < F::CA; ++i) {
#pragma unroll
    for (int m = 0; m < F::RA; ++m) {
      cplx v = czero();
      if (F::activeA(i, tj) && line_ok && lane_ok) {
        const int g = zm[F::idxA(i, m, tj)];
        if (g >= 0) v = qbase[(long long)g * a.nb];
      }
      va[i][m] = v;
    }
  }
  F::template stageA_store<NB>(va, sm, tj);
  __syncthreads();
  cplx vb[F::CB][F::RB];
  F::template stageB_load<NB>(vb, sm, tw, tj);
  if (line_ok) {
    cplx* out = a.wa + ((long long)gl * NZ * a.m.ncol + col) * NB + b;
    const long long zs = (long long)a.m.ncol * NB;
#pragma unroll
    for (int i = 0; i < F::CB; ++i) {
      if (F::activeB(i, tj)) {
#pragma unroll
        for (int m = 0; m < F::RB; ++m) out[F::idxB(i, m, tj) * zs] = vb[i][m];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// y pass, inverse: A (occupied columns) -> B (all y) on the occupied x planes.
// grid: (ceil(nxo * nz / LPC), ngroups)
template <int NY>
__global__ void __launch_bounds__(Cfg<NY>::NT, long_min_blocks(NY)) k_y_inv(PassArgs a) {
  using F = LineFFT<NY, +1>;
  using C = Cfg<NY>;
  JRB_THREAD_COORDS(NY)
  const int nz = a.m.nz;
  const int line = blockIdx.x * C::LPC + ls;
  const bool line_ok = line < a.m.nxo * nz;
  const int xo = line_ok ? line / nz : 0;
  const int z = line_ok ? line % nz : 0;
  const int gl = blockIdx.y;
  cplx* sm = smem + (size_t)ls * NY * NB + b;
  cplx tw[F::CB][F::NTW];
  F::load_twiddles(tw, a.tw, tj);

  const cplx* in = a.wa + (((long long)gl * nz + z) * a.m.ncol) * NB + b;
  const int32_t* yc = a.m.ycol + (long long)xo * NY;
  cplx va[F::CA][F::RA];
#pragma unroll
  for (int i = 0; i < F::CA; ++i) {
#pragma unroll
    for (int m = 0; m < F::RA; ++m) {
      cplx v = czero();
      if (F::activeA(i, tj) && line_ok) {
        const int col = yc[F::idxA(i, m, tj)];
        if (col >= 0) v = in[(long long)col * NB];
      }
      va[i][m] = v;
    }
  }
  F::template stageA_store<NB>(va, sm, tj);
  __syncthreads();
  cplx vb[F::CB][F::RB];
  F::template stageB_load<NB>(vb, sm, tw, tj);
  if (line_ok) {
    cplx* out = a.wb + ((((long long)gl * a.m.nxo + xo) * NY) * nz + z) * NB + b;
#pragma unroll
    for (int i = 0; i < F::CB; ++i) {
      if (F::activeB(i, tj)) {
#pragma unroll
        for (int m = 0; m < F::RB; ++m)
          out[(long long)F::idxB(i, m, tj) * nz * NB] = vb[i][m];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// x pass, inverse, fused with rho += f |psi|^2.  A CTA owns LPC (y,z) lines and loops over
// every band group of the batch, so rho is read-modified-written once per launch without
// atomics (deterministic summation order).
// grid: (ceil(ny * nz / LPC))
template <int NX>
__global__ void __launch_bounds__(Cfg<NX>::NT, long_min_blocks(NX)) k_x_inv_density(PassArgs a) {
  using F = LineFFT<NX, +1>;
  using C = Cfg<NX>;
  JRB_THREAD_COORDS(NX)
  const long long nyz = (long long)a.m.ny * a.m.nz;
  const long long yz = (long long)blockIdx.x * C::LPC + ls;
  const bool line_ok = yz < nyz;
  cplx* sm = smem + (size_t)ls * NX * NB + b;
  cplx tw[F::CB][F::NTW];
  F::load_twiddles(tw, a.tw, tj);

  // element offsets inside one group's slab (nxo * ny * nz * NB < 2^31, checked at plan creation)
  int ioff[F::CA][F::RA];
#pragma unroll
  for (int i = 0; i < F::CA; ++i) {
#pragma unroll
    for (int m = 0; m < F::RA; ++m) {
      int o = -1;
      if (F::activeA(i, tj) && line_ok) {
        const int xo = a.m.xmap[F::idxA(i, m, tj)];
        if (xo >= 0) o = (int)(((long long)xo * nyz + yz) * NB + b);
      }
      ioff[i][m] = o;
    }
  }
  double acc[F::CB][F::RB];
#pragma unroll
  for (int i = 0; i < F::CB; ++i)
#pragma unroll
    for (int m = 0; m < F::RB; ++m) acc[i][m] = 0.0;

  const long long gstride = (long long)a.m.nxo * nyz * NB;
  for (int gl = 0; gl < a.ngroups; ++gl) {
    const cplx* in = a.wb + (long long)gl * gstride;
    cplx va[F::CA][F::RA];
#pragma unroll
    for (int i = 0; i < F::CA; ++i)
#pragma unroll
      for (int m = 0; m < F::RA; ++m) va[i][m] = ioff[i][m] >= 0 ? in[ioff[i][m]] : czero();
    const double fw = a.focc[(long long)(a.g0 + gl) * NB + b];
    F::template stageA_store<NB>(va, sm, tj);
    __syncthreads();
    cplx vb[F::CB][F::RB];
    F::template stageB_load<NB>(vb, sm, tw, tj);
#pragma unroll
    for (int i = 0; i < F::CB; ++i)
#pragma unroll
      for (int m = 0; m < F::RB; ++m)
        if (F::activeB(i, tj))
          acc[i][m] += fw * (vb[i][m].x * vb[i][m].x + vb[i][m].y * vb[i][m].y);
    __syncthreads();
  }
  // sum over the NB band lanes (adjacent lanes of one warp), then one lane per entry adds
  // into rho.
#pragma unroll
  for (int i = 0; i < F::CB; ++i) {
#pragma unroll
    for (int m = 0; m < F::RB; ++m) {
      double v = acc[i][m];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      if (F::activeB(i, tj) && line_ok && ((i * F::RB + m) % NB) == b) {
        a.rho[(long long)F::idxB(i, m, tj) * nyz + yz] += v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// x pass of the Hamiltonian apply: inverse transform, multiply by v_eff(r) / N, forward
// transform, keep the occupied x planes.  In place on B.  The inverse stage-B outputs are
// the forward stage-A inputs of the same thread (see fft_lines.cuh), so psi(r) and
// v_eff psi(r) live in registers only.
// grid: (ceil(ny * nz / LPC)); dynamic smem: 2 exchange buffers
template <int NX>
__global__ void __launch_bounds__(Cfg<NX>::NT, long_min_blocks(NX)) k_x_vmul(PassArgs a) {
  using FI = LineFFT<NX, +1>;
  using FF = LineFFT<NX, -1>;
  using C = Cfg<NX>;
  static_assert(FI::CB == FF::CA && FI::RB == FF::RA, "register chaining contract");
  JRB_THREAD_COORDS(NX)
  const long long nyz = (long long)a.m.ny * a.m.nz;
  const long long yz = (long long)blockIdx.x * C::LPC + ls;
  const bool line_ok = yz < nyz;
  // long lines: one exchange buffer (plus a barrier) so that two CTAs fit the shared memory
  constexpr bool TWO_BUF = NX <= 96;  // == vmul_two_buffers(NX)
  cplx* sm0 = smem + (size_t)ls * NX * NB + b;
  cplx* sm1 = TWO_BUF ? sm0 + (size_t)C::LPC * NX * NB : sm0;
  cplx twi[FI::CB][FI::NTW];
  cplx twf[FF::CB][FF::NTW];
  FI::load_twiddles(twi, a.tw, tj);
  FF::load_twiddles(twf, a.tw, tj);

  // The forward stage-B outputs sit on the x indices of the inverse stage-A inputs (idxB of the
  // forward plan == idxA of the inverse plan), so one offset table serves loads and stores.
  static_assert(FF::CB == FI::CA && FF::RB == FI::RA, "x index sets of inverse-in / forward-out");
  int ioff[FI::CA][FI::RA];
#pragma unroll
  for (int i = 0; i < FI::CA; ++i) {
#pragma unroll
    for (int m = 0; m < FI::RA; ++m) {
      int o = -1;
      if (FI::activeA(i, tj) && line_ok) {
        const int xo = a.m.xmap[FI::idxA(i, m, tj)];
        if (xo >= 0) o = (int)(((long long)xo * nyz + yz) * NB + b);
      }
      ioff[i][m] = o;
    }
  }
  double vv[FI::CB][FI::RB];
#pragma unroll
  for (int i = 0; i < FI::CB; ++i) {
#pragma unroll
    for (int m = 0; m < FI::RB; ++m) {
      double v = 0.0;
      if (FI::activeB(i, tj) && line_ok)
        v = a.veff[(long long)FI::idxB(i, m, tj) * nyz + yz] * a.vscale;
      vv[i][m] = v;
    }
  }

  const long long gstride = (long long)a.m.nxo * nyz * NB;
  for (int gl = 0; gl < a.ngroups; ++gl) {
    cplx* buf = a.wb + (long long)gl * gstride;
    cplx va[FI::CA][FI::RA];
#pragma unroll
    for (int i = 0; i < FI::CA; ++i)
#pragma unroll
      for (int m = 0; m < FI::RA; ++m) va[i][m] = ioff[i][m] >= 0 ? buf[ioff[i][m]] : czero();
    FI::template stageA_store<NB>(va, sm0, tj);
    __syncthreads();
    cplx vb[FI::CB][FI::RB];
    FI::template stageB_load<NB>(vb, sm0, twi, tj);
#pragma unroll
    for (int i = 0; i < FI::CB; ++i)
#pragma unroll
      for (int m = 0; m < FI::RB; ++m) vb[i][m] = cscale(vb[i][m], vv[i][m]);
    if constexpr (!TWO_BUF) __syncthreads();
    FF::template stageA_store<NB>(vb, sm1, tj);
    __syncthreads();
    cplx vc[FF::CB][FF::RB];
    FF::template stageB_load<NB>(vc, sm1, twf, tj);
#pragma unroll
    for (int i = 0; i < FF::CB; ++i)
#pragma unroll
      for (int m = 0; m < FF::RB; ++m)
        if (ioff[i][m] >= 0) buf[ioff[i][m]] = vc[i][m];
    if constexpr (!TWO_BUF) __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// y pass, forward: B (all y) -> A (occupied columns).   grid: (ceil(nxo * nz / LPC), ngroups)
template <int NY>
__global__ void __launch_bounds__(Cfg<NY>::NT, long_min_blocks(NY)) k_y_fwd(PassArgs a) {
  using F = LineFFT<NY, -1>;
  using C = Cfg<NY>;
  JRB_THREAD_COORDS(NY)
  const int nz = a.m.nz;
  const int line = blockIdx.x * C::LPC + ls;
  const bool line_ok = line < a.m.nxo * nz;
  const int xo = line_ok ? line / nz : 0;
  const int z = line_ok ? line % nz : 0;
  const int gl = blockIdx.y;
  cplx* sm = smem + (size_t)ls * NY * NB + b;
  cplx tw[F::CB][F::NTW];
  F::load_twiddles(tw, a.tw, tj);

  const cplx* in = a.wb + ((((long long)gl * a.m.nxo + xo) * NY) * nz + z) * NB + b;
  cplx va[F::CA][F::RA];
#pragma unroll
  for (int i = 0; i < F::CA; ++i) {
#pragma unroll
    for (int m = 0; m < F::RA; ++m) {
      cplx v = czero();
      if (F::activeA(i, tj) && line_ok) v = in[(long long)F::idxA(i, m, tj) * nz * NB];
      va[i][m] = v;
    }
  }
  F::template stageA_store<NB>(va, sm, tj);
  __syncthreads();
  cplx vb[F::CB][F::RB];
  F::template stageB_load<NB>(vb, sm, tw, tj);
  if (line_ok) {
    cplx* out = a.wa + (((long long)gl * nz + z) * a.m.ncol) * NB + b;
    const int32_t* yc = a.m.ycol + (long long)xo * NY;
#pragma unroll
    for (int i = 0; i < F::CB; ++i) {
      if (F::activeB(i, tj)) {
#pragma unroll
        for (int m = 0; m < F::RB; ++m) {
          const int col = yc[F::idxB(i, m, tj)];
          if (col >= 0) out[(long long)col * NB] = vb[i][m];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// z pass, forward, fused with the gather onto the sphere and the kinetic term:
//   hq[s,k,g,b] = 1/2 |G_g + k|^2 q[s,k,g,b] + FFT_z(A)[col, z(g)]
// grid: (ceil(ncol / LPC), ngroups)
template <int NZ>
__global__ void __launch_bounds__(Cfg<NZ>::NT, z_min_blocks(NZ)) k_z_fwd_gather(PassArgs a) {
  using F = LineFFT<NZ, -1>;
  using C = Cfg<NZ>;
  JRB_THREAD_COORDS(NZ)
  const int col = blockIdx.x * C::LPC + ls;
  const int gl = blockIdx.y;
  const int gid = a.g0 + gl;
  const int sk = gid / a.ngpk;
  const int k = sk % a.nk;
  const int b0 = (gid % a.ngpk) * NB;
  const bool lane_ok = (b0 + b) < a.nb;
  const bool line_ok = col < a.m.ncol;
  cplx* sm = smem + (size_t)ls * NZ * NB + b;
  cplx tw[F::CB][F::NTW];
  F::load_twiddles(tw, a.tw, tj);

  const long long in_off = ((long long)gl * NZ * a.m.ncol + (line_ok ? col : 0)) * NB + b;
  const cplx* in = a.wa + in_off;
  const long long zs = (long long)a.m.ncol * NB;
  cplx va[F::CA][F::RA];
#pragma unroll
  for (int i = 0; i < F::CA; ++i) {
#pragma unroll
    for (int m = 0; m < F::RA; ++m) {
      cplx v = czero();
      if (F::activeA(i, tj) && line_ok) {
        v = in[F::idxA(i, m, tj) * zs];
        if (a.wa_add) v = cadd(v, a.wa_add[in_off + F::idxA(i, m, tj) * zs]);
      }
      va[i][m] = v;
    }
  }
  F::template stageA_store<NB>(va, sm, tj);
  __syncthreads();
  cplx vb[F::CB][F::RB];
  F::template stageB_load<NB>(vb, sm, tw, tj);
  if (line_ok && lane_ok) {
    const long long sbase = ((long long)sk * a.m.ng) * a.nb + b0 + b;
    const int32_t* zm = a.m.zmap + (long long)col * NZ;
    const double* gk2 = a.gk2 + (long long)k * a.m.ng;
#pragma unroll
    for (int i = 0; i < F::CB; ++i) {
      if (F::activeB(i, tj)) {
#pragma unroll
        for (int m = 0; m < F::RB; ++m) {
          const int g = zm[F::idxB(i, m, tj)];
          if (g >= 0) {
            const long long o = sbase + (long long)g * a.nb;
            const cplx c = a.q[o];
            const double t = 0.5 * gk2[g];
            a.hq[o] = cmake(vb[i][m].x + t * c.x, vb[i][m].y + t * c.y);
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Dense strided line FFT (single grids rho / v and the jrb_fft3d drop-in).  The NB lanes
// run over `lane_total` neighbouring lines (stride lane_stride); elements of a line are
// elem_stride apart.   grid: (ceil(n0 * ceil(lane_total/NB) / LPC), batch)
template <int N, int DIR>
__global__ void __launch_bounds__(Cfg<N>::NT, long_min_blocks(N)) k_dense_line(DenseArgs a) {
  using F = LineFFT<N, DIR>;
  using C = Cfg<N>;
  JRB_THREAD_COORDS(N)
  const int n1c = (a.lane_total + NB - 1) / NB;
  const long long lg = (long long)blockIdx.x * C::LPC + ls;
  const bool lg_ok = lg < (long long)a.n0 * n1c;
  const int i0 = lg_ok ? (int)(lg / n1c) : 0;
  const int i1 = lg_ok ? (int)(lg % n1c) : 0;
  const int lane = i1 * NB + b;
  const bool ok = lg_ok && lane < a.lane_total;
  const long long base =
    (long long)blockIdx.y * a.batch_stride + (long long)i0 * a.stride0 + (long long)lane * a.lane_stride;
  cplx* sm = smem + (size_t)ls * N * NB + b;
  cplx tw[F::CB][F::NTW];
  F::load_twiddles(tw, a.tw, tj);
  cplx va[F::CA][F::RA];
#pragma unroll
  for (int i = 0; i < F::CA; ++i) {
#pragma unroll
    for (int m = 0; m < F::RA; ++m) {
      cplx v = czero();
      if (F::activeA(i, tj) && ok) v = a.in[base + (long long)F::idxA(i, m, tj) * a.elem_stride];
      va[i][m] = v;
    }
  }
  F::template stageA_store<NB>(va, sm, tj);
  __syncthreads();
  cplx vb[F::CB][F::RB];
  F::template stageB_load<NB>(vb, sm, tw, tj);
  if (ok) {
#pragma unroll
    for (int i = 0; i < F::CB; ++i) {
      if (F::activeB(i, tj)) {
#pragma unroll
        for (int m = 0; m < F::RB; ++m)
          a.out[base + (long long)F::idxB(i, m, tj) * a.elem_stride] = cscale(vb[i][m], a.scale);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// host-side launchers (one instantiation per supported axis length)
template <class K>
inline int set_smem_attr(K kernel, int bytes) {
  if (bytes > 48 * 1024) {
    JRB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  }
  return 0;
}

template <int N>
int launch_pass(PassKind kind, const PassArgs& a, cudaStream_t st) {
  using C = Cfg<N>;
  const int lpc = C::LPC;
  int rc = 0;
  switch (kind) {
    case PASS_Z_INV_SCATTER: {
      static int once = set_smem_attr(k_z_inv_scatter<N>, C::SMEM);
      if (once) return once;
      dim3 grid((a.m.ncol + lpc - 1) / lpc, a.ngroups);
      k_z_inv_scatter<N><<<grid, C::NT, C::SMEM, st>>>(a);
      break;
    }
    case PASS_Y_INV: {
      static int once = set_smem_attr(k_y_inv<N>, C::SMEM);
      if (once) return once;
      dim3 grid((a.m.nxo * a.m.nz + lpc - 1) / lpc, a.ngroups);
      k_y_inv<N><<<grid, C::NT, C::SMEM, st>>>(a);
      break;
    }
    case PASS_X_DENSITY: {
      static int once = set_smem_attr(k_x_inv_density<N>, C::SMEM);
      if (once) return once;
      dim3 grid((a.m.ny * a.m.nz + lpc - 1) / lpc);
      k_x_inv_density<N><<<grid, C::NT, C::SMEM, st>>>(a);
      break;
    }
    case PASS_X_VMUL: {
      constexpr int smem = (vmul_two_buffers(N) ? 2 : 1) * C::SMEM;
      static int once = set_smem_attr(k_x_vmul<N>, smem);
      if (once) return once;
      dim3 grid((a.m.ny * a.m.nz + lpc - 1) / lpc);
      k_x_vmul<N><<<grid, C::NT, smem, st>>>(a);
      break;
    }
    case PASS_Y_FWD: {
      static int once = set_smem_attr(k_y_fwd<N>, C::SMEM);
      if (once) return once;
      dim3 grid((a.m.nxo * a.m.nz + lpc - 1) / lpc, a.ngroups);
      k_y_fwd<N><<<grid, C::NT, C::SMEM, st>>>(a);
      break;
    }
    case PASS_Z_FWD_GATHER: {
      static int once = set_smem_attr(k_z_fwd_gather<N>, C::SMEM);
      if (once) return once;
      dim3 grid((a.m.ncol + lpc - 1) / lpc, a.ngroups);
      k_z_fwd_gather<N><<<grid, C::NT, C::SMEM, st>>>(a);
      break;
    }
    default:
      set_error("unknown pass kind");
      return JRB_EINVAL;
  }
  JRB_CHECK_LAUNCH("pencil pass launch");
  return rc;
}

template <int N>
int launch_dense(const DenseArgs& a, int dir, long long batch, cudaStream_t st) {
  using C = Cfg<N>;
  const int n1c = (a.lane_total + NB - 1) / NB;
  const long long groups = (long long)a.n0 * n1c;
  dim3 grid((unsigned)((groups + C::LPC - 1) / C::LPC), (unsigned)batch);
  if (dir > 0) {
    static int once = set_smem_attr(k_dense_line<N, +1>, C::SMEM);
    if (once) return once;
    k_dense_line<N, +1><<<grid, C::NT, C::SMEM, st>>>(a);
  } else {
    static int once = set_smem_attr(k_dense_line<N, -1>, C::SMEM);
    if (once) return once;
    k_dense_line<N, -1><<<grid, C::NT, C::SMEM, st>>>(a);
  }
  JRB_CHECK_LAUNCH("dense line fft launch");
  return 0;
}

// per-translation-unit dispatch tables (fft_passes_g*.cu); return 1 if n is not in the group
int pass_group0(PassKind kind, int n, const PassArgs& a, cudaStream_t st);
int pass_group1(PassKind kind, int n, const PassArgs& a, cudaStream_t st);
int pass_group2(PassKind kind, int n, const PassArgs& a, cudaStream_t st);
int pass_group3(PassKind kind, int n, const PassArgs& a, cudaStream_t st);
int pass_group4(PassKind kind, int n, const PassArgs& a, cudaStream_t st);
int pass_group5(PassKind kind, int n, const PassArgs& a, cudaStream_t st);
int dense_group0(int n, const DenseArgs& a, int dir, long long batch, cudaStream_t st);
int dense_group1(int n, const DenseArgs& a, int dir, long long batch, cudaStream_t st);
int dense_group2(int n, const DenseArgs& a, int dir, long long batch, cudaStream_t st);
int dense_group3(int n, const DenseArgs& a, int dir, long long batch, cudaStream_t st);
int dense_group4(int n, const DenseArgs& a, int dir, long long batch, cudaStream_t st);
int dense_group5(int n, const DenseArgs& a, int dir, long long batch, cudaStream_t st);

}  // namespace jrb
