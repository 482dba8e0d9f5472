// Non-local pseudopotential term on the cut-off sphere.
//
// Reference: hamiltonian_nonlocal / energy_nonlocal (jrystal/pseudopotential/nloc.py:143-158,
// 217-236) contract the coefficients with potential_nl_psi_reciprocal Phi[k, beta, m, x, y, z]
// (= 4 pi sqrt(D) Y_lm beta(|G+k|) exp(-i (G+k).R) i^l, nloc.py:60-141) over the WHOLE FFT box:
//     F[s,k,b,p] = sum_G c[s,k,b,G] Phi[k,p,G]          (no conjugation, p = (beta, m) flattened)
//     H_nl[b1,b2] = sum_p conj(F[b1,p]) F[b2,p] / vol,   E_nl = sum_skb f_skb H_nl[b,b]
// c is zero outside the sphere, so only Phi on the sphere matters: the caller hands over
// Phi[k][p][g] (nk x nproj x ng complex, 1.3 GB at C4 instead of 68 GB of dense projectors).  Then
//     P[s,k,p,b] = sum_g Phi[k,p,g] Q[s,k,g,b]            (k_nl_project, split over row chunks)
//     E_nl       = sum f_b |P[p,b]|^2 / vol               (k_nl_energy)
//     HQ[g,b]   += sum_p conj(Phi[k,p,g]) P[p,b] / vol    (k_nl_apply)  = dE_nl/dQ* per unit f
// Both products run on the FP64 tensor cores (DMMA) through the rectangular forms of the QR
// kernels (qr_gemm.cuh: k_gram with GramRect, k_apply with ApplyRect) on Phit = conj(Phi)^T, the
// [ng][nproj] tall operand built once by jrb_set_nonlocal; the FP64-FMA tile kernels below are
// kept as the reference path (JRB_NL_FMA=1).
#include <algorithm>
#include <cstdlib>

#include "plan.h"

namespace jrb {

constexpr int NL_TP = 16;   // projectors per tile
constexpr int NL_TB = 32;   // bands per tile
constexpr int NL_TG = 32;   // sphere rows per smem slice
constexpr int NL_CHUNKS = 16;

// partial[chunk][sk][p][b] = sum_{g in chunk} Phi[k][p][g] Q[sk][g][b]
// CTA tile 32 projectors x 32 bands, thread (tx, ty) of (16, 16) owns the 2 x 2 outputs
// (ty, ty + 16) x (tx, tx + 16): 4 shared-memory loads per 4 complex FMAs.
// grid: (ceil(nproj/32) * ceil(nb/32), nsk, NL_CHUNKS), block (16, 16)
__global__ void __launch_bounds__(256)
k_nl_project(const cplx* __restrict__ phi, const cplx* __restrict__ q, long long ng, int nb,
             int nproj, int nk, int sk0, cplx* __restrict__ partial) {
  __shared__ cplx sphi[2 * NL_TP][NL_TG + 1];
  __shared__ cplx sq[NL_TG][NL_TB + 1];
  const int nbt = (nb + NL_TB - 1) / NL_TB;
  const int p0 = (blockIdx.x / nbt) * 2 * NL_TP, b0 = (blockIdx.x % nbt) * NL_TB;
  const int sk = blockIdx.y, k = (sk0 + sk) % nk;  // q / partial are indexed from sk0
  const long long rows = (ng + gridDim.z - 1) / gridDim.z;
  const long long g_lo = blockIdx.z * rows, g_hi = min(ng, g_lo + rows);
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * 16 + tx;
  cplx acc[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc[i][j] = cmake(0.0, 0.0);
  for (long long g0 = g_lo; g0 < g_hi; g0 += NL_TG) {
    for (int e = tid; e < 2 * NL_TP * NL_TG; e += 256) {  // Phi tile: 32 x 32, g fastest in memory
      const int pp = e / NL_TG, gg = e % NL_TG;
      const bool ok = p0 + pp < nproj && g0 + gg < g_hi;
      sphi[pp][gg] = ok ? phi[((long long)k * nproj + p0 + pp) * ng + g0 + gg] : cmake(0.0, 0.0);
    }
    for (int e = tid; e < NL_TG * NL_TB; e += 256) {  // Q tile: 32 x 32, b fastest
      const int gg = e / NL_TB, bb = e % NL_TB;
      const bool ok = g0 + gg < g_hi && b0 + bb < nb;
      sq[gg][bb] = ok ? q[((long long)sk * ng + g0 + gg) * nb + b0 + bb] : cmake(0.0, 0.0);
    }
    __syncthreads();
#pragma unroll 8
    for (int gg = 0; gg < NL_TG; ++gg) {
      const cplx f0 = sphi[ty][gg], f1 = sphi[ty + 16][gg];
      const cplx q0 = sq[gg][tx], q1 = sq[gg][tx + 16];
      cplx v;
      v = cmul(f0, q0); acc[0][0].x += v.x; acc[0][0].y += v.y;
      v = cmul(f0, q1); acc[0][1].x += v.x; acc[0][1].y += v.y;
      v = cmul(f1, q0); acc[1][0].x += v.x; acc[1][0].y += v.y;
      v = cmul(f1, q1); acc[1][1].x += v.x; acc[1][1].y += v.y;
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int pp = p0 + ty + 16 * i, bb = b0 + tx + 16 * j;
      if (pp < nproj && bb < nb)
        partial[(((long long)blockIdx.z * gridDim.y + sk) * nproj + pp) * nb + bb] = acc[i][j];
    }
}

// P = sum of the chunk partials (fixed order), Ps = P / vol.  grid: (ceil(nsk*nproj*nb / 256))
__global__ void k_nl_reduce(const cplx* __restrict__ partial, int nchunks, long long n,
                            cplx* __restrict__ P, cplx* __restrict__ Ps, double inv_vol) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  cplx s = cmake(0.0, 0.0);
  for (int c = 0; c < nchunks; ++c) {
    const cplx v = partial[c * n + i];
    s.x += v.x; s.y += v.y;
  }
  P[i] = s;
  Ps[i] = cmake(s.x * inv_vol, s.y * inv_vol);
}

// Phit[k][g][p] = conj(Phi[k][p][g]) through a 32 x 32 shared-memory tile.
// grid: (ceil(ng / 32), ceil(nproj / 32), nk), block (32, 8)
__global__ void __launch_bounds__(256)
k_nl_transpose(const cplx* __restrict__ phi, long long ng, int nproj, cplx* __restrict__ phit) {
  __shared__ cplx tile[32][33];
  const long long g0 = (long long)blockIdx.x * 32;
  const int p0 = blockIdx.y * 32, k = blockIdx.z;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const long long g = g0 + threadIdx.x;
    const int pp = p0 + j;
    tile[j][threadIdx.x] = (g < ng && pp < nproj) ? phi[((long long)k * nproj + pp) * ng + g] : cmake(0.0, 0.0);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const long long g = g0 + j;
    const int pp = p0 + threadIdx.x;
    if (g < ng && pp < nproj) phit[((long long)k * ng + g) * nproj + pp] = cconj(tile[threadIdx.x][j]);
  }
}

int launch_nonlocal_transpose(jrb_plan* p, cudaStream_t st) {
  dim3 grid((unsigned)((p->ng + 31) / 32), (p->nproj + 31) / 32, p->nk), block(32, 8);
  k_nl_transpose<<<grid, block, 0, st>>>(p->d_nl_phi, p->ng, p->nproj, p->d_nl_phit);
  JRB_CHECK_LAUNCH("k_nl_transpose");
  return 0;
}

static bool nl_use_fma() {
  static int v = [] {
    const char* env = std::getenv("JRB_NL_FMA");
    return env ? std::atoi(env) : 0;
  }();
  return v != 0;
}

// part[blk] = sum over a slice of (sk, p, b) of occ[sk][b] |P|^2; then one CTA adds the slices in
// order: e_out[0] += sum / vol (deterministic)
__global__ void __launch_bounds__(256)
k_nl_energy(const cplx* __restrict__ P, const double* __restrict__ occ, int nsk, int nproj, int nb,
            double* __restrict__ part) {
  __shared__ double sh[256];
  double acc = 0.0;
  const long long n = (long long)nsk * nproj * nb;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int b = (int)(i % nb);
    const int sk = (int)(i / ((long long)nproj * nb));
    const cplx v = P[i];
    acc += occ[(long long)sk * nb + b] * (v.x * v.x + v.y * v.y);
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}
__global__ void k_nl_energy_final(const double* __restrict__ part, int n, double inv_vol,
                                  double* __restrict__ e_out) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += part[i];
  e_out[0] += s * inv_vol;
}

// hq[sk][g][b] += sum_p conj(Phi[k][p][g]) P[sk][p][b] / vol
// grid: (ceil(ng/32), ceil(nb/32), nsk), block (32, 8): thread (b, g-sub), 4 rows each
__global__ void __launch_bounds__(256)
k_nl_apply(const cplx* __restrict__ phi, const cplx* __restrict__ P, long long ng, int nb,
           int nproj, int nk, int sk0, double inv_vol, cplx* __restrict__ hq) {
  __shared__ cplx sphi[NL_TP][NL_TG + 1];
  __shared__ cplx sp[NL_TP][NL_TB + 1];
  const long long g0 = (long long)blockIdx.x * NL_TG;
  const int b0 = blockIdx.y * NL_TB;
  const int sk = blockIdx.z, k = (sk0 + sk) % nk;  // P / hq are indexed from sk0
  const int tb = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * 32 + tb;
  cplx acc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] = cmake(0.0, 0.0);
  for (int p0 = 0; p0 < nproj; p0 += NL_TP) {
    for (int e = tid; e < NL_TP * NL_TG; e += 256) {
      const int pp = e / NL_TG, gg = e % NL_TG;
      const bool ok = p0 + pp < nproj && g0 + gg < ng;
      sphi[pp][gg] = ok ? phi[((long long)k * nproj + p0 + pp) * ng + g0 + gg] : cmake(0.0, 0.0);
    }
    for (int e = tid; e < NL_TP * NL_TB; e += 256) {
      const int pp = e / NL_TB, bb = e % NL_TB;
      const bool ok = p0 + pp < nproj && b0 + bb < nb;
      sp[pp][bb] = ok ? P[((long long)sk * nproj + p0 + pp) * nb + b0 + bb] : cmake(0.0, 0.0);
    }
    __syncthreads();
#pragma unroll
    for (int pp = 0; pp < NL_TP; ++pp) {
      const cplx pv = sp[pp][tb];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const cplx f = sphi[pp][ty + 8 * j];
        acc[j].x += f.x * pv.x + f.y * pv.y;   // conj(f) * pv
        acc[j].y += f.x * pv.y - f.y * pv.x;
      }
    }
    __syncthreads();
  }
  if (b0 + tb < nb) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long g = g0 + ty + 8 * j;
      if (g < ng) {
        cplx* o = hq + ((long long)sk * ng + g) * nb + b0 + tb;
        const cplx h = *o;
        *o = cmake(h.x + acc[j].x * inv_vol, h.y + acc[j].y * inv_vol);
      }
    }
  }
}

// P of the (spin,k) range [sk0, sk0 + nsk) into p->d_nl_p (indexed from sk0)
int launch_nonlocal_project(jrb_plan* p, int sk0, int nsk, const cplx* q, cudaStream_t st) {
  int nchunks = NL_CHUNKS;
  if (nl_use_fma()) {
    const int nbt = (p->nb + NL_TB - 1) / NL_TB, npt = (p->nproj + 2 * NL_TP - 1) / (2 * NL_TP);
    dim3 grid(npt * nbt, nsk, NL_CHUNKS), block(16, 16);
    k_nl_project<<<grid, block, 0, st>>>(p->d_nl_phi, q + (long long)sk0 * p->ng * p->nb, p->ng, p->nb,
                                        p->nproj, p->nk, sk0, p->d_nl_part);
    JRB_CHECK_LAUNCH("k_nl_project");
  } else {
    // P = Phi Q = conj(Phit)^T Q on DMMA; the k-point of item sk0 + i is (sk0 + i) % nk: the range
    // never wraps a spin boundary in the middle of a call unless it starts at a multiple of nk
    nchunks = (int)std::max<long long>(1, std::min<long long>(NL_CHUNKS, p->ng / 256));
    const int k0 = sk0 % p->nk;
    int rc = launch_gram_rect(p, nsk, p->d_nl_phit + (long long)k0 * p->ng * p->nproj, p->nproj,
                              p->nk - k0 >= nsk ? nsk : p->nk, q + (long long)sk0 * p->ng * p->nb,
                              nchunks, p->d_nl_part, st);
    if (rc) return rc;
  }
  const long long n = (long long)nsk * p->nproj * p->nb;
  k_nl_reduce<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
    p->d_nl_part, nchunks, n, p->d_nl_p + (long long)sk0 * p->nproj * p->nb,
    p->d_nl_ps + (long long)sk0 * p->nproj * p->nb, 1.0 / p->vol);
  JRB_CHECK_LAUNCH("k_nl_reduce");
  return 0;
}

int launch_nonlocal_energy(jrb_plan* p, const double* occ, double* e_inout, cudaStream_t st) {
  // partial sums live at the head of the (by now consumed) chunk buffer of the projection
  double* part = reinterpret_cast<double*>(p->d_nl_part);
  const int blocks = 148;
  k_nl_energy<<<blocks, 256, 0, st>>>(p->d_nl_p, occ, p->ns * p->nk, p->nproj, p->nb, part);
  JRB_CHECK_LAUNCH("k_nl_energy");
  k_nl_energy_final<<<1, 1, 0, st>>>(part, blocks, 1.0 / p->vol, e_inout);
  JRB_CHECK_LAUNCH("k_nl_energy_final");
  return 0;
}

int launch_nonlocal_apply(jrb_plan* p, int sk0, int nsk, cplx* hq, cudaStream_t st) {
  if (!nl_use_fma()) {
    // hq += Phi^H P / vol = Phit (P / vol) on DMMA
    const int k0 = sk0 % p->nk;
    return launch_apply_rect(p, nsk, p->d_nl_phit + (long long)k0 * p->ng * p->nproj, p->nproj,
                             p->nk - k0 >= nsk ? nsk : p->nk,
                             p->d_nl_ps + (long long)sk0 * p->nproj * p->nb,
                             hq + (long long)sk0 * p->ng * p->nb, st);
  }
  dim3 grid((unsigned)((p->ng + NL_TG - 1) / NL_TG), (p->nb + NL_TB - 1) / NL_TB, nsk), block(32, 8);
  k_nl_apply<<<grid, block, 0, st>>>(p->d_nl_phi, p->d_nl_p + (long long)sk0 * p->nproj * p->nb,
                                     p->ng, p->nb, p->nproj, p->nk, sk0, 1.0 / p->vol,
                                     hq + (long long)sk0 * p->ng * p->nb);
  JRB_CHECK_LAUNCH("k_nl_apply");
  return 0;
}

}  // namespace jrb
