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
k_chol_blocked(cplx* __restrict__ S, cplx* __restrict__ Rt, int nb, int* __restrict__ fail_flag) {
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
k_tri_inv_cols(const cplx* __restrict__ R, int nb, cplx* __restrict__ Rinv) {
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
static int gram_chunks(const jrb_plan* p, int tiles, int nsk) {
  // CTAs = upper super-tile pairs x chunks x (spin,k); two CTAs are resident per SM.  Pick the
  // chunk count (>= 256 rows each) whose last wave is fullest, preferring >= 2 waves.
  const int per_chunk = nsk * (tiles * (tiles + 1) / 2);
  const int slots = 2 * 148;
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
  if (tiles == 1) {
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

static int chol_and_inverse(jrb_plan* p, int nsk, int nchunks, cplx* S, cplx* Rt, cplx* Rit,
                            cudaStream_t st) {
  const int nb = p->nb;
  dim3 egrid((unsigned)(((long long)nb * nb + SMALL_T - 1) / SMALL_T), nsk);
  k_gram_reduce<<<egrid, SMALL_T, 0, st>>>(p->d_gpart, nchunks, nb, S);
  JRB_CHECK_LAUNCH("k_gram_reduce");
  const int smem = nb * (CHOL_PB + 1 + CHOL_KC + 1) * (int)sizeof(cplx);
  static int once = opt_in_smem(k_chol_blocked, 200 * 1024);
  if (once) return once;
  if (smem > 200 * 1024) {
    set_error("Cholesky panel does not fit shared memory (too many bands)");
    return JRB_EUNSUPPORTED;
  }
  int* fail = reinterpret_cast<int*>(p->d_scal + 32);
  k_chol_blocked<<<nsk, CHOL_T, smem, st>>>(S, Rt, nb, fail);
  JRB_CHECK_LAUNCH("k_chol_blocked");
  dim3 igrid((nb + 7) / 8, nsk);
  k_tri_inv_cols<<<igrid, 256, 8 * nb * (int)sizeof(cplx), st>>>(Rt, nb, Rit);
  JRB_CHECK_LAUNCH("k_tri_inv_cols");
  return 0;
}

// Cholesky-QR2 of the (spin,k) range [sk0, sk0 + nsk).  q / r are indexed from sk0.
int launch_qr_fwd_range(jrb_plan* p, int sk0, int nsk, const double* w_re, const double* w_im,
                        cplx* q, cplx* r, cudaStream_t st) {
  const int nb = p->nb;
  const long long nn = (long long)nb * nb;
  const long long nall = (long long)p->ns * p->nk * nn;
  const long long soff = (long long)sk0 * p->ng * nb, moff = (long long)sk0 * nn;
  cplx* S = p->d_small + moff;                // slot 0: Gram / Cholesky work
  cplx* Rt = p->d_small + nall + moff;        // slot 1: R of the current pass
  cplx* Rit = p->d_small + 2 * nall + moff;   // slot 2: its inverse
  cplx* R1 = p->d_small + 3 * nall + moff;    // slot 3: R1 (first pass), kept for the compose
  cplx* R1inv = p->d_small + 4 * nall + moff; // slot 4
  cplx* tmp = p->d_tmp + soff;
  cplx* rinv = p->d_rinv + moff;
  int nchunks = 0, rc = 0;
  TallMat W{w_re + soff, w_im + soff, nb};
  TallMat none{nullptr, nullptr, 0};
  // pass 1: R1, Q1 = W R1^-1
  if ((rc = run_gram(p, nsk, W, W, true, p->d_gpart, &nchunks, st))) return rc;
  if ((rc = chol_and_inverse(p, nsk, nchunks, S, R1, R1inv, st))) return rc;
  if ((rc = run_apply<0>(p, nsk, W, R1inv, TRI_UPPER, none, nullptr, TRI_FULL, 1,
                         reinterpret_cast<double*>(tmp), nullptr, st)))
    return rc;
  // pass 2: R2, Q = Q1 R2^-1; R = R2 R1, R^-1 = R1^-1 R2^-1
  TallMat Q1{reinterpret_cast<const double*>(tmp), nullptr, nb};
  if ((rc = run_gram(p, nsk, Q1, Q1, true, p->d_gpart, &nchunks, st))) return rc;
  if ((rc = chol_and_inverse(p, nsk, nchunks, S, Rt, Rit, st))) return rc;
  dim3 egrid((unsigned)((nn + SMALL_T - 1) / SMALL_T), nsk);
  k_tri_compose<<<egrid, SMALL_T, 0, st>>>(Rt, Rit, R1, R1inv, nb, r + moff, rinv);
  JRB_CHECK_LAUNCH("k_tri_compose");
  return run_apply<0>(p, nsk, Q1, Rit, TRI_UPPER, none, nullptr, TRI_FULL, 1,
                      reinterpret_cast<double*>(q + soff), nullptr, st);
}

int launch_qr_fwd(jrb_plan* p, const double* w_re, const double* w_im, cplx* q, cplx* r,
                  cudaStream_t st) {
  return launch_qr_fwd_range(p, 0, p->ns * p->nk, w_re, w_im, q, r, st);
}

// Adjoint for the (spin,k) range [sk0, sk0 + nsk); all arrays indexed from (spin,k) 0.
int launch_qr_bwd_range(jrb_plan* p, int sk0, int nsk, const cplx* q, const cplx* r,
                        const cplx* gq, const double* occ, double* g_re, double* g_im,
                        cudaStream_t st) {
  const int nb = p->nb;
  const long long nn = (long long)nb * nb;
  const long long nall = (long long)p->ns * p->nk * nn;
  const long long soff = (long long)sk0 * p->ng * nb, moff = (long long)sk0 * nn;
  const cplx* rinv = p->d_rinv + moff;  // R^-1 of the plan's own forward call
  if (r != p->d_r) {
    cplx* ri = p->d_small + 3 * nall + moff;
    dim3 igrid((nb + 7) / 8, nsk);
    k_tri_inv_cols<<<igrid, 256, 8 * nb * (int)sizeof(cplx), st>>>(r + moff, nb, ri);
    JRB_CHECK_LAUNCH("k_tri_inv_cols");
    rinv = ri;
  }
  cplx* X = p->d_small + moff;
  cplx* T1 = p->d_small + nall + moff;
  cplx* T2 = p->d_small + 2 * nall + moff;
  int nchunks = 0, rc = 0;
  TallMat Q{reinterpret_cast<const double*>(q + soff), nullptr, nb};
  TallMat G{reinterpret_cast<const double*>(gq + soff), nullptr, nb};
  if ((rc = run_gram(p, nsk, Q, G, false, p->d_gpart, &nchunks, st))) return rc;
  dim3 egrid((unsigned)((nn + SMALL_T - 1) / SMALL_T), nsk);
  k_bwd_x<<<egrid, SMALL_T, 0, st>>>(p->d_gpart, nchunks, nb, occ ? occ + (long long)sk0 * nb : nullptr,
                                     rinv, X, T1);
  JRB_CHECK_LAUNCH("k_bwd_x");
  k_bwd_t2<<<egrid, SMALL_T, 0, st>>>(X, rinv, nb, T2);
  JRB_CHECK_LAUNCH("k_bwd_t2");
  return run_apply<1>(p, nsk, G, T1, TRI_LOWER, Q, T2, TRI_FULL, 2, g_re + soff, g_im + soff, st);
}

int launch_qr_bwd(jrb_plan* p, const cplx* q, const cplx* r, const cplx* gq, const double* occ,
                  double* g_re, double* g_im, cudaStream_t st) {
  return launch_qr_bwd_range(p, 0, p->ns * p->nk, q, r, gq, occ, g_re, g_im, st);
}

}  // namespace jrb
