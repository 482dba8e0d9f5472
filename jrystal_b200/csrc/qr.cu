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
// Small (nb x nb) per-(spin,k) work: one CTA each, operating in global/L2 memory.
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) {  // a * conj(b)
  return cmake(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

// S = sum_chunks partial; in-place Cholesky S = L L^H; R = L^H; Rinv = R^-1.
// If r_prev != nullptr (second pass): r_out = R * r_prev, rinv_out = rinv_prev * Rinv.
__global__ void __launch_bounds__(1024)
k_chol_inv(const cplx* __restrict__ partial, int nchunks, int nb, cplx* __restrict__ S,
           cplx* __restrict__ Rt, cplx* __restrict__ Rit, const cplx* __restrict__ r_prev,
           const cplx* __restrict__ rinv_prev, cplx* __restrict__ r_out,
           cplx* __restrict__ rinv_out, cplx* __restrict__ scratch, int* __restrict__ fail_flag) {
  const int sk = blockIdx.x, nsk = gridDim.x;
  const long long nn = (long long)nb * nb;
  S += sk * nn; Rt += sk * nn; Rit += sk * nn;
  r_out += sk * nn; rinv_out += sk * nn; scratch += sk * nn;
  const int tid = threadIdx.x, nt = blockDim.x;
  // k_gram writes only the blocks on and above the block diagonal: mirror (Hermitian)
  for (long long e = tid; e < nn; e += nt) {
    const int i = (int)(e / nb), j = (int)(e % nb);
    const long long src = j >= i ? e : (long long)j * nb + i;
    cplx s = cmake(0.0, 0.0);
    for (int c = 0; c < nchunks; ++c) {
      const cplx v = partial[((long long)c * nsk + sk) * nn + src];
      s.x += v.x; s.y += v.y;
    }
    S[e] = j >= i ? s : cconj(s);
  }
  __syncthreads();
  // right-looking Cholesky on the lower triangle
  for (int j = 0; j < nb; ++j) {
    const double djj = S[(long long)j * nb + j].x;
    if (!(djj > 0.0)) {
      if (tid == 0) atomicExch(fail_flag, 1);
    }
    const double d = sqrt(djj > 0.0 ? djj : 1.0);
    __syncthreads();
    for (int i = j + 1 + tid; i < nb; i += nt) {
      cplx v = S[(long long)i * nb + j];
      S[(long long)i * nb + j] = cmake(v.x / d, v.y / d);
    }
    if (tid == 0) S[(long long)j * nb + j] = cmake(d, 0.0);
    __syncthreads();
    const int m = nb - j - 1;
    for (long long e = tid; e < (long long)m * m; e += nt) {
      const int i = j + 1 + (int)(e / m), k = j + 1 + (int)(e % m);
      if (k <= i) {
        const cplx p = cmulc(S[(long long)i * nb + j], S[(long long)k * nb + j]);
        cplx v = S[(long long)i * nb + k];
        S[(long long)i * nb + k] = cmake(v.x - p.x, v.y - p.y);
      }
    }
    __syncthreads();
  }
  // R = L^H (upper), zeros below
  for (long long e = tid; e < nn; e += nt) {
    const int i = (int)(e / nb), j = (int)(e % nb);
    Rt[e] = j >= i ? cconj(S[(long long)j * nb + i]) : cmake(0.0, 0.0);
    Rit[e] = cmake(0.0, 0.0);
  }
  __syncthreads();
  // Rinv by back substitution, one warp per column
  const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
  for (int j = warp; j < nb; j += nwarps) {
    if (lane == 0) Rit[(long long)j * nb + j] = cmake(1.0 / Rt[(long long)j * nb + j].x, 0.0);
    __syncwarp();
    for (int i = j - 1; i >= 0; --i) {
      double sx = 0.0, sy = 0.0;
      for (int k = i + 1 + lane; k <= j; k += 32) {
        const cplx p = cmul(Rt[(long long)i * nb + k], Rit[(long long)k * nb + j]);
        sx += p.x; sy += p.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
      }
      if (lane == 0) {
        const double rii = Rt[(long long)i * nb + i].x;
        Rit[(long long)i * nb + j] = cmake(-sx / rii, -sy / rii);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (r_prev == nullptr) {
    for (long long e = tid; e < nn; e += nt) {
      r_out[e] = Rt[e];
      rinv_out[e] = Rit[e];
    }
  } else {
    r_prev += sk * nn; rinv_prev += sk * nn;
    // both products are upper triangular
    for (long long e = tid; e < nn; e += nt) {
      const int i = (int)(e / nb), j = (int)(e % nb);
      cplx a = cmake(0.0, 0.0), b = cmake(0.0, 0.0);
      if (j >= i) {
        for (int k = i; k <= j; ++k) {
          const cplx p = cmul(Rt[(long long)i * nb + k], r_prev[(long long)k * nb + j]);
          a.x += p.x; a.y += p.y;
          const cplx q = cmul(rinv_prev[(long long)i * nb + k], Rit[(long long)k * nb + j]);
          b.x += q.x; b.y += q.y;
        }
      }
      S[e] = a;       // staged (L is dead by now): r_out may alias r_prev
      scratch[e] = b;
    }
    __syncthreads();
    for (long long e = tid; e < nn; e += nt) {
      r_out[e] = S[e];
      rinv_out[e] = scratch[e];
    }
  }
}

// Rinv = R^-1 for an upper-triangular R supplied by the caller (jrb_qr_bwd with foreign r).
__global__ void __launch_bounds__(1024)
k_tri_inv(const cplx* __restrict__ R, int nb, cplx* __restrict__ Rinv) {
  const long long nn = (long long)nb * nb;
  R += blockIdx.x * nn; Rinv += blockIdx.x * nn;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (long long e = tid; e < nn; e += nt) Rinv[e] = cmake(0.0, 0.0);
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
  for (int j = warp; j < nb; j += nwarps) {
    if (lane == 0) {
      const cplx d = R[(long long)j * nb + j];
      const double n2 = d.x * d.x + d.y * d.y;
      Rinv[(long long)j * nb + j] = cmake(d.x / n2, -d.y / n2);
    }
    __syncwarp();
    for (int i = j - 1; i >= 0; --i) {
      double sx = 0.0, sy = 0.0;
      for (int k = i + 1 + lane; k <= j; k += 32) {
        const cplx p = cmul(R[(long long)i * nb + k], Rinv[(long long)k * nb + j]);
        sx += p.x; sy += p.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
      }
      if (lane == 0) {
        const cplx d = R[(long long)i * nb + i];
        const double n2 = d.x * d.x + d.y * d.y;
        // -(s / d)
        Rinv[(long long)i * nb + j] =
          cmake(-(sx * d.x + sy * d.y) / n2, -(sy * d.x - sx * d.y) / n2);
      }
      __syncwarp();
    }
  }
}

// Backward small step: M = (sum partial) diag(f); X = -(up(M) + up(M)^H + diag Re M);
// T1 = diag(f) Rinv^H ; T2 = X Rinv^H.
__global__ void __launch_bounds__(1024)
k_bwd_small(const cplx* __restrict__ partial, int nchunks, int nb, const double* __restrict__ occ,
            const cplx* __restrict__ rinv, cplx* __restrict__ X, cplx* __restrict__ T1,
            cplx* __restrict__ T2) {
  const int sk = blockIdx.x, nsk = gridDim.x;
  const long long nn = (long long)nb * nb;
  X += sk * nn; T1 += sk * nn; T2 += sk * nn; rinv += sk * nn;
  const double* f = occ ? occ + (long long)sk * nb : nullptr;
  const int tid = threadIdx.x, nt = blockDim.x;
  // M staged in T2
  for (long long e = tid; e < nn; e += nt) {
    const int i = (int)(e / nb), j = (int)(e % nb);
    cplx s = cmake(0.0, 0.0);
    if (j >= i) {  // k_gram computes only the upper blocks; only up(M) and diag are used
      for (int c = 0; c < nchunks; ++c) {
        const cplx v = partial[((long long)c * nsk + sk) * nn + e];
        s.x += v.x; s.y += v.y;
      }
    }
    const double fj = f ? f[j] : 1.0;
    T2[e] = cmake(s.x * fj, s.y * fj);
  }
  __syncthreads();
  for (long long e = tid; e < nn; e += nt) {
    const int i = (int)(e / nb), j = (int)(e % nb);
    cplx x;
    if (j > i) {
      const cplx m = T2[(long long)i * nb + j];
      x = cmake(-m.x, -m.y);
    } else if (j < i) {
      const cplx m = T2[(long long)j * nb + i];
      x = cmake(-m.x, m.y);
    } else {
      x = cmake(-T2[e].x, 0.0);
    }
    X[e] = x;
    // T1[i][j] = f_i conj(rinv[j][i])
    const double fi = f ? f[i] : 1.0;
    const cplx r = rinv[(long long)j * nb + i];
    T1[e] = cmake(fi * r.x, -fi * r.y);
  }
  __syncthreads();
  // T2[i][j] = sum_k X[i][k] conj(rinv[j][k]); rinv upper triangular -> k >= j
  for (long long e = tid; e < nn; e += nt) {
    const int i = (int)(e / nb), j = (int)(e % nb);
    cplx s = cmake(0.0, 0.0);
    for (int k = j; k < nb; ++k) {
      const cplx p = cmulc(X[(long long)i * nb + k], rinv[(long long)j * nb + k]);
      s.x += p.x; s.y += p.y;
    }
    T2[e] = s;
  }
}

// ---------------------------------------------------------------------------------------
static int gram_chunks(const jrb_plan* p, int tiles) {
  // CTAs = upper super-tile pairs x chunks x (spin,k); two CTAs are resident per SM.  Pick the
  // chunk count (>= 256 rows each) whose last wave is fullest, preferring >= 2 waves.
  const int nsk = p->ns * p->nk;
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

int qr_gram_chunks(const jrb_plan* p) {
  const int tiles = (p->nb + QT - 1) / QT;
  return gram_chunks(p, tiles);
}

template <class K>
static int opt_in_smem(K kernel, int bytes) {
  if (bytes > 48 * 1024)
    JRB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return 0;
}

// partial = upper blocks of A^H B, split over row chunks
static int run_gram(jrb_plan* p, TallMat A, TallMat B, bool same, cplx* partial, int* nchunks_out,
                    cudaStream_t st) {
  constexpr int ST = 3;
  const int nsk = p->ns * p->nk;
  const int tiles = (p->nb + QT - 1) / QT;
  const int nchunks = gram_chunks(p, tiles);
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
static int run_apply_ncb(jrb_plan* p, TallMat in1, const cplx* t1, int tri1, TallMat in2,
                         const cplx* t2, int tri2, int nterms, double* out_a, double* out_b,
                         cudaStream_t st) {
  constexpr int ST = 2;
  const int nsk = p->ns * p->nk;
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
static int run_apply(jrb_plan* p, TallMat in1, const cplx* t1, int tri1, TallMat in2,
                     const cplx* t2, int tri2, int nterms, double* out_a, double* out_b,
                     cudaStream_t st) {
  const int nb = p->nb;
  const bool split = in1.im != nullptr;
  if (MODE == 0 && split) {
    if (nb <= 32)
      return run_apply_ncb<0, 4, true>(p, in1, t1, tri1, in2, t2, tri2, nterms, out_a, out_b, st);
    return run_apply_ncb<0, 9, true>(p, in1, t1, tri1, in2, t2, tri2, nterms, out_a, out_b, st);
  }
  if (nb <= 32)
    return run_apply_ncb<MODE, 4, false>(p, in1, t1, tri1, in2, t2, tri2, nterms, out_a, out_b, st);
  return run_apply_ncb<MODE, 9, false>(p, in1, t1, tri1, in2, t2, tri2, nterms, out_a, out_b, st);
}

// H[sk][i][j] = sum_chunks partial (upper blocks), mirrored as a Hermitian matrix
__global__ void __launch_bounds__(256)
k_sum_partials_herm(const cplx* __restrict__ partial, int nchunks, int nb, cplx* __restrict__ H) {
  const int sk = blockIdx.x, nsk = gridDim.x;
  const long long nn = (long long)nb * nb;
  for (long long e = threadIdx.x; e < nn; e += blockDim.x) {
    const int i = (int)(e / nb), j = (int)(e % nb);
    const long long src = j >= i ? e : (long long)j * nb + i;
    cplx s = cmake(0.0, 0.0);
    for (int c = 0; c < nchunks; ++c) {
      const cplx v = partial[((long long)c * nsk + sk) * nn + src];
      s.x += v.x; s.y += v.y;
    }
    H[sk * nn + e] = j >= i ? s : cconj(s);
  }
}

// H_ij = <q_i | hq_j> for Hermitian H (hamiltonian.hamiltonian_matrix): one Gram on DMMA
int launch_hamiltonian_matrix(jrb_plan* p, const cplx* q, const cplx* hq, cplx* h, cudaStream_t st) {
  int nchunks = 0, rc = 0;
  TallMat Q{reinterpret_cast<const double*>(q), nullptr, p->nb};
  TallMat G{reinterpret_cast<const double*>(hq), nullptr, p->nb};
  if ((rc = run_gram(p, Q, G, false, p->d_gpart, &nchunks, st))) return rc;
  k_sum_partials_herm<<<p->ns * p->nk, 256, 0, st>>>(p->d_gpart, nchunks, p->nb, h);
  JRB_CHECK_LAUNCH("k_sum_partials_herm");
  return 0;
}

int launch_qr_fwd(jrb_plan* p, const double* w_re, const double* w_im, cplx* q, cplx* r,
                  cudaStream_t st) {
  const int nsk = p->ns * p->nk;
  const long long nn = (long long)p->nb * p->nb;
  cplx* S = p->d_small;                 // slot 0: Gram / Cholesky work
  cplx* Rt = p->d_small + nsk * nn;     // slot 1
  cplx* Rit = p->d_small + 2 * nsk * nn;  // slot 2
  int* fail = reinterpret_cast<int*>(p->d_scal + 32);
  int nchunks = 0, rc = 0;
  TallMat W{w_re, w_im, p->nb};
  TallMat none{nullptr, nullptr, 0};
  // pass 1
  if ((rc = run_gram(p, W, W, true, p->d_gpart, &nchunks, st))) return rc;
  cplx* R2inv = p->d_small + 3 * nsk * nn;  // slot 3: staging, then R2^-1 for the last apply
  k_chol_inv<<<nsk, 1024, 0, st>>>(p->d_gpart, nchunks, p->nb, S, Rt, Rit, nullptr, nullptr, r,
                                   p->d_rinv, R2inv, fail);
  JRB_CHECK_LAUNCH("k_chol_inv");
  if ((rc = run_apply<0>(p, W, p->d_rinv, TRI_UPPER, none, nullptr, TRI_FULL, 1,
                         reinterpret_cast<double*>(p->d_tmp), nullptr, st)))
    return rc;
  // pass 2
  TallMat Q1{reinterpret_cast<const double*>(p->d_tmp), nullptr, p->nb};
  if ((rc = run_gram(p, Q1, Q1, true, p->d_gpart, &nchunks, st))) return rc;
  k_chol_inv<<<nsk, 1024, 0, st>>>(p->d_gpart, nchunks, p->nb, S, Rt, Rit, r, p->d_rinv, r,
                                   p->d_rinv, R2inv, fail);
  JRB_CHECK_LAUNCH("k_chol_inv");
  // Rit still holds R2^-1 (only Rt and S were reused for staging)
  JRB_CUDA(cudaMemcpyAsync(R2inv, Rit, sizeof(cplx) * nsk * nn, cudaMemcpyDeviceToDevice, st));
  if ((rc = run_apply<0>(p, Q1, R2inv, TRI_UPPER, none, nullptr, TRI_FULL, 1,
                         reinterpret_cast<double*>(q), nullptr, st)))
    return rc;
  return 0;
}

int launch_qr_bwd(jrb_plan* p, const cplx* q, const cplx* r, const cplx* gq, const double* occ,
                  double* g_re, double* g_im, cudaStream_t st) {
  const int nsk = p->ns * p->nk;
  const long long nn = (long long)p->nb * p->nb;
  const cplx* rinv = p->d_rinv;  // R^-1 of the plan's own forward call
  if (r != p->d_r) {
    cplx* ri = p->d_small + 3 * nsk * nn;
    k_tri_inv<<<nsk, 1024, 0, st>>>(r, p->nb, ri);
    JRB_CHECK_LAUNCH("k_tri_inv");
    rinv = ri;
  }
  cplx* X = p->d_small;
  cplx* T1 = p->d_small + nsk * nn;
  cplx* T2 = p->d_small + 2 * nsk * nn;
  int nchunks = 0, rc = 0;
  TallMat Q{reinterpret_cast<const double*>(q), nullptr, p->nb};
  TallMat G{reinterpret_cast<const double*>(gq), nullptr, p->nb};
  if ((rc = run_gram(p, Q, G, false, p->d_gpart, &nchunks, st))) return rc;
  k_bwd_small<<<nsk, 1024, 0, st>>>(p->d_gpart, nchunks, p->nb, occ, rinv, X, T1, T2);
  JRB_CHECK_LAUNCH("k_bwd_small");
  return run_apply<1>(p, G, T1, TRI_LOWER, Q, T2, TRI_FULL, 2, g_re, g_im, st);
}

}  // namespace jrb
