// Host-side orchestration of the pencil passes: batching of band groups, spin handling,
// and the dense 3-pass transform behind jrb_fft3d.
#include <algorithm>

#include <cstdlib>

#include "fft_fused.cuh"
#include "fft_passes.cuh"
#include "plan.h"

namespace jrb {

static int run_pass(PassKind kind, int n, const PassArgs& a, cudaStream_t st) {
  int rc = pass_group0(kind, n, a, st);
  if (rc == 1) rc = pass_group1(kind, n, a, st);
  if (rc == 1) rc = pass_group2(kind, n, a, st);
  if (rc == 1) rc = pass_group3(kind, n, a, st);
  if (rc == 1) rc = pass_group4(kind, n, a, st);
  if (rc == 1) rc = pass_group5(kind, n, a, st);
  if (rc == 1) {
    set_error("no compiled pencil pass for axis length " + std::to_string(n));
    return JRB_EUNSUPPORTED;
  }
  return rc;
}

static int run_dense(int n, const DenseArgs& a, int dir, long long batch, cudaStream_t st) {
  int rc = dense_group0(n, a, dir, batch, st);
  if (rc == 1) rc = dense_group1(n, a, dir, batch, st);
  if (rc == 1) rc = dense_group2(n, a, dir, batch, st);
  if (rc == 1) rc = dense_group3(n, a, dir, batch, st);
  if (rc == 1) rc = dense_group4(n, a, dir, batch, st);
  if (rc == 1) rc = dense_group5(n, a, dir, batch, st);
  if (rc == 1) {
    set_error("no compiled dense line FFT for axis length " + std::to_string(n));
    return JRB_EUNSUPPORTED;
  }
  return rc;
}

int fused128_launch(int kind, const FusedArgs& a, int ctas, cudaStream_t st);
int fused128_smem(int nxo, int ncol);
int fused128_rho_reduce(const FusedArgs& a, int ctas, double* rho, cudaStream_t st);
int fused_smem_group0(int n, int nxo, int ncol);
int fused_smem_group1(int n, int nxo, int ncol);
int fused_smem_group2(int n, int nxo, int ncol);
int fused_threads_group0(int n);
int fused_threads_group1(int n);
int fused_threads_group2(int n);
int fused_psi_plane_group0(int n, int nxo);
int fused_psi_plane_group1(int n, int nxo);
int fused_psi_plane_group2(int n, int nxo);

// complex numbers per band-plane of the psi(r) cache for axis length n (< 0: length not
// compiled, or the cached H-apply kernel does not fit shared memory)
int fused_psi_plane_elems(int n, int nxo) {
  int r = fused_psi_plane_group0(n, nxo);
  if (r == -2) r = fused_psi_plane_group1(n, nxo);
  if (r == -2) r = fused_psi_plane_group2(n, nxo);
  return r;
}

int fused_smem_need(int n, int nxo, int ncol) {
  int r = fused_smem_group0(n, nxo, ncol);
  if (r < 0) r = fused_smem_group1(n, nxo, ncol);
  if (r < 0) r = fused_smem_group2(n, nxo, ncol);
  return r;
}

// The fused y+x kernels need nx == ny, a compiled size and slabs that fit shared memory.
bool fused_available(int nx, int ny, int nxo, int ncol) {
  if (const char* env = std::getenv("JRB_NO_FUSE"))
    if (std::atoi(env) != 0) return false;
  if (nx != ny) return false;
  const int need = fused_smem_need(nx, nxo, ncol);
  // column and Y-row indices are packed in 16 bits each (fft_fused.cuh)
  return need > 0 && need <= 200 * 1024 && ncol < 65000 && nxo * (nx + 8) < 65000;
}

// The 128 x 128 variant: occupied frequencies |f| < 32 in x and y, everything in shared memory.
bool fused128_available(int nx, int ny, int nxo, int ncol, int band_limited32) {
  if (const char* env = std::getenv("JRB_NO_FUSE"))
    if (std::atoi(env) != 0) return false;
  return nx == 128 && ny == 128 && band_limited32 && nxo <= 64 && ncol < 65000 &&
         fused128_smem(nxo, ncol) <= 226 * 1024;
}

// persistent CTAs of the fused kernels: one per resident slot
int fused_cta_count(int n, int nxo, int ncol) {
  int nt = fused_threads_group0(n);
  if (nt < 0) nt = fused_threads_group1(n);
  if (nt < 0) nt = fused_threads_group2(n);
  const int smem = fused_smem_need(n, nxo, ncol);
  int per_sm = nt <= 256 ? 2 : 1;
  if (2 * (smem + 1024) > 227 * 1024) per_sm = 1;
  if (const char* env = std::getenv("JRB_FUSED_CTAS_PER_SM")) per_sm = std::max(1, std::atoi(env));
  return 148 * per_sm;
}

static int run_fused(int kind, int n, const FusedArgs& a, int ctas, cudaStream_t st) {
  int rc = fused_group0(kind, n, a, ctas, st);
  if (rc == 1) rc = fused_group1(kind, n, a, ctas, st);
  if (rc == 1) rc = fused_group2(kind, n, a, ctas, st);
  if (rc == 1) {
    set_error("no compiled fused pass for axis length " + std::to_string(n));
    return JRB_EUNSUPPORTED;
  }
  return rc;
}

// segments (distinct z planes) one persistent CTA can touch for a batch of `ngroups`
static int fused_segmax(const jrb_plan* p, int ngroups) {
  const long long W = (long long)p->nz * (p->fused == 2 ? 2 : 1) * ngroups;
  const long long items = (W + p->fused_ctas - 1) / p->fused_ctas;
  return (int)((items - 1) / ngroups + 2);
}

static FusedArgs fused_args(jrb_plan* p, const PassArgs& a) {
  FusedArgs f{};
  f.m = a.m;
  f.wa = a.wa;
  f.tw = p->d_tw_x;
  f.tw64 = p->d_tw_half;
  // fused == 2: the work space holds three copies of A (input, y-parity 0 and 1 outputs)
  f.wout[0] = p->d_ws_a + p->a_copy_elems;
  f.wout[1] = p->d_ws_a + 2 * p->a_copy_elems;
  f.focc = a.focc;
  f.veff = a.veff;
  f.rho_part = p->d_rho_part;
  f.seg_z = p->d_seg_z;
  f.segmax = fused_segmax(p, a.ngroups);
  f.nb = a.nb;
  f.ngpk = a.ngpk;
  f.g0 = a.g0;
  f.ngroups = a.ngroups;
  f.vscale = a.vscale;
  f.band_limited = p->band_limited;
  if (const char* env = std::getenv("JRB_NO_SPARSE"))
    if (std::atoi(env) != 0) f.band_limited = 0;
  // TMA staging: which tensor map covers a.wa, and the row (16 doubles = 8 bands) it starts at
  f.tmap = nullptr;
  f.row0 = 0;
  if (p->d_tmaps) {
    const char* maps = static_cast<const char*>(p->d_tmaps);
    const long long keep_elems = (long long)p->ns * p->nk * p->ngroups_per_k * p->a_group_elems;
    if (p->d_a_keep && a.wa >= p->d_a_keep && a.wa < p->d_a_keep + keep_elems) {
      f.tmap = maps + 128;
      f.row0 = (a.wa - p->d_a_keep) / NB;
    } else if (a.wa >= p->d_ws_a && a.wa < p->d_ws_a + p->a_copy_elems) {
      f.tmap = maps;
      f.row0 = (a.wa - p->d_ws_a) / NB;
    }
  }
  // psi(r) cache: written by the density sweep of jrb_eval_begin, read by the H-apply of
  // jrb_eval_finish (the keep_* flags bracket exactly those two sweeps)
  f.psi = (p->d_psi && (p->keep_write || p->keep_read)) ? p->d_psi : nullptr;
  return f;
}

static PassArgs base_args(jrb_plan* p) {
  PassArgs a{};
  a.m = p->maps;
  a.wa = p->d_ws_a;
  a.wb = p->d_ws_b;
  a.focc = p->d_focc;
  a.gk2 = p->d_gk2;
  a.nb = p->nb;
  a.nk = p->nk;
  a.ngpk = p->ngroups_per_k;
  a.vscale = 1.0 / (double)p->ngrid;
  return a;
}

// Adds sum_{groups in [ga, gb)} occ |psi|^2 of spin s to rho_spin (d_focc must be current).
static int density_groups(jrb_plan* p, const cplx* q, double* rho_spin, int s, int ga, int gb,
                          cudaStream_t st) {
  int rc = 0;
  if (p->wf) {  // the orbital grid's plan runs the passes, into its own density
    p->wf->keep_write = p->keep_write;
    rc = density_groups(p->wf, q, p->wf->d_rho_w + (size_t)s * p->wf->ngrid, s, ga, gb, st);
    p->wf->keep_write = 0;
    return rc;
  }
  const int per_spin = p->nk * p->ngroups_per_k;
  for (int g0 = ga; g0 < gb; g0 += p->batch_groups) {
    PassArgs a = base_args(p);
    a.q = q;
    a.g0 = s * per_spin + g0;
    a.ngroups = std::min(p->batch_groups, gb - g0);
    a.rho = rho_spin;
    a.tw = p->d_tw_z;
    if (p->keep_write && p->d_a_keep && !p->d_psi)
      a.wa = p->d_a_keep + (long long)a.g0 * p->a_group_elems;
    if ((rc = run_pass(PASS_Z_INV_SCATTER, p->nz, a, st))) return rc;
    if (p->fused == 2) {
      FusedArgs f = fused_args(p, a);
      if ((rc = fused128_launch(0, f, p->fused_ctas, st))) return rc;
      if ((rc = fused128_rho_reduce(f, p->fused_ctas, a.rho, st))) return rc;
      continue;
    }
    if (p->fused) {
      FusedArgs f = fused_args(p, a);
      if ((rc = run_fused(0, p->nx, f, p->fused_ctas, st))) return rc;
      dim3 grid((p->nx * p->ny + 31) / 32, p->nz), block(32, 8);
      k_rho_reduce<<<grid, block, 0, st>>>(p->d_rho_part, p->d_seg_z, p->fused_ctas, f.segmax,
                                           a.ngroups, p->nx * p->ny, p->nz, a.rho);
      JRB_CHECK_LAUNCH("k_rho_reduce");
      continue;
    }
    a.tw = p->d_tw_y;
    if ((rc = run_pass(PASS_Y_INV, p->ny, a, st))) return rc;
    a.tw = p->d_tw_x;
    if ((rc = run_pass(PASS_X_DENSITY, p->nx, a, st))) return rc;
  }
  return 0;
}

int launch_density_begin(jrb_plan* p, double* rho, cudaStream_t st) {
  if (p->wf) {
    JRB_CUDA(cudaMemsetAsync(p->wf->d_rho_w, 0, sizeof(double) * (size_t)p->ns * p->wf->ngrid, st));
  } else {
    JRB_CUDA(cudaMemsetAsync(rho, 0, sizeof(double) * (size_t)p->ns * p->ngrid, st));
  }
  return 0;
}

// Orbital grid -> the plan's grid: rho is band limited to 2 gmax, which both boxes hold, so the
// Fourier interpolation reproduces the reference's rho(r_j) on the plan's grid exactly.
int launch_density_end(jrb_plan* p, double* rho, cudaStream_t st) {
  jrb_plan* w = p->wf;
  if (!w) return 0;
  int rc = 0;
  for (int s = 0; s < p->ns; ++s) {
    if ((rc = launch_real_to_complex(w->d_rho_w + (size_t)s * w->ngrid, w->ngrid, w->d_grid, st)))
      return rc;
    if ((rc = launch_fft3d_dense(w, w->d_grid, w->d_grid, JRB_FFT_FORWARD, 1, 1.0, st))) return rc;
    if ((rc = launch_resample(w->d_grid, w->nx, w->ny, w->nz, p->d_grid, p->nx, p->ny, p->nz,
                              1.0 / (double)w->ngrid, st)))
      return rc;
    if ((rc = launch_fft3d_dense(p, p->d_grid, p->d_grid, JRB_FFT_INVERSE, 1, 1.0, st))) return rc;
    if ((rc = launch_complex_to_real(p->d_grid, p->ngrid, rho + (size_t)s * p->ngrid, st))) return rc;
  }
  return 0;
}

// v_eff on the plan's grid -> the orbital grid: only the Fourier components |f| <= 2 gmax of the
// potential reach the sphere part of v_eff psi, so truncating to the orbital box changes nothing
// the Hamiltonian apply returns.
int launch_hpsi_prepare(jrb_plan* p, const double* veff, cudaStream_t st) {
  jrb_plan* w = p->wf;
  if (!w) return 0;
  int rc = 0;
  for (int s = 0; s < p->ns; ++s) {
    if ((rc = launch_real_to_complex(veff + (size_t)s * p->ngrid, p->ngrid, p->d_grid, st))) return rc;
    if ((rc = launch_fft3d_dense(p, p->d_grid, p->d_grid, JRB_FFT_FORWARD, 1, 1.0, st))) return rc;
    if ((rc = launch_resample(p->d_grid, p->nx, p->ny, p->nz, w->d_grid, w->nx, w->ny, w->nz,
                              1.0 / (double)p->ngrid, st)))
      return rc;
    if ((rc = launch_fft3d_dense(w, w->d_grid, w->d_grid, JRB_FFT_INVERSE, 1, 1.0, st))) return rc;
    if ((rc = launch_complex_to_real(w->d_grid, w->ngrid, w->d_veff + (size_t)s * w->ngrid, st)))
      return rc;
  }
  return 0;
}

// rho[s] = sum_{k,b} occ |psi|^2  (jrb_density).
int launch_density_partial(jrb_plan* p, const cplx* q, const double* occ, double* rho, cudaStream_t st) {
  int rc = launch_focc(p, occ, st);
  if (rc) return rc;
  if ((rc = launch_density_begin(p, rho, st))) return rc;
  const int per_spin = p->nk * p->ngroups_per_k;
  for (int s = 0; s < p->ns; ++s)
    if ((rc = density_groups(p, q, rho + (size_t)s * p->ngrid, s, 0, per_spin, st))) return rc;
  return 0;
}

int launch_density(jrb_plan* p, const cplx* q, const double* occ, double* rho, cudaStream_t st) {
  int rc = launch_density_partial(p, q, occ, rho, st);
  if (rc) return rc;
  return launch_density_end(p, rho, st);
}

// the k-points [k0, k1) of spin 0 only, accumulated into rho (no memset, no focc refresh)
int launch_density_krange(jrb_plan* p, const cplx* q, double* rho, int k0, int k1, cudaStream_t st) {
  return density_groups(p, q, rho, 0, k0 * p->ngroups_per_k, k1 * p->ngroups_per_k, st);
}

// veff_spin: on the grid of the plan that runs the passes (launch_hpsi_prepare for a child)
static int hpsi_groups(jrb_plan* p, const cplx* q, const double* veff_spin, cplx* hq, int s, int ga,
                       int gb, cudaStream_t st) {
  int rc = 0;
  if (p->wf) {
    jrb_plan* w = p->wf;
    w->keep_read = p->keep_read;
    rc = hpsi_groups(w, q, w->d_veff + (size_t)s * w->ngrid, hq, s, ga, gb, st);
    w->keep_read = 0;
    if (rc) return rc;
  }
  const int per_spin = p->nk * p->ngroups_per_k;
  for (int g0 = ga; g0 < gb && !p->wf; g0 += p->batch_groups) {
    PassArgs a = base_args(p);
    a.q = q;
    a.hq = hq;
    a.g0 = s * per_spin + g0;
    a.ngroups = std::min(p->batch_groups, gb - g0);
    a.veff = veff_spin;
    a.tw = p->d_tw_z;
    const bool from_psi = p->keep_read && p->d_psi && p->fused == 1;  // psi(r) of the density sweep
    if (from_psi) {
      // nothing to recompute: the cached kernel only writes the columns
    } else if (p->keep_read && p->d_a_keep) {
      a.wa = p->d_a_keep + (long long)a.g0 * p->a_group_elems;  // columns of the density sweep
    } else {
      if ((rc = run_pass(PASS_Z_INV_SCATTER, p->nz, a, st))) return rc;
    }
    if (p->fused == 2) {
      FusedArgs f = fused_args(p, a);
      if ((rc = fused128_launch(1, f, p->fused_ctas, st))) return rc;
      a.wa = f.wout[0];
      a.wa_add = f.wout[1];
    } else if (p->fused) {
      FusedArgs f = fused_args(p, a);
      if ((rc = run_fused(from_psi ? 2 : 1, p->nx, f, p->fused_ctas, st))) return rc;
    } else {
      a.tw = p->d_tw_y;
      if ((rc = run_pass(PASS_Y_INV, p->ny, a, st))) return rc;
      a.tw = p->d_tw_x;
      if ((rc = run_pass(PASS_X_VMUL, p->nx, a, st))) return rc;
      a.tw = p->d_tw_y;
      if ((rc = run_pass(PASS_Y_FWD, p->ny, a, st))) return rc;
    }
    a.tw = p->d_tw_z;
    if ((rc = run_pass(PASS_Z_FWD_GATHER, p->nz, a, st))) return rc;
  }
  if (p->nproj > 0) {
    // non-local pseudopotential: hq += Phi^H (Phi q) / vol on the (spin,k) items of [ga, gb)
    const int sk_lo = s * p->nk + ga / p->ngroups_per_k;
    const int sk_hi = s * p->nk + (gb + p->ngroups_per_k - 1) / p->ngroups_per_k;
    if (!(p->nl_p_valid && q == p->d_q))
      if ((rc = launch_nonlocal_project(p, sk_lo, sk_hi - sk_lo, q, st))) return rc;
    if ((rc = launch_nonlocal_apply(p, sk_lo, sk_hi - sk_lo, hq, st))) return rc;
  }
  return 0;
}

// hq = 1/2|G+k|^2 q + (sqrt(Omega)/N) fftn(veff psi)|mask   (jrb_hpsi)
// veff == nullptr: the potential of the last jrb_hpsi_prepare (band mode: v_eff[rho_gs] is fixed over
// thousands of steps, so it is copied / resampled once instead of on every call)
int launch_hpsi(jrb_plan* p, const cplx* q, const double* veff, cplx* hq, cudaStream_t st) {
  int rc = 0;
  if (veff == nullptr) {
    veff = p->d_veff;
  } else if ((rc = launch_hpsi_prepare(p, veff, st))) {
    return rc;
  }
  const int per_spin = p->nk * p->ngroups_per_k;
  for (int s = 0; s < p->ns; ++s)
    if ((rc = hpsi_groups(p, q, veff + (size_t)s * p->ngrid, hq, s, 0, per_spin, st))) return rc;
  return 0;
}

int launch_hpsi_krange(jrb_plan* p, const cplx* q, const double* veff, cplx* hq, int k0, int k1,
                       cudaStream_t st) {
  return hpsi_groups(p, q, veff, hq, 0, k0 * p->ngroups_per_k, k1 * p->ngroups_per_k, st);
}

// Dense batched 3-D transform over (nx, ny, nz), C order; `scale` applied on the last pass.
int launch_fft3d_dense(jrb_plan* p, const cplx* in, cplx* out, int dir, int64_t batch,
                       double scale, cudaStream_t st) {
  const long long nx = p->nx, ny = p->ny, nz = p->nz;
  int rc = 0;
  for (int64_t b0 = 0; b0 < batch; b0 += 32768) {
    const long long nbat = std::min<int64_t>(32768, batch - b0);
    const cplx* src = in + (size_t)b0 * p->ngrid;
    cplx* dst = out + (size_t)b0 * p->ngrid;
    DenseArgs a{};
    a.batch_stride = p->ngrid;
    // z lines: lanes over y (stride nz), one line group set per x
    a.in = src; a.out = dst; a.tw = p->d_tw_z;
    a.n0 = (int)nx; a.stride0 = ny * nz; a.lane_total = (int)ny; a.lane_stride = nz;
    a.elem_stride = 1; a.scale = 1.0;
    if ((rc = run_dense((int)nz, a, dir, nbat, st))) return rc;
    // y lines: lanes over z
    a.in = dst; a.out = dst; a.tw = p->d_tw_y;
    a.n0 = (int)nx; a.stride0 = ny * nz; a.lane_total = (int)nz; a.lane_stride = 1;
    a.elem_stride = nz; a.scale = 1.0;
    if ((rc = run_dense((int)ny, a, dir, nbat, st))) return rc;
    // x lines: lanes over z
    a.in = dst; a.out = dst; a.tw = p->d_tw_x;
    a.n0 = (int)ny; a.stride0 = nz; a.lane_total = (int)nz; a.lane_stride = 1;
    a.elem_stride = ny * nz; a.scale = scale;
    if ((rc = run_dense((int)nx, a, dir, nbat, st))) return rc;
  }
  return 0;
}

}  // namespace jrb
