// extern "C" entry points declared in include/jrystal_b200.h: argument checks and
// composition of the kernels into the reference's operations.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <string>

#include "plan.h"

using namespace jrb;

#define REQUIRE(cond, msg)                           \
  do {                                               \
    if (!(cond)) {                                   \
      set_error(std::string(__func__) + ": " + msg); \
      return JRB_EINVAL;                             \
    }                                                \
  } while (0)

static inline cudaStream_t S(jrb_stream s) { return reinterpret_cast<cudaStream_t>(s); }
static inline const cplx* C(const double* p) { return reinterpret_cast<const cplx*>(p); }
static inline cplx* C(double* p) { return reinterpret_cast<cplx*>(p); }

static int enter(jrb_plan* p, bool needs_grid = true) {
  if (!p) {
    set_error("null plan");
    return JRB_EINVAL;
  }
  if (needs_grid && p->ngrid == 0) {
    set_error("this entry point needs a full plan (jrb_plan_create), not a rows-only plan");
    return JRB_EINVAL;
  }
  JRB_CUDA(cudaSetDevice(p->device));
  return 0;
}

extern "C" int jrb_set_atoms(jrb_plan* p, const double* pos_h, const double* chg_h, int32_t na,
                             jrb_stream st) {
  int rc = enter(p);
  if (rc) return rc;
  return launch_set_atoms(p, pos_h, chg_h, na, S(st));
}

extern "C" int jrb_set_nonlocal(jrb_plan* p, const double* phi, int32_t nproj, jrb_stream st) {
  int rc = enter(p);
  if (rc) return rc;
  REQUIRE(nproj >= 0 && (nproj == 0 || phi), "bad projector table");
  if (p->nproj > 0) {  // give back what the previous table was charged (the band driver sets one per k-point)
    const size_t old_phi = (size_t)p->nk * p->nproj * p->ng;
    const size_t old_p = (size_t)p->ns * p->nk * p->nproj * p->nb;
    p->ws_bytes -= (int64_t)((2 * old_phi + 18 * old_p) * sizeof(cplx));
  }
  for (cplx** q : {&p->d_nl_phi, &p->d_nl_p, &p->d_nl_part, &p->d_nl_phit, &p->d_nl_ps}) {
    if (*q) cudaFree(*q);
    *q = nullptr;
  }
  p->nproj = 0;
  p->nl_p_valid = 0;
  if (nproj == 0) return 0;
  const size_t nphi = (size_t)p->nk * nproj * p->ng;
  const size_t np = (size_t)p->ns * p->nk * nproj * p->nb;
  JRB_CUDA(cudaMalloc(&p->d_nl_phi, nphi * sizeof(cplx)));
  JRB_CUDA(cudaMalloc(&p->d_nl_p, np * sizeof(cplx)));
  JRB_CUDA(cudaMalloc(&p->d_nl_part, 16 * np * sizeof(cplx)));
  JRB_CUDA(cudaMalloc(&p->d_nl_phit, nphi * sizeof(cplx)));
  JRB_CUDA(cudaMalloc(&p->d_nl_ps, np * sizeof(cplx)));
  JRB_CUDA(cudaMemcpyAsync(p->d_nl_phi, phi, nphi * sizeof(cplx), cudaMemcpyDeviceToDevice, S(st)));
  p->nproj = nproj;
  p->ws_bytes += (int64_t)((2 * nphi + 18 * np) * sizeof(cplx));
  return launch_nonlocal_transpose(p, S(st));  // conj(Phi)^T: the tall operand of the DMMA products
}

extern "C" int jrb_nonlocal_energy(jrb_plan* p, const double* q, const double* occ, double* e_nl,
                                   jrb_stream st) {
  int rc = enter(p);
  if (rc) return rc;
  REQUIRE(q && occ && e_nl, "null array");
  REQUIRE(p->nproj > 0, "call jrb_set_nonlocal first");
  JRB_CUDA(cudaMemsetAsync(e_nl, 0, sizeof(double), S(st)));
  p->nl_p_valid = 0;
  if ((rc = launch_nonlocal_project(p, 0, p->ns * p->nk, C(q), S(st)))) return rc;
  return launch_nonlocal_energy(p, occ, e_nl, S(st));
}

extern "C" int jrb_external_position_gradient(jrb_plan* p, const double* rho, double* grad,
                                              jrb_stream st) {
  int rc = enter(p);
  if (rc) return rc;
  REQUIRE(rho && grad, "null array");
  return launch_external_position_gradient(p, rho, grad, S(st));
}

extern "C" int jrb_set_external_potential(jrb_plan* p, const double* vhat, jrb_stream st) {
  int rc = enter(p);
  if (rc) return rc;
  REQUIRE(vhat, "null array");
  JRB_CUDA(cudaMemcpyAsync(p->d_vext, vhat, sizeof(cplx) * (size_t)p->ngrid, cudaMemcpyDeviceToDevice,
                           S(st)));
  if (p->natoms <= 0) p->natoms = 1;  // "an external potential is set"
  p->atoms_on_device = 0;             // no point charges behind it
  return 0;
}

extern "C" int jrb_set_kpoints(jrb_plan* p, const double* kpts_h, jrb_stream st) {
  int rc = enter(p);
  if (rc) return rc;
  REQUIRE(kpts_h, "null array");
  return launch_set_kpoints(p, kpts_h, S(st));
}

extern "C" int jrb_qr_fwd(jrb_plan* p, const double* w_re, const double* w_im, double* q,
                          double* r, jrb_stream st) {
  int rc = enter(p, false);
  if (rc) return rc;
  REQUIRE(w_re && w_im && q && r, "null array");
  return launch_qr_fwd(p, w_re, w_im, C(q), C(r), S(st));
}

extern "C" int jrb_qr_bwd(jrb_plan* p, const double* q, const double* r, const double* gq,
                          double* g_re, double* g_im, jrb_stream st) {
  int rc = enter(p, false);
  if (rc) return rc;
  REQUIRE(q && r && gq && g_re && g_im, "null array");
  return launch_qr_bwd(p, C(q), C(r), C(gq), nullptr, g_re, g_im, S(st));
}

extern "C" int jrb_qr_rows_gram(jrb_plan* p, const double* w_re, const double* w_im, int32_t pass,
                                double* s_out, jrb_stream st) {
  int rc = enter(p, false);
  if (rc) return rc;
  REQUIRE(pass == 0 || pass == 1, "pass must be 0 or 1");
  REQUIRE(s_out && (pass == 1 || (w_re && w_im)), "null array");
  if (pass == 0) JRB_CUDA(cudaMemsetAsync(p->d_scal + 32, 0, sizeof(double), S(st)));
  return launch_qr_gram_phase(p, 0, p->ns * p->nk, w_re, w_im, pass, C(s_out), S(st));
}

extern "C" int jrb_qr_rows_apply(jrb_plan* p, const double* w_re, const double* w_im, int32_t pass,
                                 double* s_inout, double* q, double* r, jrb_stream st) {
  int rc = enter(p, false);
  if (rc) return rc;
  REQUIRE(pass == 0 || pass == 1, "pass must be 0 or 1");
  REQUIRE(s_inout && (pass == 1 ? (q && r) : (w_re && w_im)), "null array");
  return launch_qr_apply_phase(p, 0, p->ns * p->nk, w_re, w_im, pass, C(s_inout), C(q), C(r), S(st));
}

extern "C" int jrb_qr_rows_bwd_gram(jrb_plan* p, const double* q, const double* gq, double* m_out,
                                    jrb_stream st) {
  int rc = enter(p, false);
  if (rc) return rc;
  REQUIRE(q && gq && m_out, "null array");
  return launch_qr_bwd_gram_phase(p, 0, p->ns * p->nk, C(q), C(gq), C(m_out), S(st));
}

extern "C" int jrb_qr_rows_bwd_apply(jrb_plan* p, const double* q, const double* gq,
                                     const double* occ, const double* m, double* g_re,
                                     double* g_im, jrb_stream st) {
  int rc = enter(p, false);
  if (rc) return rc;
  REQUIRE(q && gq && m && g_re && g_im, "null array");
  return launch_qr_bwd_apply_phase(p, 0, p->ns * p->nk, C(q), p->d_r, C(gq), occ, C(m), g_re, g_im,
                                   S(st));
}

extern "C" int jrb_check_status(jrb_plan* p, jrb_stream st) {
  int rc = enter(p, false);
  if (rc) return rc;
  JRB_CUDA(cudaStreamSynchronize(S(st)));
  int fail = 0;
  JRB_CUDA(cudaMemcpy(&fail, reinterpret_cast<int*>(p->d_scal + 32), sizeof(int),
                      cudaMemcpyDeviceToHost));
  if (fail) {
    set_error("Cholesky-QR: Gram matrix not positive definite (rank-deficient parameters)");
    return JRB_EINVAL;
  }
  if (comm_error(p, S(st))) {
    set_error("communicator: a peer rank did not arrive within JRB_COMM_TIMEOUT_S (all-reduce over peer memory)");
    return JRB_ECUDA;
  }
  return 0;
}

extern "C" int jrb_expand(jrb_plan* p, const double* q, double* dense, jrb_stream st) {
  int rc = enter(p);
  if (rc) return rc;
  REQUIRE(q && dense, "null array");
  return launch_expand(p, C(q), C(dense), S(st));
}

extern "C" int jrb_squeeze(jrb_plan* p, const double* dense, double* q, jrb_stream st) {
  int rc = enter(p);
  if (rc) return rc;
  REQUIRE(q && dense, "null array");
  return launch_squeeze(p, C(dense), C(q), S(st));
}

extern "C" int jrb_density(jrb_plan* p, const double* q, const double* occ, double* rho,
                           jrb_stream st) {
  int rc = enter(p);
  if (rc) return rc;
  REQUIRE(q && occ && rho, "null array");
  return launch_density(p, C(q), occ, rho, S(st));
}

extern "C" int jrb_kinetic(jrb_plan* p, const double* q, double* t_skb, jrb_stream st) {
  int rc = enter(p);
  if (rc) return rc;
  REQUIRE(q && t_skb, "null array");
  return launch_kinetic(p, C(q), t_skb, S(st));
}

extern "C" int jrb_grid_potential(jrb_plan* p, const double* rho, int32_t xc_id,
                                  int32_t kohn_sham, double* energies, double* veff,
                                  jrb_stream st) {
  int rc = enter(p);
  if (rc) return rc;
  REQUIRE(rho && energies && veff, "null array");
  return launch_grid_potential(p, rho, xc_id, kohn_sham, 7, energies, veff, S(st));
}

extern "C" int jrb_potential(jrb_plan* p, const double* rho, int32_t xc_id, int32_t kohn_sham,
                             int32_t parts, double* v_out, jrb_stream st) {
  int rc = enter(p);
  if (rc) return rc;
  REQUIRE(rho && v_out, "null array");
  REQUIRE(parts >= 1 && parts <= 7, "parts must be a non-empty subset of JRB_V_HARTREE|EXTERNAL|XC");
  return launch_grid_potential(p, rho, xc_id, kohn_sham, parts | 8, nullptr, v_out, S(st));
}

extern "C" int jrb_density_reciprocal(jrb_plan* p, const double* rho, double* rho_hat,
                                      jrb_stream st) {
  int rc = enter(p);
  if (rc) return rc;
  REQUIRE(rho && rho_hat, "null array");
  return launch_density_reciprocal(p, rho, C(rho_hat), S(st));
}

extern "C" int jrb_wave_grid(jrb_plan* p, const double* q, double* psi, jrb_stream st) {
  int rc = enter(p);
  if (rc) return rc;
  REQUIRE(q && psi, "null array");
  if ((rc = launch_expand(p, C(q), C(psi), S(st)))) return rc;
  // ifftn(c) * N / sqrt(Omega)   (jrystal/_src/pw.py:209-210)
  return launch_fft3d_dense(p, C(psi), C(psi), JRB_FFT_INVERSE, (int64_t)p->ns * p->nk * p->nb,
                            1.0 / std::sqrt(p->vol), S(st));
}

extern "C" int jrb_hpsi_prepare(jrb_plan* p, const double* veff, jrb_stream st) {
  int rc = enter(p);
  if (rc) return rc;
  REQUIRE(veff, "null array");
  if (veff != p->d_veff)
    JRB_CUDA(cudaMemcpyAsync(p->d_veff, veff, sizeof(double) * (size_t)p->ns * p->ngrid,
                             cudaMemcpyDeviceToDevice, S(st)));
  if ((rc = launch_hpsi_prepare(p, p->d_veff, S(st)))) return rc;
  p->veff_prepared = 1;
  return 0;
}

extern "C" int jrb_hpsi(jrb_plan* p, const double* q, const double* veff, double* hq,
                        jrb_stream st) {
  int rc = enter(p);
  if (rc) return rc;
  REQUIRE(q && hq, "null array");
  REQUIRE(q != hq, "hq must not alias q");
  REQUIRE(veff || p->veff_prepared, "veff is null and no potential was prepared (jrb_hpsi_prepare)");
  if (veff) p->veff_prepared = 0;  // the work space now holds this call's potential
  p->nl_p_valid = 0;
  return launch_hpsi(p, C(q), veff, C(hq), S(st));
}

extern "C" int jrb_band_expect(jrb_plan* p, const double* q, const double* hq, double* eps,
                               jrb_stream st) {
  int rc = enter(p);
  if (rc) return rc;
  REQUIRE(q && hq && eps, "null array");
  return launch_band_expect(p, C(q), C(hq), eps, S(st));
}

extern "C" int jrb_hamiltonian_matrix(jrb_plan* p, const double* q, const double* hq, double* h,
                                      jrb_stream st) {
  int rc = enter(p);
  if (rc) return rc;
  REQUIRE(q && hq && h, "null array");
  return launch_hamiltonian_matrix(p, C(q), C(hq), C(h), S(st));
}

extern "C" int jrb_fft3d(jrb_plan* p, const double* in, double* out, int32_t direction,
                         int64_t batch, jrb_stream st) {
  int rc = enter(p);
  if (rc) return rc;
  REQUIRE(in && out, "null array");
  REQUIRE(direction == JRB_FFT_FORWARD || direction == JRB_FFT_INVERSE, "bad direction");
  REQUIRE(batch >= 0, "negative batch");
  if (batch == 0) return 0;
  const double scale = direction == JRB_FFT_INVERSE ? 1.0 / (double)p->ngrid : 1.0;
  return launch_fft3d_dense(p, C(in), C(out), direction, batch, scale, S(st));
}

// phase boundary i of the evaluation in flight (jrb_plan_phase_timing)
static inline void phase_mark(jrb_plan* p, int i, cudaStream_t st) {
  if (p->phase_timing) cudaEventRecord(p->ev_phase[i], st);
}

// QR, density sweep, kinetic (+ non-local) energy.  reduce: all-reduce the partial density over the
// plan's communicator where it is smallest -- on the box the sweep ran on (the orbital grid),
// before the Fourier interpolation onto the plan's grid -- together with e_kin.
static int eval_begin_impl(jrb_plan* p, const double* w_re, const double* w_im, const double* occ,
                           double* rho, double* e_kin, bool reduce, bool keep_rhohat,
                           cudaStream_t st) {
  int rc = 0;
  JRB_CUDA(cudaMemsetAsync(p->d_scal + 32, 0, sizeof(double), st));  // Cholesky failure flag
  phase_mark(p, 0, st);
  if ((rc = launch_qr_fwd(p, w_re, w_im, p->d_q, p->d_r, st))) return rc;
  phase_mark(p, 1, st);
  p->keep_write = 1;
  rc = launch_density_partial(p, p->d_q, occ, rho, st);
  p->keep_write = 0;
  p->keep_filled = rc == 0;
  if (rc) return rc;
  phase_mark(p, 2, st);
  if ((rc = launch_kinetic(p, p->d_q, p->d_tkb, st))) return rc;
  if ((rc = launch_weighted_sum(p, p->d_tkb, occ, (int64_t)p->ns * p->nk * p->nb, e_kin, st)))
    return rc;
  if (p->nproj > 0) {  // e_kin carries the sphere-local one-electron terms: kinetic + non-local
    p->nl_p_valid = 0;
    if ((rc = launch_nonlocal_project(p, 0, p->ns * p->nk, p->d_q, st))) return rc;
    if ((rc = launch_nonlocal_energy(p, occ, e_kin, st))) return rc;
    p->nl_p_valid = 1;  // the H-apply of jrb_eval_finish reuses P
  }
  if (reduce && comm_world(p) > 1) {
    double* part = p->wf ? p->wf->d_rho_w : rho;
    const long long n = (long long)p->ns * (p->wf ? p->wf->ngrid : p->ngrid);
    if ((rc = launch_comm_allreduce(p, part, n, e_kin, 1, st))) return rc;
  }
  // keep_rhohat: the interpolation leaves fftn(rho) in p->d_grid for the grid part that follows
  return keep_rhohat ? launch_density_end_hat(p, rho, st) : launch_density_end(p, rho, st);
}

extern "C" int jrb_eval_begin(jrb_plan* p, const double* w_re, const double* w_im,
                              const double* occ, double* rho, double* e_kin, jrb_stream st) {
  int rc = enter(p);
  if (rc) return rc;
  REQUIRE(w_re && w_im && occ && rho && e_kin, "null array");
  return eval_begin_impl(p, w_re, w_im, occ, rho, e_kin, false, false, S(st));
}

__global__ void k_pack_energies(const double* e_kin, const double* grid_e, double* out) {
  // reference order of total_energy(split=True): kinetic, external, hartree, xc
  out[0] = e_kin[0];
  out[1] = grid_e[1];
  out[2] = grid_e[0];
  out[3] = grid_e[2];
}

// rhohat_ready: p->d_grid holds fftn(rho) (eval_begin_impl with keep_rhohat just ran)
static int eval_finish_impl(jrb_plan* p, const double* occ, const double* rho, const double* e_kin,
                            int32_t xc_id, double* energies, double* g_re, double* g_im,
                            double* g_occ, bool rhohat_ready, jrb_stream st) {
  int rc = 0;
  double* grid_e = p->d_scal;            // E_H, E_ext, E_xc
  double* veff = p->d_veff;
  p->veff_prepared = 0;                  // d_veff is overwritten by this evaluation's potential
  phase_mark(p, 3, S(st));
  if (grid_fused_ok(p, xc_id)) {
    // energies + v_eff straight onto the orbital box (each field transformed once)
    if ((rc = launch_grid_potential_orbital(p, rho, rhohat_ready, xc_id, grid_e, S(st)))) return rc;
    veff = nullptr;                      // H-apply with the potential now in the orbital plan
  } else if ((rc = launch_grid_potential(p, rho, xc_id, 0, 7, grid_e, veff, S(st)))) {
    return rc;
  }
  phase_mark(p, 4, S(st));
  p->keep_read = p->keep_filled;
  rc = launch_hpsi(p, p->d_q, veff, p->d_hq, S(st));
  p->keep_read = 0;
  phase_mark(p, 5, S(st));
  p->keep_filled = 0;  // the fused H-apply works in place on the kept columns
  p->nl_p_valid = 0;
  if (rc) return rc;
  if (g_occ) {
    if ((rc = launch_band_expect(p, p->d_q, p->d_hq, g_occ, S(st)))) return rc;
  }
  if ((rc = launch_qr_bwd(p, p->d_q, p->d_r, p->d_hq, occ, g_re, g_im, S(st)))) return rc;
  k_pack_energies<<<1, 1, 0, S(st)>>>(e_kin, grid_e, energies);
  JRB_CHECK_LAUNCH("k_pack_energies");
  phase_mark(p, 6, S(st));
  return 0;
}

extern "C" int jrb_plan_phase_timing(jrb_plan* p, int32_t enable) {
  int rc = enter(p);
  if (rc) return rc;
  if (enable && !p->ev_phase[0])
    for (int i = 0; i < 8; ++i) JRB_CUDA(cudaEventCreate(&p->ev_phase[i]));
  p->phase_timing = enable ? 1 : 0;
  return 0;
}

extern "C" int jrb_plan_phase_times(jrb_plan* p, double* ms) {
  int rc = enter(p);
  if (rc) return rc;
  REQUIRE(ms, "null array");
  REQUIRE(p->phase_timing && p->ev_phase[0], "phase timing is off (jrb_plan_phase_timing)");
  JRB_CUDA(cudaEventSynchronize(p->ev_phase[6]));
  for (int i = 0; i < 6; ++i) {
    float t = 0.f;
    JRB_CUDA(cudaEventElapsedTime(&t, p->ev_phase[i], p->ev_phase[i + 1]));
    ms[i] = (double)t;
  }
  return 0;
}

extern "C" int jrb_eval_finish(jrb_plan* p, const double* occ, const double* rho,
                               const double* e_kin, int32_t xc_id, double* energies, double* g_re,
                               double* g_im, double* g_occ, jrb_stream st) {
  int rc = enter(p);
  if (rc) return rc;
  REQUIRE(occ && rho && e_kin && energies && g_re && g_im, "null array");
  return eval_finish_impl(p, occ, rho, e_kin, xc_id, energies, g_re, g_im, g_occ, false, st);
}

extern "C" int jrb_eval(jrb_plan* p, const double* w_re, const double* w_im, const double* occ,
                        int32_t xc_id, double* energies, double* g_re, double* g_im, double* g_occ,
                        double* rho, jrb_stream st) {
  int rc = enter(p);
  if (rc) return rc;
  REQUIRE(w_re && w_im && occ && energies && g_re && g_im && rho, "null array");
  double* e_kin = p->d_scal + 40;
  const bool hat = grid_fused_ok(p, xc_id);
  if ((rc = eval_begin_impl(p, w_re, w_im, occ, rho, e_kin, true, hat, S(st)))) return rc;
  return eval_finish_impl(p, occ, rho, e_kin, xc_id, energies, g_re, g_im, g_occ, hat, st);
}

static int ensure_host_buffers(jrb_plan* p) {
  if (p->d_wre) return 0;
  const size_t nw = (size_t)p->ns * p->nk * p->ng * p->nb;
  const size_t no = (size_t)p->ns * p->nk * p->nb;
  JRB_CUDA(cudaMalloc(&p->d_wre, nw * sizeof(double)));
  JRB_CUDA(cudaMalloc(&p->d_wim, nw * sizeof(double)));
  JRB_CUDA(cudaMalloc(&p->d_gre, nw * sizeof(double)));
  JRB_CUDA(cudaMalloc(&p->d_gim, nw * sizeof(double)));
  JRB_CUDA(cudaMalloc(&p->d_occ, no * sizeof(double)));
  JRB_CUDA(cudaMalloc(&p->d_rho, (size_t)p->ns * p->ngrid * sizeof(double)));
  JRB_CUDA(cudaMalloc(&p->d_en, 8 * sizeof(double)));
  JRB_CUDA(cudaStreamCreateWithFlags(&p->h2d_stream, cudaStreamNonBlocking));
  JRB_CUDA(cudaStreamCreateWithFlags(&p->d2h_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 16; ++i) {
    JRB_CUDA(cudaEventCreateWithFlags(&p->ev_in[i], cudaEventDisableTiming));
    JRB_CUDA(cudaEventCreateWithFlags(&p->ev_out[i], cudaEventDisableTiming));
  }
  p->ws_bytes += (int64_t)((4 * nw + no + 8) * sizeof(double) + (size_t)p->ns * p->ngrid * 8);
  return 0;
}

// k-point chunks of the host path: copies of chunk c+1 overlap the kernels of chunk c
static int host_chunks(const jrb_plan* p) {
  if (p->nproj > 0) return 1;  // the chunked pipeline does not carry the non-local energy term
  // chunking pays once a chunk carries tens of MB (measured on B200, C2: 1 chunk 24.8, 4 chunks
  // 35.3, 8 chunks 33.6, 16 chunks 27.3 eval/s end to end); tiny problems stay in one piece
  // (a k-sharded rank of an 8-GPU C2 run holds 71 MB: two to four chunks still hide most of the copies)
  const double bytes = 16.0 * p->nk * (double)p->ng * p->nb;  // w_re + w_im
  int n = p->ns == 1 ? (int)std::min(4.0, bytes / (16.0 * 1024 * 1024)) : 1;
  if (const char* env = std::getenv("JRB_HOST_CHUNKS")) n = std::atoi(env);
  if (p->ns != 1) n = 1;
  return std::max(1, std::min(std::min(n, 16), p->nk));
}

extern "C" int jrb_energy_grad_host(jrb_plan* p, const double* w_re_h, const double* w_im_h,
                                    const double* occ_h, int32_t xc_id, double* energies_h,
                                    double* g_re_h, double* g_im_h, double* rho_h) {
  int rc = enter(p);
  if (rc) return rc;
  REQUIRE(w_re_h && w_im_h && occ_h && energies_h && g_re_h && g_im_h, "null array");
  if ((rc = ensure_host_buffers(p))) return rc;
  cudaStream_t st = p->own_stream;
  const size_t no = (size_t)p->ns * p->nk * p->nb * sizeof(double);
  const size_t per_k = (size_t)p->ng * p->nb;  // parameters per k-point
  double* rho = p->d_rho;
  double* e_kin = p->d_en + 4;
  const int nch = host_chunks(p);
  JRB_CUDA(cudaMemcpyAsync(p->d_occ, occ_h, no, cudaMemcpyHostToDevice, st));
  if (nch == 1) {
    const size_t nw = (size_t)p->ns * p->nk * per_k * sizeof(double);
    JRB_CUDA(cudaMemcpyAsync(p->d_wre, w_re_h, nw, cudaMemcpyHostToDevice, st));
    JRB_CUDA(cudaMemcpyAsync(p->d_wim, w_im_h, nw, cudaMemcpyHostToDevice, st));
    const bool hat = grid_fused_ok(p, xc_id);
    if ((rc = eval_begin_impl(p, p->d_wre, p->d_wim, p->d_occ, rho, e_kin, true, hat, st))) return rc;
    if ((rc = eval_finish_impl(p, p->d_occ, rho, e_kin, xc_id, p->d_en, p->d_gre, p->d_gim, nullptr,
                               hat, st)))
      return rc;
    JRB_CUDA(cudaMemcpyAsync(g_re_h, p->d_gre, nw, cudaMemcpyDeviceToHost, st));
    JRB_CUDA(cudaMemcpyAsync(g_im_h, p->d_gim, nw, cudaMemcpyDeviceToHost, st));
  } else {
    // forward: H2D of chunk c+1 (copy stream) under QR + density of chunk c (compute stream)
    auto k_of = [&](int c) { return (int)((long long)c * p->nk / nch); };
    for (int c = 0; c < nch; ++c) {
      const size_t off = (size_t)k_of(c) * per_k, n = (size_t)(k_of(c + 1) - k_of(c)) * per_k;
      JRB_CUDA(cudaMemcpyAsync(p->d_wre + off, w_re_h + off, n * sizeof(double),
                               cudaMemcpyHostToDevice, p->h2d_stream));
      JRB_CUDA(cudaMemcpyAsync(p->d_wim + off, w_im_h + off, n * sizeof(double),
                               cudaMemcpyHostToDevice, p->h2d_stream));
      JRB_CUDA(cudaEventRecord(p->ev_in[c], p->h2d_stream));
    }
    if ((rc = launch_focc(p, p->d_occ, st))) return rc;
    JRB_CUDA(cudaMemsetAsync(p->d_scal + 32, 0, sizeof(double), st));  // Cholesky failure flag
    if ((rc = launch_density_begin(p, rho, st))) return rc;
    for (int c = 0; c < nch; ++c) {
      const int k0 = k_of(c), k1 = k_of(c + 1);
      JRB_CUDA(cudaStreamWaitEvent(st, p->ev_in[c], 0));
      if ((rc = launch_qr_fwd_range(p, k0, k1 - k0, p->d_wre, p->d_wim, p->d_q, p->d_r, st))) return rc;
      p->keep_write = 1;
      rc = launch_density_krange(p, p->d_q, rho, k0, k1, st);
      p->keep_write = 0;
      if (rc) return rc;
      if ((rc = launch_kinetic_range(p, k0, k1 - k0, p->d_q, p->d_tkb, st))) return rc;
    }
    if ((rc = launch_weighted_sum(p, p->d_tkb, p->d_occ, (int64_t)p->nk * p->nb, e_kin, st))) return rc;
    if (comm_world(p) > 1) {  // k-sharded ranks: the one collective, on the box the sweep ran on
      double* part = p->wf ? p->wf->d_rho_w : rho;
      const long long n = (long long)p->ns * (p->wf ? p->wf->ngrid : p->ngrid);
      if ((rc = launch_comm_allreduce(p, part, n, e_kin, 1, st))) return rc;
    }
    const bool hat = grid_fused_ok(p, xc_id);
    if ((rc = hat ? launch_density_end_hat(p, rho, st) : launch_density_end(p, rho, st))) return rc;
    // backward: D2H of chunk c (copy stream) under H-apply + QR adjoint of chunk c+1
    double* grid_e = p->d_scal;
    p->veff_prepared = 0;
    if (hat) {
      if ((rc = launch_grid_potential_orbital(p, rho, true, xc_id, grid_e, st))) return rc;
    } else {
      if ((rc = launch_grid_potential(p, rho, xc_id, 0, 7, grid_e, p->d_veff, st))) return rc;
      if ((rc = launch_hpsi_prepare(p, p->d_veff, st))) return rc;
    }
    for (int c = 0; c < nch; ++c) {
      const int k0 = k_of(c), k1 = k_of(c + 1);
      const size_t off = (size_t)k0 * per_k, n = (size_t)(k1 - k0) * per_k;
      p->keep_read = 1;
      rc = launch_hpsi_krange(p, p->d_q, p->d_veff, p->d_hq, k0, k1, st);
      p->keep_read = 0;
      if (rc) return rc;
      if ((rc = launch_qr_bwd_range(p, k0, k1 - k0, p->d_q, p->d_r, p->d_hq, p->d_occ, p->d_gre,
                                    p->d_gim, st)))
        return rc;
      JRB_CUDA(cudaEventRecord(p->ev_out[c], st));
      JRB_CUDA(cudaStreamWaitEvent(p->d2h_stream, p->ev_out[c], 0));
      JRB_CUDA(cudaMemcpyAsync(g_re_h + off, p->d_gre + off, n * sizeof(double),
                               cudaMemcpyDeviceToHost, p->d2h_stream));
      JRB_CUDA(cudaMemcpyAsync(g_im_h + off, p->d_gim + off, n * sizeof(double),
                               cudaMemcpyDeviceToHost, p->d2h_stream));
    }
    k_pack_energies<<<1, 1, 0, st>>>(e_kin, grid_e, p->d_en);
    JRB_CHECK_LAUNCH("k_pack_energies");
  }
  JRB_CUDA(cudaMemcpyAsync(energies_h, p->d_en, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (rho_h)
    JRB_CUDA(cudaMemcpyAsync(rho_h, rho, (size_t)p->ns * p->ngrid * sizeof(double),
                             cudaMemcpyDeviceToHost, st));
  JRB_CUDA(cudaStreamSynchronize(st));
  if (nch > 1) JRB_CUDA(cudaStreamSynchronize(p->d2h_stream));
  int fail = 0;
  JRB_CUDA(cudaMemcpy(&fail, reinterpret_cast<int*>(p->d_scal + 32), sizeof(int),
                      cudaMemcpyDeviceToHost));
  if (fail) {
    set_error("Cholesky-QR: Gram matrix not positive definite (rank-deficient parameters)");
    return JRB_EINVAL;
  }
  return 0;
}
