// Internal plan object behind the opaque `jrb_plan` of include/jrystal_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/jrystal_b200.h"
#include "dft_small.cuh"

namespace jrb {

constexpr int NB = 8;  // band lanes per group: 8 x 16 B = one 128-byte line

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

#define JRB_CUDA(expr)                                     \
  do {                                                     \
    cudaError_t e__ = (expr);                              \
    if (e__ != cudaSuccess) return ::jrb::cuda_fail(e__, #expr); \
  } while (0)

void count_launch();

#define JRB_CHECK_LAUNCH(name)                             \
  do {                                                     \
    ::jrb::count_launch();                                 \
    cudaError_t e__ = cudaGetLastError();                  \
    if (e__ != cudaSuccess) return ::jrb::cuda_fail(e__, name); \
  } while (0)

// Index maps derived from the frequency mask (all device pointers).
struct SphereMaps {
  int nx, ny, nz;
  int ncol;   // (x, y) columns that contain at least one kept plane wave
  int nxo;    // x planes that contain at least one such column
  int64_t ng;
  const int32_t* zmap;  // [ncol][nz]  compact index g of (col, z) or -1
  const int32_t* ycol;  // [nxo][ny]   column id of (x plane, y) or -1
  const int32_t* xmap;  // [nx]        x-plane id of x or -1
  const int32_t* gidx;  // [ng]        linear index into the dense (nx,ny,nz) box
};

}  // namespace jrb

struct jrb_plan {
  int nx, ny, nz, ns, nk, nb;
  int64_t ng, ngrid;
  int device;
  int ngroups_per_k;  // ceil(nb / NB)
  int batch_groups;
  double cell[9], recip[9], vol;
  jrb::SphereMaps maps;
  // owned device tables
  int32_t *d_zmap, *d_ycol, *d_xmap, *d_gidx;
  double* d_gk2;                 // [nk][ng]  |G+k|^2
  jrb::cplx *d_tw_x, *d_tw_y, *d_tw_z;  // exp(-2 pi i t / n)
  // pencil work space
  jrb::cplx* d_ws_a;  // [batch][ncol][nz][NB] (three copies when fused == 2)
  long long a_copy_elems;  // elements of one copy
  // z-transformed columns of EVERY band group, kept from the density sweep of jrb_eval_begin for
  // the H-apply of jrb_eval_finish (saves the second scatter + z pass); null when it would exceed
  // the budget (JRB_KEEP_A_MB, default 8192) or the passes are not fused
  jrb::cplx* d_a_keep;
  long long a_group_elems;
  int keep_write, keep_read, keep_filled;
  // psi(r) of EVERY orbital on the box the passes run on, in the register layout of the fused
  // x stage (fft_fused.cuh: FusedArgs::psi): stored by the density sweep of jrb_eval_begin, read
  // by the H-apply of jrb_eval_finish, which then runs only the forward transforms.  Takes
  // precedence over d_a_keep; null when the passes are not fused (fused != 1) or it exceeds the
  // budget (JRB_PSI_CACHE_MB, default 65536, and at most 40 % of the free device memory)
  jrb::cplx* d_psi;
  long long psi_group_elems;  // complex numbers per band group (nz planes x 8 bands)
  jrb::cplx* d_ws_b;  // [batch][nxo][ny][nz][NB]
  double* d_focc;     // [ns*nk*ngroups_per_k][NB] occupation / Omega, zero padded
  int fused;          // 1: y and x passes fused per z-plane (fft_fused.cuh); B slab unused;
                      // 2: the 128 x 128 variant (fft_fused128.cuh)
  int band_limited32; // occupied x and y indices all in [0, 32) u [n - 32, n)
  jrb::cplx* d_tw_half;  // exp(-2 pi i t / (nx / 2)) (fused == 2)
  // TMA tensor maps (two CUtensorMap objects in device memory) of d_ws_a and d_a_keep viewed as
  // [rows = (group, z, column)][8 bands x (re, im)] doubles; null: cp.async staging (plan.cu)
  void* d_tmaps;
  int fused_ctas;     // persistent CTAs of the fused kernels (resident slots)
  int band_limited;   // occupied x and y indices all in [0, 16) u [n - 16, n) (sparse radix-8 butterflies)
  int fused_segmax;   // partial density planes per CTA (upper bound over batch sizes)
  double* d_rho_part; // [fused_ctas * fused_segmax][nx*ny] partial density planes
  int* d_seg_z;       // [fused_ctas * fused_segmax] z of each partial plane or -1
  // grid work space
  jrb::cplx* d_grid;      // [ngrid] dense complex grid
  jrb::cplx* d_vext;      // [ngrid] V_ext(G) incl. the reference's -N/Omega factor
  double* d_partials;     // block partial sums for the grid reductions
  double* d_veff;         // [ns][ngrid] effective potential of the fused evaluation
  jrb::cplx* d_gga;       // [3][ngrid] gradient components of rho (GGA); null on an orbital-grid child
  double* d_vxc;          // [ngrid] local part of the GGA potential
  int n_partial_blocks;
  int natoms;
  double *d_pos, *d_chg;   // atoms of the last jrb_set_atoms (position gradient of E_ext)
  double* d_atom_part;     // [natoms][64][3] block partials of that gradient
  int atoms_on_device;
  // non-local pseudopotential projectors on the sphere (nonlocal.cu); nproj == 0: none
  int nproj;
  jrb::cplx* d_nl_phi;   // [nk][nproj][ng]
  jrb::cplx* d_nl_p;     // [ns*nk][nproj][nb]  P = Phi Q
  jrb::cplx* d_nl_part;  // [16 chunks][ns*nk][nproj][nb]
  jrb::cplx* d_nl_phit;  // [nk][ng][nproj] conj(Phi) transposed: the tall operand of the DMMA products
  jrb::cplx* d_nl_ps;    // [ns*nk][nproj][nb]  P / vol (the small operand of the apply)
  int nl_p_valid;        // d_nl_p holds Phi Q of the plan's own Q (jrb_eval_begin -> jrb_eval_finish)
  // evaluation work space (Q, R, R^-1, HQ, W-sized temp)
  jrb::cplx *d_q, *d_hq, *d_tmp;
  jrb::cplx *d_r, *d_rinv, *d_small;  // [ns*nk][nb][nb] each (d_small: 5 of them)
  jrb::cplx* d_gpart;                 // [chunks][ns*nk][nb][nb] split-K Gram partials
  double *d_tkb, *d_eps;              // [ns*nk*nb]
  double* d_sphere_part;              // [8 chunks][ns*nk*nb] partials of the sphere reductions
  double* d_scal;                     // small device scalars
  int* d_skip;                        // [ns*nk] second-pass Cholesky already written (k_near_identity)
  double* d_emax;                     // [ns*nk] max|Q1^H Q1 - I| of the second pass (bit pattern)
  cudaStream_t own_stream, h2d_stream, d2h_stream;
  cudaEvent_t ev_in[16], ev_out[16];  // per k-chunk events of jrb_energy_grad_host
  // jrb_eval phase timing (jrb_plan_phase_timing): events at the phase boundaries of the LAST
  // evaluation, recorded on the caller's stream; 0 = off (the default; never during capture)
  int phase_timing;
  cudaEvent_t ev_phase[8];
  // host staging for jrb_energy_grad_host
  double *d_wre, *d_wim, *d_gre, *d_gim, *d_occ, *d_rho, *d_en;
  int64_t ws_bytes;
  // Orbital grid (jrb_plan_set_orbital_grid): a child plan on a smaller, alias-free FFT box that
  // runs every per-orbital transform; this plan keeps the reference's grid for rho, the
  // potentials and the dense API.  Null: orbitals are transformed on the plan's own grid.
  jrb_plan* wf;
  int orbital_only;     // this plan IS such a child (no evaluation work space)
  double* d_rho_w;      // child: [ns][ngrid] density on the orbital grid
  int veff_prepared;    // jrb_hpsi_prepare: d_veff (and the child's) hold the caller's fixed potential
  // host copies for the child's construction and jrb_set_kpoints
  int32_t* h_freq;      // [ng][3] integer frequencies of the kept plane waves, compact order
  double* h_kpts;       // [nk][3]
  int gmax[3];          // largest |frequency| per axis on the sphere
  // communicator over NVLink peer memory (comm.cu: jrb_comm_create / jrb_comm_connect); null = none
  void* comm;
};

namespace jrb {

// fft_passes_*.cu : pencil passes, dispatched on the axis length
int launch_density(jrb_plan* p, const cplx* q, const double* occ, double* rho, cudaStream_t st);
int launch_hpsi(jrb_plan* p, const cplx* q, const double* veff, cplx* hq, cudaStream_t st);
int launch_density_krange(jrb_plan* p, const cplx* q, double* rho, int k0, int k1, cudaStream_t st);
int launch_hpsi_krange(jrb_plan* p, const cplx* q, const double* veff, cplx* hq, int k0, int k1,
                       cudaStream_t st);
int launch_kinetic_range(jrb_plan* p, int sk0, int nsk, const cplx* q, double* t_skb,
                         cudaStream_t st);
int launch_fft3d_dense(jrb_plan* p, const cplx* in, cplx* out, int dir, int64_t batch,
                       double scale, cudaStream_t st);
// density accumulation protocol of the k-chunked host path (and of launch_density itself):
// begin (zero) -> launch_density_krange ... -> end (orbital grid -> the plan's grid)
int launch_density_begin(jrb_plan* p, double* rho, cudaStream_t st);
int launch_density_end(jrb_plan* p, double* rho, cudaStream_t st);
// veff on the plan's grid -> what launch_hpsi_krange needs (resampled onto the orbital grid)
int launch_hpsi_prepare(jrb_plan* p, const double* veff, cudaStream_t st);
// grid_kernels.cu: Fourier resampling between two boxes (dst bins without a partner are zero)
int launch_resample(const cplx* src, int sx, int sy, int sz, cplx* dst, int dx, int dy, int dz,
                    double scale, cudaStream_t st);
int launch_real_to_complex(const double* in, long long n, cplx* out, cudaStream_t st);
int launch_complex_to_real(const cplx* in, long long n, double* out, cudaStream_t st);
bool line_length_supported(int n);
bool fused_available(int nx, int ny, int nxo, int ncol);
bool fused128_available(int nx, int ny, int nxo, int ncol, int band_limited32);
int fused_cta_count(int n, int nxo, int ncol);
int fused_psi_plane_elems(int n, int nxo);

// grid_kernels.cu
// fused grid part of an evaluation (orbital box, LDA, one spin): see launch_grid_potential_orbital
bool grid_fused_ok(const jrb_plan* p, int xc_id);
int launch_density_end_hat(jrb_plan* p, double* rho, cudaStream_t st);
int launch_grid_potential_orbital(jrb_plan* p, const double* rho, bool rhohat_ready, int xc_id,
                                  double* energies, cudaStream_t st);
int launch_set_atoms(jrb_plan* p, const double* pos_h, const double* chg_h, int na,
                     cudaStream_t st);
int launch_grid_potential(jrb_plan* p, const double* rho, int xc_id, int kohn_sham, int parts,
                          double* energies, double* veff, cudaStream_t st);
int launch_density_reciprocal(jrb_plan* p, const double* rho, cplx* rho_hat, cudaStream_t st);
int launch_kinetic(jrb_plan* p, const cplx* q, double* t_skb, cudaStream_t st);
int launch_band_expect(jrb_plan* p, const cplx* q, const cplx* hq, double* eps, cudaStream_t st);
int launch_weighted_sum(jrb_plan* p, const double* a, const double* w, int64_t n, double* out,
                        cudaStream_t st);
int launch_expand(jrb_plan* p, const cplx* q, cplx* dense, cudaStream_t st);
int launch_squeeze(jrb_plan* p, const cplx* dense, cplx* q, cudaStream_t st);
int launch_focc(jrb_plan* p, const double* occ, cudaStream_t st);
int launch_set_kpoints(jrb_plan* p, const double* kpts_h, cudaStream_t st);
int launch_nonlocal_project(jrb_plan* p, int sk0, int nsk, const cplx* q, cudaStream_t st);
int launch_nonlocal_energy(jrb_plan* p, const double* occ, double* e_inout, cudaStream_t st);
int launch_nonlocal_apply(jrb_plan* p, int sk0, int nsk, cplx* hq, cudaStream_t st);
int launch_external_position_gradient(jrb_plan* p, const double* rho, double* grad, cudaStream_t st);

// comm.cu: all-reduce over peer memory (in-place SUM over ranks of buf[0..n) and extra[0..m), m <= 8)
int launch_comm_allreduce(jrb_plan* p, double* buf, long long n, double* extra, int m,
                          cudaStream_t st);
int comm_world(const jrb_plan* p);   // 1 without a connected communicator
int comm_error(jrb_plan* p, cudaStream_t st);
void comm_destroy(jrb_plan* p);
// fft_dispatch.cu: occupation table + zero + every (spin, k, band group) swept into the density of
// the grid the passes run on (the orbital box when there is one): launch_density without the
// final resampling, so that a multi-GPU caller can all-reduce the smaller box in between
int launch_density_partial(jrb_plan* p, const cplx* q, const double* occ, double* rho, cudaStream_t st);

// qr.cu
int qr_gram_partial_mats(const jrb_plan* p);
// rectangular products on the DMMA Gram / apply kernels (projector products, nonlocal.cu)
int launch_gram_rect(jrb_plan* p, int nsk, const cplx* At, int nbA, int a_mod, const cplx* B,
                     int nchunks, cplx* partial, cudaStream_t st);
int launch_apply_rect(jrb_plan* p, int nsk, const cplx* In, int kdim, int in_mod, const cplx* T,
                      cplx* out, cudaStream_t st);
int launch_nonlocal_transpose(jrb_plan* p, cudaStream_t st);
// split-phase QR (row-sharded callers all-reduce S / M between the phases)
int launch_qr_gram_phase(jrb_plan* p, int sk0, int nsk, const double* w_re, const double* w_im,
                         int pass, cplx* S, cudaStream_t st);
int launch_qr_apply_phase(jrb_plan* p, int sk0, int nsk, const double* w_re, const double* w_im,
                          int pass, cplx* S, cplx* qout, cplx* r, cudaStream_t st);
int launch_qr_bwd_gram_phase(jrb_plan* p, int sk0, int nsk, const cplx* q, const cplx* gq, cplx* M,
                             cudaStream_t st);
int launch_qr_bwd_apply_phase(jrb_plan* p, int sk0, int nsk, const cplx* q, const cplx* r,
                              const cplx* gq, const double* occ, const cplx* M, double* g_re,
                              double* g_im, cudaStream_t st);
int launch_qr_fwd_range(jrb_plan* p, int sk0, int nsk, const double* w_re, const double* w_im,
                        cplx* q, cplx* r, cudaStream_t st);
int launch_qr_bwd_range(jrb_plan* p, int sk0, int nsk, const cplx* q, const cplx* r,
                        const cplx* gq, const double* occ, double* g_re, double* g_im,
                        cudaStream_t st);
int launch_qr_fwd(jrb_plan* p, const double* w_re, const double* w_im, cplx* q, cplx* r,
                  cudaStream_t st);
int launch_hamiltonian_matrix(jrb_plan* p, const cplx* q, const cplx* hq, cplx* h, cudaStream_t st);
int launch_qr_bwd(jrb_plan* p, const cplx* q, const cplx* r, const cplx* gq, const double* occ,
                  double* g_re, double* g_im, cudaStream_t st);

}  // namespace jrb
