/* jrystal_b200 -- C ABI of the B200-native energy+gradient path of sail-sg/jrystal.
 *
 * This header is the drop-in boundary (DESIGN.md "Boundary").  The reference
 * has no native interface: its boundary is the Python function API of jrystal/{pw,energy,
 * potential,hamiltonian}.py and the JAX primitive pair `ifftn_sharding` / `fftn_sharding`
 * (jrystal/_src/spmd/fft.py:79-134).  Each entry point below names the reference function(s)
 * it replaces; INTEGRATION.md shows the XLA-FFI / custom_vjp and ctypes bindings.
 *
 * Conventions
 *   - plain `extern "C"`, no C++ types, no exceptions across the boundary;
 *   - every call returns 0 on success, a negative JRB_E* code on failure; the message of the
 *     last failure on the calling thread is available from jrb_last_error();
 *   - all array arguments are DEVICE pointers owned by the caller unless the name ends in
 *     `_host`; the library never frees or retains them past the call;
 *   - calls are asynchronous on the passed CUDA stream (a `cudaStream_t` cast to void*),
 *     never synchronise and never allocate (work space belongs to the plan), except the
 *     `*_host` entry point which copies, runs and synchronises;
 *   - complex arrays are interleaved (re, im) FP64 = numpy complex128 / jnp.complex128;
 *   - array layouts are the reference's: parameters and sphere coefficients are
 *     (ns, nk, ng, nb) with the band axis contiguous (jrystal/_src/pw.py:88-91), compact
 *     index g enumerates mask==True in C order (jrystal/_src/utils.py:279-281), grids are
 *     (..., nx, ny, nz) C order.
 *   - there is NO CPU fallback: every entry point needs a CUDA device.
 */
#ifndef JRYSTAL_B200_H_
#define JRYSTAL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JRB_OK 0
#define JRB_EINVAL (-1)      /* bad argument (shape, null pointer, unsupported value) */
#define JRB_EUNSUPPORTED (-2) /* FFT length without a compiled plan, xc functional ...  */
#define JRB_ECUDA (-3)       /* CUDA runtime error                                      */
#define JRB_ENOMEM (-4)      /* device allocation failed                                */

/* xc functionals (jrystal/_src/xc.py:242-253, LDA branch).  Names joined by '+' in the
 * reference are summed; the ids below are the combinations the kernels implement. */
#define JRB_XC_LDA_X 1
#define JRB_XC_LDA_X_C_PW 2 /* "lda_x+lda_c_pw"; two spin channels: polarised PW92 (xc.py:60-61) */
/* GGA branch (xc.py:67-112: sigma = |ifftn(i G fftn(rho))|^2, eps(rho, sigma)); unpolarised.
 * The potential is the exact discrete derivative of E_xc = (Omega/N) sum rho eps (what jax.grad
 * gives the reference in energy mode); band mode (kohn_sham) uses the same potential instead of
 * xc.py:127-219, whose gradient term calls a 1-D fft on a 3-D field. */
#define JRB_XC_GGA_X_PBE 3 /* "gga_x_pbe" */
#define JRB_XC_GGA_PBE 4   /* "gga_x_pbe+gga_c_pbe" (the reference's config.yaml default) */

#define JRB_FFT_FORWARD (-1) /* jnp.fft.fftn  : exp(-i...)            (fftn_sharding)  */
#define JRB_FFT_INVERSE (+1) /* jnp.fft.ifftn : exp(+i...) and 1/N    (ifftn_sharding) */

typedef struct jrb_plan jrb_plan;
typedef void* jrb_stream; /* cudaStream_t */

typedef struct {
  int32_t nx, ny, nz;    /* FFT grid; each a 7-smooth length with a compiled line plan   */
  int32_t ns, nk, nb;    /* spins (1|2), k-points and bands held by THIS rank            */
  const uint8_t* mask;   /* host, nx*ny*nz bytes, C order, non-zero = kept plane wave    */
  const double* kpts;    /* host, [nk][3] Cartesian k-vectors (1/Bohr)                   */
  const double* cell;    /* host, [3][3] rows = lattice vectors (Bohr)                   */
  int32_t device;        /* CUDA device ordinal                                          */
  int32_t batch_groups;  /* band groups (8 bands) per pencil-pass batch; 0 = automatic   */
} jrb_plan_desc;

/* Builds everything the reference's drivers build before their loop
 * (calc/calc_ground_state_energy_all_electrons.py:93-106: grids, mask-derived index maps,
 * |G+k|^2 tables) plus twiddles and work space.  Replaces grid.g_vectors / spherical_mask
 * consumers inside the hot loop. */
int jrb_plan_create(const jrb_plan_desc* desc, jrb_plan** out);
int jrb_plan_destroy(jrb_plan* plan);
/* number of kept plane waves ng = mask.sum() (jrystal/_src/pw.py:89) */
int64_t jrb_plan_num_g(const jrb_plan* plan);
/* bytes of device work space owned by the plan */
int64_t jrb_plan_workspace_bytes(const jrb_plan* plan);

/* Orbital grid.  The reference transforms every orbital on the same (nx, ny, nz) box it uses for
 * the density and the potentials (pw.wave_grid, jrystal/_src/pw.py:208-211; utils.expand_coefficient,
 * utils.py:277-281).  psi only carries the frequencies of the cut-off sphere, |f_c| <= gmax_c, so
 * |psi|^2 and the sphere part of v_eff*psi involve |f_c| <= 2 gmax_c and any box with
 * n_c >= 4 gmax_c + 1 transforms them without aliasing: after this call the per-orbital passes
 * (jrb_density, jrb_hpsi, jrb_eval_begin/finish, jrb_energy_grad_host) run on (nxw, nyw, nzw);
 * rho is brought to the plan's grid by Fourier interpolation and v_eff to the orbital grid by
 * Fourier truncation, so every result is the reference's to rounding (tests/test_orbital_grid_gpu.py,
 * tests/test_exactness_math.py).  Everything else (grids in the arguments, XC, Hartree, jrb_fft3d,
 * jrb_wave_grid) keeps the plan's own grid.  Fails with JRB_EINVAL if a RESIZED axis violates
 * 4 gmax + 1 <= n_w < n (an axis left at n is always accepted: where the caller's own grid is too
 * coarse it aliases exactly as the reference does), JRB_EUNSUPPORTED without a compiled line
 * length.  Allocates (set-up). */
int jrb_plan_set_orbital_grid(jrb_plan* plan, int32_t nxw, int32_t nyw, int32_t nzw);
/* dims[3] = the box the per-orbital passes currently run on / the smallest alias-free box */
int jrb_plan_orbital_grid(const jrb_plan* plan, int32_t* dims);
int jrb_plan_min_orbital_grid(const jrb_plan* plan, int32_t* dims);
/* which kernels run the y and x passes on the current orbital box: 0 single pencil passes,
 * 1 fused y+x plane kernels (fft_fused.cuh), 2 the 128 x 128 fused family (fft_fused128.cuh) */
int jrb_plan_orbital_fused(const jrb_plan* plan);
/* bytes of the psi(r) cache of the current orbital box (0: none).  With it jrb_eval_begin stores
 * every orbital's psi(r) while it accumulates the density and jrb_eval_finish applies v_eff to the
 * stored values instead of repeating the inverse transforms (the passes are bound by the FP64 and
 * shared-memory pipes while HBM idles).  Allocated at plan creation when the fused plane kernels
 * run and it fits JRB_PSI_CACHE_MB (default 65536; 0 switches it off) and 40 % of the free memory */
int64_t jrb_plan_psi_cache_bytes(const jrb_plan* plan);

/* Pre-computes V_ext(G) once: potential.external_reciprocal (jrystal/_src/potential.py:
 * 153-166), which the reference re-evaluates every step although it is parameter free. */
int jrb_set_atoms(jrb_plan* plan, const double* positions_host, const double* charges_host,
                  int32_t natoms, jrb_stream stream);

/* Non-local part of a norm-conserving pseudopotential (pseudopotential/nloc.py:143-158, 217-236):
 * phi = potential_nl_psi_reciprocal restricted to the cut-off sphere, device complex
 * (nk, nproj, ng), p = (beta, m) flattened, sqrt(D) already folded in as the reference does
 * (nloc.py:60-141).  Once set (nproj > 0; nproj == 0 removes it; allocates, set-up time):
 *   - jrb_hpsi adds  Phi^H (Phi q) / vol  (hamiltonian_nonlocal applied to q), so
 *     jrb_band_expect / jrb_hamiltonian_matrix and the gradients of jrb_eval_finish include it;
 *   - jrb_eval_begin adds E_nl = sum f |Phi q|^2 / vol (energy_nonlocal) to its e_kin output,
 *     which then is "kinetic + non-local" in energies[0];
 *   - jrb_nonlocal_energy returns E_nl alone (device double).
 * With projectors set jrb_energy_grad_host runs its unchunked (non-pipelined) variant. */
int jrb_set_nonlocal(jrb_plan* plan, const double* phi, int32_t nproj, jrb_stream stream);
int jrb_nonlocal_energy(jrb_plan* plan, const double* q, const double* occ, double* e_nl,
                        jrb_stream stream);

/* Position cotangent of energy.external (energy.py:121-135 through potential.external_reciprocal,
 * potential.py:153-166) for the atoms of the last jrb_set_atoms: grad[a][c] = dE_ext / dR_a,c, what
 * jax.grad w.r.t. `position` gives the reference (docs/tutorial/differentiation.rst:128-147);
 * forces are minus this minus the Ewald part.  rho: (ns, nx, ny, nz); grad: device, [natoms][3]. */
int jrb_external_position_gradient(jrb_plan* plan, const double* rho, double* grad,
                                   jrb_stream stream);

/* Sets the reciprocal-space external potential V(G) directly (device, complex (nx, ny, nz), the
 * convention of potential.external_reciprocal: E = Re sum_G conj(V) rho_hat * Omega / N^2,
 * v(r) = ifftn(V)).  Besides replacing jrb_set_atoms this is the LOCAL part of a norm-conserving
 * pseudopotential: pseudopotential/local.py:164-187 (energy_local = reciprocal_braket(V_loc, rho_hat),
 * hamiltonian_local = <psi| ifftn(V_loc) |psi>) contracts exactly like the all-electron external
 * term, with V_loc(G) precomputed on the host from the UPF data. */
int jrb_set_external_potential(jrb_plan* plan, const double* vhat, jrb_stream stream);

/* Replaces the plan's k-points (same count nk) by rebuilding the |G+k|^2 table on the device: the
 * band-structure driver walks a k-path with one plan (calc/calc_band_structure_all_electrons.py:
 * 115-182 re-traces `update` per k-point instead).  kpts_host: [nk][3] Cartesian, 1/Bohr. */
int jrb_set_kpoints(jrb_plan* plan, const double* kpts_host, jrb_stream stream);

/* unitary_module.unitary_matrix (jrystal/_src/unitary_module.py:66-81): Q R = w_re + i w_im per
 * (spin, k).  Cholesky-QR2 on FP64 tensor cores; gauge: diag(R) real POSITIVE (LAPACK's
 * Householder Q differs by a sign per column; energies, density and gradients are invariant).
 * q: (ns,nk,ng,nb) complex, r: (ns,nk,nb,nb) complex upper triangular. */
int jrb_qr_fwd(jrb_plan* plan, const double* w_re, const double* w_im, double* q, double* r,
               jrb_stream stream);
/* reverse mode of the above (what JAX's QR AD rule does for the reference):
 * gq = dE/dQ* (complex, same layout as q); outputs dE/dw_re, dE/dw_im. */
int jrb_qr_bwd(jrb_plan* plan, const double* q, const double* r, const double* gq, double* g_re,
               double* g_im, jrb_stream stream);

/* Row-sharded orthonormalisation (Gamma-only supercells have fewer k-points than
 * GPUs, so the rows g of W are split over ranks for the QR while the bands are split for the
 * FFTs).  Same mathematics as jrb_qr_fwd / jrb_qr_bwd (unitary_module.py:66-81 and its AD rule),
 * cut where a sum over the row axis crosses ranks; the caller all-reduces (SUM) the small
 * (ns, nk, nb, nb) complex matrices between the phases:
 *     for pass in 0, 1:  jrb_qr_rows_gram(pass) -> all-reduce S -> jrb_qr_rows_apply(pass)
 *     jrb_qr_rows_bwd_gram -> all-reduce M -> jrb_qr_rows_bwd_apply
 * They work on the plan's ng rows: a full plan (then they equal jrb_qr_fwd/bwd on one rank) or a
 * rows-only plan from jrb_plan_create_rows (this rank's row block; no grid, QR entry points only).
 * w_re/w_im/q/gq/g_re/g_im: (ns, nk, nrows, nb).  pass 1 ignores w_re/w_im (it reads the Q1 of
 * pass 0 from plan work space) and writes q and r.  S is overwritten by the factorisation.
 * jrb_qr_rows_bwd_apply uses the R^-1 of the plan's own last forward call. */
int jrb_plan_create_rows(int64_t nrows, int32_t ns, int32_t nk, int32_t nb, int32_t device,
                         jrb_plan** out);
int jrb_qr_rows_gram(jrb_plan* plan, const double* w_re, const double* w_im, int32_t pass,
                     double* s_out, jrb_stream stream);
int jrb_qr_rows_apply(jrb_plan* plan, const double* w_re, const double* w_im, int32_t pass,
                      double* s_inout, double* q, double* r, jrb_stream stream);
int jrb_qr_rows_bwd_gram(jrb_plan* plan, const double* q, const double* gq, double* m_out,
                         jrb_stream stream);
int jrb_qr_rows_bwd_apply(jrb_plan* plan, const double* q, const double* gq, const double* occ,
                          const double* m, double* g_re, double* g_im, jrb_stream stream);

/* Synchronises `stream` and reports a deferred numerical failure of the asynchronous calls
 * (Cholesky breakdown on rank-deficient parameters): 0 or JRB_EINVAL. */
int jrb_check_status(jrb_plan* plan, jrb_stream stream);

/* utils.expand_coefficient (jrystal/_src/utils.py:277-281): (ns,nk,ng,nb) -> dense
 * (ns,nk,nb,nx,ny,nz).  Only for API parity; the fused paths never build the dense box. */
int jrb_expand(jrb_plan* plan, const double* q, double* coeff_dense, jrb_stream stream);
/* inverse of the above (utils.squeeze_coefficient, 284-308). */
int jrb_squeeze(jrb_plan* plan, const double* coeff_dense, double* q, jrb_stream stream);

/* pw.density_grid with occupation (jrystal/_src/pw.py:273-284) fused with pw.coeff's scatter
 * and pw.wave_grid: rho[s,x,y,z] = sum_kb occ[s,k,b] |psi_skb(r)|^2; psi never reaches HBM.
 * rho is OVERWRITTEN. */
int jrb_density(jrb_plan* plan, const double* q, const double* occ, double* rho,
                jrb_stream stream);

/* per-orbital kinetic expectation t[s,k,b] = 1/2 sum_G |G+k|^2 |c_G|^2 on the sphere
 * (energy.kinetic without occupation, jrystal/_src/energy.py:172-180; braket.expectation
 * mode='kinetic', diagonal). */
int jrb_kinetic(jrb_plan* plan, const double* q, double* t_skb, jrb_stream stream);

/* energy.hartree + energy.external + energy.xc_energy (jrystal/_src/energy.py:68-82, 121-135,
 * 204-211) and potential.effective (potential.py:255-279) in one sweep over the grid:
 *   energies[0..2] = E_hartree, E_external, E_xc   (device doubles)
 *   veff[s,x,y,z]  = dE/drho_s = Re ifftn(4 pi n(G)/G^2 + V_ext(G)) + v_xc,s(r)
 * kohn_sham != 0 evaluates the terms the way hamiltonian_matrix_trace does (un-halved Hartree
 * energy, v_xc as the xc "energy density"; potential.py:67-75, xc.py:247-250). */
int jrb_grid_potential(jrb_plan* plan, const double* rho, int32_t xc_id, int32_t kohn_sham,
                       double* energies, double* veff, jrb_stream stream);

/* potential.effective (jrystal/_src/potential.py:203-279) with the reference's own semantics,
 * real part: `parts` selects the terms (split=True returns them one by one).  Unlike
 * jrb_grid_potential's veff (= dE/drho), the Hartree term is HALVED and the xc term is eps_xc
 * unless kohn_sham != 0 (potential.py:67-75, xc.py:247-250) -- the form the reference's
 * energy_test.py:61-112 identity uses. */
#define JRB_V_HARTREE 1
#define JRB_V_EXTERNAL 2
#define JRB_V_XC 4
int jrb_potential(jrb_plan* plan, const double* rho, int32_t xc_id, int32_t kohn_sham,
                  int32_t parts, double* v_out, jrb_stream stream);

/* pw.density_grid_reciprocal (jrystal/_src/pw.py:333-334): rho_hat[s] = fftn(rho[s]), complex. */
int jrb_density_reciprocal(jrb_plan* plan, const double* rho, double* rho_hat, jrb_stream stream);

/* pw.wave_grid o pw.coeff (jrystal/_src/pw.py:208-211): psi[s,k,b,x,y,z] = ifftn(expand(q)) *
 * N / sqrt(Omega), dense (API parity / diagnostics; the hot path never materialises it). */
int jrb_wave_grid(jrb_plan* plan, const double* q, double* psi, jrb_stream stream);

/* Hamiltonian apply on the sphere: hq = 1/2|G+k|^2 q + (sqrt(Omega)/N) fftn(veff * psi)|mask.
 * This is the reverse pass of the energy w.r.t. Q (dE/dQ* = occ * hq) and the forward+reverse
 * pass of hamiltonian.hamiltonian_matrix_trace (jrystal/_src/hamiltonian.py:147-168). */
int jrb_hpsi(jrb_plan* plan, const double* q, const double* veff, double* hq, jrb_stream stream);
/* Band mode keeps v_eff[rho_gs] fixed over thousands of steps (the reference recomputes it inside
 * every step, hamiltonian.py:147-156): jrb_hpsi_prepare copies it into plan work space (and
 * resamples it onto the orbital grid once); jrb_hpsi with veff == NULL then applies the prepared
 * potential.  Any call that passes its own veff, and jrb_eval_finish / jrb_energy_grad_host,
 * replace the prepared potential. */
int jrb_hpsi_prepare(jrb_plan* plan, const double* veff, jrb_stream stream);

/* eps[s,k,b] = Re sum_G conj(q) hq: diagonal of braket.expectation(real)+(kinetic)
 * (jrystal/_src/braket.py:189-206) = dE/d occ[s,k,b]. */
int jrb_band_expect(jrb_plan* plan, const double* q, const double* hq, double* eps_skb,
                    jrb_stream stream);

/* hamiltonian.hamiltonian_matrix (jrystal/_src/hamiltonian.py:171-240; the reference takes nb
 * Hessian-vector products through AD, hessian.py:21-55): H[s,k,i,j] = <q_i| hq_j> with
 * hq = jrb_hpsi(q), one Gram on the FP64 tensor cores; H is Hermitian (upper blocks computed,
 * lower mirrored).  h: (ns, nk, nb, nb) complex. */
int jrb_hamiltonian_matrix(jrb_plan* plan, const double* q, const double* hq, double* h,
                           jrb_stream stream);

/* Dense batched 3-D C2C transform over the last three axes: exact drop-in for the
 * primitives ifftn_sharding / fftn_sharding (jrystal/_src/spmd/fft.py:68-75), numpy
 * normalisation (inverse divides by N).  in == out is allowed. */
int jrb_fft3d(jrb_plan* plan, const double* in, double* out, int32_t direction, int64_t batch,
              jrb_stream stream);

/* One energy+gradient evaluation = value_and_grad(total_energy) of the energy-mode driver
 * (calc/calc_ground_state_energy_all_electrons.py:119-137,175-181), optimiser excluded.
 * Split in two so a multi-GPU host can all-reduce rho/E_kin between the halves:
 *   begin : QR, fused scatter+IFFT+density, kinetic   -> rho (partial over this rank's k/bands),
 *           e_kin (device double, partial)
 *   finish: grid potential from the (all-reduced) rho, H-apply, QR adjoint
 *           -> energies[4] = E_kin(as passed in), E_ext, E_har, E_xc; g_re, g_im; g_occ
 *              (optional, may be NULL) = dE/d occ.
 * The Q/R/HQ intermediates live in plan-owned work space. */
int jrb_eval_begin(jrb_plan* plan, const double* w_re, const double* w_im, const double* occ,
                   double* rho, double* e_kin, jrb_stream stream);
int jrb_eval_finish(jrb_plan* plan, const double* occ, const double* rho, const double* e_kin,
                    int32_t xc_id, double* energies, double* g_re, double* g_im, double* g_occ,
                    jrb_stream stream);

/* Multi-GPU (one process per GPU): a plan-owned communicator over NVLink / NVSwitch peer memory
 * and the ONE collective of the path, the all-reduce of the partial densities.  The reference
 * shards the k-mesh over devices (calc/calc_ground_state_energy_all_electrons.py:83-91,151-158,
 * `nk % ndev == 0`, _src/spmd/uniform.py:22-24) and XLA inserts this all-reduce where
 * einsum('skb...,skb->s...') contracts the sharded k axis (_src/pw.py:278).
 *   set-up (allocates, synchronises; not on the hot path):
 *     jrb_comm_create  : allocates this rank's symmetric region (capacity doubles per buffer;
 *                        0 = enough for the plan's density + E_kin, or for the small matrices of a
 *                        rows-only plan) and writes its CUDA IPC handle (jrb_comm_handle_bytes()
 *                        bytes) to handle_out; the caller exchanges the handles of all ranks by
 *                        any means (torch.distributed, MPI, a file);
 *     jrb_comm_connect : handles = world x jrb_comm_handle_bytes() bytes in rank order; maps the
 *                        peers' regions.  world == 1 needs no connect.
 *   data path (asynchronous on the stream, no host round trip, CUDA-graph capturable; the same
 *   call order on every rank; every rank ends with BIT-IDENTICAL sums, reduced in rank order):
 *     jrb_allreduce     : in-place SUM over the ranks of buf[0..n) (device doubles);
 *     jrb_allreduce_rho : the same for rho (ns, nx, ny, nz) and e_kin[1] (may be NULL) between
 *                         jrb_eval_begin and jrb_eval_finish.
 * A peer that never arrives raises an error that jrb_check_status reports (bounded spin). */
int jrb_comm_handle_bytes(void);
int jrb_comm_create(jrb_plan* plan, int32_t rank, int32_t world, int64_t capacity, void* handle_out);
int jrb_comm_connect(jrb_plan* plan, const void* handles);
int jrb_comm_world(const jrb_plan* plan); /* 1 without a connected communicator */
int jrb_allreduce(jrb_plan* plan, double* buf, int64_t n, jrb_stream stream);
int jrb_allreduce_rho(jrb_plan* plan, double* rho, double* e_kin, jrb_stream stream);

/* jrb_eval_begin + the all-reduce + jrb_eval_finish in one call: with a connected communicator the
 * partial density is reduced INSIDE, on the box the orbitals were transformed on (the orbital
 * grid: 4x fewer bytes than the 128^3 grid of the diamond-64 configuration), before its Fourier
 * interpolation onto the plan's grid; without one it is the single-GPU evaluation.  w_re, w_im,
 * occ, g_re, g_im, g_occ (may be NULL): this rank's k-points; rho (ns, nx, ny, nz) and
 * energies[4] = E_kin, E_ext, E_har, E_xc: the whole system's, identical on every rank. */
int jrb_eval(jrb_plan* plan, const double* w_re, const double* w_im, const double* occ,
             int32_t xc_id, double* energies, double* g_re, double* g_im, double* g_occ,
             double* rho, jrb_stream stream);

/* Phase timing of jrb_eval / jrb_eval_begin + jrb_eval_finish (a measurement aid for bench.py: the
 * phases are timed where they run, inside the evaluation, by CUDA events on the caller's stream).
 * enable != 0: every following evaluation records events at its phase boundaries (do not enable
 * while the stream is being captured into a CUDA graph).  jrb_plan_phase_times waits for the last
 * evaluation and returns ms[6] = QR forward, density sweep, density reduction + interpolation +
 * kinetic / non-local energy, grid potential, H-apply sweep, QR adjoint. */
int jrb_plan_phase_timing(jrb_plan* plan, int32_t enable);
int jrb_plan_phase_times(jrb_plan* plan, double* ms);

/* Same evaluation through HOST buffers (the reference-facing call a non-CUDA host makes):
 * copies w_re/w_im/occ in (k-point chunks, overlapped with the kernels), runs the evaluation,
 * copies energies[4], g_re, g_im (and rho_host if non-NULL) back and synchronises.  With a
 * connected communicator (jrb_comm_connect) every rank passes its own k-points and the partial
 * densities are all-reduced inside, as in jrb_eval. */
int jrb_energy_grad_host(jrb_plan* plan, const double* w_re_host, const double* w_im_host,
                         const double* occ_host, int32_t xc_id, double* energies_host,
                         double* g_re_host, double* g_im_host, double* rho_host);

/* optax.adam update of one parameter array on the device: the optimiser of the energy-mode driver
 * (calc/calc_ground_state_energy_all_electrons.py:175-181, opt_utils.py:153-168; default
 * lr 0.01, b1 0.9, b2 0.99, eps 1e-8, config.py).  `state` is 4 device doubles, zero-initialised by
 * the caller: step count and the two bias corrections, advanced ONCE per optimisation step by
 * jrb_adam_tick and read by every jrb_adam_apply of that step -- no host value changes between
 * steps, so begin + finish + tick + apply can be captured in one CUDA graph and replayed.
 * No plan needed; arrays 16-byte aligned; m, v zero-initialised by the caller. */
int jrb_adam_tick(double* state, double b1, double b2, jrb_stream stream);
int jrb_adam_apply(int64_t n, double* param, const double* grad, double* m, double* v, double lr,
                   double b1, double b2, double eps, const double* state, jrb_stream stream);

const char* jrb_last_error(void);
int jrb_version(void);
/* number of CUDA kernels this library has launched in the calling process (bench evidence) */
int64_t jrb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* JRYSTAL_B200_H_ */
