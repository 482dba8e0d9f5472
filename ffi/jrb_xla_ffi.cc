// XLA-FFI handlers over the C ABI of include/jrystal_b200.h: the custom calls that
// ffi/jrystal_b200_jax.py registers with jax.ffi and wraps in jax.custom_vjp, so that jax.grad of
// jrystal's own loss (calc/calc_ground_state_energy_all_electrons.py:119-137,175-181) and its
// primitive pair ifftn_sharding / fftn_sharding (jrystal/_src/spmd/fft.py:79-134) run on
// libjrystal_b200.so.
//
// Build (needs jax >= 0.5.3 for the header; JAX is NOT in this repo's build image, where
// tests/test_ffi_shim.py compiles this file against a stand-in for the XLA header instead):
//   ffi/build_ffi.sh      ->  ffi/libjrb_xla_ffi.so
//
// Conventions: one handler per C-ABI call; the plan pointer travels as an int64 attribute (plans
// are created from Python through ctypes at set-up time, they are not traced ops); buffers in
// header order; every call is asynchronous on XLA's stream, never allocates and never
// synchronises, hence command-buffer (CUDA graph) safe.  A non-zero return code becomes an
// XLA_FFI_Error carrying jrb_last_error().
#include <cstdint>

#include "xla/ffi/api/ffi.h"

#include "jrystal_b200.h"

namespace ffi = xla::ffi;

namespace {

using F64 = ffi::Buffer<ffi::F64>;
using C128 = ffi::Buffer<ffi::C128>;
using RF64 = ffi::ResultBuffer<ffi::F64>;
using RC128 = ffi::ResultBuffer<ffi::C128>;
using Stream = ffi::PlatformStream<cudaStream_t>;

inline jrb_plan* P(int64_t handle) { return reinterpret_cast<jrb_plan*>(handle); }
inline const double* D(const F64& b) { return b.typed_data(); }
inline double* D(RF64& b) { return b->typed_data(); }
// complex128 buffers are interleaved (re, im) doubles on both sides
inline const double* D(const C128& b) { return reinterpret_cast<const double*>(b.typed_data()); }
inline double* D(RC128& b) { return reinterpret_cast<double*>(b->typed_data()); }

inline ffi::Error Check(int rc) {
  if (rc == JRB_OK) return ffi::Error::Success();
  const char* msg = jrb_last_error();
  if (rc == JRB_EINVAL || rc == JRB_EUNSUPPORTED)
    return ffi::Error::InvalidArgument(msg ? msg : "jrystal_b200: invalid argument");
  return ffi::Error::Internal(msg ? msg : "jrystal_b200: internal error");
}

// ---- unitary_module.unitary_matrix (unitary_module.py:66-81) and its AD rule -------------------
ffi::Error QrFwd(cudaStream_t st, int64_t plan, F64 w_re, F64 w_im, RC128 q, RC128 r) {
  return Check(jrb_qr_fwd(P(plan), D(w_re), D(w_im), D(q), D(r), st));
}
ffi::Error QrBwd(cudaStream_t st, int64_t plan, C128 q, C128 r, C128 gq, RF64 g_re, RF64 g_im) {
  return Check(jrb_qr_bwd(P(plan), D(q), D(r), D(gq), D(g_re), D(g_im), st));
}

// ---- pw.density_grid with occupation (pw.py:273-284), fused with the scatter and the IFFT ------
ffi::Error Density(cudaStream_t st, int64_t plan, C128 q, F64 occ, RF64 rho) {
  return Check(jrb_density(P(plan), D(q), D(occ), D(rho), st));
}

// ---- pw.wave_grid o pw.coeff (pw.py:208-211), dense -------------------------------------------
ffi::Error WaveGrid(cudaStream_t st, int64_t plan, C128 q, RC128 psi) {
  return Check(jrb_wave_grid(P(plan), D(q), D(psi), st));
}

// ---- utils.expand_coefficient / squeeze_coefficient (utils.py:277-308) ------------------------
ffi::Error Expand(cudaStream_t st, int64_t plan, C128 q, RC128 dense) {
  return Check(jrb_expand(P(plan), D(q), D(dense), st));
}
ffi::Error Squeeze(cudaStream_t st, int64_t plan, C128 dense, RC128 q) {
  return Check(jrb_squeeze(P(plan), D(dense), D(q), st));
}

// ---- energy.kinetic per orbital (energy.py:172-180) -------------------------------------------
ffi::Error Kinetic(cudaStream_t st, int64_t plan, C128 q, RF64 t_skb) {
  return Check(jrb_kinetic(P(plan), D(q), D(t_skb), st));
}

// ---- energy.hartree / external / xc_energy + potential.effective in one sweep -----------------
ffi::Error GridPotential(cudaStream_t st, int64_t plan, int32_t xc_id, int32_t kohn_sham, F64 rho,
                         RF64 energies, RF64 veff) {
  return Check(jrb_grid_potential(P(plan), D(rho), xc_id, kohn_sham, D(energies), D(veff), st));
}

// ---- potential.effective with the reference's own semantics (potential.py:203-279) ------------
ffi::Error Potential(cudaStream_t st, int64_t plan, int32_t xc_id, int32_t kohn_sham,
                     int32_t parts, F64 rho, RF64 v) {
  return Check(jrb_potential(P(plan), D(rho), xc_id, kohn_sham, parts, D(v), st));
}

// ---- pw.density_grid_reciprocal (pw.py:333-334) ------------------------------------------------
ffi::Error DensityReciprocal(cudaStream_t st, int64_t plan, F64 rho, RC128 rho_hat) {
  return Check(jrb_density_reciprocal(P(plan), D(rho), D(rho_hat), st));
}

// ---- Hamiltonian apply: reverse pass of the energy, forward + reverse of the band-mode loss ---
ffi::Error Hpsi(cudaStream_t st, int64_t plan, C128 q, F64 veff, RC128 hq) {
  return Check(jrb_hpsi(P(plan), D(q), D(veff), D(hq), st));
}
// band mode: v_eff[rho_gs] fixed over thousands of steps (hamiltonian.py:147-156 recomputes it)
ffi::Error HpsiPrepare(cudaStream_t st, int64_t plan, F64 veff, RF64 token) {
  (void)token;  // a 1-element output only orders this call before the HpsiPrepared calls
  return Check(jrb_hpsi_prepare(P(plan), D(veff), st));
}
ffi::Error HpsiPrepared(cudaStream_t st, int64_t plan, C128 q, F64 token, RC128 hq) {
  (void)token;
  return Check(jrb_hpsi(P(plan), D(q), nullptr, D(hq), st));
}

// ---- braket.expectation diagonal (braket.py:189-206) = dE / d occupation ----------------------
ffi::Error BandExpect(cudaStream_t st, int64_t plan, C128 q, C128 hq, RF64 eps) {
  return Check(jrb_band_expect(P(plan), D(q), D(hq), D(eps), st));
}

// ---- hamiltonian.hamiltonian_matrix (hamiltonian.py:171-240) ----------------------------------
ffi::Error HamiltonianMatrix(cudaStream_t st, int64_t plan, C128 q, C128 hq, RC128 h) {
  return Check(jrb_hamiltonian_matrix(P(plan), D(q), D(hq), D(h), st));
}

// ---- dense drop-in for the primitive pair ifftn_sharding / fftn_sharding (spmd/fft.py:68-75) --
ffi::Error Fft3d(cudaStream_t st, int64_t plan, int32_t direction, C128 x, RC128 y) {
  const auto dims = x.dimensions();
  if (dims.size() < 3)
    return ffi::Error::InvalidArgument("Input must have at least 3 dimensions");  // fft.py:47-48
  int64_t batch = 1;
  for (size_t i = 0; i + 3 < dims.size(); ++i) batch *= dims[i];
  return Check(jrb_fft3d(P(plan), D(x), D(y), direction, batch, st));
}

// ---- the fused evaluation: value_and_grad(total_energy), optimiser excluded -------------------
ffi::Error EvalBegin(cudaStream_t st, int64_t plan, F64 w_re, F64 w_im, F64 occ, RF64 rho,
                     RF64 e_kin) {
  return Check(jrb_eval_begin(P(plan), D(w_re), D(w_im), D(occ), D(rho), D(e_kin), st));
}
ffi::Error EvalFinish(cudaStream_t st, int64_t plan, int32_t xc_id, F64 occ, F64 rho, F64 e_kin,
                      RF64 energies, RF64 g_re, RF64 g_im, RF64 g_occ) {
  return Check(jrb_eval_finish(P(plan), D(occ), D(rho), D(e_kin), xc_id, D(energies), D(g_re),
                               D(g_im), D(g_occ), st));
}
// one call; with a connected communicator (jrb_comm_connect) the all-reduce of the partial
// densities of a k-sharded mesh runs inside, over NVLink peer memory
ffi::Error Eval(cudaStream_t st, int64_t plan, int32_t xc_id, F64 w_re, F64 w_im, F64 occ,
                RF64 energies, RF64 g_re, RF64 g_im, RF64 g_occ, RF64 rho) {
  return Check(jrb_eval(P(plan), D(w_re), D(w_im), D(occ), xc_id, D(energies), D(g_re), D(g_im),
                        D(g_occ), D(rho), st));
}
// rho = psum over the k mesh as a custom call (for callers that keep begin / finish separate)
ffi::Error AllreduceRho(cudaStream_t st, int64_t plan, F64 rho_in, F64 e_kin_in, RF64 rho,
                        RF64 e_kin) {
  // XLA gives distinct output buffers unless aliased: copy, then reduce in place
  const size_t nrho = rho_in.element_count() * sizeof(double);
  if (rho->typed_data() != rho_in.typed_data() &&
      cudaMemcpyAsync(rho->typed_data(), rho_in.typed_data(), nrho, cudaMemcpyDeviceToDevice, st) !=
        cudaSuccess)
    return ffi::Error::Internal("cudaMemcpyAsync failed");
  if (e_kin->typed_data() != e_kin_in.typed_data() &&
      cudaMemcpyAsync(e_kin->typed_data(), e_kin_in.typed_data(), sizeof(double),
                      cudaMemcpyDeviceToDevice, st) != cudaSuccess)
    return ffi::Error::Internal("cudaMemcpyAsync failed");
  return Check(jrb_allreduce_rho(P(plan), D(rho), D(e_kin), st));
}

// ---- optax.adam on the device (opt_utils.py:153-168) ------------------------------------------
ffi::Error AdamTick(cudaStream_t st, double b1, double b2, F64 state_in, RF64 state) {
  if (state->typed_data() != state_in.typed_data() &&
      cudaMemcpyAsync(state->typed_data(), state_in.typed_data(), 4 * sizeof(double),
                      cudaMemcpyDeviceToDevice, st) != cudaSuccess)
    return ffi::Error::Internal("cudaMemcpyAsync failed");
  return Check(jrb_adam_tick(D(state), b1, b2, st));
}

}  // namespace

#define JRB_PLAN_BINDING() \
  ffi::Ffi::Bind().Ctx<Stream>().Attr<int64_t>("plan")

XLA_FFI_DEFINE_HANDLER_SYMBOL(JrbQrFwd, QrFwd,
                              JRB_PLAN_BINDING().Arg<F64>().Arg<F64>().Ret<C128>().Ret<C128>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JrbQrBwd, QrBwd,
                              JRB_PLAN_BINDING().Arg<C128>().Arg<C128>().Arg<C128>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JrbDensity, Density,
                              JRB_PLAN_BINDING().Arg<C128>().Arg<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JrbWaveGrid, WaveGrid, JRB_PLAN_BINDING().Arg<C128>().Ret<C128>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JrbExpand, Expand, JRB_PLAN_BINDING().Arg<C128>().Ret<C128>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JrbSqueeze, Squeeze, JRB_PLAN_BINDING().Arg<C128>().Ret<C128>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JrbKinetic, Kinetic, JRB_PLAN_BINDING().Arg<C128>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JrbGridPotential, GridPotential,
                              JRB_PLAN_BINDING().Attr<int32_t>("xc_id").Attr<int32_t>("kohn_sham")
                                .Arg<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JrbPotential, Potential,
                              JRB_PLAN_BINDING().Attr<int32_t>("xc_id").Attr<int32_t>("kohn_sham")
                                .Attr<int32_t>("parts").Arg<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JrbDensityReciprocal, DensityReciprocal,
                              JRB_PLAN_BINDING().Arg<F64>().Ret<C128>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JrbHpsi, Hpsi, JRB_PLAN_BINDING().Arg<C128>().Arg<F64>().Ret<C128>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JrbHpsiPrepare, HpsiPrepare, JRB_PLAN_BINDING().Arg<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JrbHpsiPrepared, HpsiPrepared,
                              JRB_PLAN_BINDING().Arg<C128>().Arg<F64>().Ret<C128>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JrbBandExpect, BandExpect,
                              JRB_PLAN_BINDING().Arg<C128>().Arg<C128>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JrbHamiltonianMatrix, HamiltonianMatrix,
                              JRB_PLAN_BINDING().Arg<C128>().Arg<C128>().Ret<C128>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JrbFft3d, Fft3d,
                              JRB_PLAN_BINDING().Attr<int32_t>("direction").Arg<C128>().Ret<C128>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JrbEvalBegin, EvalBegin,
                              JRB_PLAN_BINDING().Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JrbEvalFinish, EvalFinish,
                              JRB_PLAN_BINDING().Attr<int32_t>("xc_id").Arg<F64>().Arg<F64>().Arg<F64>()
                                .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JrbEval, Eval,
                              JRB_PLAN_BINDING().Attr<int32_t>("xc_id").Arg<F64>().Arg<F64>().Arg<F64>()
                                .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JrbAllreduceRho, AllreduceRho,
                              JRB_PLAN_BINDING().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JrbAdamTick, AdamTick,
                              ffi::Ffi::Bind().Ctx<Stream>().Attr<double>("b1").Attr<double>("b2")
                                .Arg<F64>().Ret<F64>());
