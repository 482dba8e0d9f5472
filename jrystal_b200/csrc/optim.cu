// Adam update of the plane-wave parameters on the device (the energy-mode driver's optimiser,
// calc/calc_ground_state_energy_all_electrons.py:175-181 with optax.adam, opt_utils.py:153-168).
//
// optax.adam(lr, b1, b2, eps):  m = b1 m + (1 - b1) g,  v = b2 v + (1 - b2) g^2,
//   p -= lr * (m / (1 - b1^t)) / (sqrt(v / (1 - b2^t)) + eps),  t counted from 1.
// The step counter and the two bias corrections live in device memory (state[0..2]) so that the
// whole optimisation step can be replayed as a CUDA graph: jrb_adam_tick advances them once per
// step, jrb_adam_apply streams one parameter array (HBM bound: 4 reads + 3 writes of 8 B per
// element, 128-bit accesses, grid sized to the SM count).
#include <algorithm>
#include <cstdint>

#include "plan.h"

namespace jrb {

__global__ void k_adam_tick(double* __restrict__ state, double b1, double b2) {
  const double t = state[0] + 1.0;
  state[0] = t;
  state[1] = 1.0 / (1.0 - pow(b1, t));
  state[2] = 1.0 / (1.0 - pow(b2, t));
}

__global__ void __launch_bounds__(256)
k_adam_apply(long long n, double* __restrict__ p, const double* __restrict__ g,
             double* __restrict__ m, double* __restrict__ v, double lr, double b1, double b2,
             double eps, const double* __restrict__ state) {
  const double c1 = state[1], c2 = state[2];
  const long long n2 = n >> 1;
  double2* p2 = reinterpret_cast<double2*>(p);
  const double2* g2 = reinterpret_cast<const double2*>(g);
  double2* m2 = reinterpret_cast<double2*>(m);
  double2* v2 = reinterpret_cast<double2*>(v);
  auto upd = [&](double& pp, double gg, double& mm, double& vv) {
    mm = b1 * mm + (1.0 - b1) * gg;
    vv = b2 * vv + (1.0 - b2) * gg * gg;
    pp -= lr * (mm * c1) / (sqrt(vv * c2) + eps);
  };
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2;
       i += (long long)gridDim.x * blockDim.x) {
    double2 pp = p2[i], mm = m2[i], vv = v2[i];
    const double2 gg = g2[i];
    upd(pp.x, gg.x, mm.x, vv.x);
    upd(pp.y, gg.y, mm.y, vv.y);
    p2[i] = pp; m2[i] = mm; v2[i] = vv;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const long long i = n - 1;
    double pp = p[i], mm = m[i], vv = v[i];
    upd(pp, g[i], mm, vv);
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}

}  // namespace jrb

using namespace jrb;

extern "C" int jrb_adam_tick(double* state, double b1, double b2, jrb_stream st) {
  if (!state) {
    set_error("jrb_adam_tick: null state");
    return JRB_EINVAL;
  }
  k_adam_tick<<<1, 1, 0, reinterpret_cast<cudaStream_t>(st)>>>(state, b1, b2);
  JRB_CHECK_LAUNCH("k_adam_tick");
  return 0;
}

extern "C" int jrb_adam_apply(int64_t n, double* param, const double* grad, double* m, double* v,
                              double lr, double b1, double b2, double eps, const double* state,
                              jrb_stream st) {
  if (n < 0 || !param || !grad || !m || !v || !state) {
    set_error("jrb_adam_apply: bad argument");
    return JRB_EINVAL;
  }
  if (n == 0) return 0;
  if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
       reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) {
    set_error("jrb_adam_apply: arrays must be 16-byte aligned");
    return JRB_EINVAL;
  }
  const long long want = ((n >> 1) + 255) / 256;
  const int blocks = (int)std::max<long long>(1, std::min<long long>(want, 148 * 8));
  k_adam_apply<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(st)>>>(n, param, grad, m, v, lr, b1,
                                                                      b2, eps, state);
  JRB_CHECK_LAUNCH("k_adam_apply");
  return 0;
}
