// Exchange-correlation functionals evaluated per grid point by the fused grid kernels
// (grid_kernels.cu: k_veff_xc, k_gga_local): closed-form eps_xc and its derivatives.
// Device code (__device__ __forceinline__ under nvcc); the same source compiles as plain inline
// functions for the HOST so that tests/host/test_xc_functionals.cpp can check the arithmetic on
// the CPU against the oracle (tests/test_kernel_math_host.py).
#pragma once
#include <cmath>

#include "../../include/jrystal_b200.h"

#if defined(__CUDACC__)
#define JRB_XC_FN __device__ __forceinline__
#else
#define JRB_XC_FN inline
#endif

namespace jrb {

// LDA energy density per particle and its derivative for the unpolarised gas.
JRB_XC_FN void lda_eps(int xc_id, double n, double& eps, double& deps) {
  const double thr = 1e-15;  // LibXC dens_threshold
  eps = 0.0;
  deps = 0.0;
  if (!(n > thr)) return;
  const double cx = -0.73855876638202240588;  // -3/4 (3/pi)^(1/3)
  const double c13 = cbrt(n);
  eps = cx * c13;
  deps = eps / (3.0 * n);
  if (xc_id == JRB_XC_LDA_X_C_PW) {
    const double A = 0.031091, a1 = 0.21370, b1 = 7.5957, b2 = 3.5876, b3 = 1.6382, b4 = 0.49294;
    const double rs = cbrt(3.0 / (4.0 * M_PI * n));
    const double sr = sqrt(rs);
    const double Q = 2.0 * A * (b1 * sr + b2 * rs + b3 * rs * sr + b4 * rs * rs);
    const double dQ = 2.0 * A * (0.5 * b1 / sr + b2 + 1.5 * b3 * sr + 2.0 * b4 * rs);
    const double L = log1p(1.0 / Q);
    const double ec = -2.0 * A * (1.0 + a1 * rs) * L;
    const double dec_drs = -2.0 * A * a1 * L + 2.0 * A * (1.0 + a1 * rs) * dQ / (Q * Q + Q);
    eps += ec;
    deps += dec_drs * (-rs / (3.0 * n));
  }
}


// ---------------------------------------------------------------------------------------
// GGA (PBE exchange and correlation, unpolarised).  The reference evaluates
// eps_xc(rho, sigma) per grid point with sigma = sum_j |d_j|^2, d_j = ifftn(i G_j fftn(rho))
// (xc.py:67-112, 242-253) and lets jax.grad differentiate E_xc = (Omega/N) sum rho eps through
// both arguments.  Here the same energy is accumulated and its exact discrete derivative is
// formed by hand:  dE/d rho(r) = de/d rho - 2 Re ifftn( sum_j i G_j fftn( de/d sigma * d_j ) )
// with e = rho eps (the adjoint of D_j = ifftn i G_j fftn on the FFT grid is u -> -conj(D_j conj u),
// Nyquist bins included, so the result matches autograd to rounding).  de/d rho and de/d sigma
// come from forward-mode duals of the closed-form eps(rho, sigma).
struct Dual {  // value, d/d rho, d/d sigma
  double v, r, s;
};
JRB_XC_FN Dual dmk(double v, double r = 0.0, double s = 0.0) {
  Dual d;
  d.v = v; d.r = r; d.s = s;
  return d;
}
JRB_XC_FN Dual operator+(Dual a, Dual b) { return dmk(a.v + b.v, a.r + b.r, a.s + b.s); }
JRB_XC_FN Dual operator-(Dual a, Dual b) { return dmk(a.v - b.v, a.r - b.r, a.s - b.s); }
JRB_XC_FN Dual operator*(Dual a, Dual b) {
  return dmk(a.v * b.v, a.r * b.v + a.v * b.r, a.s * b.v + a.v * b.s);
}
JRB_XC_FN Dual operator*(double c, Dual a) { return dmk(c * a.v, c * a.r, c * a.s); }
JRB_XC_FN Dual operator+(double c, Dual a) { return dmk(c + a.v, a.r, a.s); }
JRB_XC_FN Dual operator/(Dual a, Dual b) {
  const double q = a.v / b.v, ib = 1.0 / b.v;
  return dmk(q, (a.r - q * b.r) * ib, (a.s - q * b.s) * ib);
}
JRB_XC_FN Dual dchain(Dual a, double f, double df) { return dmk(f, df * a.r, df * a.s); }
JRB_XC_FN Dual dsqrt(Dual a) { const double f = sqrt(a.v); return dchain(a, f, 0.5 / f); }
JRB_XC_FN Dual dcbrt(Dual a) { const double f = cbrt(a.v); return dchain(a, f, f / (3.0 * a.v)); }
JRB_XC_FN Dual dlog1p(Dual a) { return dchain(a, log1p(a.v), 1.0 / (1.0 + a.v)); }
JRB_XC_FN Dual dexpm1(Dual a) { const double f = expm1(a.v); return dchain(a, f, f + 1.0); }

// eps_xc(rho, sigma) of gga_x_pbe (+ gga_c_pbe) as LibXC defines them (jax_xc 0.0.8 translates
// LibXC's maple sources; not vendored -> parity unpinned)
JRB_XC_FN Dual pbe_eps(int xc_id, double rho, double sigma) {
  const Dual n = dmk(rho, 1.0, 0.0), sg = dmk(sigma, 0.0, 1.0);
  Dual eps = dmk(0.0);
  const Dual n13 = dcbrt(n);
  if (rho > 1e-15) {  // gga_x_pbe dens_threshold
    const double kappa = 0.8040, mu = 0.2195149727645171;
    const double cx = -0.73855876638202240588;                 // -3/4 (3/pi)^(1/3)
    const double c_s2 = 1.0 / (4.0 * 9.5707800006273513);     // 1 / (4 (3 pi^2)^(2/3))
    const Dual n83 = (n * n) * (n13 * n13);
    const Dual s2 = c_s2 * (sg / n83);
    const Dual fx = (1.0 + kappa) + (-kappa * kappa) * (dmk(1.0) / (kappa + mu * s2));
    eps = eps + (cx * n13) * fx;
  }
  if (xc_id == JRB_XC_GGA_PBE && rho > 1e-12) {  // gga_c_pbe dens_threshold
    // PW92 with LibXC's pw_mod parameters (what gga_c_pbe includes), zeta = 0
    const double A = 0.0310907, a1 = 0.21370, b1 = 7.5957, b2 = 3.5876, b3 = 1.6382, b4 = 0.49294;
    const double beta = 0.06672455060314922, gamma = 0.031090690869654895;  // (1 - ln 2) / pi^2
    const Dual rs = 0.62035049089940001667 * (dmk(1.0) / n13);  // (3 / (4 pi))^(1/3) n^(-1/3)
    const Dual sr = dsqrt(rs);
    const Dual q = (2.0 * A) * (b1 * sr + b2 * rs + b3 * (rs * sr) + b4 * (rs * rs));
    const Dual ec = (-2.0 * A) * ((1.0 + a1 * rs) * dlog1p(dmk(1.0) / q));
    // t^2 = sigma / (4 ks^2 n^2), ks^2 = 4 kF / pi, kF = (3 pi^2 n)^(1/3)
    const double c_t2 = M_PI / (16.0 * 3.0936677262801360);  // pi / (16 (3 pi^2)^(1/3))
    const Dual t2 = c_t2 * (sg / ((n * n) * n13));
    const Dual aa = (beta / gamma) * (dmk(1.0) / dexpm1((-1.0 / gamma) * ec));
    const Dual f1 = t2 + aa * (t2 * t2);
    const Dual f2 = (beta / gamma) * (f1 / (1.0 + aa * f1));
    eps = eps + ec + gamma * dlog1p(f2);
  }
  return eps;
}

// x^(4/3) with its derivative, regular at x = 0 (fully polarised points)
JRB_XC_FN Dual dpow43(Dual a) {
  const double c = cbrt(a.v > 0.0 ? a.v : 0.0);
  return dchain(a, a.v * c, (4.0 / 3.0) * c);
}
JRB_XC_FN Dual pw92_g(Dual rs, Dual sr, double a, double a1, double b1, double b2, double b3,
                      double b4) {
  const Dual q = (2.0 * a) * (b1 * sr + b2 * rs + b3 * (rs * sr) + b4 * (rs * rs));
  return (-2.0 * a) * ((1.0 + a1 * rs) * dlog1p(dmk(1.0) / q));
}

// Two spin channels (xc.py:54-64): eps per particle and d eps / d rho_up, d eps / d rho_dn.
//   exchange:    eps_x = 1/2 [eps_x(2 rho_up) + eps_x(2 rho_dn)]  (the reference's spin-scaling call)
//   correlation: LibXC lda_c_pw, polarised (xc.py:60-61): PW92 with the spin interpolation
//                g1 + zeta^4 f (g2 - g1 + g3 / f''(0)) - f g3 / f''(0); Dual slots r, s = d/d rho_up,
//                d/d rho_dn
JRB_XC_FN void lda_pol_eps(int xc_id, double ru, double rd, double& eps, double& deu, double& ded) {
  double eu, ed;
  lda_eps(JRB_XC_LDA_X, 2.0 * ru, eu, deu);
  lda_eps(JRB_XC_LDA_X, 2.0 * rd, ed, ded);
  eps = 0.5 * (eu + ed);
  if (xc_id != JRB_XC_LDA_X_C_PW) return;
  const double nt = ru + rd;
  if (!(nt > 1e-15)) return;
  const Dual up = dmk(ru, 1.0, 0.0), dn = dmk(rd, 0.0, 1.0);
  const Dual n = up + dn;
  const Dual rs = 0.62035049089940001667 * (dmk(1.0) / dcbrt(n));
  const Dual sr = dsqrt(rs);
  Dual zeta = (up - dn) / n;
  if (zeta.v > 1.0) zeta = dmk(1.0);
  if (zeta.v < -1.0) zeta = dmk(-1.0);
  const Dual g1 = pw92_g(rs, sr, 0.031091, 0.21370, 7.5957, 3.5876, 1.6382, 0.49294);
  const Dual g2 = pw92_g(rs, sr, 0.015545, 0.20548, 14.1189, 6.1977, 3.3662, 0.62517);
  const Dual g3 = pw92_g(rs, sr, 0.016887, 0.11125, 10.357, 3.6231, 0.88026, 0.49671);
  const double fz20 = 1.709921, fden = 1.0 / (2.5198420997897463295 - 2.0);  // 2^(4/3) - 2
  const Dual f = fden * (-2.0 + (dpow43(1.0 + zeta) + dpow43(1.0 + (-1.0) * zeta)));
  const Dual z2 = zeta * zeta, z4 = z2 * z2;
  const Dual ec = g1 + (z4 * f) * ((g2 - g1) + (1.0 / fz20) * g3) - (1.0 / fz20) * (f * g3);
  eps += ec.v;
  deu += ec.r;
  ded += ec.s;
}

}  // namespace jrb
