// Host build of jrystal_b200/csrc/xc_functionals.cuh: prints eps_xc and its derivatives for a
// table of densities / gradient norms; tests/test_kernel_math_host.py compares the numbers with
// the oracle (oracle/reference_port.py: _eps_lda_x, _eps_lda_c_pw, _eps_gga_x_pbe, _eps_gga_c_pbe
// differentiated by torch autograd).  One line per point:
//   lda <xc_id> <n> <eps> <deps/dn>
//   gga <xc_id> <rho> <sigma> <eps> <deps/drho> <deps/dsigma>
//   pol <xc_id> <rho_up> <rho_dn> <eps> <deps/drho_up> <deps/drho_dn>
#include <cstdio>
#include <initializer_list>
#include "../../jrystal_b200/csrc/xc_functionals.cuh"

int main() {
  const double dens[] = {0.0, 1e-16, 3e-13, 1e-9, 1e-6, 1e-3, 0.02, 0.37, 1.0, 4.2, 55.0, 900.0};
  const double sig_rel[] = {0.0, 1e-6, 0.01, 0.5, 3.0, 40.0};  // sigma = rel * rho^(8/3) scale
  for (int id : {JRB_XC_LDA_X, JRB_XC_LDA_X_C_PW})
    for (double n : dens) {
      double e, de;
      jrb::lda_eps(id, n, e, de);
      std::printf("lda %d %.17g %.17g %.17g\n", id, n, e, de);
    }
  for (int id : {JRB_XC_GGA_X_PBE, JRB_XC_GGA_PBE})
    for (double n : dens)
      for (double r : sig_rel) {
        if (n == 0.0) continue;  // the kernel never differentiates at rho = 0 (dcbrt divides by rho)
        const double sigma = r * std::pow(n, 8.0 / 3.0) * 30.0;
        const jrb::Dual e = jrb::pbe_eps(id, n, sigma);
        std::printf("gga %d %.17g %.17g %.17g %.17g %.17g\n", id, n, sigma, e.v, e.r, e.s);
      }
  // two spin channels: from unpolarised (zeta = 0) over partial to full polarisation (zeta = +-1)
  const double ntot[] = {1e-9, 1e-4, 0.02, 0.37, 4.2, 55.0};
  const double zeta[] = {0.0, 0.1, -0.35, 0.8, 0.999, -1.0, 1.0};
  for (int id : {JRB_XC_LDA_X, JRB_XC_LDA_X_C_PW})
    for (double n : ntot)
      for (double z : zeta) {
        const double ru = 0.5 * n * (1.0 + z), rd = 0.5 * n * (1.0 - z);
        double e, du, dd;
        jrb::lda_pol_eps(id, ru, rd, e, du, dd);
        std::printf("pol %d %.17g %.17g %.17g %.17g %.17g\n", id, ru, rd, e, du, dd);
      }
  return 0;
}
