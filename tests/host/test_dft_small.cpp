// Host-side check of the in-register DFT templates against a naive O(R^2) DFT.
#include <cmath>
#include <cstdio>
#include <complex>
#include <vector>
#include "../../jrystal_b200/csrc/dft_small.cuh"
using namespace jrb;
template <int R, int DIR>
double check() {
  cplx v[R];
  std::vector<std::complex<double>> x(R);
  for (int i = 0; i < R; ++i) {
    x[i] = {std::sin(1.3 * i + 0.2) + 0.1 * i, std::cos(0.7 * i * i + 1.1)};
    v[i] = cmake(x[i].real(), x[i].imag());
  }
  Dft<R, DIR>::run(v);
  double err = 0, nrm = 0;
  for (int k = 0; k < R; ++k) {
    std::complex<long double> acc = 0;
    for (int j = 0; j < R; ++j) {
      long double ang = DIR * 2.0L * M_PIl * ((j * k) % R) / R;
      acc += std::complex<long double>(x[j]) * std::complex<long double>(cosl(ang), sinl(ang));
    }
    err = std::fmax(err, std::abs(std::complex<double>(acc) - std::complex<double>(v[k].x, v[k].y)));
    nrm = std::fmax(nrm, (double)std::abs(acc));
  }
  return err / nrm;
}
template <int R>
int one() {
  double e1 = check<R, -1>(), e2 = check<R, 1>();
  std::printf("R=%2d fwd %.2e inv %.2e\n", R, e1, e2);
  return (e1 < 1e-14 && e2 < 1e-14) ? 0 : 1;
}
int main() {
  int bad = 0;
  bad += one<1>(); bad += one<2>(); bad += one<3>(); bad += one<4>(); bad += one<5>();
  bad += one<6>(); bad += one<7>(); bad += one<8>(); bad += one<9>(); bad += one<10>();
  bad += one<12>(); bad += one<14>(); bad += one<15>(); bad += one<16>();
  std::printf(bad ? "FAIL\n" : "OK\n");
  return bad;
}
