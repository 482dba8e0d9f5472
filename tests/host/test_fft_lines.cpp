// Host-side emulation of the two-stage line FFT (threads run sequentially per phase)
// against a naive DFT, for every planned length, both directions, plus the
// inverse -> forward register chaining used by the H-apply x-pass.
#include <cmath>
#include <cstdio>
#include <complex>
#include <vector>
#include "../../jrystal_b200/csrc/fft_lines.cuh"
using namespace jrb;
typedef std::complex<double> cd;

static std::vector<cd> naive(const std::vector<cd>& x, int dir) {
  int n = x.size();
  std::vector<cd> out(n);
  for (int k = 0; k < n; ++k) {
    std::complex<long double> acc = 0;
    for (int j = 0; j < n; ++j) {
      long double ang = dir * 2.0L * M_PIl * ((long long)j * k % n) / n;
      acc += std::complex<long double>(x[j]) * std::complex<long double>(cosl(ang), sinl(ang));
    }
    out[k] = cd(acc);
  }
  return out;
}

template <int N, int DIR>
std::vector<cd> run_line(const std::vector<cd>& x) {
  using F = LineFFT<N, DIR>;
  std::vector<cplx> table(N), sm(N);
  for (int t = 0; t < N; ++t) table[t] = cmake(std::cos(2 * M_PI * t / N), -std::sin(2 * M_PI * t / N));
  std::vector<cd> out(N);
  for (int tj = 0; tj < F::TPL; ++tj) {
    cplx va[F::CA][F::RA];
    for (int i = 0; i < F::CA; ++i)
      for (int m = 0; m < F::RA; ++m)
        if (F::activeA(i, tj)) { cd v = x[F::idxA(i, m, tj)]; va[i][m] = cmake(v.real(), v.imag()); }
    F::template stageA_store<1>(va, sm.data(), tj);
  }
  for (int tj = 0; tj < F::TPL; ++tj) {
    cplx tw[F::CB][F::NTW];
    cplx vb[F::CB][F::RB];
    F::load_twiddles(tw, table.data(), tj);
    F::template stageB_load<1>(vb, sm.data(), tw, tj);
    for (int i = 0; i < F::CB; ++i)
      for (int m = 0; m < F::RB; ++m)
        if (F::activeB(i, tj)) out[F::idxB(i, m, tj)] = cd(vb[i][m].x, vb[i][m].y);
  }
  return out;
}

template <int N>
int one() {
  std::vector<cd> x(N);
  for (int i = 0; i < N; ++i) x[i] = cd(std::sin(1.3 * i + 0.2) + 0.01 * i, std::cos(0.7 * i * i + 1.1));
  int bad = 0;
  double e[2];
  for (int d = 0; d < 2; ++d) {
    auto ref = naive(x, d ? 1 : -1);
    auto got = d ? run_line<N, 1>(x) : run_line<N, -1>(x);
    double err = 0, nrm = 0;
    for (int k = 0; k < N; ++k) { err = std::fmax(err, std::abs(ref[k] - got[k])); nrm = std::fmax(nrm, std::abs(ref[k])); }
    e[d] = err / nrm;
    if (!(e[d] < 5e-15)) bad = 1;
  }
  // chaining contract
  using FI = LineFFT<N, 1>;
  using FF = LineFFT<N, -1>;
  static_assert(FI::CB == FF::CA && FI::RB == FF::RA, "chain shape");
  for (int tj = 0; tj < FI::TPL; ++tj)
    for (int i = 0; i < FI::CB; ++i)
      for (int m = 0; m < FI::RB; ++m)
        if (FI::idxB(i, m, tj) != FF::idxA(i, m, tj) || FI::activeB(i, tj) != FF::activeA(i, tj)) bad = 1;
  std::printf("N=%3d (%2dx%2d tpl %2d) fwd %.2e inv %.2e %s\n", N, LinePlan<N>::r1, LinePlan<N>::r2,
              LinePlan<N>::tpl, e[0], e[1], bad ? "BAD" : "ok");
  return bad;
}

int main() {
  int bad = 0;
#define T(n) bad += one<n>();
  T(2) T(3) T(4) T(5) T(6) T(7) T(8) T(9) T(10) T(12) T(14) T(15) T(16) T(18) T(20) T(24) T(25) T(27) T(28)
  T(30) T(32) T(36) T(40) T(45) T(48) T(49) T(50) T(54) T(56) T(60) T(64) T(72) T(80) T(81) T(90) T(96) T(100)
  T(108) T(112) T(120) T(128) T(144) T(160) T(192) T(256)
  std::printf(bad ? "FAIL\n" : "OK\n");
  return bad;
}
