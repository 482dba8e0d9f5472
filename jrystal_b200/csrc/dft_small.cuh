// In-register complex-FP64 DFTs of length R <= 16 (the butterfly radices of the pencil
// FFT passes).  Everything is resolved at compile time: composite lengths are
// Cooley-Tukey products of the prime kernels with immediate twiddles (trivial twiddles
// 1, -1, +-i cost nothing), so a thread holds the whole butterfly in registers.
//
// Replaces: the per-axis transforms inside jnp.fft.ifftn / fftn as called by
// jrystal/_src/spmd/fft.py:68-75 (XLA lowers those to cuFFT / pocketfft).
//
// DIR = -1: forward  (exp(-2 pi i jk/R), numpy fftn convention)
// DIR = +1: inverse  (exp(+2 pi i jk/R), unnormalised)
#pragma once
#include "twiddle_consts.h"

#if defined(__CUDACC__)
#define JRB_HD __host__ __device__ __forceinline__
#else
#define JRB_HD inline
#endif

namespace jrb {

struct alignas(16) cplx {
  double x, y;
};

JRB_HD cplx cmake(double x, double y) { cplx r; r.x = x; r.y = y; return r; }
JRB_HD cplx cadd(cplx a, cplx b) { return cmake(a.x + b.x, a.y + b.y); }
JRB_HD cplx csub(cplx a, cplx b) { return cmake(a.x - b.x, a.y - b.y); }
JRB_HD cplx cmul(cplx a, cplx b) {
  return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
JRB_HD cplx cscale(cplx a, double s) { return cmake(a.x * s, a.y * s); }
JRB_HD cplx cconj(cplx a) { return cmake(a.x, -a.y); }
// a * (+i) and a * (-i)
JRB_HD cplx cmul_pi(cplx a) { return cmake(-a.y, a.x); }
JRB_HD cplx cmul_mi(cplx a) { return cmake(a.y, -a.x); }

template <int I>
struct Int {
  static constexpr int value = I;
};

template <int I, int N, class F>
JRB_HD void static_for(F&& f) {
  if constexpr (I < N) {
    f(Int<I>{});
    static_for<I + 1, N>(f);
  }
}

// a * exp(DIR * 2 pi i T / R) with the trivial cases free.
template <int R, int T_, int DIR>
JRB_HD cplx mul_root(cplx a) {
  constexpr int T = ((T_ % R) + R) % R;
  if constexpr (T == 0) {
    return a;
  } else if constexpr (2 * T == R) {
    return cmake(-a.x, -a.y);
  } else if constexpr (4 * T == R) {
    return DIR > 0 ? cmul_pi(a) : cmul_mi(a);
  } else if constexpr (4 * T == 3 * R) {
    return DIR > 0 ? cmul_mi(a) : cmul_pi(a);
  } else {
    constexpr double c = cos2pi(R, T);
    constexpr double s = (DIR > 0 ? 1.0 : -1.0) * sin2pi(R, T);
    return cmake(a.x * c - a.y * s, a.x * s + a.y * c);
  }
}

constexpr int smallest_split(int r) {
  // first Cooley-Tukey factor A (R = A * B); 4 before 2 keeps the radix-4 structure.
  if (r % 4 == 0 && r > 4) return 4;
  if (r % 2 == 0 && r > 2) return 2;
  if (r % 3 == 0 && r > 3) return 3;
  if (r % 5 == 0 && r > 5) return 5;
  if (r % 7 == 0 && r > 7) return 7;
  return r;
}

template <int R, int DIR>
struct Dft;

template <int DIR>
struct Dft<1, DIR> {
  static JRB_HD void run(cplx (&v)[1]) {}
};

template <int DIR>
struct Dft<2, DIR> {
  static JRB_HD void run(cplx (&v)[2]) {
    cplx a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  }
};

template <int DIR>
struct Dft<4, DIR> {
  static JRB_HD void run(cplx (&v)[4]) {
    cplx t0 = cadd(v[0], v[2]);
    cplx t1 = csub(v[0], v[2]);
    cplx t2 = cadd(v[1], v[3]);
    cplx d = csub(v[1], v[3]);
    cplx t3 = DIR > 0 ? cmul_pi(d) : cmul_mi(d);
    v[0] = cadd(t0, t2);
    v[1] = cadd(t1, t3);
    v[2] = csub(t0, t2);
    v[3] = csub(t1, t3);
  }
};

// Odd primes: symmetric / antisymmetric split, (P-1)^2/2 real*complex products.
template <int P, int DIR>
struct DftOddPrime {
  static constexpr int H = (P - 1) / 2;
  static JRB_HD void run(cplx (&v)[P]) {
    cplx s[H], d[H];
    static_for<0, H>([&](auto j_) {
      constexpr int j = decltype(j_)::value;
      s[j] = cadd(v[j + 1], v[P - 1 - j]);
      d[j] = csub(v[j + 1], v[P - 1 - j]);
    });
    cplx a0 = v[0];
    cplx sum = a0;
    static_for<0, H>([&](auto j_) { sum = cadd(sum, s[decltype(j_)::value]); });
    v[0] = sum;
    static_for<1, H + 1>([&](auto k_) {
      constexpr int k = decltype(k_)::value;
      cplx A = a0;
      cplx B = cmake(0.0, 0.0);
      static_for<0, H>([&](auto j_) {
        constexpr int j = decltype(j_)::value;
        constexpr int t = ((j + 1) * k) % P;
        constexpr double c = cos2pi(P, t);
        constexpr double sn = sin2pi(P, t);
        A = cmake(A.x + c * s[j].x, A.y + c * s[j].y);
        B = cmake(B.x + sn * d[j].x, B.y + sn * d[j].y);
      });
      // forward: X_k = A - iB, X_{P-k} = A + iB ; inverse: swapped.
      cplx iB = cmul_pi(B);
      if (DIR > 0) {
        v[k] = cadd(A, iB);
        v[P - k] = csub(A, iB);
      } else {
        v[k] = csub(A, iB);
        v[P - k] = cadd(A, iB);
      }
    });
  }
};

template <int DIR>
struct Dft<3, DIR> : DftOddPrime<3, DIR> {};
template <int DIR>
struct Dft<5, DIR> : DftOddPrime<5, DIR> {};
template <int DIR>
struct Dft<7, DIR> : DftOddPrime<7, DIR> {};

// Radix-8 butterflies for band-limited data.  With stride-8 decimation of a 64-point line whose
// non-zero entries are the frequencies |f| < 16 (indices [0,16) u [48,64)), a first-stage
// butterfly only sees inputs m in {0, 1, 6, 7}, and a last-stage butterfly only has to produce
// those outputs.  Exploiting the zeros by hand saves 20 / 8 of the 56 FP64 instructions.
//
// Inputs v[0], v[1], v[6], v[7] (others ignored) -> all 8 outputs.  36 FP64 instructions.
template <int DIR>
JRB_HD void dft8_sparse_in(cplx (&v)[8]) {
  constexpr double r = 0.70710678118654752440;
  const cplx a0 = v[0], a1 = v[1], a6 = v[6], a7 = v[7];
  auto rot = [](cplx z) { return DIR > 0 ? cmul_pi(z) : cmul_mi(z); };  // z * (DIR * i)
  const cplx s = cadd(a1, a7), d = csub(a1, a7);
  const cplx S = cscale(s, r), D = rot(cscale(d, r)), E = rot(d), J = rot(a6);
  const cplx P = cadd(a0, a6), M = csub(a0, a6), Q1 = csub(a0, J), Q3 = cadd(a0, J);
  const cplx SpD = cadd(S, D), DmS = csub(D, S);
  v[0] = cadd(P, s);
  v[4] = csub(P, s);
  v[2] = cadd(M, E);
  v[6] = csub(M, E);
  v[1] = cadd(Q1, SpD);
  v[5] = csub(Q1, SpD);
  v[3] = cadd(Q3, DmS);
  v[7] = csub(Q3, DmS);
}

// All 8 inputs -> outputs v[0], v[1], v[6], v[7] only (others left unspecified).  48 FP64
// instructions.
template <int DIR>
JRB_HD void dft8_sparse_out(cplx (&v)[8]) {
  constexpr double r = 0.70710678118654752440;
  auto rot = [](cplx z) { return DIR > 0 ? cmul_pi(z) : cmul_mi(z); };  // z * (DIR * i)
  const cplx e0 = cadd(v[0], v[4]), e1 = cadd(v[1], v[5]), e2 = cadd(v[2], v[6]), e3 = cadd(v[3], v[7]);
  const cplx o0 = csub(v[0], v[4]), o1 = csub(v[1], v[5]), o2 = csub(v[2], v[6]), o3 = csub(v[3], v[7]);
  v[0] = cadd(cadd(e0, e2), cadd(e1, e3));
  // w^6 = -DIR i
  v[6] = csub(csub(e0, e2), rot(csub(e1, e3)));
  const cplx u = cscale(csub(o1, o3), r), w = rot(cscale(cadd(o1, o3), r));
  const cplx g = rot(o2);
  const cplx A = cadd(cadd(o0, g), u), B = cadd(csub(o0, g), u);
  v[1] = cadd(A, w);
  v[7] = csub(B, w);
}

// Composite: R = A * B, decimation in time, natural-order output.
//   X[k1 + A k2] = sum_{n2<B} w_B^{n2 k2} w_R^{n2 k1} sum_{n1<A} x[n1 B + n2] w_A^{n1 k1}
template <int R, int DIR>
struct Dft {
  static constexpr int A = smallest_split(R);
  static constexpr int B = R / A;
  static_assert(A < R, "prime radix without a specialisation");
  static JRB_HD void run(cplx (&v)[R]) {
    cplx t[B][A];
    static_for<0, B>([&](auto n2_) {
      constexpr int n2 = decltype(n2_)::value;
      cplx u[A];
      static_for<0, A>([&](auto n1_) {
        constexpr int n1 = decltype(n1_)::value;
        u[n1] = v[n1 * B + n2];
      });
      Dft<A, DIR>::run(u);
      static_for<0, A>([&](auto k1_) {
        constexpr int k1 = decltype(k1_)::value;
        t[n2][k1] = mul_root<R, n2 * k1, DIR>(u[k1]);
      });
    });
    static_for<0, A>([&](auto k1_) {
      constexpr int k1 = decltype(k1_)::value;
      cplx u[B];
      static_for<0, B>([&](auto n2_) {
        constexpr int n2 = decltype(n2_)::value;
        u[n2] = t[n2][k1];
      });
      Dft<B, DIR>::run(u);
      static_for<0, B>([&](auto k2_) {
        constexpr int k2 = decltype(k2_)::value;
        v[k1 + A * k2] = u[k2];
      });
    });
  }
};

}  // namespace jrb
