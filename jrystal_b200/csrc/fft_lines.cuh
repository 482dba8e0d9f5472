// Two-stage 1-D line FFT shared by every pencil pass (z, y, x; inverse and forward).
//
// A line of length N = RA * RB is transformed by TPL cooperating threads:
//   stage A: RB butterflies of radix RA on the strided inputs  x[jA + m RB]   (m < RA)
//            -> exchange buffer y[jA RA + q]
//   stage B: RA butterflies of radix RB on y[jB + m RA] * w_N^{jB m}            (m < RB)
//            -> natural-order outputs X[jB + m RA]
// (Stockham autosort; one shared-memory exchange per line.)  The forward transform uses
// the radices in the opposite order of the inverse one, so that the outputs a thread
// holds after an inverse stage B are exactly the inputs of its forward stage A: the
// x-pass of the Hamiltonian apply (inverse -> * v_eff -> forward) chains in registers.
//
// Data layout contract: the element index runs with stride NB (complex numbers) in both
// shared and global memory and the NB "band lanes" are adjacent, so a quarter warp always
// touches 8 x 16 B = 128 contiguous bytes: coalesced in HBM/L2 and conflict free in smem.
//
// Lengths N <= 16 are one in-register butterfly (RB = 1 or RA = 1 degenerate stage).
#pragma once
#include "dft_small.cuh"

namespace jrb {

template <int N>
struct LinePlan;  // r1, r2: inverse-order radices (r1 * r2 == N); tpl: threads per line

#define JRB_LINE_PLAN(N_, R1_, R2_, TPL_)                      \
  template <>                                                  \
  struct LinePlan<N_> {                                        \
    static constexpr int n = N_, r1 = R1_, r2 = R2_, tpl = TPL_; \
    static_assert(R1_ * R2_ == N_, "bad plan");               \
  };

// single butterfly
JRB_LINE_PLAN(2, 2, 1, 1)
JRB_LINE_PLAN(3, 3, 1, 1)
JRB_LINE_PLAN(4, 4, 1, 1)
JRB_LINE_PLAN(5, 5, 1, 1)
JRB_LINE_PLAN(6, 6, 1, 1)
JRB_LINE_PLAN(7, 7, 1, 1)
JRB_LINE_PLAN(8, 8, 1, 1)
JRB_LINE_PLAN(9, 9, 1, 1)
JRB_LINE_PLAN(10, 10, 1, 1)
JRB_LINE_PLAN(12, 12, 1, 1)
JRB_LINE_PLAN(14, 14, 1, 1)
JRB_LINE_PLAN(15, 15, 1, 1)
JRB_LINE_PLAN(16, 16, 1, 1)
// two stages; tpl = gcd(r1, r2) keeps every thread busy in both stages
JRB_LINE_PLAN(18, 3, 6, 3)
JRB_LINE_PLAN(20, 2, 10, 2)
JRB_LINE_PLAN(24, 4, 6, 2)
JRB_LINE_PLAN(25, 5, 5, 5)
JRB_LINE_PLAN(27, 3, 9, 3)
JRB_LINE_PLAN(28, 2, 14, 2)
JRB_LINE_PLAN(30, 3, 10, 1)
JRB_LINE_PLAN(32, 4, 8, 4)
JRB_LINE_PLAN(36, 6, 6, 6)
JRB_LINE_PLAN(40, 4, 10, 2)
JRB_LINE_PLAN(45, 3, 15, 3)
JRB_LINE_PLAN(48, 4, 12, 4)
JRB_LINE_PLAN(49, 7, 7, 7)
JRB_LINE_PLAN(50, 5, 10, 5)
JRB_LINE_PLAN(54, 6, 9, 3)
JRB_LINE_PLAN(56, 4, 14, 2)
JRB_LINE_PLAN(60, 6, 10, 2)
JRB_LINE_PLAN(64, 8, 8, 8)
JRB_LINE_PLAN(72, 6, 12, 6)
JRB_LINE_PLAN(80, 8, 10, 2)
JRB_LINE_PLAN(81, 9, 9, 9)
JRB_LINE_PLAN(90, 6, 15, 3)
JRB_LINE_PLAN(96, 8, 12, 4)
JRB_LINE_PLAN(100, 10, 10, 10)
JRB_LINE_PLAN(108, 9, 12, 3)
JRB_LINE_PLAN(112, 8, 14, 2)
JRB_LINE_PLAN(120, 10, 12, 2)
JRB_LINE_PLAN(128, 16, 8, 8)
JRB_LINE_PLAN(144, 12, 12, 12)
JRB_LINE_PLAN(160, 10, 16, 2)
JRB_LINE_PLAN(192, 12, 16, 4)
JRB_LINE_PLAN(256, 16, 16, 16)

template <int N, int DIR>
struct LineFFT {
  using P = LinePlan<N>;
  static constexpr int RA = DIR > 0 ? P::r1 : P::r2;  // first-stage radix
  static constexpr int RB = N / RA;                    // second-stage radix
  static constexpr int TPL = P::tpl;
  static constexpr int CA = (RB + TPL - 1) / TPL;  // stage-A butterflies per thread
  static constexpr int CB = (RA + TPL - 1) / TPL;  // stage-B butterflies per thread
  // Long lines keep their inter-stage twiddles in the (L1-resident) table instead of registers:
  // 16 elements per thread leave no room for CB x (RB - 1) complex twiddles next to them, and
  // the register count decides how many CTAs are resident (profiles/r02_ncu_c3a_baseline.md).
  // In that mode tw[0][0] only carries the table pointer.
  static constexpr bool TW_TABLE = N > 96;
  static constexpr int NTW = TW_TABLE ? 1 : (RB > 1 ? RB - 1 : 1);

  static JRB_HD const cplx* tw_table_ptr(const cplx& slot) {
    union { double d; const cplx* p; } u;
    u.d = slot.x;
    return u.p;
  }
  // w_N^{DIR * jB * m} from the table
  template <bool CONJ_TW>
  static JRB_HD cplx tw_lookup(const cplx* table, int jB, int m) {
#if defined(__CUDA_ARCH__)
    const double2 t = __ldg(reinterpret_cast<const double2*>(table) + (jB * m) % N);
    const cplx w = cmake(t.x, t.y);
#else
    const cplx w = table[(jB * m) % N];
#endif
    return ((DIR > 0) != CONJ_TW) ? cconj(w) : w;
  }

  // element index a thread touches: input of stage A / output of stage B
  static JRB_HD int idxA(int i, int m, int tj) { return (tj + i * TPL) + m * RB; }
  static JRB_HD int idxB(int i, int m, int tj) { return (tj + i * TPL) + m * RA; }
  static JRB_HD bool activeA(int i, int tj) { return (RB % TPL == 0) || (tj + i * TPL < RB); }
  static JRB_HD bool activeB(int i, int tj) { return (RA % TPL == 0) || (tj + i * TPL < RA); }

  // tw[i][m-1] = w_N^{DIR * jB_i * m}; `table` holds exp(-2 pi i t / N), t < N.
  static JRB_HD void load_twiddles(cplx (&tw)[CB][NTW], const cplx* __restrict__ table,
                                   int tj) {
    if constexpr (RA == 1) return;  // single butterfly: all twiddles are 1
    if constexpr (TW_TABLE) {
      union { double d; const cplx* p; } u;
      u.d = 0.0;
      u.p = table;
      tw[0][0] = cmake(u.d, 0.0);
      return;
    }
#pragma unroll
    for (int i = 0; i < CB; ++i) {
      int jB = tj + i * TPL;
      if (!activeB(i, tj)) jB = 0;
#pragma unroll
      for (int m = 1; m < RB; ++m) {
        cplx w = table[(jB * m) % N];
        tw[i][m - 1] = DIR > 0 ? cconj(w) : w;
      }
    }
  }

  // Stage A: va[i][m] holds x[idxA(i, m)]; butterflies, then scatter into the exchange
  // buffer `sm` (element stride S complex numbers, already offset to this line/lane).
  template <int S>
  static JRB_HD void stageA_store(cplx (&va)[CA][RA], cplx* sm, int tj) {
#pragma unroll
    for (int i = 0; i < CA; ++i) {
      if (activeA(i, tj)) {
        Dft<RA, DIR>::run(va[i]);
        const int jA = tj + i * TPL;
#pragma unroll
        for (int q = 0; q < RA; ++q) sm[(jA * RA + q) * S] = va[i][q];
      }
    }
  }

  // Band-limited variants (radix-8 stages only, see dft_small.cuh): stage A whose butterfly
  // inputs are non-zero only for m in {0, 1, 6, 7}; stage B that only produces those outputs.
  static constexpr bool SPARSE_OK = (RA == 8 && RB == 8);
  template <int S>
  static JRB_HD void stageA_store_sparse(cplx (&va)[CA][RA], cplx* sm, int tj) {
    static_assert(RA == 8, "sparse butterflies are radix 8");
#pragma unroll
    for (int i = 0; i < CA; ++i) {
      if (activeA(i, tj)) {
        dft8_sparse_in<DIR>(va[i]);
        const int jA = tj + i * TPL;
#pragma unroll
        for (int q = 0; q < RA; ++q) sm[(jA * RA + q) * S] = va[i][q];
      }
    }
  }
  template <int S, bool CONJ_TW = false>
  static JRB_HD void stageB_load_sparse(cplx (&vb)[CB][RB], const cplx* sm,
                                        const cplx (&tw)[CB][NTW], int tj) {
    static_assert(RB == 8, "sparse butterflies are radix 8");
#pragma unroll
    for (int i = 0; i < CB; ++i) {
      if (activeB(i, tj)) {
        const int jB = tj + i * TPL;
        vb[i][0] = sm[jB * S];
#pragma unroll
        for (int m = 1; m < RB; ++m) {
          if constexpr (CONJ_TW) {
            vb[i][m] = cmul(sm[(jB + m * RA) * S], cconj(tw[i][m - 1]));
          } else {
            vb[i][m] = cmul(sm[(jB + m * RA) * S], tw[i][m - 1]);
          }
        }
        dft8_sparse_out<DIR>(vb[i]);
      }
    }
  }

  // Stage B: gather from the exchange buffer, twiddle, butterflies; vb[i][m] then holds
  // X[idxB(i, m)].
  // CONJ_TW: multiply by the conjugates of `tw` (lets a forward transform reuse the inverse
  // transform's twiddle registers when both stages have the same radix).
  template <int S, bool CONJ_TW = false>
  static JRB_HD void stageB_load(cplx (&vb)[CB][RB], const cplx* sm,
                                 const cplx (&tw)[CB][NTW], int tj) {
#pragma unroll
    for (int i = 0; i < CB; ++i) {
      if (activeB(i, tj)) {
        const int jB = tj + i * TPL;
        vb[i][0] = sm[jB * S];
#pragma unroll
        for (int m = 1; m < RB; ++m) {
          if constexpr (RA == 1) {
            vb[i][m] = sm[(jB + m * RA) * S];
          } else if constexpr (TW_TABLE) {
            vb[i][m] = cmul(sm[(jB + m * RA) * S],
                            tw_lookup<CONJ_TW>(tw_table_ptr(tw[0][0]), jB, m));
          } else if constexpr (CONJ_TW) {
            vb[i][m] = cmul(sm[(jB + m * RA) * S], cconj(tw[i][m - 1]));
          } else {
            vb[i][m] = cmul(sm[(jB + m * RA) * S], tw[i][m - 1]);
          }
        }
        Dft<RB, DIR>::run(vb[i]);
      }
    }
  }
};

}  // namespace jrb
