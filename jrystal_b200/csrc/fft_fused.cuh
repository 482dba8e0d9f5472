// Fused y+x pencil passes on z-planes (hand-written sm_100a kernels).
//
// After the z pass the only global intermediate is A[group][z][col][NB] (occupied (x,y)
// columns, 2-5 % of the box).  Work item = (z-plane, band group); the CTAs are persistent and
// split the z-major item list evenly (grid = resident CTA slots, so every SM is busy to the
// end).  Band by band a CTA runs
//   stage  : the NEXT band's plane columns arrive in a shared-memory staging buffer while the
//            current band is being transformed (no exposed global latency): one elected thread
//            issues TMA tensor copies (cp.async.bulk.tensor.2d: a box of {one band = 16 bytes,
//            256 columns} of the 2-D view [rows = (group, z, column)][8 bands] of the column work
//            space lands as 256 consecutive complex numbers, i.e. the band is de-interleaved on
//            the way in) that complete on an mbarrier; the fallback (JRB_NO_TMA=1, or no tensor
//            map) is one cp.async (LDGSTS) per column and thread
//   y stage: y-transforms the occupied x planes into a shared-memory slab Y[xo][y]
//   x stage: zero-pads the occupied x entries to nx, x-transforms every y line and
//     k_yx_density : accumulates f |psi|^2 in registers (a thread owns the same (x, y) points
//                    for every band of a plane) -> one partial-plane write per (CTA, z segment),
//                    summed in a fixed order by k_rho_reduce (deterministic, no atomics);
//     k_yx_vmul    : multiplies by v_eff(r) / N, transforms back along x in registers, keeps the
//                    occupied x planes in Y, then y-transforms back and scatters into A in place.
//     k_x_vmul_cached : the H-apply of jrb_eval when the plan holds the psi(r) cache: reads
//                    psi(r) as k_yx_density stored it (FusedArgs::psi), multiplies, and runs only
//                    the forward x and y transforms (no stage, no inverse).
// v psi(r) and the half-transformed slab never reach global memory; psi(r) does only as the
// optional cache (written once by the density sweep, read once by the H-apply).  Per 64 x 64
// band-plane the recomputing pair needs ~1.7 k cycles of the SM's FP64 pipe and ~1.8 k cycles of
// its shared-memory pipe per transform direction (every element crosses the Stockham exchange
// once per line): bound by those two on-chip pipes with HBM idle, which is what the cache trades
// against.
//
// Thread layout: t = lane + 8 (tj + TPL slot); the 8 "lanes" are 8 ADJACENT LINES of the
// dimension that is not being transformed (x planes in the y stage, y lines in the x stage), so
// every shared-memory access of a quarter warp is 8 consecutive complex numbers (conflict
// free).  Y is double buffered: the x stage of band b and the y stage of band b+1 run without
// a CTA-wide barrier in between (only the 2-warp slot barriers of the exchange buffers).
#pragma once
#include "fft_passes.cuh"

namespace jrb {

struct FusedArgs {
  SphereMaps m;
  cplx* wa;            // A[group][z][col][NB]
  cplx* wout[2];       // fft_fused128.cuh: per-y-parity output copies of A
  const cplx* tw;      // exp(-2 pi i t / n), n = nx = ny
  const cplx* tw64;    // exp(-2 pi i t / (n / 2)) (128-point planes, fft_fused128.cuh)
  const double* focc;  // [groups][NB] occupation / Omega
  const double* veff;  // [nx*ny*nz] of the current spin
  double* rho_part;    // [gridDim.x * segmax][nx*ny] partial density planes of this launch
  int* seg_z;          // [gridDim.x * segmax] z of each partial plane, -1 = unused
  int segmax;
  int nb, ngpk;
  int g0, ngroups;     // first global group id, number of groups in the batch
  double vscale;
  int band_limited;    // every occupied x and y index lies in [0, 2 RB) u [n - 2 RB, n)
  // TMA staging: tensor map (in global memory) of the buffer `wa` points into, viewed as
  // [rows of 8 bands][16 doubles], and the row of wa[0]; null = cp.async staging
  const void* tmap;
  long long row0;
  // psi(r) cache of the whole evaluation (null: none): k_yx_density stores every band-plane it
  // transforms, k_x_vmul_cached reads it back instead of repeating the inverse transforms.
  // Entry of (global group g, z, band): psi_plane complex numbers in the x-stage register layout
  // [round][butterfly][element][thread] of the kernel (both kernels share the thread mapping)
  cplx* psi;
};

// butterfly slots that can be non-zero for band-limited data (see dft_small.cuh)
#define JRB_SPARSE_M(m) ((m) < 2 || (m) >= 6)

constexpr int fused_slots(int tpl) {
  // 81-point lines (9 threads each): 6 slots = 432 threads.  The slots of this length are not
  // whole warps, so the exchanges use CTA-wide barriers either way; six slots cover the 5 line
  // groups of the y stage in one round and the 11 of the x stage in two (four slots: 2 + 3 rounds)
  if (tpl == 9) return 6;
  int s = 256 / (NB * tpl);
  if (s < 1) s = 1;
  while ((NB * tpl * s) % 32 != 0) ++s;
  return s;
}

template <int N>
struct FCfg {
  static constexpr int TPL = LinePlan<N>::tpl;
  static constexpr int SLOT_THREADS = NB * TPL;
  static constexpr int SLOTS = fused_slots(TPL);
  static constexpr int NT = SLOT_THREADS * SLOTS;
  static constexpr bool WARP_SLOTS = SLOT_THREADS % 32 == 0;  // slot = whole warps -> named barriers
  static constexpr int NG = (N + NB - 1) / NB;         // x-stage line groups (8 y lines each)
  static constexpr int NR = (NG + SLOTS - 1) / SLOTS;  // x-stage rounds
  static constexpr int SX = N + ((N % 8 == 1) ? 0 : (9 - N % 8) % 8);  // Y row stride == 1 (mod 8)
  static constexpr int EXCH = SLOTS * N * NB;          // exchange buffers (complex)
  // psi(r) cache entry of one band-plane (complex): x-stage registers of every thread
  static constexpr int PSI_PLANE = NR * LineFFT<N, +1>::CB * LineFFT<N, +1>::RB * NT;
  // shared memory (complex numbers): 2 Y slabs, exchange buffers, 2 staging buffers
  static JRB_HD int ybuf_elems(int nxo) { return nxo * SX; }
  // staging buffers hold whole TMA boxes of 256 columns (the tail box is filled past ncol)
  static JRB_HD int stage_elems(int ncol) { return (ncol + 255) / 256 * 256; }
  static JRB_HD int smem_bytes(int nxo, int ncol) {
    return (2 * ybuf_elems(nxo) + EXCH + 2 * stage_elems(ncol)) * (int)sizeof(cplx);
  }
  // k_x_vmul_cached: one Y slab, the exchange buffers and v_eff of the current plane (doubles,
  // one private slot per x-stage register of every thread); no staging buffers
  static JRB_HD int smem_bytes_cached(int nxo) {
    return (ybuf_elems(nxo) + EXCH) * (int)sizeof(cplx) + PSI_PLANE * (int)sizeof(double);
  }
};

// streaming (evict-first) accesses of the psi(r) cache: written once, read once, far larger than
// L2, so it should not displace the column work space and the tables.  Two complex numbers per
// 256-bit access (LDG.256 / STG.256 of sm_100): half the instructions on the LSU queue.
// Layout of round r of a band-plane entry, E = CB * RB registers per thread, NT threads:
//   pair j (registers 2j, 2j+1) of thread t at (r E + 2 j) NT + 2 t, an odd E's last register
//   at (r E + E - 1) NT + t
__device__ __forceinline__ void psi_store2(cplx* p, const cplx& a, const cplx& b) {
  asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "d"(a.x), "d"(a.y), "d"(b.x),
               "d"(b.y)
               : "memory");
}
__device__ __forceinline__ void psi_load2(const cplx* p, cplx& a, cplx& b) {
  asm volatile("ld.global.cs.v4.f64 {%0, %1, %2, %3}, [%4];\n"
               : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y)
               : "l"(p));
}
__device__ __forceinline__ void psi_store(cplx* p, const cplx& v) {
  __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y));
}
__device__ __forceinline__ cplx psi_load(const cplx* p) {
  const double2 t = __ldcs(reinterpret_cast<const double2*>(p));
  return cmake(t.x, t.y);
}
// registers v[CB][RB] of round r <-> the entry at `base` (thread offset NOT included)
template <int CB, int RB, int NT>
__device__ __forceinline__ void psi_store_round(cplx* base, int r, int t, const cplx (&v)[CB][RB]) {
  constexpr int E = CB * RB;
  cplx* p = base + (long long)r * E * NT;
#pragma unroll
  for (int j = 0; j < E / 2; ++j)
    psi_store2(p + 2 * j * NT + 2 * t, v[(2 * j) / RB][(2 * j) % RB], v[(2 * j + 1) / RB][(2 * j + 1) % RB]);
  if constexpr (E % 2 == 1) psi_store(p + (E - 1) * NT + t, v[CB - 1][RB - 1]);
}
template <int CB, int RB, int NT>
__device__ __forceinline__ void psi_load_round(const cplx* base, int r, int t, cplx (&v)[CB][RB]) {
  constexpr int E = CB * RB;
  const cplx* p = base + (long long)r * E * NT;
#pragma unroll
  for (int j = 0; j < E / 2; ++j)
    psi_load2(p + 2 * j * NT + 2 * t, v[(2 * j) / RB][(2 * j) % RB], v[(2 * j + 1) / RB][(2 * j + 1) % RB]);
  if constexpr (E % 2 == 1) v[CB - 1][RB - 1] = psi_load(p + (E - 1) * NT + t);
}
__device__ __forceinline__ long long psi_entry(const FusedArgs& a, int gl, int z, int band) {
  return ((long long)(a.g0 + gl) * a.m.nz + z) * NB + band;
}

template <int N>
__device__ __forceinline__ void slot_barrier(int slot) {
  if constexpr (FCfg<N>::WARP_SLOTS && FCfg<N>::SLOTS > 1 && FCfg<N>::SLOTS < 15) {
    asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "n"(FCfg<N>::SLOT_THREADS) : "memory");
  } else {
    __syncthreads();
  }
}

__device__ __forceinline__ void fused_cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void fused_cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void fused_cp_wait_all() {
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

// Position in the flat (item, band) list of a CTA: item w = z * ngroups + gl.  Stepping is
// incremental (no divisions in the band loop).
struct BandPos {
  int w, z, gl, gmod, band;  // gmod = (g0 + gl) % ngpk: group index within its (spin, k)
};
__device__ __forceinline__ int band_count(const FusedArgs& a, const BandPos& p) {
  return min(NB, a.nb - p.gmod * NB);
}
__device__ __forceinline__ BandPos band_first(const FusedArgs& a, int w) {
  BandPos p;
  p.w = w;
  p.z = w / a.ngroups;
  p.gl = w - p.z * a.ngroups;
  p.gmod = (a.g0 + p.gl) % a.ngpk;
  p.band = 0;
  return p;
}
__device__ __forceinline__ BandPos band_next(const FusedArgs& a, BandPos p, int w_end, int gmod0) {
  if (p.w < w_end && ++p.band >= band_count(a, p)) {
    p.band = 0;
    ++p.w;
    ++p.gl;
    if (++p.gmod == a.ngpk) p.gmod = 0;
    if (p.gl == a.ngroups) {
      p.gl = 0;
      ++p.z;
      p.gmod = gmod0;
    }
  }
  return p;
}
__device__ __forceinline__ cplx* band_plane(const FusedArgs& a, const BandPos& p) {
  return a.wa + (long long)(p.gl * a.m.nz + p.z) * (a.m.ncol * NB) + p.band;
}

// ---- mbarrier + TMA (cp.async.bulk.tensor) helpers ---------------------------------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(b), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(b) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "WAIT_%=:\n"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
    "@p bra DONE_%=;\n"
    "bra WAIT_%=;\n"
    "DONE_%=:\n"
    "}\n" ::"r"(b),
    "r"(parity)
    : "memory");
}
// box {2 doubles, 256 rows} at (c0, c1) of the 2-D tensor -> 256 consecutive complex numbers
__device__ __forceinline__ void tma_load_box(void* smem_dst, const void* tmap, int c0, int c1,
                                             unsigned long long* bar) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
    "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, "
    "{%2, %3}], [%4];\n" ::"r"(d),
    "l"(tmap), "r"(c0), "r"(c1), "r"(b)
    : "memory");
}

// stage the plane columns of one band: stage[col] = A[gl][z][col][band].  TMA: thread 0 arms the
// buffer's mbarrier with the byte count and issues one box per 256 columns (nothing to stage:
// a plain arrive, so that every stage is matched by exactly one completed phase).
template <int NT>
__device__ __forceinline__ void fused_stage(const FusedArgs& a, const BandPos& p, int w_end,
                                            cplx* stage, unsigned long long* bar) {
  if (a.tmap) {
    if (threadIdx.x == 0) {
      if (p.w < w_end) {
        const int nbox = (a.m.ncol + 255) >> 8;
        mbar_arrive_expect_tx(bar, (unsigned)(nbox * 256 * sizeof(cplx)));
        const int row = (int)(a.row0 + (long long)(p.gl * a.m.nz + p.z) * a.m.ncol);
        for (int b = 0; b < nbox; ++b) tma_load_box(stage + 256 * b, a.tmap, 2 * p.band, row + 256 * b, bar);
      } else {
        mbar_arrive(bar);
      }
    }
    return;
  }
  if (p.w < w_end) {
    const cplx* src = band_plane(a, p);
    for (int c = threadIdx.x; c < a.m.ncol; c += NT)
      fused_cp_async16(stage + c, src + (long long)c * NB);
  }
  fused_cp_commit();
}
// the staged band of `bar`'s buffer has landed (TMA: phase `parity` of its mbarrier)
__device__ __forceinline__ void fused_stage_wait(const FusedArgs& a, unsigned long long* bar,
                                                 unsigned parity) {
  if (a.tmap) mbar_wait(bar, parity);
  else fused_cp_wait_all();
}

// y stage, inverse: staged columns of one band-plane -> Y[xo][y].  All threads call it.
// pk: per-thread packed indices, low 16 bits = column + 1 of (x plane slot * 8 + lane, y =
// idxA) for the first y-stage iteration, high 16 bits = Y row offset + 1 of x = idxA (x stage).
template <int N, bool ONE_ITER, bool SP>
__device__ __forceinline__ void fused_y_inverse(
  const FusedArgs& a, const cplx* stage, cplx* ybuf, cplx* ex,
  const cplx (&tw)[LineFFT<N, +1>::CB][LineFFT<N, +1>::NTW],
  const unsigned (&pk)[LineFFT<N, +1>::CA][LineFFT<N, +1>::RA], int lane, int tj, int slot) {
  using F = LineFFT<N, +1>;
  using C = FCfg<N>;
  const int ngx = (a.m.nxo + NB - 1) / NB;
  const int grp_end = ONE_ITER ? slot + 1 : (ngx + C::SLOTS - 1) / C::SLOTS * C::SLOTS;
  for (int grp = slot; grp < grp_end; grp += C::SLOTS) {
    const int xo = grp * NB + lane;
    const bool ok = grp < ngx && xo < a.m.nxo;
    cplx va[F::CA][F::RA];
    if (ONE_ITER || grp == slot) {
#pragma unroll
      for (int i = 0; i < F::CA; ++i)
#pragma unroll
        for (int m = 0; m < F::RA; ++m) {
          if (SP && !JRB_SPARSE_M(m)) continue;  // structurally zero, never read
          const int col = (int)(pk[i][m] & 0xffffu) - 1;
          va[i][m] = col >= 0 ? stage[col] : czero();
        }
    } else {
      const int32_t* yc = a.m.ycol + (long long)(ok ? xo : 0) * N;
#pragma unroll
      for (int i = 0; i < F::CA; ++i) {
#pragma unroll
        for (int m = 0; m < F::RA; ++m) {
          cplx v = czero();
          if (F::activeA(i, tj) && ok) {
            const int col = yc[F::idxA(i, m, tj)];
            if (col >= 0) v = stage[col];
          }
          va[i][m] = v;
        }
      }
    }
    if constexpr (SP) {
      F::template stageA_store_sparse<NB>(va, ex, tj);
    } else {
      F::template stageA_store<NB>(va, ex, tj);
    }
    slot_barrier<N>(slot);
    cplx vb[F::CB][F::RB];
    F::template stageB_load<NB>(vb, ex, tw, tj);
    if (ok) {
      cplx* out = ybuf + (long long)xo * C::SX;
#pragma unroll
      for (int i = 0; i < F::CB; ++i) {
        if (F::activeB(i, tj)) {
#pragma unroll
          for (int m = 0; m < F::RB; ++m) out[F::idxB(i, m, tj)] = vb[i][m];
        }
      }
    }
    slot_barrier<N>(slot);
  }
}

// ---------------------------------------------------------------------------------------
// grid: (G) persistent CTAs; dynamic smem: FCfg<N>::smem_bytes(nxo, ncol)
template <int N, bool ONE_ITER, bool SP>
__global__ void __launch_bounds__(FCfg<N>::NT, (FCfg<N>::NT <= 256 ? 2 : 1))
k_yx_density(FusedArgs a) {
  using F = LineFFT<N, +1>;
  using C = FCfg<N>;
  extern __shared__ __align__(128) unsigned char smem_fused_[];
  __shared__ __align__(8) unsigned long long sbar[2];
  // staging buffers first: TMA destinations (whole 4 KB boxes from a 128-byte aligned base)
  cplx* stage0 = reinterpret_cast<cplx*>(smem_fused_);
  const int ssz = C::stage_elems(a.m.ncol);
  cplx* ybuf0 = stage0 + 2 * ssz;
  const int ysz = C::ybuf_elems(a.m.nxo);
  cplx* exbase = ybuf0 + 2 * ysz;
  const int t = threadIdx.x;
  const int lane = t % NB;
  const int tj = (t / NB) % C::TPL;
  const int slot = t / C::SLOT_THREADS;
  cplx* ex = exbase + (size_t)slot * N * NB + lane;
  if (a.tmap && t == 0) {
    mbar_init(&sbar[0], 1);
    mbar_init(&sbar[1], 1);
    mbar_fence_init();
  }
  unsigned sph0 = 0, sph1 = 0;  // phase of each staging buffer's mbarrier
  cplx tw[F::CB][F::NTW];
  F::load_twiddles(tw, a.tw, tj);

  // packed per-thread indices (see fused_y_inverse): x-stage Y rows and y-stage columns
  unsigned pk[F::CA][F::RA];
  {
    const int xo0 = slot * NB + lane;
#pragma unroll
    for (int i = 0; i < F::CA; ++i)
#pragma unroll
      for (int m = 0; m < F::RA; ++m) {
        unsigned lo = 0, hi = 0;
        if (F::activeA(i, tj)) {
          const int xo = a.m.xmap[F::idxA(i, m, tj)];
          if (xo >= 0) hi = (unsigned)(xo * C::SX + 1);
          if (xo0 < a.m.nxo) lo = (unsigned)(a.m.ycol[(long long)xo0 * N + F::idxA(i, m, tj)] + 1);
        }
        pk[i][m] = lo | (hi << 16);
      }
  }
  double acc[C::NR][F::CB][F::RB];
#pragma unroll
  for (int r = 0; r < C::NR; ++r)
#pragma unroll
    for (int i = 0; i < F::CB; ++i)
#pragma unroll
      for (int m = 0; m < F::RB; ++m) acc[r][i][m] = 0.0;

  const long long W = (long long)a.m.nz * a.ngroups;
  const int c = blockIdx.x, G = gridDim.x;
  const int w_end = (int)((c + 1) * W / G);
  const int gmod0 = a.g0 % a.ngpk;
  BandPos cur = band_first(a, (int)(c * W / G));
  int seg = 0;
  auto flush = [&](int z) {
    double* out = a.rho_part + ((long long)c * a.segmax + seg) * N * N;
#pragma unroll
    for (int r = 0; r < C::NR; ++r) {
      const int y = (r * C::SLOTS + slot) * NB + lane;
#pragma unroll
      for (int i = 0; i < F::CB; ++i)
#pragma unroll
        for (int m = 0; m < F::RB; ++m) {
          if (F::activeB(i, tj) && y < N) out[(long long)F::idxB(i, m, tj) * N + y] = acc[r][i][m];
          acc[r][i][m] = 0.0;
        }
    }
    if (t == 0) a.seg_z[c * a.segmax + seg] = z;
    ++seg;
  };

  __syncthreads();  // mbarriers initialised
  if (cur.w < w_end) {
    // prologue: stage band 0, run its y stage, stage band 1
    fused_stage<C::NT>(a, cur, w_end, stage0, &sbar[0]);
    fused_stage_wait(a, &sbar[0], sph0);
    sph0 ^= 1;
    __syncthreads();
    fused_stage<C::NT>(a, band_next(a, cur, w_end, gmod0), w_end, stage0 + ssz, &sbar[1]);
    fused_y_inverse<N, ONE_ITER, SP>(a, stage0, ybuf0, ex, tw, pk, lane, tj, slot);
    int par = 0;  // parity of the current band: it lives in Y[par]
    int cur_z = cur.z;
    double fw_next = a.focc[(a.g0 + cur.gl) * NB + cur.band];
    while (cur.w < w_end) {
      // Y[par] complete (y stage of the current band by every slot); staged data of the next
      // band landed and visible; everybody is done with Y[par ^ 1] and stage[par]
      if (par) { fused_stage_wait(a, &sbar[0], sph0); sph0 ^= 1; }
      else { fused_stage_wait(a, &sbar[1], sph1); sph1 ^= 1; }
      __syncthreads();
      // band b+1 is staged in stage[par ^ 1]; prefetch band b+2 into stage[par]
      const BandPos nxt = band_next(a, cur, w_end, gmod0);
      fused_stage<C::NT>(a, band_next(a, nxt, w_end, gmod0), w_end, stage0 + par * ssz, &sbar[par]);
      if (cur.z != cur_z) {
        flush(cur_z);
        cur_z = cur.z;
      }
      const double fw = fw_next;
      if (nxt.w < w_end) fw_next = a.focc[(a.g0 + nxt.gl) * NB + nxt.band];
      const cplx* ybuf = ybuf0 + par * ysz;
#pragma unroll
      for (int r = 0; r < C::NR; ++r) {
        const int y = (r * C::SLOTS + slot) * NB + lane;
        const bool ok = y < N;
        cplx va[F::CA][F::RA];
#pragma unroll
        for (int i = 0; i < F::CA; ++i)
#pragma unroll
          for (int m = 0; m < F::RA; ++m) {
            if (SP && !JRB_SPARSE_M(m)) continue;
            va[i][m] = ((pk[i][m] >> 16) != 0 && ok) ? ybuf[(int)(pk[i][m] >> 16) - 1 + y] : czero();
          }
        if constexpr (SP) {
          F::template stageA_store_sparse<NB>(va, ex, tj);
        } else {
          F::template stageA_store<NB>(va, ex, tj);
        }
        slot_barrier<N>(slot);
        cplx vb[F::CB][F::RB];
        F::template stageB_load<NB>(vb, ex, tw, tj);
        if (a.psi && ok) {
          // entry address recomputed per round: a pointer kept across the round costs registers
          cplx* pc = a.psi + psi_entry(a, cur.gl, cur.z, cur.band) * C::PSI_PLANE;
          if constexpr (F::RA % F::TPL != 0) {  // idle butterfly slots: defined values
#pragma unroll
            for (int i = 0; i < F::CB; ++i)
              if (!F::activeB(i, tj))
#pragma unroll
                for (int m = 0; m < F::RB; ++m) vb[i][m] = czero();
          }
          psi_store_round<F::CB, F::RB, C::NT>(pc, r, t, vb);
        }
#pragma unroll
        for (int i = 0; i < F::CB; ++i)
#pragma unroll
          for (int m = 0; m < F::RB; ++m)
            if (F::activeB(i, tj))
              acc[r][i][m] += fw * (vb[i][m].x * vb[i][m].x + vb[i][m].y * vb[i][m].y);
        slot_barrier<N>(slot);
      }
      // y stage of the next band into the other Y buffer
      if (nxt.w < w_end)
        fused_y_inverse<N, ONE_ITER, SP>(a, stage0 + (par ^ 1) * ssz, ybuf0 + (par ^ 1) * ysz, ex, tw, pk, lane,
                           tj, slot);
      cur = nxt;
      par ^= 1;
    }
    if (!a.tmap) fused_cp_wait_all();
    flush(cur_z);
  }
  if (t == 0)
    for (int s = seg; s < a.segmax; ++s) a.seg_z[c * a.segmax + s] = -1;
}

// rho[x][y][z] += sum of the partial planes tagged z, in slot order (deterministic)
// grid: (ceil(nx*ny / 32), nz), block 32 x 8: thread (tx, ty) sums slots ty, ty+8, ...
static __global__ void __launch_bounds__(256)
k_rho_reduce(const double* __restrict__ part, const int* __restrict__ seg_z, int nctas, int segmax,
             int ngroups, int nxy, int nz, double* __restrict__ rho) {
  __shared__ double sh[8][33];
  const int xy = blockIdx.x * 32 + threadIdx.x;
  const int z = blockIdx.y;
  // CTA c covers items [c W / G, (c + 1) W / G), W = nz ngroups: candidates for plane z
  const long long W = (long long)nz * ngroups;
  const int c_lo = max(0, (int)(((long long)z * ngroups * nctas) / W) - 1);
  const int c_hi = min(nctas - 1, (int)((((long long)z + 1) * ngroups * nctas) / W) + 1);
  double s = 0.0;
  if (xy < nxy)
    for (int k = c_lo * segmax + threadIdx.y; k < (c_hi + 1) * segmax; k += 8)
      if (seg_z[k] == z) s += part[(long long)k * nxy + xy];
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && xy < nxy) {
    double r = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) r += sh[j][threadIdx.x];
    rho[(long long)xy * nz + z] += r;
  }
}

// ---------------------------------------------------------------------------------------
// Hamiltonian-apply middle: A (plane z) -> y inverse -> x inverse -> * v_eff / N -> x forward
// -> y forward -> A (in place).   grid: (G) persistent CTAs; dynamic smem as above.
template <int N, bool ONE_ITER, bool SP>
__global__ void __launch_bounds__(FCfg<N>::NT, (FCfg<N>::NT <= 256 ? 2 : 1))
k_yx_vmul(FusedArgs a) {
  using FI = LineFFT<N, +1>;
  using FF = LineFFT<N, -1>;
  using C = FCfg<N>;
  static_assert(FI::CB == FF::CA && FI::RB == FF::RA, "register chaining contract");
  extern __shared__ __align__(128) unsigned char smem_fused_[];
  __shared__ __align__(8) unsigned long long sbar[2];
  cplx* stage0 = reinterpret_cast<cplx*>(smem_fused_);
  const int ssz = C::stage_elems(a.m.ncol);
  cplx* ybuf0 = stage0 + 2 * ssz;
  const int ysz = C::ybuf_elems(a.m.nxo);
  cplx* exbase = ybuf0 + 2 * ysz;
  const int t = threadIdx.x;
  const int lane = t % NB;
  const int tj = (t / NB) % C::TPL;
  const int slot = t / C::SLOT_THREADS;
  cplx* ex = exbase + (size_t)slot * N * NB + lane;
  if (a.tmap && t == 0) {
    mbar_init(&sbar[0], 1);
    mbar_init(&sbar[1], 1);
    mbar_fence_init();
  }
  unsigned sph0 = 0, sph1 = 0;
  const long long nyz = (long long)N * a.m.nz;
  // equal radices: the forward twiddles are the conjugates of the inverse ones -> one register set
  constexpr bool SHARE_TW = LinePlan<N>::r1 == LinePlan<N>::r2;
  cplx twi[FI::CB][FI::NTW];
  cplx twf[SHARE_TW ? 1 : FF::CB][SHARE_TW ? 1 : FF::NTW];
  FI::load_twiddles(twi, a.tw, tj);
  if constexpr (!SHARE_TW) FF::load_twiddles(twf, a.tw, tj);
  // forward stage B; band-limited data: only the outputs m in {0, 1, 6, 7} are produced
  auto fwd_stageB = [&](cplx (&v)[FF::CB][FF::RB]) {
    if constexpr (SP) {
      static_assert(!SP || SHARE_TW, "band-limited variant needs equal radices");
      FF::template stageB_load_sparse<NB, true>(v, ex, twi, tj);
    } else if constexpr (SHARE_TW) {
      FF::template stageB_load<NB, true>(v, ex, twi, tj);
    } else {
      FF::template stageB_load<NB, false>(v, ex, twf, tj);
    }
  };

  // packed per-thread indices (see fused_y_inverse): x-stage Y rows and y-stage columns
  unsigned pk[FI::CA][FI::RA];
  {
    const int xo0 = slot * NB + lane;
#pragma unroll
    for (int i = 0; i < FI::CA; ++i)
#pragma unroll
      for (int m = 0; m < FI::RA; ++m) {
        unsigned lo = 0, hi = 0;
        if (FI::activeA(i, tj)) {
          const int xo = a.m.xmap[FI::idxA(i, m, tj)];
          if (xo >= 0) hi = (unsigned)(xo * C::SX + 1);
          if (xo0 < a.m.nxo) lo = (unsigned)(a.m.ycol[(long long)xo0 * N + FI::idxA(i, m, tj)] + 1);
        }
        pk[i][m] = lo | (hi << 16);
      }
  }
  // forward stage-B outputs land on the same x indices as the inverse stage-A inputs
  // (idxB of the forward plan == idxA of the inverse plan), so pk's Y-row field serves both
  static_assert(FF::CB == FI::CA && FF::RB == FI::RA, "x index sets of inverse-in / forward-out");
  // v_eff(x, y, z) / N at the points this thread owns in the x stage (reloaded when z changes)
  double vv[C::NR][FI::CB][FI::RB];
  auto load_v = [&](int z) {
#pragma unroll
    for (int r = 0; r < C::NR; ++r) {
      const int y = (r * C::SLOTS + slot) * NB + lane;
#pragma unroll
      for (int i = 0; i < FI::CB; ++i)
#pragma unroll
        for (int m = 0; m < FI::RB; ++m) {
          double v = 0.0;
          if (FI::activeB(i, tj) && y < N)
            v = a.veff[(long long)FI::idxB(i, m, tj) * nyz + (long long)y * a.m.nz + z] * a.vscale;
          vv[r][i][m] = v;
        }
    }
  };

  const int ngx = (a.m.nxo + NB - 1) / NB;
  const long long W = (long long)a.m.nz * a.ngroups;
  const int c = blockIdx.x, G = gridDim.x;
  const int w_end = (int)((c + 1) * W / G);
  const int gmod0 = a.g0 % a.ngpk;
  BandPos cur = band_first(a, (int)(c * W / G));
  __syncthreads();  // mbarriers initialised
  if (cur.w >= w_end) return;
  fused_stage<C::NT>(a, cur, w_end, stage0, &sbar[0]);
  fused_stage_wait(a, &sbar[0], sph0);
  sph0 ^= 1;
  __syncthreads();
  fused_stage<C::NT>(a, band_next(a, cur, w_end, gmod0), w_end, stage0 + ssz, &sbar[1]);
  fused_y_inverse<N, ONE_ITER, SP>(a, stage0, ybuf0, ex, twi, pk, lane, tj, slot);
  int par = 0;
  int cur_z = cur.z;
  load_v(cur_z);
  while (cur.w < w_end) {
    if (par) { fused_stage_wait(a, &sbar[0], sph0); sph0 ^= 1; }
    else { fused_stage_wait(a, &sbar[1], sph1); sph1 ^= 1; }
    __syncthreads();
    const BandPos nxt = band_next(a, cur, w_end, gmod0);
    fused_stage<C::NT>(a, band_next(a, nxt, w_end, gmod0), w_end, stage0 + par * ssz, &sbar[par]);
    if (cur.z != cur_z) {
      cur_z = cur.z;
      load_v(cur_z);
    }
    cplx* ybuf = ybuf0 + par * ysz;
    // x stage: inverse, multiply, forward; occupied x planes written back into Y
#pragma unroll
    for (int r = 0; r < C::NR; ++r) {
      const int y = (r * C::SLOTS + slot) * NB + lane;
      const bool ok = y < N;
      cplx va[FI::CA][FI::RA];
#pragma unroll
      for (int i = 0; i < FI::CA; ++i)
#pragma unroll
        for (int m = 0; m < FI::RA; ++m) {
          if (SP && !JRB_SPARSE_M(m)) continue;
          va[i][m] = ((pk[i][m] >> 16) != 0 && ok) ? ybuf[(int)(pk[i][m] >> 16) - 1 + y] : czero();
        }
      if constexpr (SP) {
        FI::template stageA_store_sparse<NB>(va, ex, tj);
      } else {
        FI::template stageA_store<NB>(va, ex, tj);
      }
      slot_barrier<N>(slot);
      cplx vb[FI::CB][FI::RB];
      FI::template stageB_load<NB>(vb, ex, twi, tj);
#pragma unroll
      for (int i = 0; i < FI::CB; ++i)
#pragma unroll
        for (int m = 0; m < FI::RB; ++m) vb[i][m] = cscale(vb[i][m], vv[r][i][m]);
      slot_barrier<N>(slot);
      FF::template stageA_store<NB>(vb, ex, tj);
      slot_barrier<N>(slot);
      cplx vc[FF::CB][FF::RB];
      fwd_stageB(vc);
      if (ok) {
#pragma unroll
        for (int i = 0; i < FF::CB; ++i)
#pragma unroll
          for (int m = 0; m < FF::RB; ++m) {
            if (SP && !JRB_SPARSE_M(m)) continue;
            if ((pk[i][m] >> 16) != 0) ybuf[(int)(pk[i][m] >> 16) - 1 + y] = vc[i][m];
          }
      }
      slot_barrier<N>(slot);
    }
    __syncthreads();
    // y stage, forward: Y -> occupied columns of A (global, in place)
    cplx* dst = band_plane(a, cur);
    const int grp_end = ONE_ITER ? slot + 1 : (ngx + C::SLOTS - 1) / C::SLOTS * C::SLOTS;
    for (int grp = slot; grp < grp_end; grp += C::SLOTS) {
      const int xo = grp * NB + lane;
      const bool ok = grp < ngx && xo < a.m.nxo;
      const cplx* in = ybuf + (long long)(ok ? xo : 0) * C::SX;
      cplx va[FF::CA][FF::RA];
#pragma unroll
      for (int i = 0; i < FF::CA; ++i)
#pragma unroll
        for (int m = 0; m < FF::RA; ++m)
          va[i][m] = (FF::activeA(i, tj) && ok) ? in[FF::idxA(i, m, tj)] : czero();
      FF::template stageA_store<NB>(va, ex, tj);
      slot_barrier<N>(slot);
      cplx vb[FF::CB][FF::RB];
      fwd_stageB(vb);
      if (ok) {
        // forward stage-B outputs sit on the y indices of the inverse stage-A inputs, so in the
        // single-pass case the packed column indices of pk apply (no map lookups)
        const int32_t* yc = a.m.ycol + (long long)xo * N;
#pragma unroll
        for (int i = 0; i < FF::CB; ++i) {
          if (FF::activeB(i, tj)) {
#pragma unroll
            for (int m = 0; m < FF::RB; ++m) {
              if (SP && !JRB_SPARSE_M(m)) continue;
              const int col = ONE_ITER ? (int)(pk[i][m] & 0xffffu) - 1 : yc[FF::idxB(i, m, tj)];
              if (col >= 0) dst[(long long)col * NB] = vb[i][m];
            }
          }
        }
      }
      slot_barrier<N>(slot);
    }
    if (nxt.w < w_end)
      fused_y_inverse<N, ONE_ITER, SP>(a, stage0 + (par ^ 1) * ssz, ybuf0 + (par ^ 1) * ysz, ex, twi, pk, lane,
                         tj, slot);
    cur = nxt;
    par ^= 1;
  }
  if (!a.tmap) fused_cp_wait_all();
}

// ---------------------------------------------------------------------------------------
// Hamiltonian-apply middle from the psi(r) cache: psi(r) of the density sweep (global memory,
// streamed once) -> * v_eff / N -> x forward -> y forward -> A.  Half the line transforms of
// k_yx_vmul (the FP64 / shared-memory pipes bind that kernel, HBM idles at ~10 %), paid for
// with one streaming read of the cache.  The loads of the next x-stage round are issued right
// after the current round's registers are consumed, so their latency hides behind the round's
// two exchanges; v_eff of the plane sits in shared memory (thread-private slots, conflict free)
// to leave the registers to that prefetch.
// grid: (G) persistent CTAs; dynamic smem: FCfg<N>::smem_bytes_cached(nxo)
template <int N, bool ONE_ITER, bool SP>
__global__ void __launch_bounds__(FCfg<N>::NT, (FCfg<N>::NT <= 256 ? 2 : 1))
k_x_vmul_cached(FusedArgs a) {
  using FI = LineFFT<N, +1>;
  using FF = LineFFT<N, -1>;
  using C = FCfg<N>;
  static_assert(FI::CB == FF::CA && FI::RB == FF::RA, "register chaining contract");
  static_assert(FF::CB == FI::CA && FF::RB == FI::RA, "x index sets of inverse-in / forward-out");
  extern __shared__ __align__(128) unsigned char smem_fused_[];
  cplx* ybuf = reinterpret_cast<cplx*>(smem_fused_);
  cplx* exbase = ybuf + C::ybuf_elems(a.m.nxo);
  const int t = threadIdx.x;
  double* vs = reinterpret_cast<double*>(exbase + C::EXCH) + t;  // [round][butterfly][element][thread]
  const int lane = t % NB;
  const int tj = (t / NB) % C::TPL;
  const int slot = t / C::SLOT_THREADS;
  cplx* ex = exbase + (size_t)slot * N * NB + lane;
  constexpr bool SHARE_TW = LinePlan<N>::r1 == LinePlan<N>::r2;
  // forward twiddles; equal radices: held as the inverse plan's set and conjugated on use, like
  // k_yx_vmul, so the band-limited butterflies are shared
  cplx twi[SHARE_TW ? FI::CB : 1][SHARE_TW ? FI::NTW : 1];
  cplx twf[SHARE_TW ? 1 : FF::CB][SHARE_TW ? 1 : FF::NTW];
  if constexpr (SHARE_TW) FI::load_twiddles(twi, a.tw, tj);
  else FF::load_twiddles(twf, a.tw, tj);
  auto fwd_stageB = [&](cplx (&v)[FF::CB][FF::RB]) {
    if constexpr (SP) {
      static_assert(!SP || SHARE_TW, "band-limited variant needs equal radices");
      FF::template stageB_load_sparse<NB, true>(v, ex, twi, tj);
    } else if constexpr (SHARE_TW) {
      FF::template stageB_load<NB, true>(v, ex, twi, tj);
    } else {
      FF::template stageB_load<NB, false>(v, ex, twf, tj);
    }
  };
  // packed per-thread indices (see fused_y_inverse): Y rows of the x-stage outputs and the
  // columns of the single-pass y stage
  unsigned pk[FI::CA][FI::RA];
  {
    const int xo0 = slot * NB + lane;
#pragma unroll
    for (int i = 0; i < FI::CA; ++i)
#pragma unroll
      for (int m = 0; m < FI::RA; ++m) {
        unsigned lo = 0, hi = 0;
        if (FI::activeA(i, tj)) {
          const int xo = a.m.xmap[FI::idxA(i, m, tj)];
          if (xo >= 0) hi = (unsigned)(xo * C::SX + 1);
          if (xo0 < a.m.nxo) lo = (unsigned)(a.m.ycol[(long long)xo0 * N + FI::idxA(i, m, tj)] + 1);
        }
        pk[i][m] = lo | (hi << 16);
      }
  }
  // v_eff(x, y, z) / N at the points this thread owns in the x stage -> its shared-memory slots
  auto load_v = [&](int z) {
    const long long nyz = (long long)N * a.m.nz;
#pragma unroll
    for (int r = 0; r < C::NR; ++r) {
      const int y = (r * C::SLOTS + slot) * NB + lane;
#pragma unroll
      for (int i = 0; i < FI::CB; ++i)
#pragma unroll
        for (int m = 0; m < FI::RB; ++m) {
          double v = 0.0;
          if (FI::activeB(i, tj) && y < N)
            v = a.veff[(long long)FI::idxB(i, m, tj) * nyz + (long long)y * a.m.nz + z] * a.vscale;
          vs[((r * FI::CB + i) * FI::RB + m) * C::NT] = v;
        }
    }
  };
  // prefetch registers: psi(r) of the next x-stage round
  cplx pf[FI::CB][FI::RB];
  auto issue = [&](const BandPos& p, int r) {
    const int y = (r * C::SLOTS + slot) * NB + lane;
    const cplx* src = a.psi + psi_entry(a, p.gl, p.z, p.band) * C::PSI_PLANE;
    if (y < N) {
      psi_load_round<FI::CB, FI::RB, C::NT>(src, r, t, pf);
    } else {
#pragma unroll
      for (int i = 0; i < FI::CB; ++i)
#pragma unroll
        for (int m = 0; m < FI::RB; ++m) pf[i][m] = czero();
    }
  };

  const int ngx = (a.m.nxo + NB - 1) / NB;
  const long long W = (long long)a.m.nz * a.ngroups;
  const int c = blockIdx.x, G = gridDim.x;
  const int w_end = (int)((c + 1) * W / G);
  const int gmod0 = a.g0 % a.ngpk;
  BandPos cur = band_first(a, (int)(c * W / G));
  if (cur.w >= w_end) return;
  issue(cur, 0);
  int cur_z = cur.z;
  load_v(cur_z);  // thread-private slots: no barrier needed
  while (cur.w < w_end) {
    const BandPos nxt = band_next(a, cur, w_end, gmod0);
    if (cur.z != cur_z) {
      cur_z = cur.z;
      load_v(cur_z);
    }
#pragma unroll
    for (int r = 0; r < C::NR; ++r) {
      const int y = (r * C::SLOTS + slot) * NB + lane;
      const bool ok = y < N;
      cplx vb[FI::CB][FI::RB];
#pragma unroll
      for (int i = 0; i < FI::CB; ++i)
#pragma unroll
        for (int m = 0; m < FI::RB; ++m)
          vb[i][m] = cscale(pf[i][m], vs[((r * FI::CB + i) * FI::RB + m) * C::NT]);
      if (r + 1 < C::NR) issue(cur, r + 1);
      else if (nxt.w < w_end) issue(nxt, 0);
      FF::template stageA_store<NB>(vb, ex, tj);
      slot_barrier<N>(slot);
      cplx vc[FF::CB][FF::RB];
      fwd_stageB(vc);
      if (ok) {
#pragma unroll
        for (int i = 0; i < FF::CB; ++i)
#pragma unroll
          for (int m = 0; m < FF::RB; ++m) {
            if (SP && !JRB_SPARSE_M(m)) continue;
            if ((pk[i][m] >> 16) != 0) ybuf[(int)(pk[i][m] >> 16) - 1 + y] = vc[i][m];
          }
      }
      slot_barrier<N>(slot);
    }
    __syncthreads();  // Y complete
    // y stage, forward: Y -> occupied columns of A (global)
    cplx* dst = band_plane(a, cur);
    const int grp_end = ONE_ITER ? slot + 1 : (ngx + C::SLOTS - 1) / C::SLOTS * C::SLOTS;
    for (int grp = slot; grp < grp_end; grp += C::SLOTS) {
      const int xo = grp * NB + lane;
      const bool ok = grp < ngx && xo < a.m.nxo;
      const cplx* in = ybuf + (long long)(ok ? xo : 0) * C::SX;
      cplx va[FF::CA][FF::RA];
#pragma unroll
      for (int i = 0; i < FF::CA; ++i)
#pragma unroll
        for (int m = 0; m < FF::RA; ++m)
          va[i][m] = (FF::activeA(i, tj) && ok) ? in[FF::idxA(i, m, tj)] : czero();
      FF::template stageA_store<NB>(va, ex, tj);
      slot_barrier<N>(slot);
      cplx vb[FF::CB][FF::RB];
      fwd_stageB(vb);
      if (ok) {
        const int32_t* yc = a.m.ycol + (long long)xo * N;
#pragma unroll
        for (int i = 0; i < FF::CB; ++i) {
          if (FF::activeB(i, tj)) {
#pragma unroll
            for (int m = 0; m < FF::RB; ++m) {
              if (SP && !JRB_SPARSE_M(m)) continue;
              const int col = ONE_ITER ? (int)(pk[i][m] & 0xffffu) - 1 : yc[FF::idxB(i, m, tj)];
              if (col >= 0) dst[(long long)col * NB] = vb[i][m];
            }
          }
        }
      }
      slot_barrier<N>(slot);
    }
    __syncthreads();  // everybody is done reading Y before the next band's x stage rewrites it
    cur = nxt;
  }
}

// ---------------------------------------------------------------------------------------
template <int N>
int launch_fused(int kind, const FusedArgs& a, int ctas, cudaStream_t st) {
  using C = FCfg<N>;
  const int smem = C::smem_bytes(a.m.nxo, a.m.ncol);
  // all occupied x planes fit one pass of the slots (the common case): leaner y stages;
  // band-limited radix-8 x radix-8 lines additionally use the sparse butterflies
  const bool one = (a.m.nxo + NB - 1) / NB <= C::SLOTS;
  constexpr bool SP_OK = LineFFT<N, +1>::SPARSE_OK && LineFFT<N, -1>::SPARSE_OK;
  const bool sp = SP_OK && one && a.band_limited;
  if (kind == 0) {
    static int once = set_smem_attr(k_yx_density<N, true, false>, 200 * 1024) |
                      set_smem_attr(k_yx_density<N, false, false>, 200 * 1024) |
                      set_smem_attr(k_yx_density<N, true, SP_OK>, 200 * 1024);
    if (once) return once;
    if (sp) k_yx_density<N, true, SP_OK><<<ctas, C::NT, smem, st>>>(a);
    else if (one) k_yx_density<N, true, false><<<ctas, C::NT, smem, st>>>(a);
    else k_yx_density<N, false, false><<<ctas, C::NT, smem, st>>>(a);
  } else if (kind == 2) {
    static int once = set_smem_attr(k_x_vmul_cached<N, true, false>, 200 * 1024) |
                      set_smem_attr(k_x_vmul_cached<N, false, false>, 200 * 1024) |
                      set_smem_attr(k_x_vmul_cached<N, true, SP_OK>, 200 * 1024);
    if (once) return once;
    const int smem_c = C::smem_bytes_cached(a.m.nxo);
    if (sp) k_x_vmul_cached<N, true, SP_OK><<<ctas, C::NT, smem_c, st>>>(a);
    else if (one) k_x_vmul_cached<N, true, false><<<ctas, C::NT, smem_c, st>>>(a);
    else k_x_vmul_cached<N, false, false><<<ctas, C::NT, smem_c, st>>>(a);
  } else {
    static int once = set_smem_attr(k_yx_vmul<N, true, false>, 200 * 1024) |
                      set_smem_attr(k_yx_vmul<N, false, false>, 200 * 1024) |
                      set_smem_attr(k_yx_vmul<N, true, SP_OK>, 200 * 1024);
    if (once) return once;
    if (sp) k_yx_vmul<N, true, SP_OK><<<ctas, C::NT, smem, st>>>(a);
    else if (one) k_yx_vmul<N, true, false><<<ctas, C::NT, smem, st>>>(a);
    else k_yx_vmul<N, false, false><<<ctas, C::NT, smem, st>>>(a);
  }
  JRB_CHECK_LAUNCH("fused yx pass launch");
  return 0;
}

template <int N>
int fused_smem_bytes(int nxo, int ncol) {
  return FCfg<N>::smem_bytes(nxo, ncol);
}
template <int N>
int fused_threads() {
  return FCfg<N>::NT;
}
// psi(r) cache: complex numbers per band-plane, or -1 when k_x_vmul_cached does not fit shared
// memory for this many occupied x planes
template <int N>
int fused_psi_plane(int nxo) {
  return FCfg<N>::smem_bytes_cached(nxo) <= 200 * 1024 ? FCfg<N>::PSI_PLANE : -1;
}

// per-translation-unit dispatch (fft_fused_g*.cu); return 1 if n is not in the group
int fused_group0(int kind, int n, const FusedArgs& a, int ctas, cudaStream_t st);
int fused_group1(int kind, int n, const FusedArgs& a, int ctas, cudaStream_t st);
int fused_group2(int kind, int n, const FusedArgs& a, int ctas, cudaStream_t st);
// shared memory the fused kernels need for axis length n (or -1 if not compiled)
int fused_smem_need(int n, int nxo, int ncol);

}  // namespace jrb
