// Fused y+x pencil passes for 128 x 128 planes (hand-written sm_100a kernels).
//
// A 128-point line does not fit the scheme of fft_fused.cuh: 16 elements per thread plus the
// per-plane density accumulators need more registers than two resident CTAs have, and a
// 128 x 128 slab pair does not fit shared memory.  What the plane-wave sphere offers instead:
// the occupied frequencies satisfy |f| < 32 (alias-free grids have 4 g_max + 1 <= n), so ONE
// radix-2 step is free of butterflies.  Decimation in frequency of the inverse transform,
//     psi[2 n' + p] = sum_{f'} ( c[f(f')] w_128^{+f p} ) w_64^{+f' n'},   f' = f mod 64,
// turns a 128-point line with band-limited input into TWO independent 64-point lines (even and
// odd outputs) of the same folded input, the odd one pre-multiplied by a phase; no two occupied
// frequencies fold onto the same f'.  The forward transform is the mirror image
//     c[f] = E[f'] + w_128^{-f} O[f'],   E/O = 64-point transforms of the even / odd samples,
// needed only on the occupied f.  So the whole plane is processed with the 8 x 8 line plan of
// the 64^3 kernels (8 elements per thread, twiddles in registers):
//   * the y parity splits the plane in two independent half planes Y_py[xo][y'] (y = 2y' + py):
//     39 x 64 complex = 40 KB, double buffered;
//   * in the x stage the two x parities of a line run side by side on two 64-thread half slots.
// One CTA of 512 threads per SM (16 warps, <= 128 registers): 8 half slots of 8 lanes x 8
// threads; half slot h owns x planes 8h..8h+7 in the y stages and (line group h / 2, x parity
// h % 2) in the x stages.
//   k_yx128_density : work item = (z, py, band group); f |psi|^2 accumulated in registers per
//                     (z, py) half plane -> partial half planes -> k_rho_reduce128.
//   k_yx128_vmul    : same work items; y inverse -> x inverse -> * v_eff / N -> x forward (the two
//                     x parities are summed in the slab) -> y forward; each y parity writes its
//                     share of the columns to its own buffer, k_z_fwd_gather adds the two.
#pragma once
#include "fft_fused.cuh"

namespace jrb {

struct F128 {
  static constexpr int N = 128, M = 64;
  static constexpr int HS_THREADS = 64;  // half slot: 8 lanes x 8 threads
  static constexpr int HSLOTS = 8;
  static constexpr int NT = HS_THREADS * HSLOTS;  // 512
  static constexpr int SX = M + 1;                // Y row stride (complex), == 1 (mod 8)
  static constexpr int EXCH = HSLOTS * M * NB;    // exchange buffers (complex)
  static JRB_HD int ybuf_elems(int nxo) { return nxo * SX; }
  static JRB_HD int stage_elems(int ncol) { return (ncol + 7) / 8 * 8; }
  static JRB_HD int smem_bytes(int nxo, int ncol, bool) {
    return (2 * ybuf_elems(nxo) + EXCH + 2 * stage_elems(ncol)) * (int)sizeof(cplx);
  }
};

__device__ __forceinline__ void hs_barrier(int hs) {
  asm volatile("bar.sync %0, 64;" ::"r"(hs + 1) : "memory");
}
__device__ __forceinline__ void slot128_barrier(int slot) {
  asm volatile("bar.sync %0, 128;" ::"r"(slot + 9) : "memory");
}

// frequency of folded index f' (|f| < 32): f' < 32 -> f', else f' + 64
__device__ __forceinline__ int unfold128(int fp) { return fp < 32 ? fp : fp + 64; }

// v[m] *= w_128^{DIR f},  f = unfold(tj + 8 m):  w_128^{DIR tj} (thread constant c) times
// w_16^{DIR s(m)}, s = m for m < 4 and m + 8 above (compile-time roots).
template <int DIR>
__device__ __forceinline__ void phase128(cplx (&v)[8], cplx c) {
  const cplx cd = DIR > 0 ? c : cconj(c);
  static_for<0, 8>([&](auto m_) {
    constexpr int m = decltype(m_)::value;
    constexpr int s = m < 4 ? m : m + 8;
    v[m] = mul_root<16, s, DIR>(cmul(v[m], cd));
  });
}

// position in the (item, band) list of a CTA; item w = plane * ngroups + gl
struct Pos128 {
  int w, plane, gl, gmod, band;
};
__device__ __forceinline__ int pos_bands(const FusedArgs& a, const Pos128& p) {
  return min(NB, a.nb - p.gmod * NB);
}
__device__ __forceinline__ Pos128 pos_first(const FusedArgs& a, int w) {
  Pos128 p;
  p.w = w;
  p.plane = w / a.ngroups;
  p.gl = w - p.plane * a.ngroups;
  p.gmod = (a.g0 + p.gl) % a.ngpk;
  p.band = 0;
  return p;
}
__device__ __forceinline__ Pos128 pos_next(const FusedArgs& a, Pos128 p, int w_end, int gmod0) {
  if (p.w < w_end && ++p.band >= pos_bands(a, p)) {
    p.band = 0;
    ++p.w;
    ++p.gl;
    if (++p.gmod == a.ngpk) p.gmod = 0;
    if (p.gl == a.ngroups) {
      p.gl = 0;
      ++p.plane;
      p.gmod = gmod0;
    }
  }
  return p;
}
__device__ __forceinline__ cplx* pos_columns(const FusedArgs& a, int gl, int z, int band) {
  return a.wa + (long long)(gl * a.m.nz + z) * (a.m.ncol * NB) + band;
}
__device__ __forceinline__ void stage128(const FusedArgs& a, const cplx* src, bool valid,
                                         cplx* stage) {
  if (valid)
    for (int c = threadIdx.x; c < a.m.ncol; c += F128::NT)
      fused_cp_async16(stage + c, src + (long long)c * NB);
  fused_cp_commit();
}

// per-thread constants shared by both kernels
struct Thr128 {
  int lane, tj, hs;
  cplx c;            // w_128^{+tj}
  unsigned pk[8];    // low 16: column + 1 of (x plane hs * 8 + lane, fy' = tj + 8 m);
                     // high 16: Y row offset + 1 of fx' = tj + 8 m
};
__device__ __forceinline__ Thr128 thr128_init(const FusedArgs& a) {
  Thr128 t;
  const int tid = threadIdx.x;
  t.lane = tid % NB;
  t.tj = (tid / NB) % 8;
  t.hs = tid / F128::HS_THREADS;
  const cplx w = a.tw[t.tj];  // exp(-2 pi i tj / 128)
  t.c = cconj(w);
  const int xo0 = t.hs * NB + t.lane;
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int f = unfold128(t.tj + 8 * m);
    unsigned lo = 0, hi = 0;
    const int xo = a.m.xmap[f];
    if (xo >= 0) hi = (unsigned)(xo * F128::SX + 1);
    if (xo0 < a.m.nxo) lo = (unsigned)(a.m.ycol[(long long)xo0 * F128::N + f] + 1);
    t.pk[m] = lo | (hi << 16);
  }
  return t;
}

// y stage, inverse, one y parity: staged columns -> Y[xo][y'] (y = 2 y' + py)
__device__ __forceinline__ void y_inverse128(const FusedArgs& a, const Thr128& t, int py,
                                             const cplx* stage, cplx* ybuf, cplx* ex,
                                             const cplx (&tw)[1][7]) {
  using F = LineFFT<64, +1>;
  if (t.hs * NB >= a.m.nxo) return;  // whole half slot without an x plane (barriers are per half slot)
  const int xo = t.hs * NB + t.lane;
  cplx va[1][8];
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int col = (int)(t.pk[m] & 0xffffu) - 1;
    va[0][m] = col >= 0 ? stage[col] : czero();
  }
  if (py) phase128<+1>(va[0], t.c);
  F::template stageA_store<NB>(va, ex, t.tj);
  hs_barrier(t.hs);
  cplx vb[1][8];
  F::template stageB_load<NB>(vb, ex, tw, t.tj);
  if (xo < a.m.nxo) {
    cplx* out = ybuf + xo * F128::SX;
#pragma unroll
    for (int m = 0; m < 8; ++m) out[t.tj + 8 * m] = vb[0][m];
  }
  hs_barrier(t.hs);
}

// ---------------------------------------------------------------------------------------
// grid: persistent CTAs (one per SM); dynamic smem: F128::smem_bytes(nxo, ncol, false)
// rho_part: [gridDim.x * segmax][128 * 64] partial half planes, index x * 64 + y';
// seg_z: plane tag 2 z + py of each partial, -1 = unused
__global__ void __launch_bounds__(F128::NT, 1) k_yx128_density(FusedArgs a) {
  using F = LineFFT<64, +1>;
  extern __shared__ __align__(16) unsigned char smem_raw_[];
  cplx* ybuf0 = reinterpret_cast<cplx*>(smem_raw_);
  const int ysz = F128::ybuf_elems(a.m.nxo);
  cplx* exbase = ybuf0 + 2 * ysz;
  cplx* stage0 = exbase + F128::EXCH;
  const int ssz = F128::stage_elems(a.m.ncol);
  const Thr128 t = thr128_init(a);
  cplx* ex = exbase + (size_t)t.hs * F128::M * NB + t.lane;
  cplx tw[1][7];
  F::load_twiddles(tw, a.tw64, t.tj);
  const int slot = t.hs >> 1, px = t.hs & 1;

  double acc[2][8];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int m = 0; m < 8; ++m) acc[r][m] = 0.0;

  const long long W = (long long)a.m.nz * 2 * a.ngroups;
  const int c = blockIdx.x, G = gridDim.x;
  const int w_end = (int)((c + 1) * W / G);
  const int gmod0 = a.g0 % a.ngpk;
  Pos128 cur = pos_first(a, (int)(c * W / G));
  int seg = 0;
  auto flush = [&](int plane) {
    double* out = a.rho_part + ((long long)c * a.segmax + seg) * (F128::N * F128::M);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int yp = (r * 4 + slot) * NB + t.lane;
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int x = 2 * (t.tj + 8 * m) + px;
        out[x * F128::M + yp] = acc[r][m];
        acc[r][m] = 0.0;
      }
    }
    if (threadIdx.x == 0) a.seg_z[c * a.segmax + seg] = plane;
    ++seg;
  };
  auto columns = [&](const Pos128& p) { return pos_columns(a, p.gl, p.plane >> 1, p.band); };

  if (cur.w < w_end) {
    stage128(a, columns(cur), true, stage0);
    fused_cp_wait_all();
    __syncthreads();
    {
      const Pos128 n1 = pos_next(a, cur, w_end, gmod0);
      stage128(a, columns(n1), n1.w < w_end, stage0 + ssz);
    }
    y_inverse128(a, t, cur.plane & 1, stage0, ybuf0, ex, tw);
    int par = 0;
    int cur_plane = cur.plane;
    double fw_next = a.focc[(a.g0 + cur.gl) * NB + cur.band];
    while (cur.w < w_end) {
      fused_cp_wait_all();
      __syncthreads();
      const Pos128 nxt = pos_next(a, cur, w_end, gmod0);
      {
        const Pos128 n2 = pos_next(a, nxt, w_end, gmod0);
        stage128(a, columns(n2), n2.w < w_end, stage0 + par * ssz);
      }
      if (cur.plane != cur_plane) {
        flush(cur_plane);
        cur_plane = cur.plane;
      }
      const double fw = fw_next;
      if (nxt.w < w_end) fw_next = a.focc[(a.g0 + nxt.gl) * NB + nxt.band];
      const cplx* ybuf = ybuf0 + par * ysz;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int yp = (r * 4 + slot) * NB + t.lane;
        cplx va[1][8];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const unsigned row = t.pk[m] >> 16;
          va[0][m] = row != 0 ? ybuf[(int)row - 1 + yp] : czero();
        }
        if (px) phase128<+1>(va[0], t.c);
        F::template stageA_store<NB>(va, ex, t.tj);
        hs_barrier(t.hs);
        cplx vb[1][8];
        F::template stageB_load<NB>(vb, ex, tw, t.tj);
#pragma unroll
        for (int m = 0; m < 8; ++m)
          acc[r][m] += fw * (vb[0][m].x * vb[0][m].x + vb[0][m].y * vb[0][m].y);
        hs_barrier(t.hs);
      }
      if (nxt.w < w_end)
        y_inverse128(a, t, nxt.plane & 1, stage0 + (par ^ 1) * ssz, ybuf0 + (par ^ 1) * ysz, ex, tw);
      cur = nxt;
      par ^= 1;
    }
    fused_cp_wait_all();
    flush(cur_plane);
  }
  if (threadIdx.x == 0)
    for (int s = seg; s < a.segmax; ++s) a.seg_z[c * a.segmax + s] = -1;
}

// rho[x][y][z] += sum of the partial half planes tagged 2 z + py, in slot order (deterministic)
// grid: (128 * 64 / 32, 2 nz), block 32 x 8
static __global__ void __launch_bounds__(256)
k_rho_reduce128(const double* __restrict__ part, const int* __restrict__ seg_z, int nctas,
                int segmax, int ngroups, int nz, double* __restrict__ rho) {
  __shared__ double sh[8][33];
  constexpr int NXY = F128::N * F128::M;
  const int e = blockIdx.x * 32 + threadIdx.x;  // x * 64 + y'
  const int plane = blockIdx.y;
  const long long W = (long long)nz * 2 * ngroups;
  const int c_lo = max(0, (int)(((long long)plane * ngroups * nctas) / W) - 1);
  const int c_hi = min(nctas - 1, (int)((((long long)plane + 1) * ngroups * nctas) / W) + 1);
  double s = 0.0;
  for (int k = c_lo * segmax + threadIdx.y; k < (c_hi + 1) * segmax; k += 8)
    if (seg_z[k] == plane) s += part[(long long)k * NXY + e];
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0) {
    double r = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) r += sh[j][threadIdx.x];
    const int x = e / F128::M, y = 2 * (e % F128::M) + (plane & 1), z = plane >> 1;
    rho[((long long)x * F128::N + y) * nz + z] += r;
  }
}

// ---------------------------------------------------------------------------------------
// Hamiltonian-apply middle on 128 x 128 planes.  Work item = (z, py, band group) as in the density
// kernel, so v_eff(x, y = 2y' + py, z) / N stays in registers over many bands.  The y parities are
// independent all the way back to the columns: parity py writes its share of the y-forward
// result to wout[py][group][z][col][NB]; k_z_fwd_gather adds the two shares when it loads.
// grid: persistent CTAs; dynamic smem: F128::smem_bytes(nxo, ncol, false)
__global__ void __launch_bounds__(F128::NT, 1) k_yx128_vmul(FusedArgs a) {
  using FI = LineFFT<64, +1>;
  using FF = LineFFT<64, -1>;
  extern __shared__ __align__(16) unsigned char smem_raw_[];
  cplx* ybuf0 = reinterpret_cast<cplx*>(smem_raw_);
  const int ysz = F128::ybuf_elems(a.m.nxo);
  cplx* exbase = ybuf0 + 2 * ysz;
  cplx* stage0 = exbase + F128::EXCH;
  const int ssz = F128::stage_elems(a.m.ncol);
  const Thr128 t = thr128_init(a);
  cplx* ex = exbase + (size_t)t.hs * F128::M * NB + t.lane;
  cplx tw[1][7];  // inverse twiddles; the forward ones are their conjugates
  FI::load_twiddles(tw, a.tw64, t.tj);
  const int slot = t.hs >> 1, px = t.hs & 1;
  const long long nyz = (long long)F128::N * a.m.nz;

  // v_eff(x, y, z) / N at the points of this thread's two x-stage lines
  double vv[2][8];
  auto load_v = [&](int plane) {
    const int z = plane >> 1, py = plane & 1;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int y = 2 * ((r * 4 + slot) * NB + t.lane) + py;
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int x = 2 * (t.tj + 8 * m) + px;
        vv[r][m] = __ldg(a.veff + (long long)x * nyz + (long long)y * a.m.nz + z) * a.vscale;
      }
    }
  };

  const long long W = (long long)a.m.nz * 2 * a.ngroups;
  const int c = blockIdx.x, G = gridDim.x;
  const int w_end = (int)((c + 1) * W / G);
  const int gmod0 = a.g0 % a.ngpk;
  Pos128 cur = pos_first(a, (int)(c * W / G));
  if (cur.w >= w_end) return;
  auto columns = [&](const Pos128& p) { return pos_columns(a, p.gl, p.plane >> 1, p.band); };

  stage128(a, columns(cur), true, stage0);
  fused_cp_wait_all();
  __syncthreads();
  {
    const Pos128 n1 = pos_next(a, cur, w_end, gmod0);
    stage128(a, columns(n1), n1.w < w_end, stage0 + ssz);
  }
  y_inverse128(a, t, cur.plane & 1, stage0, ybuf0, ex, tw);
  int par = 0;
  int cur_plane = cur.plane;
  load_v(cur_plane);
  while (cur.w < w_end) {
    fused_cp_wait_all();
    __syncthreads();
    const Pos128 nxt = pos_next(a, cur, w_end, gmod0);
    {
      const Pos128 n2 = pos_next(a, nxt, w_end, gmod0);
      stage128(a, columns(n2), n2.w < w_end, stage0 + par * ssz);
    }
    if (cur.plane != cur_plane) {
      cur_plane = cur.plane;
      load_v(cur_plane);
    }
    const int py = cur.plane & 1;
    cplx* ybuf = ybuf0 + par * ysz;
    cplx* yother = ybuf0 + (par ^ 1) * ysz;
    // x stage: the two x parities of a line side by side (half slots 2 s and 2 s + 1)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int yp = (r * 4 + slot) * NB + t.lane;
      cplx va[1][8];
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const unsigned row = t.pk[m] >> 16;
        va[0][m] = row != 0 ? ybuf[(int)row - 1 + yp] : czero();
      }
      slot128_barrier(slot);  // both x parities hold their inputs
      if (px) phase128<+1>(va[0], t.c);
      FI::template stageA_store<NB>(va, ex, t.tj);
      hs_barrier(t.hs);
      cplx vb[1][8];
      FI::template stageB_load<NB>(vb, ex, tw, t.tj);
#pragma unroll
      for (int m = 0; m < 8; ++m) vb[0][m] = cscale(vb[0][m], vv[r][m]);
      hs_barrier(t.hs);
      FF::template stageA_store<NB>(vb, ex, t.tj);
      hs_barrier(t.hs);
      cplx vc[1][8];
      FF::template stageB_load<NB, true>(vc, ex, tw, t.tj);
      if (px) phase128<-1>(vc[0], t.c);
      // x parity 0 goes back in place (safe: the slot-wide barrier above ordered it after the odd
      // half slot's input loads), parity 1 into the other Y buffer, which is idle right now; the
      // y-forward stage adds the two when it loads
      cplx* dsty = px ? yother : ybuf;
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const unsigned row = t.pk[m] >> 16;
        if (row != 0) dsty[(int)row - 1 + yp] = vc[0][m];
      }
    }
    __syncthreads();
    // y stage, forward: Y -> this parity's share of the columns
    if (t.hs * NB < a.m.nxo) {
      const int xo = t.hs * NB + t.lane;
      const bool ok = xo < a.m.nxo;
      const cplx* in = ybuf + (ok ? xo : 0) * F128::SX;
      const cplx* in1 = yother + (ok ? xo : 0) * F128::SX;
      cplx va[1][8];
#pragma unroll
      for (int m = 0; m < 8; ++m)
        va[0][m] = ok ? cadd(in[t.tj + 8 * m], in1[t.tj + 8 * m]) : czero();
      FF::template stageA_store<NB>(va, ex, t.tj);
      hs_barrier(t.hs);
      cplx vb[1][8];
      FF::template stageB_load<NB, true>(vb, ex, tw, t.tj);
      if (py) phase128<-1>(vb[0], t.c);
      cplx* dst = a.wout[py] + (long long)(cur.gl * a.m.nz + (cur.plane >> 1)) * (a.m.ncol * NB) + cur.band;
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int col = (int)(t.pk[m] & 0xffffu) - 1;
        if (col >= 0) dst[(long long)col * NB] = vb[0][m];
      }
      hs_barrier(t.hs);
    }
    if (nxt.w < w_end)
      y_inverse128(a, t, nxt.plane & 1, stage0 + (par ^ 1) * ssz, ybuf0 + (par ^ 1) * ysz, ex, tw);
    cur = nxt;
    par ^= 1;
  }
  fused_cp_wait_all();
}

// ---------------------------------------------------------------------------------------
inline int launch_fused128(int kind, const FusedArgs& a, int ctas, cudaStream_t st) {
  if (kind == 0) {
    const int smem = F128::smem_bytes(a.m.nxo, a.m.ncol, false);
    static int once = set_smem_attr(k_yx128_density, 227 * 1024);
    if (once) return once;
    k_yx128_density<<<ctas, F128::NT, smem, st>>>(a);
  } else {
    const int smem = F128::smem_bytes(a.m.nxo, a.m.ncol, false);
    static int once = set_smem_attr(k_yx128_vmul, 227 * 1024);
    if (once) return once;
    k_yx128_vmul<<<ctas, F128::NT, smem, st>>>(a);
  }
  JRB_CHECK_LAUNCH("fused 128 yx pass launch");
  return 0;
}

}  // namespace jrb
