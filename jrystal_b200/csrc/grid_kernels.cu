// Fused elementwise + reduction kernels on the FFT grid and on the plane-wave sphere:
// Hartree / external / LDA-XC energies and the effective potential in one sweep each
// (jrystal/_src/potential.py:58-77,153-166,255-279; energy.py:68-82,121-135,204-211;
// xc.py:54-64,109-124), the per-orbital kinetic and band expectations on the sphere
// (energy.py:172-180; braket.py:189-206) and the scatter/gather API-parity helpers
// (utils.py:277-308).  All HBM-bound, FP64, deterministic two-stage reductions.
#include <algorithm>
#include <cmath>
#include <vector>

#include "plan.h"
#include "xc_functionals.cuh"

namespace jrb {

constexpr int RED_THREADS = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of up to NV values per thread; result valid on thread 0
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* out /* [NV] */) {
  __shared__ double sh[NV][RED_THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double s = warp_sum(v[i]);
    if (lane == 0) sh[i][w] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double s = 0;
      for (int j = 0; j < (int)(blockDim.x >> 5); ++j) s += sh[i][j];
      out[i] = s;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ int fftfreq_int(int i, int n) { return i < (n + 1) / 2 ? i : i - n; }

struct GridGeom {
  int nx, ny, nz;
  long long n;
  double b[9];  // rows b1, b2, b3
  double vol;
};

__device__ __forceinline__ void g_of(const GridGeom& g, long long lin, double& gx, double& gy,
                                     double& gz) {
  const int z = (int)(lin % g.nz);
  const int y = (int)((lin / g.nz) % g.ny);
  const int x = (int)(lin / ((long long)g.nz * g.ny));
  const double fx = fftfreq_int(x, g.nx), fy = fftfreq_int(y, g.ny), fz = fftfreq_int(z, g.nz);
  gx = fx * g.b[0] + fy * g.b[3] + fz * g.b[6];
  gy = fx * g.b[1] + fy * g.b[4] + fz * g.b[7];
  gz = fx * g.b[2] + fy * g.b[5] + fz * g.b[8];
}

static GridGeom geom_of(const jrb_plan* p) {
  GridGeom g;
  g.nx = p->nx; g.ny = p->ny; g.nz = p->nz; g.n = p->ngrid; g.vol = p->vol;
  for (int i = 0; i < 9; ++i) g.b[i] = p->recip[i];
  return g;
}

// ---------------------------------------------------------------------------------------
// V_ext(G) = -(N/Omega) 4 pi sum_a Z_a exp(-i G.R_a) / (|G|^2 + 1e-10), V_ext(0) = 0
__global__ void k_vext(GridGeom g, const double* __restrict__ pos, const double* __restrict__ chg,
                       int na, cplx* __restrict__ vext) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < g.n;
       i += (long long)gridDim.x * blockDim.x) {
    double gx, gy, gz;
    g_of(g, i, gx, gy, gz);
    const double g2 = gx * gx + gy * gy + gz * gz;
    double re = 0, im = 0;
    if (i != 0) {
      for (int a = 0; a < na; ++a) {
        const double ph = gx * pos[3 * a] + gy * pos[3 * a + 1] + gz * pos[3 * a + 2];
        double s, c;
        sincos(ph, &s, &c);
        const double vi = chg[a] / (g2 + 1e-10) * (4.0 * M_PI);
        re += vi * c;
        im -= vi * s;
      }
    }
    const double f = -(double)g.n / g.vol;
    vext[i] = cmake(re * f, im * f);
  }
}

int launch_set_atoms(jrb_plan* p, const double* pos_h, const double* chg_h, int na,
                     cudaStream_t st) {
  if (na <= 0 || !pos_h || !chg_h) {
    set_error("jrb_set_atoms: need natoms > 0 and non-null positions/charges");
    return JRB_EINVAL;
  }
  if (p->d_pos) cudaFree(p->d_pos);
  if (p->d_chg) cudaFree(p->d_chg);
  if (p->d_atom_part) cudaFree(p->d_atom_part);
  p->d_pos = p->d_chg = p->d_atom_part = nullptr;
  p->atoms_on_device = 0;
  JRB_CUDA(cudaMalloc(&p->d_pos, sizeof(double) * 3 * na));
  JRB_CUDA(cudaMalloc(&p->d_chg, sizeof(double) * na));
  JRB_CUDA(cudaMalloc(&p->d_atom_part, sizeof(double) * 3 * 64 * na));
  JRB_CUDA(cudaMemcpyAsync(p->d_pos, pos_h, sizeof(double) * 3 * na, cudaMemcpyHostToDevice, st));
  JRB_CUDA(cudaMemcpyAsync(p->d_chg, chg_h, sizeof(double) * na, cudaMemcpyHostToDevice, st));
  const int blocks = (int)std::min<long long>((p->ngrid + 255) / 256, 148 * 8);
  k_vext<<<blocks, 256, 0, st>>>(geom_of(p), p->d_pos, p->d_chg, na, p->d_vext);
  JRB_CHECK_LAUNCH("k_vext");
  JRB_CUDA(cudaStreamSynchronize(st));  // the host arrays may go away after the call
  p->natoms = na;
  p->atoms_on_device = 1;
  return 0;
}

// dE_ext / dR_a = (4 pi Z_a / N) sum_G G Im(exp(+i G.R_a) rho_hat(G)) / (|G|^2 + 1e-10), G != 0:
// the position cotangent of energy.external (energy.py:121-135 through potential.py:153-166; the
// reference gets it from jax.grad, docs/tutorial/differentiation.rst:128-147).
// grid: (64, natoms), block 256; partials [atom][block][3], summed in block order (deterministic)
__global__ void __launch_bounds__(RED_THREADS)
k_ext_pos_grad(GridGeom g, const cplx* __restrict__ rho_hat, const double* __restrict__ pos,
               const double* __restrict__ chg, double* __restrict__ part) {
  const int a = blockIdx.y;
  const double rx = pos[3 * a], ry = pos[3 * a + 1], rz = pos[3 * a + 2];
  double acc[3] = {0.0, 0.0, 0.0};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < g.n;
       i += (long long)gridDim.x * blockDim.x) {
    if (i == 0) continue;
    double gx, gy, gz;
    g_of(g, i, gx, gy, gz);
    double sn, cs;
    sincos(gx * rx + gy * ry + gz * rz, &sn, &cs);
    const cplx r = rho_hat[i];
    const double w = (cs * r.y + sn * r.x) / (gx * gx + gy * gy + gz * gz + 1e-10);
    acc[0] += gx * w;
    acc[1] += gy * w;
    acc[2] += gz * w;
  }
  const double f = 4.0 * M_PI * chg[a] / (double)g.n;
  double out[3];
  block_sum<3>(acc, out);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) part[((long long)a * gridDim.x + blockIdx.x) * 3 + c] = out[c] * f;
  }
}

__global__ void k_ext_pos_grad_reduce(const double* __restrict__ part, int nblocks, int n3,
                                      double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // atom * 3 + component
  if (i >= n3) return;
  const int a = i / 3, c = i % 3;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += part[((long long)a * nblocks + b) * 3 + c];
  out[i] = s;
}

__global__ void k_rho_to_complex(const double* __restrict__ rho, int ns, long long n,
                                 cplx* __restrict__ grid);

int launch_external_position_gradient(jrb_plan* p, const double* rho, double* grad,
                                      cudaStream_t st) {
  if (!p->atoms_on_device) {
    set_error("jrb_external_position_gradient: call jrb_set_atoms first");
    return JRB_EINVAL;
  }
  const GridGeom g = geom_of(p);
  const int blocks = (int)std::min<long long>((p->ngrid + RED_THREADS - 1) / RED_THREADS,
                                              (long long)p->n_partial_blocks);
  k_rho_to_complex<<<blocks, RED_THREADS, 0, st>>>(rho, p->ns, p->ngrid, p->d_grid);
  JRB_CHECK_LAUNCH("k_rho_to_complex");
  int rc = launch_fft3d_dense(p, p->d_grid, p->d_grid, JRB_FFT_FORWARD, 1, 1.0, st);
  if (rc) return rc;
  k_ext_pos_grad<<<dim3(64, p->natoms), RED_THREADS, 0, st>>>(g, p->d_grid, p->d_pos, p->d_chg,
                                                              p->d_atom_part);
  JRB_CHECK_LAUNCH("k_ext_pos_grad");
  k_ext_pos_grad_reduce<<<(3 * p->natoms + 127) / 128, 128, 0, st>>>(p->d_atom_part, 64,
                                                                   3 * p->natoms, grad);
  JRB_CHECK_LAUNCH("k_ext_pos_grad_reduce");
  return 0;
}

// ---------------------------------------------------------------------------------------
__global__ void k_rho_to_complex(const double* __restrict__ rho, int ns, long long n,
                                 cplx* __restrict__ grid) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    double v = rho[i];
    if (ns == 2) v += rho[n + i];
    grid[i] = cmake(v, 0.0);
  }
}

// n(G) -> E_H, E_ext partial sums and v_H(G) + V_ext(G) in place.
__global__ void __launch_bounds__(RED_THREADS)
k_hartree_ext(GridGeom g, cplx* __restrict__ grid, const cplx* __restrict__ vext, int kohn_sham,
              int parts, double* __restrict__ partials) {
  // potential written back: parts bit 0 = Hartree, bit 1 = external; bit 3 = the reference's
  // potential.effective semantics (Hartree potential halved unless kohn_sham, potential.py:67-75)
  const double hs = (parts & 1) ? (((parts & 8) && !kohn_sham) ? 0.5 : 1.0) : 0.0;
  const double es = (parts & 2) ? 1.0 : 0.0;
  double acc[2] = {0.0, 0.0};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < g.n;
       i += (long long)gridDim.x * blockDim.x) {
    double gx, gy, gz;
    g_of(g, i, gx, gy, gz);
    const double g2 = gx * gx + gy * gy + gz * gz;
    const cplx nG = grid[i];
    const cplx ve = vext[i];
    cplx vh = cmake(0.0, 0.0);
    if (i != 0) {
      const double f = 4.0 * M_PI / g2;
      vh = cmake(nG.x * f, nG.y * f);
    }
    // Re conj(v) n
    acc[0] += vh.x * nG.x + vh.y * nG.y;
    acc[1] += ve.x * nG.x + ve.y * nG.y;
    grid[i] = cmake(hs * vh.x + es * ve.x, hs * vh.y + es * ve.y);
  }
  double out[2];
  block_sum<2>(acc, out);
  if (threadIdx.x == 0) {
    const double w = g.vol / ((double)g.n * (double)g.n);
    partials[blockIdx.x * 4 + 0] = out[0] * w * (kohn_sham ? 1.0 : 0.5);
    partials[blockIdx.x * 4 + 1] = out[1] * w;
  }
}

// w_j(G) = i G_j rho_hat(G) / N  (the 1/N of ifftn folded in)
__global__ void k_gga_grad(GridGeom g, const cplx* __restrict__ rho_hat, cplx* __restrict__ w0,
                           cplx* __restrict__ w1, cplx* __restrict__ w2) {
  const double inv_n = 1.0 / (double)g.n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < g.n;
       i += (long long)gridDim.x * blockDim.x) {
    double gx, gy, gz;
    g_of(g, i, gx, gy, gz);
    const cplx r = rho_hat[i];
    const double a = -r.y * inv_n, b = r.x * inv_n;  // i * rho_hat / N
    w0[i] = cmake(gx * a, gx * b);
    w1[i] = cmake(gy * a, gy * b);
    w2[i] = cmake(gz * a, gz * b);
  }
}

// per grid point: sigma, eps, derivatives; vxc_loc = de/d rho (or eps), w_j <- de/d sigma * d_j
__global__ void __launch_bounds__(RED_THREADS)
k_gga_local(GridGeom g, const double* __restrict__ rho, int xc_id, int kohn_sham, int as_eps,
            cplx* __restrict__ w0, cplx* __restrict__ w1, cplx* __restrict__ w2,
            double* __restrict__ vxc_loc, double* __restrict__ partials) {
  double acc[1] = {0.0};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < g.n;
       i += (long long)gridDim.x * blockDim.x) {
    const double n = rho[i];
    const cplx d0 = w0[i], d1 = w1[i], d2 = w2[i];
    const double sigma = d0.x * d0.x + d0.y * d0.y + d1.x * d1.x + d1.y * d1.y + d2.x * d2.x +
                         d2.y * d2.y;
    const Dual e = pbe_eps(xc_id, n, sigma);
    const double vrho = e.v + n * e.r, vsig = n * e.s;  // derivatives of n * eps
    vxc_loc[i] = as_eps ? e.v : vrho;
    const double f = as_eps ? 0.0 : vsig;
    w0[i] = cmake(f * d0.x, f * d0.y);
    w1[i] = cmake(f * d1.x, f * d1.y);
    w2[i] = cmake(f * d2.x, f * d2.y);
    // E_xc: rho eps; band mode integrates the potential (energy.py:204-211 with kohn_sham):
    // sum rho v_xc = sum (rho de/d rho + 2 sigma de/d sigma)
    acc[0] += kohn_sham ? (vrho * n + 2.0 * vsig * sigma) : e.v * n;
  }
  double out[1];
  block_sum<1>(acc, out);
  if (threadIdx.x == 0) partials[blockIdx.x * 4 + 2] = out[0] * g.vol / (double)g.n;
}

// grid(G) += -2 i sum_j G_j fftn(de/d sigma d_j)(G)
__global__ void k_gga_add(GridGeom g, const cplx* __restrict__ w0, const cplx* __restrict__ w1,
                          const cplx* __restrict__ w2, cplx* __restrict__ grid) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < g.n;
       i += (long long)gridDim.x * blockDim.x) {
    double gx, gy, gz;
    g_of(g, i, gx, gy, gz);
    const cplx a = w0[i], b = w1[i], c = w2[i];
    const double re = gx * a.x + gy * b.x + gz * c.x, im = gx * a.y + gy * b.y + gz * c.y;
    cplx v = grid[i];
    v.x += 2.0 * im;
    v.y -= 2.0 * re;
    grid[i] = v;
  }
}

// v_HE(r) (complex, already / N) + xc -> veff[s], E_xc partial sums.
__global__ void __launch_bounds__(RED_THREADS)
k_veff_xc(GridGeom g, const cplx* __restrict__ grid, const double* __restrict__ rho, int ns,
          int xc_id, int kohn_sham, int parts, const double* __restrict__ vxc_pre,
          double* __restrict__ veff, double* __restrict__ partials) {
  // parts bit 2 = add the xc term; bit 3 = reference semantics: eps_xc unless kohn_sham
  // (xc.py:247-250), otherwise the functional derivative v_xc = eps + rho eps'
  const double xs = (parts & 4) ? 1.0 : 0.0;
  const bool as_eps = (parts & 8) && !kohn_sham;
  double acc[1] = {0.0};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < g.n;
       i += (long long)gridDim.x * blockDim.x) {
    const double vhe = grid[i].x;
    if (vxc_pre) {  // GGA: the local part comes from k_gga_local, the gradient part is in vhe
      veff[i] = (parts & 4) ? vhe + vxc_pre[i] : vhe;  // without the xc part vxc_pre is stale: not even read
    } else if (ns == 1) {
      const double n = rho[i];
      double e, de;
      lda_eps(xc_id, n, e, de);
      const double vxc = e + n * de;
      veff[i] = vhe + xs * (as_eps ? e : vxc);
      acc[0] += (kohn_sham ? vxc : e) * n;
    } else {
      // exchange: eps = 1/2 [eps_x(2 rho_up) + eps_x(2 rho_dn)]  (xc.py:56-59); correlation: the
      // functional's spin-polarised form (xc.py:60-61)
      const double ru = rho[i], rd = rho[g.n + i];
      double e, deu, ded;
      lda_pol_eps(xc_id, ru, rd, e, deu, ded);
      const double n = ru + rd;
      // deu, ded = d eps / d rho_s
      if (kohn_sham) {
        // reference vxc_lda (xc.py:114-122): v_s = eps + rho_s d eps/d rho_s
        const double vu = e + ru * deu, vd = e + rd * ded;
        veff[i] = vhe + xs * vu;
        veff[g.n + i] = vhe + xs * vd;
        acc[0] += vu * ru + vd * rd;
      } else {
        veff[i] = vhe + xs * (as_eps ? e : e + n * deu);
        veff[g.n + i] = vhe + xs * (as_eps ? e : e + n * ded);
        acc[0] += e * n;
      }
    }
  }
  double out[1];
  block_sum<1>(acc, out);
  if (threadIdx.x == 0 && !vxc_pre) partials[blockIdx.x * 4 + 2] = out[0] * g.vol / (double)g.n;
}

__global__ void k_reduce_partials(const double* __restrict__ partials, int nblocks, int ncomp,
                                  double* __restrict__ out) {
  // one CTA of 256 threads per component: 8 warps, fixed order -> deterministic
  __shared__ double red[8];
  const int c = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double s = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 256) s += partials[i * 4 + c];
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    out[c] = t;
  }
}

int launch_grid_potential(jrb_plan* p, const double* rho, int xc_id, int kohn_sham, int parts,
                          double* energies, double* veff, cudaStream_t st) {
  const bool gga = xc_id == JRB_XC_GGA_X_PBE || xc_id == JRB_XC_GGA_PBE;
  if (xc_id != JRB_XC_LDA_X && xc_id != JRB_XC_LDA_X_C_PW && !gga) {
    set_error("jrb_grid_potential: unsupported xc id (lda_x, lda_x+lda_c_pw, gga_x_pbe, "
              "gga_x_pbe+gga_c_pbe)");
    return JRB_EUNSUPPORTED;
  }
  if (p->ns == 2 && gga) {
    set_error("jrb_grid_potential: spin-polarised GGA is not implemented");
    return JRB_EUNSUPPORTED;
  }
  if (p->natoms <= 0) {
    set_error("jrb_grid_potential: call jrb_set_atoms first");
    return JRB_EINVAL;
  }
  const GridGeom g = geom_of(p);
  const int blocks = (int)std::min<long long>((p->ngrid + RED_THREADS - 1) / RED_THREADS,
                                              (long long)p->n_partial_blocks);
  k_rho_to_complex<<<blocks, RED_THREADS, 0, st>>>(rho, p->ns, p->ngrid, p->d_grid);
  JRB_CHECK_LAUNCH("k_rho_to_complex");
  int rc = launch_fft3d_dense(p, p->d_grid, p->d_grid, JRB_FFT_FORWARD, 1, 1.0, st);
  if (rc) return rc;
  const bool gga_on = gga && (parts & 4);
  const bool as_eps = (parts & 8) && !kohn_sham;
  cplx* w[3] = {p->d_gga, p->d_gga + p->ngrid, p->d_gga + 2 * p->ngrid};
  if (gga_on) {
    // gradient of rho from rho_hat (still in d_grid), then the local GGA sweep
    k_gga_grad<<<blocks, RED_THREADS, 0, st>>>(g, p->d_grid, w[0], w[1], w[2]);
    JRB_CHECK_LAUNCH("k_gga_grad");
    if ((rc = launch_fft3d_dense(p, w[0], w[0], JRB_FFT_INVERSE, 3, 1.0, st))) return rc;
    k_gga_local<<<blocks, RED_THREADS, 0, st>>>(g, rho, xc_id, kohn_sham, as_eps ? 1 : 0, w[0], w[1],
                                               w[2], p->d_vxc, p->d_partials);
    JRB_CHECK_LAUNCH("k_gga_local");
    if (!as_eps)
      if ((rc = launch_fft3d_dense(p, w[0], w[0], JRB_FFT_FORWARD, 3, 1.0, st))) return rc;
  }
  k_hartree_ext<<<blocks, RED_THREADS, 0, st>>>(g, p->d_grid, p->d_vext, kohn_sham, parts,
                                               p->d_partials);
  JRB_CHECK_LAUNCH("k_hartree_ext");
  if (gga_on && !as_eps) {
    k_gga_add<<<blocks, RED_THREADS, 0, st>>>(g, w[0], w[1], w[2], p->d_grid);
    JRB_CHECK_LAUNCH("k_gga_add");
  }
  rc = launch_fft3d_dense(p, p->d_grid, p->d_grid, JRB_FFT_INVERSE, 1, 1.0 / (double)p->ngrid, st);
  if (rc) return rc;
  k_veff_xc<<<blocks, RED_THREADS, 0, st>>>(g, p->d_grid, rho, p->ns, xc_id, kohn_sham, parts,
                                           gga ? p->d_vxc : nullptr, veff, p->d_partials);
  JRB_CHECK_LAUNCH("k_veff_xc");
  if (energies) {
    k_reduce_partials<<<3, 256, 0, st>>>(p->d_partials, blocks, 3, energies);
    JRB_CHECK_LAUNCH("k_reduce_partials");
  }
  return 0;
}

// out[s] = complex(rho[s]) for every spin, then forward FFT: pw.density_grid_reciprocal
__global__ void k_real_to_complex(const double* __restrict__ in, long long n, cplx* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = cmake(in[i], 0.0);
}

int launch_real_to_complex(const double* in, long long n, cplx* out, cudaStream_t st) {
  const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
  k_real_to_complex<<<blocks, 256, 0, st>>>(in, n, out);
  JRB_CHECK_LAUNCH("k_real_to_complex");
  return 0;
}

__global__ void k_complex_to_real(const cplx* __restrict__ in, long long n, double* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = in[i].x;
}

int launch_complex_to_real(const cplx* in, long long n, double* out, cudaStream_t st) {
  const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
  k_complex_to_real<<<blocks, 256, 0, st>>>(in, n, out);
  JRB_CHECK_LAUNCH("k_complex_to_real");
  return 0;
}

// Fourier resampling between two FFT boxes: dst(f) = scale * src(f) for the frequencies both boxes
// represent symmetrically (2|f_c| < min(n_src, n_dst) on every axis: an even box's unpaired Nyquist
// bin is dropped so real fields stay real), zero elsewhere.  One thread per dst element.
__global__ void k_resample(const cplx* __restrict__ src, const cplx* __restrict__ src2, int sx,
                           int sy, int sz, cplx* __restrict__ dst, int dx, int dy, int dz,
                           double scale) {
  const long long n = (long long)dx * dy * dz;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int z = (int)(i % dz), y = (int)((i / dz) % dy), x = (int)(i / ((long long)dz * dy));
    // an axis both boxes share is copied bin by bin (Nyquist included: a grid the caller chose too
    // coarse aliases there exactly as the reference does); a resized axis keeps 2|f| < min(n)
    const int fx = fftfreq_int(x, dx), fy = fftfreq_int(y, dy), fz = fftfreq_int(z, dz);
    const int mx = min(sx, dx), my = min(sy, dy), mz = min(sz, dz);
    const bool okx = sx == dx || 2 * abs(fx) < mx, oky = sy == dy || 2 * abs(fy) < my,
               okz = sz == dz || 2 * abs(fz) < mz;
    cplx v = cmake(0.0, 0.0);
    if (okx && oky && okz) {
      const int px = sx == dx ? x : (fx >= 0 ? fx : fx + sx);
      const int py = sy == dy ? y : (fy >= 0 ? fy : fy + sy);
      const int pz = sz == dz ? z : (fz >= 0 ? fz : fz + sz);
      const long long o = ((long long)px * sy + py) * sz + pz;
      cplx t = src[o];
      if (src2) {  // sum of two fields of the source box
        const cplx u = src2[o];
        t = cmake(t.x + u.x, t.y + u.y);
      }
      v = cmake(t.x * scale, t.y * scale);
    }
    dst[i] = v;
  }
}

int launch_resample(const cplx* src, int sx, int sy, int sz, cplx* dst, int dx, int dy, int dz,
                    double scale, cudaStream_t st) {
  const long long n = (long long)dx * dy * dz;
  const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
  k_resample<<<blocks, 256, 0, st>>>(src, nullptr, sx, sy, sz, dst, dx, dy, dz, scale);
  JRB_CHECK_LAUNCH("k_resample");
  return 0;
}

// ---------------------------------------------------------------------------------------
// The grid part of one evaluation when the orbitals live on their own box (LDA, one spin).  The
// separate calls (launch_density_end -> launch_grid_potential -> launch_hpsi_prepare) transform
// the same fields back and forth: rho goes to the plan's grid (inverse FFT) and is transformed
// forward again for the Hartree term; v_H + v_ext goes to real space, gets v_xc added and is
// transformed forward again to be truncated onto the orbital box.  Here each field is transformed
// once: rho_hat comes from the interpolation itself, and v_H(G) + v_ext(G) joins fftn(v_xc) in
// G space inside the truncating resample (linearity) -- 4 dense single-grid FFTs instead of 6
// (plus the 2 on the orbital box) and 6 launches fewer.  This replicated work is what a rank of an
// 8-GPU run cannot shrink.
bool grid_fused_ok(const jrb_plan* p, int xc_id) {
  static int off = [] {
    const char* env = std::getenv("JRB_NO_GRID_FUSE");
    return env ? std::atoi(env) : 0;
  }();
  return !off && p->wf && p->ns == 1 && p->natoms > 0 &&
         (xc_id == JRB_XC_LDA_X || xc_id == JRB_XC_LDA_X_C_PW);
}

// rho on the orbital box -> rho on the plan's grid AND rho_hat = fftn(rho) left in p->d_grid
int launch_density_end_hat(jrb_plan* p, double* rho, cudaStream_t st) {
  jrb_plan* w = p->wf;
  int rc = 0;
  if ((rc = launch_real_to_complex(w->d_rho_w, w->ngrid, w->d_grid, st))) return rc;
  if ((rc = launch_fft3d_dense(w, w->d_grid, w->d_grid, JRB_FFT_FORWARD, 1, 1.0, st))) return rc;
  if ((rc = launch_resample(w->d_grid, w->nx, w->ny, w->nz, p->d_grid, p->nx, p->ny, p->nz,
                            (double)p->ngrid / (double)w->ngrid, st)))
    return rc;
  cplx* tmp = p->d_gga + p->ngrid;  // out of place: rho_hat stays in d_grid
  if ((rc = launch_fft3d_dense(p, p->d_grid, tmp, JRB_FFT_INVERSE, 1, 1.0 / (double)p->ngrid, st)))
    return rc;
  return launch_complex_to_real(tmp, p->ngrid, rho, st);
}

// complex(v_xc(r)) and the E_xc partial sums (LDA, one spin; v_xc = eps + rho eps')
__global__ void __launch_bounds__(RED_THREADS)
k_vxc_to_complex(GridGeom g, const double* __restrict__ rho, int xc_id, cplx* __restrict__ out,
                 double* __restrict__ partials) {
  double acc[1] = {0.0};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < g.n;
       i += (long long)gridDim.x * blockDim.x) {
    const double n = rho[i];
    double e, de;
    lda_eps(xc_id, n, e, de);
    out[i] = cmake(e + n * de, 0.0);
    acc[0] += e * n;
  }
  double sum[1];
  block_sum<1>(acc, sum);
  if (threadIdx.x == 0) partials[blockIdx.x * 4 + 2] = sum[0] * g.vol / (double)g.n;
}

// energies[3] = E_H, E_ext, E_xc from rho (plan grid) and v_eff = dE/drho ON THE ORBITAL BOX
// (p->wf->d_veff); rhohat_ready: p->d_grid already holds fftn(rho) (launch_density_end_hat)
int launch_grid_potential_orbital(jrb_plan* p, const double* rho, bool rhohat_ready, int xc_id,
                                  double* energies, cudaStream_t st) {
  jrb_plan* w = p->wf;
  const GridGeom g = geom_of(p);
  const int blocks = (int)std::min<long long>((p->ngrid + RED_THREADS - 1) / RED_THREADS,
                                              (long long)p->n_partial_blocks);
  int rc = 0;
  if (!rhohat_ready) {
    k_rho_to_complex<<<blocks, RED_THREADS, 0, st>>>(rho, 1, p->ngrid, p->d_grid);
    JRB_CHECK_LAUNCH("k_rho_to_complex");
    if ((rc = launch_fft3d_dense(p, p->d_grid, p->d_grid, JRB_FFT_FORWARD, 1, 1.0, st))) return rc;
  }
  k_hartree_ext<<<blocks, RED_THREADS, 0, st>>>(g, p->d_grid, p->d_vext, 0, 3, p->d_partials);
  JRB_CHECK_LAUNCH("k_hartree_ext");
  cplx* vx = p->d_gga;
  k_vxc_to_complex<<<blocks, RED_THREADS, 0, st>>>(g, rho, xc_id, vx, p->d_partials);
  JRB_CHECK_LAUNCH("k_vxc_to_complex");
  if ((rc = launch_fft3d_dense(p, vx, vx, JRB_FFT_FORWARD, 1, 1.0, st))) return rc;
  // fftn(v_eff) = fftn(v_xc) + v_H(G) + v_ext(G), truncated onto the orbital box
  {
    const long long n = w->ngrid;
    const int rb = (int)std::min<long long>((n + 255) / 256, 148 * 8);
    k_resample<<<rb, 256, 0, st>>>(vx, p->d_grid, p->nx, p->ny, p->nz, w->d_grid, w->nx, w->ny, w->nz,
                                   1.0 / (double)p->ngrid);
    JRB_CHECK_LAUNCH("k_resample");
  }
  if ((rc = launch_fft3d_dense(w, w->d_grid, w->d_grid, JRB_FFT_INVERSE, 1, 1.0, st))) return rc;
  if ((rc = launch_complex_to_real(w->d_grid, w->ngrid, w->d_veff, st))) return rc;
  if (energies) {
    k_reduce_partials<<<3, 256, 0, st>>>(p->d_partials, blocks, 3, energies);
    JRB_CHECK_LAUNCH("k_reduce_partials");
  }
  return 0;
}

int launch_density_reciprocal(jrb_plan* p, const double* rho, cplx* rho_hat, cudaStream_t st) {
  const long long n = (long long)p->ns * p->ngrid;
  const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
  k_real_to_complex<<<blocks, 256, 0, st>>>(rho, n, rho_hat);
  JRB_CHECK_LAUNCH("k_real_to_complex");
  return launch_fft3d_dense(p, rho_hat, rho_hat, JRB_FFT_FORWARD, p->ns, 1.0, st);
}

// ---------------------------------------------------------------------------------------
// sphere reductions: out[sk][b] = sum_g w(g) * f(q, hq)
//   MODE 0: 1/2 |G+k|^2 |q|^2        (kinetic)
//   MODE 1: Re conj(q) hq            (band expectation)
// The rows g are split over gridDim.z chunks (enough CTAs to stream Q at HBM speed); the chunk
// partials are summed in a fixed order by k_sum_chunks (deterministic).
// grid: (ceil(nb / 32), nsk, nchunks), block (32, 8)
constexpr int SPHERE_CHUNKS = 8;
template <int MODE>
__global__ void __launch_bounds__(256)
k_sphere_reduce(const cplx* __restrict__ q, const cplx* __restrict__ hq,
                const double* __restrict__ gk2, long long ng, int nb, int nk, int sk0,
                double* __restrict__ part) {
  const int b = blockIdx.x * 32 + threadIdx.x;
  const int sk = sk0 + blockIdx.y;
  const int k = sk % nk;
  const long long rows = (ng + gridDim.z - 1) / gridDim.z;
  const long long g_lo = blockIdx.z * rows, g_hi = min(ng, g_lo + rows);
  double acc = 0.0;
  if (b < nb) {
    const cplx* qp = q + ((long long)sk * ng) * nb + b;
    const cplx* hp = MODE == 1 ? hq + ((long long)sk * ng) * nb + b : nullptr;
    for (long long g = g_lo + threadIdx.y; g < g_hi; g += blockDim.y) {
      const cplx c = qp[g * nb];
      if (MODE == 0) {
        acc += gk2[(long long)k * ng + g] * (c.x * c.x + c.y * c.y);
      } else {
        const cplx h = hp[g * nb];
        acc += c.x * h.x + c.y * h.y;
      }
    }
  }
  __shared__ double sh[8][33];
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && b < nb) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += sh[j][threadIdx.x];
    // part[chunk][local sk][b]
    part[((long long)blockIdx.z * gridDim.y + blockIdx.y) * nb + b] = MODE == 0 ? 0.5 * s : s;
  }
}

// out[i] = sum_chunks part[chunk][i]
__global__ void k_sum_chunks(const double* __restrict__ part, int nchunks, long long n,
                             double* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int c = 0; c < nchunks; ++c) s += part[c * n + i];
  out[i] = s;
}

template <int MODE>
static int sphere_reduce(jrb_plan* p, int sk0, int nsk, const cplx* q, const cplx* hq, double* out,
                         cudaStream_t st) {
  // few (spin, k) items (Gamma-only supercells): more row chunks so that ~4 CTAs per SM stream Q
  // (d_sphere_part holds (8 ns nk + 640) nb partials, plan.cu)
  const int nbx = (p->nb + 31) / 32;
  const int chunks = std::max(SPHERE_CHUNKS, std::min(64, 592 / (nbx * nsk)));
  dim3 grid(nbx, nsk, chunks), block(32, 8);
  k_sphere_reduce<MODE><<<grid, block, 0, st>>>(q, hq, p->d_gk2, p->ng, p->nb, p->nk, sk0,
                                               p->d_sphere_part);
  JRB_CHECK_LAUNCH("k_sphere_reduce");
  const long long n = (long long)nsk * p->nb;
  k_sum_chunks<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p->d_sphere_part, chunks, n,
                                                            out + (long long)sk0 * p->nb);
  JRB_CHECK_LAUNCH("k_sum_chunks");
  return 0;
}

int launch_kinetic_range(jrb_plan* p, int sk0, int nsk, const cplx* q, double* t_skb,
                         cudaStream_t st) {
  return sphere_reduce<0>(p, sk0, nsk, q, nullptr, t_skb, st);
}

int launch_kinetic(jrb_plan* p, const cplx* q, double* t_skb, cudaStream_t st) {
  return launch_kinetic_range(p, 0, p->ns * p->nk, q, t_skb, st);
}

int launch_band_expect(jrb_plan* p, const cplx* q, const cplx* hq, double* eps, cudaStream_t st) {
  return sphere_reduce<1>(p, 0, p->ns * p->nk, q, hq, eps, st);
}

// out[0] = sum_i a[i] w[i]   (single block, fixed order)
__global__ void __launch_bounds__(RED_THREADS)
k_weighted_sum(const double* __restrict__ a, const double* __restrict__ w, long long n,
               double* __restrict__ out) {
  double acc[1] = {0.0};
  for (long long i = threadIdx.x; i < n; i += blockDim.x) acc[0] += a[i] * w[i];
  double r[1];
  block_sum<1>(acc, r);
  if (threadIdx.x == 0) out[0] = r[0];
}

int launch_weighted_sum(jrb_plan* p, const double* a, const double* w, int64_t n, double* out,
                        cudaStream_t st) {
  (void)p;
  k_weighted_sum<<<1, RED_THREADS, 0, st>>>(a, w, n, out);
  JRB_CHECK_LAUNCH("k_weighted_sum");
  return 0;
}

// gk2[g] = |G_g + k|^2 on the sphere for one k-point (same arithmetic order as the host table of
// jrb_plan_create: G from integer frequencies, then the sum of squares over x, y, z)
__global__ void k_gk2(GridGeom geo, const int32_t* __restrict__ gidx, long long ng, double kx,
                      double ky, double kz, double* __restrict__ gk2) {
  const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (g >= ng) return;
  double gx, gy, gz;
  g_of(geo, gidx[g], gx, gy, gz);
  const double vx = gx + kx, vy = gy + ky, vz = gz + kz;
  double s = 0.0;
  s += vx * vx;
  s += vy * vy;
  s += vz * vz;
  gk2[g] = s;
}

int launch_set_kpoints(jrb_plan* p, const double* kpts_h, cudaStream_t st) {
  const GridGeom g = geom_of(p);
  for (int k = 0; k < p->nk; ++k) {
    k_gk2<<<(unsigned)((p->ng + 255) / 256), 256, 0, st>>>(
      g, p->d_gidx, p->ng, kpts_h[3 * k], kpts_h[3 * k + 1], kpts_h[3 * k + 2],
      p->d_gk2 + (long long)k * p->ng);
    JRB_CHECK_LAUNCH("k_gk2");
  }
  if (p->h_kpts) std::copy(kpts_h, kpts_h + 3 * (size_t)p->nk, p->h_kpts);
  if (p->wf) return launch_set_kpoints(p->wf, kpts_h, st);  // its z pass adds 1/2|G+k|^2 q
  return 0;
}

// focc[group][lane] = occ[sk][b0 + lane] / Omega, zero padded
__global__ void k_focc(const double* __restrict__ occ, int nb, int ngpk, int total_groups,
                       double inv_vol, double* __restrict__ focc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_groups * NB) return;
  const int gid = i / NB, lane = i % NB;
  const int sk = gid / ngpk, b = (gid % ngpk) * NB + lane;
  focc[i] = b < nb ? occ[(long long)sk * nb + b] * inv_vol : 0.0;
}

int launch_focc(jrb_plan* p, const double* occ, cudaStream_t st) {
  const int total = p->ns * p->nk * p->ngroups_per_k;
  k_focc<<<(total * NB + 255) / 256, 256, 0, st>>>(occ, p->nb, p->ngroups_per_k, total,
                                                  1.0 / p->vol, p->d_focc);
  JRB_CHECK_LAUNCH("k_focc");
  if (p->wf) return launch_focc(p->wf, occ, st);  // the plan that runs the density sweep
  return 0;
}

// dense[(sk*nb + b)*N + gidx[g]] = q[(sk*ng + g)*nb + b]   (dense pre-zeroed)
__global__ void k_expand(const cplx* __restrict__ q, const int32_t* __restrict__ gidx,
                         long long ng, int nb, long long n, long long total,
                         cplx* __restrict__ dense) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i % nb);
    const long long g = (i / nb) % ng;
    const long long sk = i / ((long long)nb * ng);
    dense[(sk * nb + b) * n + gidx[g]] = q[i];
  }
}

__global__ void k_squeeze(const cplx* __restrict__ dense, const int32_t* __restrict__ gidx,
                          long long ng, int nb, long long n, long long total,
                          cplx* __restrict__ q) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i % nb);
    const long long g = (i / nb) % ng;
    const long long sk = i / ((long long)nb * ng);
    q[i] = dense[(sk * nb + b) * n + gidx[g]];
  }
}

int launch_expand(jrb_plan* p, const cplx* q, cplx* dense, cudaStream_t st) {
  const long long total = (long long)p->ns * p->nk * p->ng * p->nb;
  JRB_CUDA(cudaMemsetAsync(dense, 0, sizeof(cplx) * (size_t)p->ns * p->nk * p->nb * p->ngrid, st));
  const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  k_expand<<<blocks, 256, 0, st>>>(q, p->d_gidx, p->ng, p->nb, p->ngrid, total, dense);
  JRB_CHECK_LAUNCH("k_expand");
  return 0;
}

int launch_squeeze(jrb_plan* p, const cplx* dense, cplx* q, cudaStream_t st) {
  const long long total = (long long)p->ns * p->nk * p->ng * p->nb;
  const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  k_squeeze<<<blocks, 256, 0, st>>>(dense, p->d_gidx, p->ng, p->nb, p->ngrid, total, q);
  JRB_CHECK_LAUNCH("k_squeeze");
  return 0;
}

}  // namespace jrb
