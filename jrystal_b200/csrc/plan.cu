// Plan construction: mask -> pruning index maps, |G+k|^2 tables, twiddles, work space.
// Host-side restatement of what the reference's drivers prepare before the loop
// (jrystal/calc/calc_ground_state_energy_all_electrons.py:93-106 via grid.g_vectors,
// jrystal/_src/grid.py:93-149, and the C-order mask enumeration of utils.py:279-281).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fft_passes.cuh"
#include <cuda.h>

#include "plan.h"

namespace jrb {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char* what) {
  g_last_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
  return JRB_ECUDA;
}

const char* last_error_cstr() { return g_last_error.c_str(); }

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

bool line_length_supported(int n) {
  switch (n) {
    case 7: case 8: case 9: case 12: case 16: case 24: case 32: case 48: case 64: case 72:
    case 96: case 128:
    // pencil passes and dense lines only (no fused y+x kernels): lengths an orbital grid can take
    case 36: case 40: case 45: case 49: case 50: case 54: case 56: case 60: case 80: case 81: case 90:
    case 100: case 112:
      return true;
    default:
      return false;
  }
}

// np.fft.fftfreq(n, 1/n): 0..ceil(n/2)-1, -floor(n/2)..-1   (jrystal/_src/grid.py:115-117)
static inline int fftfreq_int(int i, int n) { return i < (n + 1) / 2 ? i : i - n; }

static void invert3x3(const double* a, double* inv, double* det_out) {
  const double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) +
                     a[2] * (a[3] * a[7] - a[4] * a[6]);
  *det_out = det;
  inv[0] = (a[4] * a[8] - a[5] * a[7]) / det;
  inv[1] = (a[2] * a[7] - a[1] * a[8]) / det;
  inv[2] = (a[1] * a[5] - a[2] * a[4]) / det;
  inv[3] = (a[5] * a[6] - a[3] * a[8]) / det;
  inv[4] = (a[0] * a[8] - a[2] * a[6]) / det;
  inv[5] = (a[2] * a[3] - a[0] * a[5]) / det;
  inv[6] = (a[3] * a[7] - a[4] * a[6]) / det;
  inv[7] = (a[1] * a[6] - a[0] * a[7]) / det;
  inv[8] = (a[0] * a[4] - a[1] * a[3]) / det;
}

template <class T>
static int dev_alloc(T** p, size_t count, int64_t* total) {
  *p = nullptr;
  if (count == 0) count = 1;
  cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
  if (e != cudaSuccess) {
    set_error(std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
    return JRB_ENOMEM;
  }
  *total += (int64_t)(count * sizeof(T));
  return 0;
}

template <class T>
static int upload(T** dptr, const std::vector<T>& h, int64_t* total) {
  int rc = dev_alloc(dptr, h.size(), total);
  if (rc) return rc;
  if (!h.empty()) JRB_CUDA(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

static std::vector<cplx> twiddle_table(int n) {
  std::vector<cplx> t(n);
  for (int i = 0; i < n; ++i) {
    // exact quadrant reduction keeps the table accurate to < 1 ulp
    const long double ang = 2.0L * 3.141592653589793238462643383279502884L * (long double)i / n;
    t[i] = cmake((double)cosl(ang), (double)(-sinl(ang)));
    if ((4 * i) % n == 0) {
      const int q = (4 * i) / n;
      const double c[4] = {1, 0, -1, 0}, s[4] = {0, -1, 0, 1};
      t[i] = cmake(c[q], s[q]);
    }
  }
  return t;
}

}  // namespace jrb

using namespace jrb;

extern "C" const char* jrb_last_error(void) { return jrb::last_error_cstr(); }
extern "C" int jrb_version(void) { return 100; }
namespace jrb { long long launch_count(); }
extern "C" int64_t jrb_launch_count(void) { return jrb::launch_count(); }

extern "C" int64_t jrb_plan_num_g(const jrb_plan* p) { return p ? p->ng : -1; }
extern "C" int64_t jrb_plan_workspace_bytes(const jrb_plan* p) { return p ? p->ws_bytes : -1; }


// ---- TMA tensor maps of the column work space ------------------------------------------------
// The fused y+x kernels stage one band of one z-plane: the 16-byte entries [col][band] of the
// plane's [ncol][8 bands] block.  As a 2-D FP64 tensor {16 doubles, rows} with a 128-byte row pitch
// that is the box {2 doubles, 256 rows}, which cp.async.bulk.tensor.2d lands as 256 consecutive
// complex numbers.  cuTensorMapEncodeTiled comes from the driver through the runtime's entry-point
// query, so the library still links libcudart only.
typedef CUresult (*jrb_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                        const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool encode_column_map(void* base, unsigned long long rows, CUtensorMap* out) {
  static jrb_encode_tiled_fn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<jrb_encode_tiled_fn>(f);
  }();
  if (!fn || !base || rows == 0 || rows >= (1ull << 31)) return false;
  const cuuint64_t dims[2] = {16, rows};
  const cuuint64_t strides[1] = {128};  // bytes between rows
  const cuuint32_t box[2] = {2, 256};
  const cuuint32_t estr[2] = {1, 1};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// d_tmaps[0]: d_ws_a, d_tmaps[1]: d_a_keep (fused == 1 plans; JRB_NO_TMA=1 keeps cp.async)
static int make_tensor_maps(jrb_plan* p, long long ws_rows, long long keep_rows) {
  if (const char* env = std::getenv("JRB_NO_TMA"))
    if (std::atoi(env) != 0) return 0;
  if (p->fused != 1) return 0;
  alignas(64) CUtensorMap maps[2];
  std::memset(maps, 0, sizeof(maps));
  if (!encode_column_map(p->d_ws_a, (unsigned long long)ws_rows, &maps[0])) return 0;
  if (p->d_a_keep && !encode_column_map(p->d_a_keep, (unsigned long long)keep_rows, &maps[1])) return 0;
  JRB_CUDA(cudaMalloc(&p->d_tmaps, sizeof(maps)));
  JRB_CUDA(cudaMemcpy(p->d_tmaps, maps, sizeof(maps), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int jrb_plan_destroy(jrb_plan* p) {
  if (!p) return 0;
  cudaSetDevice(p->device);
  if (p->wf) jrb_plan_destroy(p->wf);
  p->wf = nullptr;
  comm_destroy(p);
  if (p->d_rho_w) cudaFree(p->d_rho_w);
  if (p->d_tmaps) cudaFree(p->d_tmaps);
  delete[] p->h_freq;
  delete[] p->h_kpts;
  void* ptrs[] = {p->d_zmap, p->d_ycol, p->d_xmap, p->d_gidx, p->d_gk2, p->d_tw_x, p->d_tw_y,
                  p->d_tw_z, p->d_tw_half, p->d_a_keep, p->d_psi, p->d_pos, p->d_chg, p->d_atom_part, p->d_nl_phi, p->d_nl_p, p->d_nl_part, p->d_nl_phit, p->d_nl_ps, p->d_ws_a, p->d_ws_b, p->d_rho_part, p->d_seg_z, p->d_focc, p->d_grid, p->d_vext,
                  p->d_partials, p->d_veff, p->d_gga, p->d_vxc, p->d_q, p->d_hq, p->d_tmp, p->d_r, p->d_rinv, p->d_small, p->d_gpart,
                  p->d_tkb, p->d_eps, p->d_sphere_part, p->d_scal, p->d_skip, p->d_emax, p->d_wre, p->d_wim, p->d_gre, p->d_gim,
                  p->d_occ, p->d_rho, p->d_en};
  for (void* q : ptrs)
    if (q) cudaFree(q);
  if (p->own_stream) cudaStreamDestroy(p->own_stream);
  if (p->h2d_stream) cudaStreamDestroy(p->h2d_stream);
  if (p->d2h_stream) cudaStreamDestroy(p->d2h_stream);
  for (int i = 0; i < 16; ++i) {
    if (i < 8 && p->ev_phase[i]) cudaEventDestroy(p->ev_phase[i]);
    if (p->ev_in[i]) cudaEventDestroy(p->ev_in[i]);
    if (p->ev_out[i]) cudaEventDestroy(p->ev_out[i]);
  }
  delete p;
  return 0;
}

// QR-only plan over `nrows` rows of the coefficient matrices (row-sharded orthonormalisation of
// Gamma-only supercells): no grid, no index maps; only the jrb_qr_rows_* entry
// points accept it.
extern "C" int jrb_plan_create_rows(int64_t nrows, int32_t ns, int32_t nk, int32_t nb,
                                    int32_t device, jrb_plan** out) {
  if (!out || nrows <= 0 || nk <= 0 || nb <= 0 || (ns != 1 && ns != 2)) {
    set_error("jrb_plan_create_rows: bad sizes (need nrows,nk,nb > 0 and ns in {1,2})");
    return JRB_EINVAL;
  }
  int ndev = 0;
  JRB_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) {
    set_error("jrb_plan_create_rows: bad device ordinal");
    return JRB_EINVAL;
  }
  JRB_CUDA(cudaSetDevice(device));
  jrb_plan* p = new jrb_plan();
  std::memset(p, 0, sizeof(*p));
  p->ns = ns; p->nk = nk; p->nb = nb;
  p->ng = nrows;
  p->device = device;
  p->ngroups_per_k = (nb + NB - 1) / NB;
  int rc = 0;
  int64_t tot = 0;
#define TRY(x)                 \
  do {                         \
    rc = (x);                  \
    if (rc) {                  \
      jrb_plan_destroy(p);     \
      return rc;               \
    }                          \
  } while (0)
  const size_t nsphere = (size_t)ns * nk * nrows * nb;
  const size_t nsmall = (size_t)ns * nk * nb * nb;
  TRY(dev_alloc(&p->d_tmp, nsphere, &tot));
  TRY(dev_alloc(&p->d_r, nsmall, &tot));
  TRY(dev_alloc(&p->d_rinv, nsmall, &tot));
  TRY(dev_alloc(&p->d_small, nsmall * 5, &tot));
  TRY(dev_alloc(&p->d_gpart, (size_t)nb * nb * (size_t)qr_gram_partial_mats(p), &tot));
  TRY(dev_alloc(&p->d_scal, 64, &tot));
  JRB_CUDA(cudaMemset(p->d_scal, 0, 64 * sizeof(double)));
  TRY(dev_alloc(&p->d_skip, (size_t)p->ns * p->nk, &tot));
  TRY(dev_alloc(&p->d_emax, (size_t)p->ns * p->nk, &tot));
#undef TRY
  p->ws_bytes = tot;
  *out = p;
  return 0;
}

static int plan_create_impl(const jrb_plan_desc* d, bool orbital_only, jrb_plan** out);

extern "C" int jrb_plan_create(const jrb_plan_desc* d, jrb_plan** out) {
  return plan_create_impl(d, false, out);
}

static int plan_create_impl(const jrb_plan_desc* d, bool orbital_only, jrb_plan** out) {
  if (!d || !out || !d->mask || !d->kpts || !d->cell) {
    set_error("jrb_plan_create: null argument");
    return JRB_EINVAL;
  }
  if (d->nx <= 0 || d->ny <= 0 || d->nz <= 0 || d->nk <= 0 || d->nb <= 0 ||
      (d->ns != 1 && d->ns != 2)) {
    set_error("jrb_plan_create: bad sizes (need nx,ny,nz,nk,nb > 0 and ns in {1,2})");
    return JRB_EINVAL;
  }
  const int dims[3] = {d->nx, d->ny, d->nz};
  for (int i = 0; i < 3; ++i) {
    if (!line_length_supported(dims[i])) {
      set_error("jrb_plan_create: FFT length " + std::to_string(dims[i]) +
                " has no compiled line plan (supported: 7 8 9 12 16 24 32 36 40 45 48 49 50 54 56 60 "
                "64 72 80 81 90 96 100 112 128)");
      return JRB_EUNSUPPORTED;
    }
  }
  int ndev = 0;
  JRB_CUDA(cudaGetDeviceCount(&ndev));
  if (d->device < 0 || d->device >= ndev) {
    set_error("jrb_plan_create: bad device ordinal");
    return JRB_EINVAL;
  }
  JRB_CUDA(cudaSetDevice(d->device));

  jrb_plan* p = new jrb_plan();
  std::memset(p, 0, sizeof(*p));
  p->nx = d->nx; p->ny = d->ny; p->nz = d->nz;
  p->ns = d->ns; p->nk = d->nk; p->nb = d->nb;
  p->device = d->device;
  p->orbital_only = orbital_only ? 1 : 0;
  p->ngrid = (int64_t)d->nx * d->ny * d->nz;
  p->ngroups_per_k = (d->nb + NB - 1) / NB;
  std::memcpy(p->cell, d->cell, sizeof(double) * 9);
  double inv[9], det;
  invert3x3(p->cell, inv, &det);
  p->vol = std::fabs(det);
  if (!(p->vol > 0)) {
    delete p;
    set_error("jrb_plan_create: singular cell");
    return JRB_EINVAL;
  }
  // B = 2 pi inv(A)^T, rows b_i
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) p->recip[i * 3 + j] = 2.0 * M_PI * inv[j * 3 + i];

  const int nx = p->nx, ny = p->ny, nz = p->nz;
  // --- index maps -------------------------------------------------------------------
  std::vector<int32_t> colid((size_t)nx * ny, -1), xmap(nx, -1), gidx;
  std::vector<int32_t> zmap, ycol;
  int ncol = 0, nxo = 0;
  int64_t ng = 0;
  for (int x = 0; x < nx; ++x) {
    bool any_x = false;
    for (int y = 0; y < ny; ++y) {
      bool any = false;
      for (int z = 0; z < nz; ++z)
        if (d->mask[((size_t)x * ny + y) * nz + z]) { any = true; break; }
      if (any) { colid[(size_t)x * ny + y] = ncol++; any_x = true; }
    }
    if (any_x) xmap[x] = nxo++;
  }
  zmap.assign((size_t)std::max(ncol, 1) * nz, -1);
  ycol.assign((size_t)std::max(nxo, 1) * ny, -1);
  for (int x = 0; x < nx; ++x)
    for (int y = 0; y < ny; ++y) {
      const int c = colid[(size_t)x * ny + y];
      if (c < 0) continue;
      ycol[(size_t)xmap[x] * ny + y] = c;
      for (int z = 0; z < nz; ++z)
        if (d->mask[((size_t)x * ny + y) * nz + z]) {
          zmap[(size_t)c * nz + z] = (int32_t)ng++;
          gidx.push_back((int32_t)(((size_t)x * ny + y) * nz + z));
        }
    }
  // band-limited in x and y: |frequency| < 16 (what the sparse radix-8 butterflies assume)
  p->band_limited = 1;
  for (int x = 0; x < nx; ++x)
    for (int y = 0; y < ny; ++y)
      if (colid[(size_t)x * ny + y] >= 0 &&
          !((x < 16 || x >= nx - 16) && (y < 16 || y >= ny - 16)))
        p->band_limited = 0;
  p->band_limited32 = 1;
  for (int x = 0; x < nx; ++x)
    for (int y = 0; y < ny; ++y)
      if (colid[(size_t)x * ny + y] >= 0 &&
          !((x < 32 || x >= nx - 32) && (y < 32 || y >= ny - 32)))
        p->band_limited32 = 0;
  if (ng == 0) {
    delete p;
    set_error("jrb_plan_create: empty mask");
    return JRB_EINVAL;
  }
  p->ng = ng;
  p->h_freq = new int32_t[(size_t)ng * 3];
  p->h_kpts = new double[(size_t)p->nk * 3];
  std::memcpy(p->h_kpts, d->kpts, sizeof(double) * 3 * (size_t)p->nk);
  for (int64_t g = 0; g < ng; ++g) {
    const int lin = gidx[g];
    const int f[3] = {fftfreq_int(lin / (nz * ny), nx), fftfreq_int((lin / nz) % ny, ny),
                      fftfreq_int(lin % nz, nz)};
    for (int c = 0; c < 3; ++c) {
      p->h_freq[g * 3 + c] = f[c];
      p->gmax[c] = std::max(p->gmax[c], std::abs(f[c]));
    }
  }
  // --- |G+k|^2 on the sphere --------------------------------------------------------
  std::vector<double> gk2((size_t)p->nk * ng);
  {
    std::vector<double> gv((size_t)ng * 3);
    for (int64_t g = 0; g < ng; ++g) {
      const int lin = gidx[g];
      const int z = lin % nz, y = (lin / nz) % ny, x = lin / (nz * ny);
      const double fx = fftfreq_int(x, nx), fy = fftfreq_int(y, ny), fz = fftfreq_int(z, nz);
      for (int c = 0; c < 3; ++c)
        gv[g * 3 + c] = fx * p->recip[0 + c] + fy * p->recip[3 + c] + fz * p->recip[6 + c];
    }
    for (int k = 0; k < p->nk; ++k)
      for (int64_t g = 0; g < ng; ++g) {
        double s = 0;
        for (int c = 0; c < 3; ++c) {
          const double v = gv[g * 3 + c] + d->kpts[k * 3 + c];
          s += v * v;
        }
        gk2[(size_t)k * ng + g] = s;
      }
  }

  int rc = 0;
  int64_t tot = 0;
#define TRY(x)                 \
  do {                         \
    rc = (x);                  \
    if (rc) {                  \
      jrb_plan_destroy(p);     \
      return rc;               \
    }                          \
  } while (0)
  TRY(upload(&p->d_zmap, zmap, &tot));
  TRY(upload(&p->d_ycol, ycol, &tot));
  TRY(upload(&p->d_xmap, xmap, &tot));
  TRY(upload(&p->d_gidx, gidx, &tot));
  TRY(upload(&p->d_gk2, gk2, &tot));
  TRY(upload(&p->d_tw_x, twiddle_table(nx), &tot));
  TRY(upload(&p->d_tw_y, twiddle_table(ny), &tot));
  TRY(upload(&p->d_tw_z, twiddle_table(nz), &tot));
  TRY(upload(&p->d_tw_half, twiddle_table(nx % 2 == 0 ? nx / 2 : nx), &tot));
  p->maps.nx = nx; p->maps.ny = ny; p->maps.nz = nz;
  p->maps.ncol = ncol; p->maps.nxo = nxo; p->maps.ng = ng;
  p->maps.zmap = p->d_zmap; p->maps.ycol = p->d_ycol; p->maps.xmap = p->d_xmap;
  p->maps.gidx = p->d_gidx;

  // --- pencil work space ------------------------------------------------------------
  const int total_groups = p->ns * p->nk * p->ngroups_per_k;
  const size_t b_per_group = (size_t)nxo * ny * nz * NB;  // complex numbers
  const size_t a_per_group = (size_t)ncol * nz * NB;
  p->fused = fused_available(nx, ny, nxo, ncol) ? 1 : 0;
  if (!p->fused && fused128_available(nx, ny, nxo, ncol, p->band_limited32)) p->fused = 2;
  int bg = d->batch_groups;
  if (const char* env = std::getenv("JRB_BATCH_GROUPS")) bg = std::atoi(env);
  if (bg <= 0) {
    // automatic: fat launches (the passes are bound by on-chip pipes, not by HBM).  Unfused: up
    // to 128 groups within a 2 GiB B slab; fused (only the column buffer A exists):
    if (p->fused) {
      // whole spin in one launch when the column buffer stays within 4 GiB (measured: C2 84.1 ->
      // 85.9 eval/s against 256-group batches with a 64-group tail), else equal batches
      const int per_spin = p->nk * p->ngroups_per_k;
      int cap = (int)(4096.0 * 1024 * 1024 / ((double)a_per_group * sizeof(cplx)));
      if (cap < 2) cap = 2;
      const int nbatch = (per_spin + cap - 1) / cap;
      bg = (per_spin + nbatch - 1) / nbatch;
    } else {
      bg = (int)(2048.0 * 1024 * 1024 / ((double)b_per_group * sizeof(cplx)));
      if (bg > 128) bg = 128;
    }
    if (bg < 2) bg = 2;
  }
  if (bg > p->nk * p->ngroups_per_k) bg = p->nk * p->ngroups_per_k;  // never straddle a spin
  if (bg > 65535) bg = 65535;
  p->batch_groups = bg;
  // fused kernels: persistent CTAs; a CTA touches at most fused_segmax z-planes per launch
  // (worst case over batch sizes 1..bg: one group per batch)
  p->fused_ctas = p->fused == 2 ? 148 : (p->fused ? fused_cta_count(nx, nxo, ncol) : 0);
  // partial density planes per CTA: whole planes (fused == 1) or y-parity half planes (fused == 2)
  const int nplanes = p->fused == 2 ? 2 * nz : nz;
  p->fused_segmax = p->fused ? (nplanes + p->fused_ctas - 1) / p->fused_ctas + 2 : 0;
  p->a_copy_elems = (long long)(a_per_group * bg);
  p->a_group_elems = (long long)a_per_group;
  // psi(r) cache of a whole evaluation (plan.h: d_psi): HBM3e capacity traded for half of the
  // H-apply's line transforms.  Optional: over budget or out of memory falls back to d_a_keep.
  p->psi_group_elems = 0;
  if (p->fused == 1) {
    double cap_mb = 65536.0;
    if (const char* env = std::getenv("JRB_PSI_CACHE_MB")) cap_mb = std::atof(env);
    const long long plane = fused_psi_plane_elems(nx, nxo);
    const size_t per_group = (size_t)plane * nz * NB;
    const double need_mb = (double)total_groups * per_group * sizeof(cplx) / (1024.0 * 1024.0);
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = 0;
    if (plane > 0 && need_mb <= cap_mb && need_mb * 1024.0 * 1024.0 <= 0.4 * (double)free_b) {
      void* q = nullptr;
      if (cudaMalloc(&q, per_group * total_groups * sizeof(cplx)) == cudaSuccess) {
        p->d_psi = static_cast<cplx*>(q);
        p->psi_group_elems = (long long)per_group;
        tot += (long long)(per_group * total_groups * sizeof(cplx));
      } else {
        cudaGetLastError();  // not an error: the cache is optional
      }
    }
  }
  if (!p->d_psi) {
    double cap_mb = 8192.0;
    if (const char* env = std::getenv("JRB_KEEP_A_MB")) cap_mb = std::atof(env);
    const double need_mb = (double)total_groups * a_per_group * sizeof(cplx) / (1024.0 * 1024.0);
    if (p->fused && need_mb <= cap_mb) TRY(dev_alloc(&p->d_a_keep, a_per_group * total_groups, &tot));
  }
  TRY(dev_alloc(&p->d_ws_a, a_per_group * bg * (p->fused == 2 ? 3 : 1), &tot));
  TRY(make_tensor_maps(p, (long long)bg * nz * ncol, (long long)total_groups * nz * ncol));
  TRY(dev_alloc(&p->d_ws_b, p->fused ? 1 : b_per_group * bg, &tot));
  TRY(dev_alloc(&p->d_rho_part,
                p->fused ? (size_t)p->fused_ctas * p->fused_segmax * nx * ny / (p->fused == 2 ? 2 : 1) : 1,
                &tot));
  TRY(dev_alloc(&p->d_seg_z, p->fused ? (size_t)p->fused_ctas * p->fused_segmax : 1, &tot));
  TRY(dev_alloc(&p->d_focc, (size_t)total_groups * NB, &tot));
  // --- grid work space --------------------------------------------------------------
  TRY(dev_alloc(&p->d_grid, (size_t)p->ngrid, &tot));
  TRY(dev_alloc(&p->d_vext, (size_t)p->ngrid, &tot));
  JRB_CUDA(cudaMemset(p->d_vext, 0, (size_t)p->ngrid * sizeof(cplx)));
  TRY(dev_alloc(&p->d_veff, (size_t)p->ns * p->ngrid, &tot));
  TRY(dev_alloc(&p->d_gga, orbital_only ? 1 : 3 * (size_t)p->ngrid, &tot));
  TRY(dev_alloc(&p->d_vxc, orbital_only ? 1 : (size_t)p->ngrid, &tot));
  p->n_partial_blocks = 1024;
  TRY(dev_alloc(&p->d_partials, (size_t)p->n_partial_blocks * 4, &tot));
  // --- evaluation work space --------------------------------------------------------
  // (a child plan on the orbital grid only runs the pencil passes: no sphere-sized buffers)
  const size_t nsphere = orbital_only ? 1 : (size_t)p->ns * p->nk * ng * p->nb;
  const size_t nsmall = orbital_only ? 1 : (size_t)p->ns * p->nk * p->nb * p->nb;
  if (orbital_only) TRY(dev_alloc(&p->d_rho_w, (size_t)p->ns * p->ngrid, &tot));
  TRY(dev_alloc(&p->d_q, nsphere, &tot));
  TRY(dev_alloc(&p->d_hq, nsphere, &tot));
  TRY(dev_alloc(&p->d_tmp, nsphere, &tot));
  TRY(dev_alloc(&p->d_r, nsmall, &tot));
  TRY(dev_alloc(&p->d_rinv, nsmall, &tot));
  TRY(dev_alloc(&p->d_small, nsmall * 5, &tot));
  TRY(dev_alloc(&p->d_gpart,
                orbital_only ? 1 : (size_t)p->nb * p->nb * (size_t)qr_gram_partial_mats(p), &tot));
  TRY(dev_alloc(&p->d_tkb, (size_t)p->ns * p->nk * p->nb, &tot));
  TRY(dev_alloc(&p->d_eps, (size_t)p->ns * p->nk * p->nb, &tot));
  TRY(dev_alloc(&p->d_sphere_part, ((size_t)8 * p->ns * p->nk + 640) * p->nb, &tot));
  TRY(dev_alloc(&p->d_scal, 64, &tot));
  JRB_CUDA(cudaMemset(p->d_scal, 0, 64 * sizeof(double)));
  TRY(dev_alloc(&p->d_skip, (size_t)p->ns * p->nk, &tot));
  TRY(dev_alloc(&p->d_emax, (size_t)p->ns * p->nk, &tot));
  JRB_CUDA(cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking));
#undef TRY
  p->ws_bytes = tot;
  *out = p;
  return 0;
}

// Orbital grid: every per-orbital transform (density sweep, Hamiltonian apply) runs on a smaller
// box (nxw, nyw, nzw) while rho, the potentials and the dense API keep the plan's grid.  Exact, not
// an approximation: psi carries frequencies |f_c| <= gmax_c, so |psi|^2 and the part of v_eff psi
// that lands on the sphere only involve |f_c| <= 2 gmax_c, which a box with n_c >= 4 gmax_c + 1
// represents without aliasing (the usual coarse wave-function grid of plane-wave codes; SURVEY.md
// 8d notes that every shipped configuration over-satisfies it).  rho moves to the plan's grid by
// Fourier interpolation, v_eff to the orbital grid by Fourier truncation (four single-grid FFTs
// per evaluation).  Allocates: set-up time only.
extern "C" int jrb_plan_set_orbital_grid(jrb_plan* p, int32_t nxw, int32_t nyw, int32_t nzw) {
  if (!p || p->ngrid == 0 || p->orbital_only) {
    set_error("jrb_plan_set_orbital_grid: needs a full plan");
    return JRB_EINVAL;
  }
  JRB_CUDA(cudaSetDevice(p->device));
  JRB_CUDA(cudaDeviceSynchronize());
  if (p->wf) {
    p->ws_bytes -= p->wf->ws_bytes;
    jrb_plan_destroy(p->wf);
    p->wf = nullptr;
  }
  const int nw[3] = {nxw, nyw, nzw}, n[3] = {p->nx, p->ny, p->nz};
  if (nxw == p->nx && nyw == p->ny && nzw == p->nz) {
    if (!p->d_ws_a) {
      set_error("jrb_plan_set_orbital_grid: the plan's own pencil work space was released by an "
                "earlier call; create a new plan to go back to the full grid");
      return JRB_EINVAL;
    }
    return 0;
  }
  for (int c = 0; c < 3; ++c) {
    // an unchanged axis is always allowed (if the caller's grid aliases there, so does the reference)
    if (nw[c] != n[c] && (nw[c] > n[c] || nw[c] < 4 * p->gmax[c] + 1)) {
      set_error("jrb_plan_set_orbital_grid: axis " + std::to_string(c) + " needs 4*gmax+1 = " +
                std::to_string(4 * p->gmax[c] + 1) + " <= n_w <= " + std::to_string(n[c]) +
                ", got " + std::to_string(nw[c]));
      return JRB_EINVAL;
    }
  }
  std::vector<uint8_t> mask((size_t)nxw * nyw * nzw, 0);
  for (int64_t g = 0; g < p->ng; ++g) {
    int idx[3];
    for (int c = 0; c < 3; ++c) {
      const int f = p->h_freq[g * 3 + c];
      idx[c] = f >= 0 ? f : f + nw[c];
    }
    mask[((size_t)idx[0] * nyw + idx[1]) * nzw + idx[2]] = 1;
  }
  jrb_plan_desc d{};
  d.nx = nxw; d.ny = nyw; d.nz = nzw;
  d.ns = p->ns; d.nk = p->nk; d.nb = p->nb;
  d.mask = mask.data();
  d.kpts = p->h_kpts;
  d.cell = p->cell;
  d.device = p->device;
  d.batch_groups = 0;
  jrb_plan* w = nullptr;
  int rc = plan_create_impl(&d, true, &w);
  if (rc) return rc;
  // the compact order must be the plan's own (both enumerate non-negative then negative
  // frequencies in C order)
  bool same = w->ng == p->ng;
  for (int64_t i = 0; same && i < p->ng * 3; ++i) same = w->h_freq[i] == p->h_freq[i];
  if (!same) {
    jrb_plan_destroy(w);
    set_error("jrb_plan_set_orbital_grid: internal error, sphere enumeration differs");
    return JRB_EINVAL;
  }
  // this plan delegates its pencil passes from now on: release its own pencil work space
  {
    const long long total_groups = (long long)p->ns * p->nk * p->ngroups_per_k;
    if (p->d_a_keep) p->ws_bytes -= total_groups * p->a_group_elems * (long long)sizeof(cplx);
    if (p->d_psi) p->ws_bytes -= total_groups * p->psi_group_elems * (long long)sizeof(cplx);
    if (p->d_ws_a) p->ws_bytes -= p->a_copy_elems * (p->fused == 2 ? 3 : 1) * (long long)sizeof(cplx);
  }
  for (cplx** q : {&p->d_a_keep, &p->d_psi, &p->d_ws_a, &p->d_ws_b}) {
    if (*q) cudaFree(*q);
    *q = nullptr;
  }
  if (p->d_rho_part) cudaFree(p->d_rho_part);
  p->d_rho_part = nullptr;
  p->wf = w;
  p->ws_bytes += w->ws_bytes;
  return 0;
}

extern "C" int jrb_plan_orbital_grid(const jrb_plan* p, int32_t* dims) {
  if (!p || !dims) return JRB_EINVAL;
  const jrb_plan* o = p->wf ? p->wf : p;
  dims[0] = o->nx; dims[1] = o->ny; dims[2] = o->nz;
  return 0;
}

/* 0: single pencil passes, 1: fused y+x plane kernels, 2: the 128 x 128 fused family */
extern "C" int jrb_plan_orbital_fused(const jrb_plan* p) {
  if (!p) return JRB_EINVAL;
  return (p->wf ? p->wf : p)->fused;
}

extern "C" int64_t jrb_plan_psi_cache_bytes(const jrb_plan* p) {
  if (!p) return 0;
  const jrb_plan* o = p->wf ? p->wf : p;
  if (!o->d_psi) return 0;
  return (int64_t)o->ns * o->nk * o->ngroups_per_k * o->psi_group_elems * (int64_t)sizeof(cplx);
}

/* smallest alias-free orbital box per axis: 4 gmax + 1 */
extern "C" int jrb_plan_min_orbital_grid(const jrb_plan* p, int32_t* dims) {
  if (!p || !dims || p->ngrid == 0) return JRB_EINVAL;
  for (int c = 0; c < 3; ++c) dims[c] = 4 * p->gmax[c] + 1;
  return 0;
}
