// Plan-owned communicator: all-reduce of the partial densities (and of the small QR matrices of
// the row-sharded layout) over NVLink / NVSwitch PEER MEMORY, written as kernels of this library.
//
// The reference shards the k-mesh over devices and lets XLA insert the all-reduce of rho where
// einsum('skb...,skb->s...') contracts the sharded k axis (jrystal/_src/pw.py:278 under the
// NamedSharding of calc/calc_ground_state_energy_all_electrons.py:83-91,151-158).  Here that one
// collective of the path is the SURVEY 8b entry point jrb_allreduce_rho.
//
// One process per GPU.  Every rank owns a SYMMETRIC region (cudaMalloc, exported with
// cudaIpcGetMemHandle, mapped by every peer with cudaIpcOpenMemHandle):
//     ctrl   : 64 x 8 bytes   epoch, three CTA counters, error word, flagA[16], flagB[16]
//     data   : [2][capacity]  the rank's published partial, slot = epoch & 1
//     result : [2][capacity]  the reduced buffer, slot = epoch & 1
// Two-step all-reduce (reduce-scatter + all-gather; per rank (W-1)/W n doubles in and out instead
// of the (W-1) n of a one-shot read), three kernels on the caller's stream, no host round trip:
//   k_comm_publish : local buffer -> own data slot; the last CTA to finish stores the epoch into
//                    flagA[rank] of EVERY rank's region (data ordered before the flag by
//                    __threadfence_system)
//   k_comm_reduce  : waits until all W flagA of the LOCAL region show the epoch, then sums ITS
//                    slice over the W data slots in RANK ORDER -- the same order for every element
//                    on every rank, so all ranks end with bit-identical sums, which the replicated
//                    grid part and the replicated Cholesky of the row-sharded layout rely on --
//                    and stores the sums into the result slot of every rank; last CTA: flagB
//   k_comm_collect : waits for all W flagB, copies the result slot to the destination; its last
//                    CTA advances the device-resident epoch (so a captured CUDA graph replays)
// Slot reuse: epoch e + 2 reuses the slots of e.  A rank publishes e + 2 after its collect of
// e + 1, which waited for every peer's flagB(e + 1), stored after that peer's reduce of e + 1 and
// hence (stream order) after its reads of data slot e.  A peer pushes results of e + 2 after it
// saw this rank's flagA(e + 2), stored after this rank's collect of e.  Calls must be issued in
// the same order on every rank (SPMD evaluation).  Spins are bounded (JRB_COMM_TIMEOUT_S,
// default 20 s of globaltimer): a missing peer raises the error word that jrb_check_status
// reports instead of hanging the GPU.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "plan.h"

namespace jrb {

constexpr int COMM_MAX = 16;
constexpr int CTRL_WORDS = 64;
// ctrl word indices
constexpr int C_EPOCH = 0, C_CNT_A = 1, C_CNT_B = 2, C_CNT_C = 3, C_ERR = 4, C_FLAG_A = 16,
              C_FLAG_B = 32;

struct PeerComm {
  int rank, world;
  long long capacity;   // doubles per slot (even)
  double* local;
  double* base[COMM_MAX];
  bool opened[COMM_MAX];
  bool connected;
  unsigned long long timeout_ns;
};

struct CommPtrs {
  double* base[COMM_MAX];
};

__device__ __forceinline__ unsigned long long* ctrl_of(double* base) {
  return reinterpret_cast<unsigned long long*>(base);
}
__device__ __forceinline__ double* data_of(double* base, long long cap, int slot) {
  return base + CTRL_WORDS + (long long)slot * cap;
}
__device__ __forceinline__ double* result_of(double* base, long long cap, int slot) {
  return base + CTRL_WORDS + (2 + (long long)slot) * cap;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// threads 0 .. world-1 wait for flag[r] >= epoch in the LOCAL region, then the CTA proceeds
__device__ __forceinline__ void wait_flags(unsigned long long* ctrl, int first, int world,
                                           unsigned long long epoch, unsigned long long timeout_ns) {
  if ((int)threadIdx.x < world) {
    const unsigned long long* f = ctrl + first + threadIdx.x;
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(f) < epoch) {
      if (globaltimer_ns() - t0 > timeout_ns) {
        ctrl[C_ERR] = 1;
        break;
      }
    }
  }
  __syncthreads();
}

// the last CTA of the grid to arrive gets true (counter reset for the next use)
__device__ __forceinline__ bool last_cta(unsigned long long* counter) {
  __shared__ bool last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long done = atomicAdd(counter, 1ull);
    last = done == (unsigned long long)gridDim.x - 1;
    if (last) *counter = 0;
  }
  __syncthreads();
  return last;
}

// n doubles of src (+ m <= 8 of extra) -> data slot of this rank
__global__ void __launch_bounds__(256)
k_comm_publish(CommPtrs p, int rank, int world, long long cap, const double* __restrict__ src,
               long long n, const double* __restrict__ extra, int m) {
  unsigned long long* ctrl = ctrl_of(p.base[rank]);
  const unsigned long long e = ctrl[C_EPOCH];
  double* slot = data_of(p.base[rank], cap, (int)(e & 1));
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long n2 = n / 2;
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const double2* s2 = reinterpret_cast<const double2*>(src);
    double2* d2 = reinterpret_cast<double2*>(slot);
    for (long long i = i0; i < n2; i += stride) d2[i] = s2[i];
    if (i0 == 0 && (n & 1)) slot[n - 1] = src[n - 1];
  } else {
    for (long long i = i0; i < n; i += stride) slot[i] = src[i];
  }
  if (i0 < m) slot[n + i0] = extra[i0];
  if (i0 == m && ((n + m) & 1)) slot[n + m] = 0.0;  // pad to an even count
  if (last_cta(ctrl + C_CNT_A) && (int)threadIdx.x < world)
    st_release_sys(ctrl_of(p.base[threadIdx.x]) + C_FLAG_A + rank, e);
}

// tot (even) doubles; this rank owns the double2 range [tot2 rank / W, tot2 (rank + 1) / W)
__global__ void __launch_bounds__(256)
k_comm_reduce(CommPtrs p, int rank, int world, long long cap, long long tot,
              unsigned long long timeout_ns) {
  unsigned long long* ctrl = ctrl_of(p.base[rank]);
  const unsigned long long e = ctrl[C_EPOCH];
  wait_flags(ctrl, C_FLAG_A, world, e, timeout_ns);
  const int slot = (int)(e & 1);
  const long long tot2 = tot / 2;
  const long long lo = tot2 * rank / world, hi = tot2 * (rank + 1) / world;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hi; i += stride) {
    double2 s = __ldcg(reinterpret_cast<const double2*>(data_of(p.base[0], cap, slot)) + i);
    for (int r = 1; r < world; ++r) {
      const double2 v = __ldcg(reinterpret_cast<const double2*>(data_of(p.base[r], cap, slot)) + i);
      s.x += v.x;
      s.y += v.y;
    }
    for (int r = 0; r < world; ++r)
      reinterpret_cast<double2*>(result_of(p.base[r], cap, slot))[i] = s;
  }
  if (last_cta(ctrl + C_CNT_B) && (int)threadIdx.x < world)
    st_release_sys(ctrl_of(p.base[threadIdx.x]) + C_FLAG_B + rank, e);
}

__global__ void __launch_bounds__(256)
k_comm_collect(CommPtrs p, int rank, int world, long long cap, double* __restrict__ dst, long long n,
               double* __restrict__ extra, int m, unsigned long long timeout_ns) {
  unsigned long long* ctrl = ctrl_of(p.base[rank]);
  const unsigned long long e = ctrl[C_EPOCH];
  wait_flags(ctrl, C_FLAG_B, world, e, timeout_ns);
  const double* res = result_of(p.base[rank], cap, (int)(e & 1));
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const long long n2 = n / 2;
    const double2* s2 = reinterpret_cast<const double2*>(res);
    double2* d2 = reinterpret_cast<double2*>(dst);
    for (long long i = i0; i < n2; i += stride) d2[i] = __ldcg(s2 + i);
    if (i0 == 0 && (n & 1)) dst[n - 1] = __ldcg(res + n - 1);
  } else {
    for (long long i = i0; i < n; i += stride) dst[i] = __ldcg(res + i);
  }
  if (i0 < m) extra[i0] = __ldcg(res + n + i0);
  if (last_cta(ctrl + C_CNT_C) && threadIdx.x == 0) ctrl[C_EPOCH] = e + 1;
}

static PeerComm* comm_of(jrb_plan* p) { return static_cast<PeerComm*>(p->comm); }

int comm_world(const jrb_plan* p) {
  const PeerComm* c = static_cast<const PeerComm*>(p->comm);
  return (c && c->connected) ? c->world : 1;
}

void comm_destroy(jrb_plan* p) {
  PeerComm* c = comm_of(p);
  if (!c) return;
  for (int r = 0; r < COMM_MAX; ++r)
    if (c->opened[r]) cudaIpcCloseMemHandle(c->base[r]);
  if (c->local) cudaFree(c->local);
  delete c;
  p->comm = nullptr;
}

// in-place SUM over the ranks of buf[0..n) and extra[0..m), asynchronous on st
int launch_comm_allreduce(jrb_plan* p, double* buf, long long n, double* extra, int m,
                          cudaStream_t st) {
  PeerComm* c = comm_of(p);
  if (!c || !c->connected) {
    set_error("no communicator: call jrb_comm_create + jrb_comm_connect first");
    return JRB_EINVAL;
  }
  if (c->world == 1 || (n == 0 && m == 0)) return 0;
  if (m < 0 || m > 8 || n < 0 || n + m + 1 > c->capacity) {
    set_error("jrb_allreduce: buffer larger than the communicator's capacity");
    return JRB_EINVAL;
  }
  CommPtrs ptrs;
  for (int r = 0; r < COMM_MAX; ++r) ptrs.base[r] = r < c->world ? c->base[r] : nullptr;
  const long long tot = (n + m + 1) / 2 * 2;
  const int blocks = (int)std::max<long long>(1, std::min<long long>((n / 2 + 255) / 256, 148 * 2));
  const long long slice2 = tot / 2 / c->world + 1;
  const int rblocks = (int)std::max<long long>(1, std::min<long long>((slice2 + 255) / 256, 148 * 2));
  k_comm_publish<<<blocks, 256, 0, st>>>(ptrs, c->rank, c->world, c->capacity, buf, n, extra, m);
  JRB_CHECK_LAUNCH("k_comm_publish");
  k_comm_reduce<<<rblocks, 256, 0, st>>>(ptrs, c->rank, c->world, c->capacity, tot, c->timeout_ns);
  JRB_CHECK_LAUNCH("k_comm_reduce");
  k_comm_collect<<<blocks, 256, 0, st>>>(ptrs, c->rank, c->world, c->capacity, buf, n, extra, m,
                                         c->timeout_ns);
  JRB_CHECK_LAUNCH("k_comm_collect");
  return 0;
}

// 1 if a bounded spin of this plan's communicator ran out (a peer never arrived)
int comm_error(jrb_plan* p, cudaStream_t st) {
  PeerComm* c = comm_of(p);
  if (!c) return 0;
  unsigned long long err = 0;
  cudaMemcpyAsync(&err, reinterpret_cast<unsigned long long*>(c->local) + C_ERR, sizeof(err),
                  cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  return err != 0;
}

}  // namespace jrb

using namespace jrb;

extern "C" int jrb_comm_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

extern "C" int jrb_comm_create(jrb_plan* p, int32_t rank, int32_t world, int64_t capacity,
                               void* handle_out) {
  if (!p || !handle_out) {
    set_error("jrb_comm_create: null argument");
    return JRB_EINVAL;
  }
  if (world < 1 || world > COMM_MAX || rank < 0 || rank >= world) {
    set_error("jrb_comm_create: need 0 <= rank < world <= 16");
    return JRB_EINVAL;
  }
  JRB_CUDA(cudaSetDevice(p->device));
  comm_destroy(p);
  if (capacity <= 0) {
    // the density of the plan's grid + E_kin, or the small matrices of a rows-only plan
    const long long grid = (long long)p->ns * p->ngrid;
    const long long small = 2ll * p->ns * p->nk * p->nb * p->nb;
    capacity = std::max(grid, small) + 16;
  }
  capacity = (capacity + 63) / 64 * 64;
  PeerComm* c = new PeerComm();
  std::memset(c, 0, sizeof(*c));
  c->rank = rank;
  c->world = world;
  c->capacity = capacity;
  double t = 20.0;
  if (const char* env = std::getenv("JRB_COMM_TIMEOUT_S")) t = std::atof(env);
  c->timeout_ns = (unsigned long long)(t * 1e9);
  const size_t bytes = (size_t)(CTRL_WORDS + 4 * capacity) * sizeof(double);
  cudaError_t e = cudaMalloc(&c->local, bytes);
  if (e != cudaSuccess) {
    delete c;
    return cuda_fail(e, "cudaMalloc (communicator region)");
  }
  p->comm = c;
  JRB_CUDA(cudaMemset(c->local, 0, bytes));
  const unsigned long long one = 1;  // epoch 0 would match the zeroed flags
  JRB_CUDA(cudaMemcpy(c->local, &one, sizeof(one), cudaMemcpyHostToDevice));
  JRB_CUDA(cudaDeviceSynchronize());
  c->base[rank] = c->local;
  cudaIpcMemHandle_t h;
  JRB_CUDA(cudaIpcGetMemHandle(&h, c->local));
  std::memcpy(handle_out, &h, sizeof(h));
  p->ws_bytes += (int64_t)bytes;
  if (world == 1) c->connected = true;
  return 0;
}

extern "C" int jrb_comm_connect(jrb_plan* p, const void* handles) {
  if (!p || !p->comm || !handles) {
    set_error("jrb_comm_connect: call jrb_comm_create first");
    return JRB_EINVAL;
  }
  JRB_CUDA(cudaSetDevice(p->device));
  PeerComm* c = comm_of(p);
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank || c->opened[r]) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const char*>(handles) + (size_t)r * sizeof(h), sizeof(h));
    void* ptr = nullptr;
    JRB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    c->base[r] = static_cast<double*>(ptr);
    c->opened[r] = true;
  }
  c->connected = true;
  return 0;
}

extern "C" int jrb_comm_world(const jrb_plan* p) { return p ? comm_world(p) : 0; }

extern "C" int jrb_allreduce(jrb_plan* p, double* buf, int64_t n, jrb_stream st) {
  if (!p || !buf) {
    set_error("jrb_allreduce: null argument");
    return JRB_EINVAL;
  }
  JRB_CUDA(cudaSetDevice(p->device));
  return launch_comm_allreduce(p, buf, n, nullptr, 0, reinterpret_cast<cudaStream_t>(st));
}

extern "C" int jrb_allreduce_rho(jrb_plan* p, double* rho, double* e_kin, jrb_stream st) {
  if (!p || !rho) {
    set_error("jrb_allreduce_rho: null argument");
    return JRB_EINVAL;
  }
  if (p->ngrid == 0) {
    set_error("jrb_allreduce_rho needs a full plan (jrb_plan_create)");
    return JRB_EINVAL;
  }
  JRB_CUDA(cudaSetDevice(p->device));
  return launch_comm_allreduce(p, rho, (long long)p->ns * p->ngrid, e_kin, e_kin ? 1 : 0,
                               reinterpret_cast<cudaStream_t>(st));
}
