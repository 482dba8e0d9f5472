// Measures the FP64 issue ceilings that bound this path on the box it runs on:
//   DFMA  (vector FP64 FMA pipe: the pencil FFT butterflies)
//   DMMA  (mma.sync.m8n8k4.f64 tensor pipe: the Cholesky-QR GEMMs)
//   LDS/STS.128 shared-memory exchange bandwidth and an L2-resident copy.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fp64_peak.cu -o tools/fp64_peak.bin
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int CHAINS>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
  double v[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) v[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) v[i] = fma(v[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += v[i];
  if (s == 123.456) out[0] = s;
}

template <int CHAINS>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a, double b) {
  double c[CHAINS][2];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

// each thread stores and loads 16-byte elements with a transposing pattern (stride 8 x 16 B)
__global__ void __launch_bounds__(256) k_smem(double* out, int iters) {
  extern __shared__ double2 sm[];
  const int t = threadIdx.x;
  double2 v = make_double2(t, t);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[j * 256 + t] = v;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      double2 w = sm[((t + j * 32) & 255) * 8 + ((t >> 5) + j) % 8];
      v.x += w.x; v.y += w.y;
    }
    __syncthreads();
  }
  if (v.x == 123.456) out[0] = v.x;
}

__global__ void __launch_bounds__(256) k_copy(const double2* __restrict__ in, double2* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) out[i] = in[i];
}

template <class F>
static float time_ms(F f, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  printf("device %s, %d SMs\n", p.name, sms);
  double* out; CK(cudaMalloc(&out, 1024));
  const int iters = 4096;
  for (int bps : {1, 2, 4}) {
    const int blocks = sms * bps;
    {
      float ms = time_ms([&] { k_dfma<16><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
      double fl = 2.0 * 16 * iters * 256.0 * blocks;
      printf("DFMA  %d CTA/SM x256 thr, 16 chains: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms / 1e9);
    }
    {
      float ms = time_ms([&] { k_dmma<8><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
      double fl = 2.0 * 8 * 8 * 4 * 8.0 * iters * 8 * blocks;  // per warp-instr 2*8*8*4 flop, 8 chains, 8 warps
      printf("DMMA  %d CTA/SM x8 warps, 8 chains:  %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms / 1e9);
    }
  }
  {
    CK(cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    const int blocks = sms * 4;
    float ms = time_ms([&] { k_smem<<<blocks, 256, 32768>>>(out, 2048); }, 5);
    double bytes = 2.0 * 8 * 16 * 256.0 * 2048 * blocks;
    printf("SMEM  4 CTA/SM LDS+STS.128: %.3f ms  %.2f TB/s (st+ld)\n", ms, bytes / ms / 1e9);
  }
  for (long long mb : {16, 48, 96, 512, 4096}) {
    const long long n = mb * 1024 * 1024 / 16 / 2;  // in + out = mb MiB
    double2 *a, *b; CK(cudaMalloc(&a, n * 16)); CK(cudaMalloc(&b, n * 16));
    CK(cudaMemset(a, 0, n * 16));
    float ms = time_ms([&] { k_copy<<<sms * 8, 256>>>(a, b, n); }, 10);
    printf("COPY  working set %5lld MiB: %.4f ms  %.2f TB/s (rd+wr)\n", mb, ms, 2.0 * n * 16 / ms / 1e9);
    CK(cudaFree(a)); CK(cudaFree(b));
  }
  return 0;
}
