// Do the FP64 FMA pipe (FFT butterflies) and the FP64 tensor pipe (DMMA, the QR GEMMs) of one SM
// run concurrently?  Half of the warps of every CTA issue DFMA chains, the other half DMMA chains;
// compared with each half running alone (the other half idle).  If the combined time is close to
// the slower of the two, QR and FFT work could be co-scheduled on the same SMs.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fp64_dmma_dfma_mix.cu -o tools/fp64_mix.bin
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

// mode bit 0: even warps run DFMA; bit 1: odd warps run DMMA
__global__ void __launch_bounds__(256) k_mix(double* out, int iters, int mode, double a, double b) {
  const int warp = threadIdx.x >> 5;
  double s = 0;
  if ((warp & 1) == 0) {
    if (mode & 1) {
      double v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = threadIdx.x * 1e-3 + i;
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fma(v[i], a, b);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) s += v[i];
    }
  } else {
    if (mode & 2) {
      double c[8][2];
#pragma unroll
      for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
      for (int it = 0; it < iters / 4; ++it) {  // one DMMA = 8 FMA per thread: same flops per warp
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    }
  }
  if (s == 123.456) out[0] = s;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, 1024));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 8192;
  for (int bps : {2, 4}) {
    for (int mode : {1, 2, 3}) {
      float best = 1e30f;
      for (int r = 0; r < 6; ++r) {
        CK(cudaEventRecord(e0));
        k_mix<<<sms * bps, 256>>>(out, iters, mode, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0 && ms < best) best = ms;
      }
      // flops: DFMA warps: 4 warps x 32 thr x 16 x 2 x iters; DMMA warps: 4 warps x (iters/4*16) x 512
      const double f_dfma = (mode & 1) ? 4.0 * 32 * 16 * 2 * iters * sms * bps : 0;
      const double f_dmma = (mode & 2) ? 4.0 * (iters / 4 * 16.0) * 512 * sms * bps : 0;
      printf("%d CTA/SM mode %s: %.3f ms  DFMA %.1f TF/s  DMMA %.1f TF/s  sum %.1f TF/s\n", bps,
             mode == 1 ? "DFMA only" : mode == 2 ? "DMMA only" : "both     ", best, f_dfma / best / 1e9,
             f_dmma / best / 1e9, (f_dfma + f_dmma) / best / 1e9);
    }
  }
  return 0;
}
