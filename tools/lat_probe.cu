// Latency probe (tuning aid): dependent-issue latencies on the SM of this GPU, in SM clocks.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/lat_probe.cu -o tools/lat_probe.bin
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(long long* out, double* sink, int nthreads_active) {
  __shared__ double sh[1024];
  __shared__ double2 sh2[1024];
  const int t = threadIdx.x;
  sh[t % 1024] = 1.0 + 1e-9 * t;
  sh2[t % 1024] = make_double2(1.0 + 1e-9 * t, 0.5);
  __syncthreads();
  double x = 1.0 + 1e-12 * t, y = 1.0000001;
  long long c0, c1;
  constexpr int N = 512;
  // 1. dependent DFMA
  c0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, y, 1e-9);
  c1 = clock64();
  if (t == 0) out[0] = (c1 - c0);
  // 2. dependent DADD
  c0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x + y;
  c1 = clock64();
  if (t == 0) out[1] = (c1 - c0);
  // 3. dependent LDS.64 (pointer chase through values)
  int idx = t % 1024;
  c0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { double v = sh[idx]; idx = (idx + (int)v) & 1023; }
  c1 = clock64();
  if (t == 0) out[2] = (c1 - c0);
  x += idx;
  // 4. __syncthreads loop
  c0 = clock64();
  for (int i = 0; i < N; ++i) { __syncthreads(); }
  c1 = clock64();
  if (t == 0) out[3] = (c1 - c0);
  // 5. rcp.approx.ftz.f64 dependent
  c0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + 1.5; }
  c1 = clock64();
  if (t == 0) out[4] = (c1 - c0);
  // 6. full division dependent
  c0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = 1.0 / x + 1.5;
  c1 = clock64();
  if (t == 0) out[5] = (c1 - c0);
  // 7. sqrt dependent
  c0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = sqrt(x) + 1.5;
  c1 = clock64();
  if (t == 0) out[6] = (c1 - c0);
  // 8. STS then barrier then LDS (publish/consume round trip)
  c0 = clock64();
  for (int i = 0; i < N; ++i) {
    if (t == (i & 31)) sh[7] = x;
    __syncthreads();
    x = fma(sh[7], 1.0000001, 1e-9);
  }
  c1 = clock64();
  if (t == 0) out[7] = (c1 - c0);
  // 9. LDS.128 dependent
  idx = t % 1024;
  c0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { double2 v = sh2[idx]; idx = (idx + (int)v.x) & 1023; }
  c1 = clock64();
  if (t == 0) out[8] = (c1 - c0);
  sink[t] = x + idx;
}
int main() {
  long long* d; double* s;
  cudaMalloc(&d, 16 * sizeof(long long)); cudaMalloc(&s, 1024 * sizeof(double));
  const char* names[] = {"DFMA dep", "DADD dep", "LDS.64 dep", "__syncthreads", "rcp.approx.f64+DADD", "1/x + DADD", "sqrt + DADD", "STS+bar+LDS+DFMA", "LDS.128 dep"};
  for (int nt : {32, 128, 512}) {
    probe<<<1, nt>>>(d, s, nt);
    probe<<<1, nt>>>(d, s, nt);
    long long h[16];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("threads %d:", nt);
    for (int i = 0; i < 9; ++i) printf("  %s %.1f", names[i], h[i] / 512.0);
    printf("\n");
  }
  return 0;
}
