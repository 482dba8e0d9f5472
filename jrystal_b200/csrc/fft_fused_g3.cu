// The fused y+x kernels for 128 x 128 planes (fft_fused128.cuh) and their dispatch helpers.
#include "fft_fused128.cuh"

namespace jrb {

int fused128_launch(int kind, const FusedArgs& a, int ctas, cudaStream_t st) {
  return launch_fused128(kind, a, ctas, st);
}

int fused128_smem(int nxo, int ncol) { return F128::smem_bytes(nxo, ncol, true); }

int fused128_rho_reduce(const FusedArgs& a, int ctas, double* rho, cudaStream_t st) {
  dim3 grid(F128::N * F128::M / 32, 2 * a.m.nz), block(32, 8);
  k_rho_reduce128<<<grid, block, 0, st>>>(a.rho_part, a.seg_z, ctas, a.segmax, a.ngroups, a.m.nz,
                                          rho);
  JRB_CHECK_LAUNCH("k_rho_reduce128");
  return 0;
}

}  // namespace jrb
