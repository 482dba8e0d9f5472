// Instantiations of the fused y+x kernels (fft_fused.cuh) for one group of axis lengths.
#include "fft_fused.cuh"

#define JRB_SIZES(X) X(7) X(8) X(9) X(12) X(16) X(24) X(32) X(36)

namespace jrb {

int fused_group0(int kind, int n, const FusedArgs& a, int ctas, cudaStream_t st) {
  switch (n) {
#define X(N_) \
  case N_:    \
    return launch_fused<N_>(kind, a, ctas, st);
    JRB_SIZES(X)
#undef X
    default:
      return 1;
  }
}

int fused_smem_group0(int n, int nxo, int ncol) {
  switch (n) {
#define X(N_) \
  case N_:    \
    return fused_smem_bytes<N_>(nxo, ncol);
    JRB_SIZES(X)
#undef X
    default:
      return -1;
  }
}

int fused_threads_group0(int n) {
  switch (n) {
#define X(N_) \
  case N_:    \
    return fused_threads<N_>();
    JRB_SIZES(X)
#undef X
    default:
      return -1;
  }
}

int fused_psi_plane_group0(int n, int nxo) {
  switch (n) {
#define X(N_) \
  case N_:    \
    return fused_psi_plane<N_>(nxo);
    JRB_SIZES(X)
#undef X
    default:
      return -2;
  }
}

}  // namespace jrb
