// Instantiations of the fused y+x kernels (fft_fused.cuh) for one group of axis lengths.
#include "fft_fused.cuh"

#define JRB_SIZES(X) X(72) X(81) X(96)

namespace jrb {

int fused_group2(int kind, int n, const FusedArgs& a, int ctas, cudaStream_t st) {
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

int fused_smem_group2(int n, int nxo, int ncol) {
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

int fused_threads_group2(int n) {
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

int fused_psi_plane_group2(int n, int nxo) {
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
