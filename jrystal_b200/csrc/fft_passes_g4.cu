// Instantiations of the pencil passes for one group of axis lengths (split over several
// translation units so nvcc can build them in parallel).
#include "fft_passes.cuh"

#define JRB_SIZES(X) X(36) X(40) X(45) X(49) X(50) X(54) X(56) X(60)

namespace jrb {

int pass_group4(PassKind kind, int n, const PassArgs& a, cudaStream_t st) {
  switch (n) {
#define X(N_) \
  case N_:    \
    return launch_pass<N_>(kind, a, st);
    JRB_SIZES(X)
#undef X
    default:
      return 1;
  }
}

int dense_group4(int n, const DenseArgs& a, int dir, long long batch, cudaStream_t st) {
  switch (n) {
#define X(N_) \
  case N_:    \
    return launch_dense<N_>(a, dir, batch, st);
    JRB_SIZES(X)
#undef X
    default:
      return 1;
  }
}

}  // namespace jrb
