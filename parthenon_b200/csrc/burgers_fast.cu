// burgers_fast.cu — same kernels with FMA contraction enabled (nvcc default -fmad=true).
#define PB2_NS fast
#include "burgers_impl.cuh"
namespace pb2 {
int burgers_fluxes_fast(const pb2_burgers_args *a, cudaStream_t s) {
  return fast::launch_fluxes(a, s);
}
int burgers_update_fast(const pb2_burgers_args *a, cudaStream_t s) {
  return fast::launch_update(a, s);
}
} // namespace pb2
