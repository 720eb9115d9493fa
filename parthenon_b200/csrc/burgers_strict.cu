// burgers_strict.cu — reference-order arithmetic, compiled with -fmad=false.
#define PB2_NS strict
#include "burgers_impl.cuh"
namespace pb2 {
int burgers_fluxes_strict(const pb2_burgers_args *a, cudaStream_t s) {
  return strict::launch_fluxes(a, s);
}
int burgers_update_strict(const pb2_burgers_args *a, cudaStream_t s) {
  return strict::launch_update(a, s);
}
} // namespace pb2
