// which of the L2 cache-hint forms runs on sm_100a?  (nvcc -arch=sm_100a l2hint_probe.cu && ./a.out)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ uint64_t pol_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ uint64_t pol_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__global__ void k_policy(uint64_t *out) { out[0] = pol_last(); out[1] = pol_first(); }
__global__ void k_ld_hint(const double *in, double *out) {
  double v;
  const uint64_t p = pol_last();
  asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(in + threadIdx.x), "l"(p));
  out[threadIdx.x] = v;
}
__global__ void k_st_hint(double *out) {
  const uint64_t p = pol_first();
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(out + threadIdx.x), "d"(1.0 * threadIdx.x), "l"(p) : "memory");
}
__global__ void k_cpasync8_hint(const double *in, double *out) {
  __shared__ double s[32];
  const uint64_t p = pol_last();
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(&s[threadIdx.x]);
  asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;" ::"r"(sa), "l"(in + threadIdx.x), "l"(p) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  out[threadIdx.x] = s[threadIdx.x];
}
__global__ void k_cpasync16_hint(const double *in, double *out) {
  __shared__ __align__(16) double s[64];
  const uint64_t p = pol_last();
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(&s[2 * threadIdx.x]);
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(sa), "l"(in + 2 * threadIdx.x), "l"(p) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  out[threadIdx.x] = s[2 * threadIdx.x];
}
#define RUN(name, ...)                                                         \
  do {                                                                         \
    name<<<1, 32>>>(__VA_ARGS__);                                              \
    cudaError_t e = cudaDeviceSynchronize();                                   \
    printf("%-18s %s\n", #name, e == cudaSuccess ? "ok" : cudaGetErrorString(e)); \
    if (e != cudaSuccess) return 1;                                            \
  } while (0)
int main() {
  double *in, *out;
  uint64_t *po;
  cudaMalloc(&in, 1024);
  cudaMalloc(&out, 1024);
  cudaMalloc(&po, 64);
  cudaMemset(in, 0, 1024);
  RUN(k_policy, po);
  uint64_t h[2];
  cudaMemcpy(h, po, 16, cudaMemcpyDeviceToHost);
  printf("policies %016llx %016llx\n", (unsigned long long)h[0], (unsigned long long)h[1]);
  RUN(k_ld_hint, in, out);
  RUN(k_st_hint, out);
  RUN(k_cpasync16_hint, in, out);
  RUN(k_cpasync8_hint, in, out);
  return 0;
}
