// common.cuh — shared helpers for the C-ABI implementation (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "parthenon_b200.h"

namespace pb2 {

void set_error(const char *fmt, ...);
extern std::atomic<int64_t> g_launches;

inline cudaStream_t as_stream(pb2_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

#define PB2_CUDA_CHECK(expr)                                                              \
  do {                                                                                    \
    cudaError_t err__ = (expr);                                                           \
    if (err__ != cudaSuccess) {                                                           \
      ::pb2::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                 \
                       cudaGetErrorString(err__));                                        \
      return (err__ == cudaErrorNoDevice || err__ == cudaErrorInsufficientDriver)         \
                 ? PB2_ERR_NO_DEVICE                                                      \
                 : PB2_ERR_CUDA;                                                          \
    }                                                                                     \
  } while (0)

#define PB2_REQUIRE(cond, msg)                                                            \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      ::pb2::set_error("%s:%d: %s", __FILE__, __LINE__, msg);                             \
      return PB2_ERR_INVALID;                                                             \
    }                                                                                     \
  } while (0)

// call after every kernel launch
#define PB2_LAUNCH_CHECK()                                                                \
  do {                                                                                    \
    ::pb2::g_launches.fetch_add(1, std::memory_order_relaxed);                            \
    PB2_CUDA_CHECK(cudaGetLastError());                                                   \
  } while (0)

int require_device();
// Device memory of region tables (exchange.cu, prores.cu).  An adaptive run destroys and
// re-creates dozens of small tables on every remesh; cudaMalloc / cudaFree for each (the free
// synchronises the device) is replaced by a pool of size classes (4 KB x 4^k) per device.
// Semantics of cudaMalloc / cudaFree are kept: a pointer handed out is not referenced by any
// work still in flight — a block freed since the pool last synchronised its device triggers ONE
// cudaDeviceSynchronize on reuse, which clears every block freed before it.
cudaError_t table_alloc(void **ptr, size_t bytes);
void table_free(void *ptr);

// ---- optional per-kernel timing with CUDA events (pb2_profile_* of the C ABI) ----------
enum KernelId {
  K_PACK = 0, K_UNPACK, K_COPY, K_RESTRICT, K_PROLONGATE, K_WEIGHTED_SUM, K_FLUX_DIV,
  K_FLUX_X, K_FLUX_Y, K_FLUX_Z, K_UPDATE, K_DERIVED_DT, K_HISTORY, K_SWEEP_X, K_SWEEP_Y, K_SWEEP_Z, K_INTERIOR, K_HALO_UNIFORM, K_FLUX_CORRECT, K_ADVECTION_FLUX, K_APPLY_BC, K_SWEEP_XPAIR,
  K_SWEEP_CHUNK_Y, K_SWEEP_CHUNK_Z, K_COUNT
};
extern std::atomic<int> g_profile_on;
void profile_begin(int id, cudaStream_t s, void **token, double work);
void profile_end(void *token, cudaStream_t s);
// `work`: what the launch processes in the kernel's own unit (zones for the stencil kernels,
// values for the copy / pack / unpack / restrict / prolongate kernels), summed per kernel so
// that bytes per launch follow from what was actually launched (pb2_profile_get_work)
struct ProfScope {
  void *token = nullptr;
  cudaStream_t s;
  ProfScope(int id, cudaStream_t st, double work = 0.0) : s(st) {
    if (g_profile_on.load(std::memory_order_relaxed)) profile_begin(id, st, &token, work);
  }
  ~ProfScope() {
    if (token) profile_end(token, s);
  }
};

// Division by a runtime-constant 32-bit divisor with a precomputed magic number
// (n must be < 2^31).  q = (umulhi(n, magic) + n) >> shift.
struct FastDiv {
  uint32_t d, magic, shift;
  __host__ void init(uint32_t div) {
    d = div;
    if (div <= 1) {
      magic = 0;
      shift = 0;
      d = 1;
      return;
    }
    uint32_t s = 0;
    while ((1u << s) < div) ++s;
    shift = s;
    magic = static_cast<uint32_t>(((1ull << 32) * ((1ull << s) - div)) / div + 1);
  }
  __device__ __forceinline__ uint32_t div(uint32_t n) const {
    return (__umulhi(n, magic) + n) >> shift;
  }
  __device__ __forceinline__ void divmod(uint32_t n, uint32_t &q, uint32_t &r) const {
    q = div(n);
    r = n - q * d;
  }
};

} // namespace pb2
