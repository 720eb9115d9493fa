// device.hpp — RAII helpers over the plumbing entry points of the C ABI
// (include/parthenon_b200.h).  The host library never links the CUDA runtime itself.
#pragma once
#include <cstddef>

#include "parthenon_b200.h"
#include "types.hpp"

namespace parthenon {

// RAII device allocation through the C ABI (zero-initialised like a Kokkos::View)
class DeviceBuffer {
 public:
  DeviceBuffer() = default;
  DeviceBuffer(const DeviceBuffer &) = delete;
  DeviceBuffer &operator=(const DeviceBuffer &) = delete;
  ~DeviceBuffer() { Free(); }
  void Allocate(size_t bytes, pb2_stream_t stream) {
    Free();
    PB2_CHECK(pb2_malloc(&p_, bytes));
    bytes_ = bytes;
    PB2_CHECK(pb2_memset(p_, 0, bytes, stream));
  }
  void Free() {
    if (p_) pb2_free(p_);
    p_ = nullptr;
    bytes_ = 0;
  }
  template <typename T = void>
  T *get() const { return static_cast<T *>(p_); }
  size_t bytes() const { return bytes_; }
  explicit operator bool() const { return p_ != nullptr; }

 private:
  void *p_ = nullptr;
  size_t bytes_ = 0;
};

} // namespace parthenon
