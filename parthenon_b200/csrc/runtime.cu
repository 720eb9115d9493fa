// runtime.cu — plumbing entry points of the C ABI (device, memory, streams, events).
#include <cstdarg>

#include "common.cuh"

namespace pb2 {
static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error("no CUDA device available (this library has no CPU fallback)");
    return PB2_ERR_NO_DEVICE;
  }
  return PB2_OK;
}
} // namespace pb2

using namespace pb2;

extern "C" {

int pb2_version(void) { return 100; }
const char *pb2_last_error(void) { return g_err; }
int64_t pb2_launch_count(void) { return g_launches.load(); }

int pb2_device_count(int *count) {
  PB2_REQUIRE(count, "null argument");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  *count = n;
  return PB2_OK;
}
int pb2_set_device(int device) {
  if (int rc = require_device()) return rc;
  PB2_CUDA_CHECK(cudaSetDevice(device));
  return PB2_OK;
}
int pb2_device_sm_count(int *count) {
  PB2_REQUIRE(count, "null argument");
  if (int rc = require_device()) return rc;
  int dev = 0;
  PB2_CUDA_CHECK(cudaGetDevice(&dev));
  PB2_CUDA_CHECK(cudaDeviceGetAttribute(count, cudaDevAttrMultiProcessorCount, dev));
  return PB2_OK;
}
int pb2_malloc(void **ptr, size_t bytes) {
  PB2_REQUIRE(ptr, "null argument");
  if (int rc = require_device()) return rc;
  PB2_CUDA_CHECK(cudaMalloc(ptr, bytes ? bytes : 1));
  return PB2_OK;
}
int pb2_free(void *ptr) {
  if (!ptr) return PB2_OK;
  PB2_CUDA_CHECK(cudaFree(ptr));
  return PB2_OK;
}
int pb2_host_alloc(void **ptr, size_t bytes) {
  PB2_REQUIRE(ptr, "null argument");
  if (int rc = require_device()) return rc;
  PB2_CUDA_CHECK(cudaMallocHost(ptr, bytes ? bytes : 1));
  return PB2_OK;
}
int pb2_host_free(void *ptr) {
  if (!ptr) return PB2_OK;
  PB2_CUDA_CHECK(cudaFreeHost(ptr));
  return PB2_OK;
}
int pb2_memset(void *ptr, int value, size_t bytes, pb2_stream_t stream) {
  PB2_CUDA_CHECK(cudaMemsetAsync(ptr, value, bytes, as_stream(stream)));
  return PB2_OK;
}
int pb2_memcpy_h2d(void *dst, const void *src, size_t bytes, pb2_stream_t stream) {
  PB2_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, as_stream(stream)));
  return PB2_OK;
}
int pb2_memcpy_d2h(void *dst, const void *src, size_t bytes, pb2_stream_t stream) {
  PB2_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
  return PB2_OK;
}
int pb2_memcpy_d2d(void *dst, const void *src, size_t bytes, pb2_stream_t stream) {
  PB2_CUDA_CHECK(
      cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
  return PB2_OK;
}
int pb2_stream_create(pb2_stream_t *stream) {
  PB2_REQUIRE(stream, "null argument");
  if (int rc = require_device()) return rc;
  cudaStream_t s;
  PB2_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *stream = s;
  return PB2_OK;
}
int pb2_stream_destroy(pb2_stream_t stream) {
  PB2_CUDA_CHECK(cudaStreamDestroy(as_stream(stream)));
  return PB2_OK;
}
int pb2_stream_sync(pb2_stream_t stream) {
  PB2_CUDA_CHECK(cudaStreamSynchronize(as_stream(stream)));
  return PB2_OK;
}
int pb2_device_sync(void) {
  PB2_CUDA_CHECK(cudaDeviceSynchronize());
  return PB2_OK;
}
int pb2_event_create(pb2_event_t *ev) {
  PB2_REQUIRE(ev, "null argument");
  if (int rc = require_device()) return rc;
  cudaEvent_t e;
  PB2_CUDA_CHECK(cudaEventCreate(&e));
  *ev = e;
  return PB2_OK;
}
int pb2_event_destroy(pb2_event_t ev) {
  PB2_CUDA_CHECK(cudaEventDestroy(reinterpret_cast<cudaEvent_t>(ev)));
  return PB2_OK;
}
int pb2_event_record(pb2_event_t ev, pb2_stream_t stream) {
  PB2_CUDA_CHECK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(ev), as_stream(stream)));
  return PB2_OK;
}
int pb2_event_sync(pb2_event_t ev) {
  PB2_CUDA_CHECK(cudaEventSynchronize(reinterpret_cast<cudaEvent_t>(ev)));
  return PB2_OK;
}
int pb2_event_query(pb2_event_t ev) {
  cudaError_t e = cudaEventQuery(reinterpret_cast<cudaEvent_t>(ev));
  if (e == cudaSuccess) return 0;
  if (e == cudaErrorNotReady) return 1;
  PB2_CUDA_CHECK(e);
  return PB2_ERR_CUDA;
}
int pb2_stream_wait_event(pb2_stream_t stream, pb2_event_t ev) {
  PB2_CUDA_CHECK(
      cudaStreamWaitEvent(as_stream(stream), reinterpret_cast<cudaEvent_t>(ev), 0));
  return PB2_OK;
}
int pb2_event_elapsed_ms(pb2_event_t start, pb2_event_t stop, float *ms) {
  PB2_REQUIRE(ms, "null argument");
  PB2_CUDA_CHECK(cudaEventElapsedTime(ms, reinterpret_cast<cudaEvent_t>(start),
                                      reinterpret_cast<cudaEvent_t>(stop)));
  return PB2_OK;
}

} // extern "C"
