// runtime.cu — plumbing entry points of the C ABI (device, memory, streams, events).
#include <cstdarg>
#include <mutex>
#include <unordered_map>
#include <map>
#include <vector>

#include <cstdlib>

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

std::atomic<int> g_profile_on{0};
namespace {
struct ProfPair {
  int id;
  double work;
  cudaEvent_t a, b;
};
std::mutex g_prof_mu;
std::vector<ProfPair *> g_prof_pending, g_prof_free;
double g_prof_ms[K_COUNT];
int64_t g_prof_n[K_COUNT];
double g_prof_work[K_COUNT];
const char *kKernelNames[K_COUNT] = {
    "pack_kernel", "unpack_kernel", "copy_kernel", "restrict_kernel", "prolongate_kernel",
    "weighted_sum_kernel", "flux_div_kernel", "flux_x_kernel", "flux_march_kernel<y>",
    "flux_march_kernel<z>", "update_kernel", "derived_dt_kernel", "history_kernel",
    "sweep_x_kernel", "sweep_march_kernel<y>", "sweep_march_kernel<z>",
    "interior_kernel", "halo_uniform_kernel", "flux_correct_kernel", "advection_flux_kernel", "apply_bc_kernel",
    "sweep_xpair_kernel", "sweep_chunk_kernel<y>", "sweep_chunk_kernel<z>"};
} // namespace

void profile_begin(int id, cudaStream_t s, void **token, double work) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfPair *p = nullptr;
  if (!g_prof_free.empty()) {
    p = g_prof_free.back();
    g_prof_free.pop_back();
  } else {
    p = new ProfPair();
    if (cudaEventCreate(&p->a) != cudaSuccess || cudaEventCreate(&p->b) != cudaSuccess) {
      cudaGetLastError();
      delete p;
      return;
    }
  }
  p->id = id;
  p->work = work;
  cudaEventRecord(p->a, s);
  *token = p;
}
void profile_end(void *token, cudaStream_t s) {
  ProfPair *p = static_cast<ProfPair *>(token);
  cudaEventRecord(p->b, s);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_pending.push_back(p);
}
static void profile_collect() {
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (ProfPair *p : g_prof_pending) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, p->a, p->b) == cudaSuccess) {
      g_prof_ms[p->id] += ms;
      g_prof_n[p->id] += 1;
      g_prof_work[p->id] += p->work;
    } else {
      cudaGetLastError();
    }
    g_prof_free.push_back(p);
  }
  g_prof_pending.clear();
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

namespace pb2 {
// FP64 FMA throughput microbenchmark: 8 independent dependency chains per thread
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
         x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b);
    x1 = fma(x1, a, b);
    x2 = fma(x2, a, b);
    x3 = fma(x3, a, b);
    x4 = fma(x4, a, b);
    x5 = fma(x5, a, b);
    x6 = fma(x6, a, b);
    x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
} // namespace pb2

// ---- pool behind table_alloc / table_free (common.cuh) ----
namespace pb2 {
namespace {
struct TableBlock {
  void *ptr;
  uint64_t freed_epoch;
};
struct TableLive {
  int device;
  size_t cls;
};
std::mutex g_tp_mu;
std::map<std::pair<int, size_t>, std::vector<TableBlock>> g_tp_free; // (device, class bytes)
std::unordered_map<void *, TableLive> g_tp_live;
std::unordered_map<int, uint64_t> g_tp_epoch; // per device: syncs the pool has done
size_t g_tp_bytes = 0;
constexpr size_t kTablePoolMaxBytes = size_t(2) << 30;
// size classes 4 KB x 4^k: a table that grows with an adaptive mesh changes class (= one real
// cudaMalloc, measured at tens of ms next to tens of GB of live slabs) once per factor of four
size_t table_class(size_t bytes) {
  size_t c = 4096;
  while (c < bytes) c <<= 2;
  return c;
}
} // namespace

cudaError_t table_alloc(void **ptr, size_t bytes) {
  const size_t cls = table_class(bytes ? bytes : 1);
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  {
    std::lock_guard<std::mutex> lk(g_tp_mu);
    auto it = g_tp_free.find({dev, cls});
    if (it != g_tp_free.end() && !it->second.empty()) {
      const TableBlock b = it->second.back();
      it->second.pop_back();
      g_tp_bytes -= cls;
      uint64_t &epoch = g_tp_epoch[dev];
      if (b.freed_epoch == epoch) { // freed since the last sync: kernels may still read it
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) return e;
        ++epoch;
      }
      g_tp_live[b.ptr] = TableLive{dev, cls};
      *ptr = b.ptr;
      return cudaSuccess;
    }
  }
  e = cudaMalloc(ptr, cls);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(g_tp_mu);
  g_tp_live[*ptr] = TableLive{dev, cls};
  return cudaSuccess;
}

void table_free(void *ptr) {
  if (!ptr) return;
  std::lock_guard<std::mutex> lk(g_tp_mu);
  auto it = g_tp_live.find(ptr);
  if (it == g_tp_live.end()) { // not ours
    cudaFree(ptr);
    return;
  }
  const TableLive live = it->second;
  g_tp_live.erase(it);
  if (g_tp_bytes + live.cls > kTablePoolMaxBytes) {
    cudaFree(ptr);
    return;
  }
  g_tp_free[{live.device, live.cls}].push_back(TableBlock{ptr, g_tp_epoch[live.device]});
  g_tp_bytes += live.cls;
}
} // namespace pb2

using namespace pb2;

extern "C" {

int pb2_measure_fp64_peak(double *tflops) {
  PB2_REQUIRE(tflops, "null argument");
  if (int rc = require_device()) return rc;
  int dev = 0, sms = 0;
  PB2_CUDA_CHECK(cudaGetDevice(&dev));
  PB2_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int ctas = sms * 8, iters = 1 << 15;
  double *out = nullptr;
  PB2_CUDA_CHECK(cudaMalloc(&out, sizeof(double) * ctas * 256));
  cudaEvent_t e0, e1;
  PB2_CUDA_CHECK(cudaEventCreate(&e0));
  PB2_CUDA_CHECK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    PB2_CUDA_CHECK(cudaEventRecord(e0));
    fp64_peak_kernel<<<ctas, 256>>>(out, iters, 0.999999, 1e-6);
    PB2_CUDA_CHECK(cudaEventRecord(e1));
    PB2_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0;
    PB2_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = 2.0 * 8.0 * iters * double(ctas) * 256.0 / (best * 1e-3) / 1e12;
  return PB2_OK;
}

int pb2_version(void) { return 100; }
const char *pb2_last_error(void) { return g_err; }
int64_t pb2_launch_count(void) { return g_launches.load(); }

int pb2_profile_enable(int on) {
  g_profile_on.store(on ? 1 : 0);
  return PB2_OK;
}
int pb2_profile_reset(void) {
  profile_collect();
  for (int i = 0; i < K_COUNT; ++i) {
    g_prof_ms[i] = 0;
    g_prof_n[i] = 0;
    g_prof_work[i] = 0;
  }
  return PB2_OK;
}
int pb2_profile_kernels(void) { return K_COUNT; }
int pb2_profile_get(int id, const char **name, double *total_ms, int64_t *launches) {
  PB2_REQUIRE(id >= 0 && id < K_COUNT, "bad kernel id");
  profile_collect();
  if (name) *name = kKernelNames[id];
  if (total_ms) *total_ms = g_prof_ms[id];
  if (launches) *launches = g_prof_n[id];
  return PB2_OK;
}

int pb2_profile_get_work(int id, double *work) {
  PB2_REQUIRE(id >= 0 && id < K_COUNT && work, "bad kernel id");
  profile_collect();
  *work = g_prof_work[id];
  return PB2_OK;
}

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
// Large device allocations are recycled through an exact-size cache: an adaptive run re-creates
// its multi-GB field slabs on every remesh, and cudaFree of such slabs (unmapping) was measured
// at up to half a second per remesh.  Callers round slab capacities (host: Variable), so sizes
// repeat.  The synchronous semantics of cudaMalloc / cudaFree are kept: a pointer is usable by
// any stream on return, and a free waits for the device before the memory can be handed out
// again.  The cache is bounded; beyond the bound memory really is returned.
namespace {
constexpr size_t kCacheMinBytes = size_t(1) << 20;
constexpr size_t kCacheMaxEntries = 96;
// upper bound of what the cache may hold back from other users of the device (torch, NCCL):
// PB2_ALLOC_CACHE_GB, default 48
size_t cache_max_bytes() {
  static const size_t v = [] {
    const char *e = std::getenv("PB2_ALLOC_CACHE_GB");
    const double gb = e ? std::atof(e) : 48.0;
    return static_cast<size_t>((gb < 0 ? 0 : gb) * double(size_t(1) << 30));
  }();
  return v;
}
struct Live {
  size_t bytes;
  int device;
};
std::mutex g_cache_mu;
std::multimap<std::pair<int, size_t>, void *> g_cache; // (device, size) -> free pointer
std::unordered_map<void *, Live> g_live;               // every live pointer of pb2_malloc
size_t g_cache_bytes = 0;

void cache_release_all_locked() {
  int cur = 0;
  cudaGetDevice(&cur);
  for (auto &kv : g_cache) {
    cudaSetDevice(kv.first.first);
    cudaFree(kv.second);
  }
  cudaSetDevice(cur);
  g_cache.clear();
  g_cache_bytes = 0;
}
} // namespace

int pb2_malloc(void **ptr, size_t bytes) {
  PB2_REQUIRE(ptr, "null argument");
  if (int rc = require_device()) return rc;
  if (bytes == 0) bytes = 1;
  int dev = 0;
  PB2_CUDA_CHECK(cudaGetDevice(&dev));
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    auto it = g_cache.find({dev, bytes});
    if (it != g_cache.end()) {
      *ptr = it->second;
      g_cache_bytes -= bytes;
      g_cache.erase(it);
      g_live[*ptr] = Live{bytes, dev};
      return PB2_OK;
    }
  }
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e == cudaErrorMemoryAllocation) { // give the cache back and try once more
    cudaGetLastError();
    std::lock_guard<std::mutex> lk(g_cache_mu);
    cache_release_all_locked();
    e = cudaMalloc(ptr, bytes);
  }
  PB2_CUDA_CHECK(e);
  std::lock_guard<std::mutex> lk(g_cache_mu);
  g_live[*ptr] = Live{bytes, dev};
  return PB2_OK;
}
int pb2_free(void *ptr) {
  if (!ptr) return PB2_OK;
  Live live{0, -1};
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    auto it = g_live.find(ptr);
    if (it != g_live.end()) {
      live = it->second;
      g_live.erase(it);
    }
  }
  if (live.bytes >= kCacheMinBytes && live.device >= 0) {
    // what cudaFree would have waited for: the device that OWNS the pointer
    int cur = 0;
    PB2_CUDA_CHECK(cudaGetDevice(&cur));
    if (cur != live.device) PB2_CUDA_CHECK(cudaSetDevice(live.device));
    const cudaError_t e = cudaDeviceSynchronize();
    if (cur != live.device) cudaSetDevice(cur);
    PB2_CUDA_CHECK(e);
    std::lock_guard<std::mutex> lk(g_cache_mu);
    if (g_cache_bytes + live.bytes <= cache_max_bytes() && g_cache.size() < kCacheMaxEntries) {
      g_cache.emplace(std::make_pair(live.device, live.bytes), ptr);
      g_cache_bytes += live.bytes;
      return PB2_OK;
    }
  }
  PB2_CUDA_CHECK(cudaFree(ptr));
  return PB2_OK;
}
int pb2_cache_trim(void) {
  std::lock_guard<std::mutex> lk(g_cache_mu);
  cache_release_all_locked();
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
int pb2_stream_create_priority(pb2_stream_t *stream, int high) {
  PB2_REQUIRE(stream, "null argument");
  if (int rc = require_device()) return rc;
  int lo = 0, hi = 0; // numerically lower = higher priority
  PB2_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  cudaStream_t s;
  PB2_CUDA_CHECK(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, high ? hi : lo));
  *stream = s;
  return PB2_OK;
}
namespace pb2 {
// one thread polls a device counter: the stream it is launched on waits, on the device, for work
// that another stream's kernel reports from inside (no host, no kernel boundary in between)
__global__ void wait_value_kernel(const volatile int *counter, const int target, int *timed_out) {
  long long spins = 0;
  while (*counter < target) {
    __nanosleep(500);
    if (++spins > 40000000ll) { // ~20 s: a producer that never reports must not hang the device
      if (timed_out) *timed_out = 1;
      break;
    }
  }
  __threadfence();
}
} // namespace pb2
int pb2_stream_wait_value(pb2_stream_t stream, const int32_t *counter, int32_t target) {
  PB2_REQUIRE(counter, "null counter");
  if (int rc = require_device()) return rc;
  pb2::wait_value_kernel<<<1, 1, 0, as_stream(stream)>>>(counter, target, nullptr);
  PB2_LAUNCH_CHECK();
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
