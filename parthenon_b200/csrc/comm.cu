// comm.cu — inter-GPU halo slabs over NCCL (NVLink 5 / NVSwitch).
//
// Replaces the per-(block pair, variable) MPI_Isend/MPI_Irecv channels of CommBuffer
// (reference src/utils/communication_buffer.hpp:209-406): the pack kernel writes every
// non-local region at a precomputed offset inside ONE contiguous slab per peer, and one
// grouped ncclSend/ncclRecv pair per peer ships it (<= 14 operations per exchange at 8 GPUs
// instead of thousands of messages).  The dt reduction (MPI_Allreduce, driver.cpp:237)
// becomes an in-place ncclAllReduce(min) on one double.
//
// NCCL is bound at run time with dlopen so that the library also loads on boxes (and CPU
// test containers) that have no NCCL; the torch-bundled libnccl.so.2 is already resident
// in processes that imported torch.
#include <dlfcn.h>

#include <cstdlib>
#include <string>

#include "common.cuh"

namespace pb2 {

typedef struct {
  char internal[128];
} ncclUniqueId_t;
typedef void *ncclComm_p;
enum { kNcclFloat64 = 8, kNcclSum = 0, kNcclMin = 3 };

struct NcclApi {
  void *handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId_t *) = nullptr;
  int (*CommInitRank)(ncclComm_p *, int, ncclUniqueId_t, int) = nullptr;
  int (*CommDestroy)(ncclComm_p) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void *, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};

static NcclApi g_nccl;

static int load_nccl() {
  if (g_nccl.handle) return PB2_OK;
  const char *names[] = {getenv("PB2_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *n : names) {
    if (!n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    set_error("cannot dlopen libnccl.so.2 (set PB2_NCCL_LIB): %s", dlerror());
    return PB2_ERR_NCCL;
  }
#define PB2_SYM(field, name)                                                              \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));                \
  if (!g_nccl.field) {                                                                    \
    set_error("libnccl lacks %s", name);                                                  \
    return PB2_ERR_NCCL;                                                                  \
  }
  PB2_SYM(GetUniqueId, "ncclGetUniqueId")
  PB2_SYM(CommInitRank, "ncclCommInitRank")
  PB2_SYM(CommDestroy, "ncclCommDestroy")
  PB2_SYM(GroupStart, "ncclGroupStart")
  PB2_SYM(GroupEnd, "ncclGroupEnd")
  PB2_SYM(Send, "ncclSend")
  PB2_SYM(Recv, "ncclRecv")
  PB2_SYM(AllReduce, "ncclAllReduce")
  PB2_SYM(GetErrorString, "ncclGetErrorString")
#undef PB2_SYM
  g_nccl.handle = h;
  return PB2_OK;
}

#define PB2_NCCL_CHECK(expr)                                                              \
  do {                                                                                    \
    int r__ = (expr);                                                                     \
    if (r__ != 0) {                                                                       \
      set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                        \
                g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?");                \
      return PB2_ERR_NCCL;                                                                \
    }                                                                                     \
  } while (0)

} // namespace pb2

struct pb2_comm {
  pb2::ncclComm_p comm;
  int rank, nranks;
  double *barrier_scratch = nullptr; // one device double of this communicator's device
};

// ---- peer push (include/parthenon_b200.h): CUDA IPC mappings and the flag kernels --------------
#include <map>
#include <mutex>

namespace pb2 {
struct IpcMapping {
  void *base = nullptr;
  int refs = 0;
};
static std::mutex g_ipc_mutex;
static std::map<std::string, IpcMapping> g_ipc_map; // key: the 64 handle bytes

// thread i: "ready for exchange seq" into peer i's flags, then wait for peer i's own ready flag
__global__ void peer_handshake_kernel(int32_t *const *__restrict__ peer_flags,
                                      const int32_t *my_flags, const int32_t *__restrict__ peers,
                                      int npeers, int me, int32_t seq) {
  const int i = threadIdx.x;
  if (i >= npeers) return;
  __threadfence_system();
  *reinterpret_cast<volatile int32_t *>(peer_flags[i] + me) = seq;
  const volatile int32_t *f = my_flags + peers[i];
  long long spins = 0;
  while (*f < seq) {
    __nanosleep(200);
    if (++spins > 100000000ll) break; // ~20 s: never hang the device on a lost peer
  }
  __threadfence_system();
}

__global__ void peer_signal_kernel(int32_t *const *__restrict__ peer_flags, int npeers, int slot,
                                   int32_t seq) {
  const int i = threadIdx.x;
  if (i >= npeers) return;
  __threadfence_system();
  *reinterpret_cast<volatile int32_t *>(peer_flags[i] + slot) = seq;
}

__global__ void peer_wait_kernel(const int32_t *my_flags, const int32_t *__restrict__ peers,
                                 int npeers, int32_t seq) {
  const int i = threadIdx.x;
  if (i >= npeers) return;
  const volatile int32_t *f = my_flags + peers[i];
  long long spins = 0;
  while (*f < seq) {
    __nanosleep(200);
    if (++spins > 100000000ll) break;
  }
  __threadfence_system();
}
} // namespace pb2

using namespace pb2;

extern "C" {

int pb2_comm_unique_id(uint8_t id[PB2_NCCL_UNIQUE_ID_BYTES]) {
  PB2_REQUIRE(id, "null argument");
  if (int rc = load_nccl()) return rc;
  ncclUniqueId_t u;
  PB2_NCCL_CHECK(g_nccl.GetUniqueId(&u));
  memcpy(id, u.internal, PB2_NCCL_UNIQUE_ID_BYTES);
  return PB2_OK;
}

int pb2_comm_create(pb2_comm **comm, int rank, int nranks,
                    const uint8_t id[PB2_NCCL_UNIQUE_ID_BYTES]) {
  PB2_REQUIRE(comm && id && nranks >= 1 && rank >= 0 && rank < nranks, "bad arguments");
  if (int rc = require_device()) return rc;
  if (int rc = load_nccl()) return rc;
  ncclUniqueId_t u;
  memcpy(u.internal, id, PB2_NCCL_UNIQUE_ID_BYTES);
  ncclComm_p c = nullptr;
  PB2_NCCL_CHECK(g_nccl.CommInitRank(&c, nranks, u, rank));
  *comm = new pb2_comm{c, rank, nranks};
  return PB2_OK;
}

int pb2_comm_destroy(pb2_comm *comm) {
  if (!comm) return PB2_OK;
  if (g_nccl.CommDestroy) g_nccl.CommDestroy(comm->comm);
  if (comm->barrier_scratch) cudaFree(comm->barrier_scratch);
  delete comm;
  return PB2_OK;
}

int pb2_comm_exchange(pb2_comm *comm, const double *send_slab, const int64_t *send_off,
                      double *recv_slab, const int64_t *recv_off, pb2_stream_t stream) {
  PB2_REQUIRE(comm && send_off && recv_off, "bad arguments");
  PB2_NCCL_CHECK(g_nccl.GroupStart());
  // an error inside the group must still close it, or every later call on this thread hangs
  int bad = 0;
  const char *what = "";
  for (int p = 0; p < comm->nranks && !bad; ++p) {
    if (p == comm->rank) continue;
    const int64_t ns = send_off[p + 1] - send_off[p], nr = recv_off[p + 1] - recv_off[p];
    if (ns > 0 && (bad = g_nccl.Send(send_slab + send_off[p], static_cast<size_t>(ns),
                                     kNcclFloat64, p, comm->comm, as_stream(stream))) != 0)
      what = "ncclSend";
    if (!bad && nr > 0 &&
        (bad = g_nccl.Recv(recv_slab + recv_off[p], static_cast<size_t>(nr), kNcclFloat64, p,
                           comm->comm, as_stream(stream))) != 0)
      what = "ncclRecv";
  }
  const int end = g_nccl.GroupEnd();
  if (bad || end) {
    const int r = bad ? bad : end;
    set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, bad ? what : "ncclGroupEnd",
              g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    return PB2_ERR_NCCL;
  }
  return PB2_OK;
}

int pb2_comm_allreduce_min(pb2_comm *comm, double *dev_value, pb2_stream_t stream) {
  PB2_REQUIRE(comm && dev_value, "bad arguments");
  PB2_NCCL_CHECK(g_nccl.AllReduce(dev_value, dev_value, 1, kNcclFloat64, kNcclMin, comm->comm,
                                  as_stream(stream)));
  return PB2_OK;
}

int pb2_comm_allreduce_sum(pb2_comm *comm, double *dev_values, int64_t n, pb2_stream_t stream) {
  PB2_REQUIRE(comm && dev_values && n >= 0, "bad arguments");
  if (n == 0) return PB2_OK;
  PB2_NCCL_CHECK(g_nccl.AllReduce(dev_values, dev_values, static_cast<size_t>(n), kNcclFloat64,
                                  kNcclSum, comm->comm, as_stream(stream)));
  return PB2_OK;
}

int pb2_comm_barrier(pb2_comm *comm, pb2_stream_t stream) {
  PB2_REQUIRE(comm, "bad arguments");
  if (!comm->barrier_scratch) {
    PB2_CUDA_CHECK(cudaMalloc(&comm->barrier_scratch, sizeof(double)));
    PB2_CUDA_CHECK(cudaMemset(comm->barrier_scratch, 0, sizeof(double)));
  }
  PB2_NCCL_CHECK(g_nccl.AllReduce(comm->barrier_scratch, comm->barrier_scratch, 1, kNcclFloat64,
                                  kNcclMin, comm->comm, as_stream(stream)));
  // a barrier for the HOST: every rank's work enqueued before it on `stream` has completed
  PB2_CUDA_CHECK(cudaStreamSynchronize(as_stream(stream)));
  return PB2_OK;
}

int pb2_ipc_export(const void *ptr, pb2_ipc_handle *handle) {
  PB2_REQUIRE(ptr && handle, "bad arguments");
  if (int rc = require_device()) return rc;
  static_assert(sizeof(cudaIpcMemHandle_t) == sizeof(handle->bytes), "handle size");
  cudaIpcMemHandle_t h;
  PB2_CUDA_CHECK(cudaIpcGetMemHandle(&h, const_cast<void *>(ptr)));
  memcpy(handle->bytes, &h, sizeof(h));
  // the handle names the whole allocation: find the pointer's offset inside it
  typedef int (*range_fn)(unsigned long long *, size_t *, unsigned long long);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  PB2_CUDA_CHECK(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr));
  PB2_REQUIRE(fn != nullptr, "cuMemGetAddressRange is not available");
  unsigned long long base = 0;
  size_t size = 0;
  const int rc = reinterpret_cast<range_fn>(fn)(&base, &size,
                                                reinterpret_cast<unsigned long long>(ptr));
  PB2_REQUIRE(rc == 0, "cuMemGetAddressRange failed");
  handle->offset = static_cast<int64_t>(reinterpret_cast<unsigned long long>(ptr) - base);
  return PB2_OK;
}

int pb2_ipc_open(const pb2_ipc_handle *handle, void **ptr) {
  PB2_REQUIRE(handle && ptr, "bad arguments");
  if (int rc = require_device()) return rc;
  std::lock_guard<std::mutex> lock(g_ipc_mutex);
  const std::string key(reinterpret_cast<const char *>(handle->bytes), sizeof(handle->bytes));
  IpcMapping &m = g_ipc_map[key];
  if (m.refs == 0) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle->bytes, sizeof(h));
    void *base = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      g_ipc_map.erase(key);
      (void)cudaGetLastError();
      set_error("cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
      return PB2_ERR_CUDA;
    }
    m.base = base;
  }
  m.refs++;
  *ptr = static_cast<char *>(m.base) + handle->offset;
  return PB2_OK;
}

int pb2_ipc_close(const pb2_ipc_handle *handle) {
  PB2_REQUIRE(handle, "bad arguments");
  std::lock_guard<std::mutex> lock(g_ipc_mutex);
  const std::string key(reinterpret_cast<const char *>(handle->bytes), sizeof(handle->bytes));
  auto it = g_ipc_map.find(key);
  if (it == g_ipc_map.end()) return PB2_OK;
  if (--it->second.refs <= 0) {
    PB2_CUDA_CHECK(cudaIpcCloseMemHandle(it->second.base));
    g_ipc_map.erase(it);
  }
  return PB2_OK;
}

int pb2_peer_handshake(int32_t *const *peer_flags, const int32_t *my_flags, const int32_t *peers,
                       int npeers, int me, int nranks, int32_t seq, pb2_stream_t stream) {
  PB2_REQUIRE(npeers >= 0 && npeers <= 1024 && me >= 0 && me < nranks, "bad arguments");
  if (npeers == 0) return PB2_OK;
  PB2_REQUIRE(peer_flags && my_flags && peers, "null argument");
  if (int rc = require_device()) return rc;
  peer_handshake_kernel<<<1, ((npeers + 31) / 32) * 32, 0, as_stream(stream)>>>(
      peer_flags, my_flags, peers, npeers, me, seq);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

int pb2_peer_signal(int32_t *const *peer_flags, int npeers, int me, int nranks, int32_t seq,
                    pb2_stream_t stream) {
  PB2_REQUIRE(npeers >= 0 && npeers <= 1024 && me >= 0 && me < nranks, "bad arguments");
  if (npeers == 0) return PB2_OK;
  PB2_REQUIRE(peer_flags, "null argument");
  if (int rc = require_device()) return rc;
  peer_signal_kernel<<<1, ((npeers + 31) / 32) * 32, 0, as_stream(stream)>>>(peer_flags, npeers,
                                                                           nranks + me, seq);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

int pb2_peer_wait(const int32_t *my_flags, const int32_t *peers, int npeers, int nranks,
                  int32_t seq, pb2_stream_t stream) {
  PB2_REQUIRE(npeers >= 0 && npeers <= 1024 && nranks >= 1, "bad arguments");
  if (npeers == 0) return PB2_OK;
  PB2_REQUIRE(my_flags && peers, "null argument");
  if (int rc = require_device()) return rc;
  peer_wait_kernel<<<1, ((npeers + 31) / 32) * 32, 0, as_stream(stream)>>>(my_flags + nranks, peers,
                                                                         npeers, seq);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

} // extern "C"
