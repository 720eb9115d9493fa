// exchange.cu — ghost-zone pack / unpack / fused same-device copy.
//
// Replaces the Kokkos TeamPolicy kernels of SendBoundBufs
// (reference src/bvals/comms/boundary_communication.cpp:95-140) and SetBounds (:273-334).
// Design (B200): one launch per table covers every region of a MeshData.  Regions are cut
// on the host into fixed-size chunks (one CTA each) so that a 360 KB face region and a
// 5.6 KB corner region load the SMs evenly; each thread moves 16-byte vectors, issues all
// its loads before its stores (4 independent 16 B loads in flight per thread), and turns
// the flat buffer index into (c,k,j,i) with precomputed multiply-shift divisions.  The
// buffer side is fully coalesced; the array side touches whole 32 B sectors (ghost rows of
// 4 doubles are sector aligned because is, ie+1 and the row pitch are multiples of 4).
#include <mutex>
#include <vector>

#include "common.cuh"
#include "tables.cuh"

namespace pb2 {

__device__ __forceinline__ void decompose(const DevRegion &r, uint32_t v, uint32_t &c,
                                          uint32_t &k, uint32_t &j, uint32_t &iv) {
  uint32_t t;
  r.dni.divmod(v, t, iv);
  uint32_t t2;
  r.dnj.divmod(t, t2, j);
  r.dnk.divmod(t2, c, k);
}

template <int V>
struct Vec;
template <>
struct Vec<1> {
  using type = double;
};
template <>
struct Vec<2> {
  using type = double2;
};

__device__ __forceinline__ bool above(double x, double thr) { return fabs(x) >= thr; }
__device__ __forceinline__ bool above(double2 x, double thr) {
  return fabs(x.x) >= thr || fabs(x.y) >= thr;
}
__device__ __forceinline__ void splat(double &x, double v) { x = v; }
__device__ __forceinline__ void splat(double2 &x, double v) { x.x = x.y = v; }

// ---- pack: buf <- var ------------------------------------------------------------------
template <int V>
__device__ __forceinline__ void pack_chunk(const DevRegion &r, uint32_t first, double *buf,
                                           int32_t *flags) {
  using T = typename Vec<V>::type;
  T val[kUnroll];
  bool ok[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const uint32_t v = first + u * kThreads + threadIdx.x;
    ok[u] = v < r.total_vec;
    if (ok[u]) {
      uint32_t c, k, j, iv;
      decompose(r, v, c, k, j, iv);
      const int64_t off = (int64_t)c * r.sc + (int64_t)(k + r.s[2]) * r.sk +
                          (int64_t)(j + r.s[1]) * r.sj + (r.s[0] + (int32_t)iv * V);
      val[u] = __ldg(reinterpret_cast<const T *>(r.var + off));
    }
  }
  bool nz = false;
  T *out = reinterpret_cast<T *>(buf + r.buf_off);
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const uint32_t v = first + u * kThreads + threadIdx.x;
    if (ok[u]) {
      __stcs(out + v, val[u]);
      nz = nz || above(val[u], r.value);
    }
  }
  if (flags != nullptr && r.flag_slot >= 0) {
    if (__syncthreads_or(nz) && threadIdx.x == 0) atomicOr(flags + r.flag_slot, 1);
  }
}

__global__ void __launch_bounds__(kThreads)
    pack_kernel(const DevRegion *__restrict__ regions, const Chunk *__restrict__ chunks,
                double *__restrict__ buf, int32_t *__restrict__ flags) {
  const Chunk ch = chunks[blockIdx.x];
  const DevRegion &r = regions[ch.region];
  // unallocated / same_to_same regions send nothing (boundary_communication.cpp:100-104)
  if (!(r.status & PB2_REGION_ALLOCATED) || (r.status & PB2_REGION_SAME_TO_SAME)) return;
  if (r.vec == 2)
    pack_chunk<2>(r, ch.first_vec, buf, flags);
  else
    pack_chunk<1>(r, ch.first_vec, buf, flags);
}

// ---- unpack: var <- buf (or the sparse default when the message was null) ----------------
template <int V>
__device__ __forceinline__ void unpack_chunk(const DevRegion &r, uint32_t first,
                                             const double *buf, bool has_data) {
  using T = typename Vec<V>::type;
  T val[kUnroll];
  const T *in = reinterpret_cast<const T *>(buf + r.buf_off);
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const uint32_t v = first + u * kThreads + threadIdx.x;
    if (v < r.total_vec) {
      if (has_data)
        val[u] = __ldcs(in + v);
      else
        splat(val[u], r.value);
    }
  }
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const uint32_t v = first + u * kThreads + threadIdx.x;
    if (v < r.total_vec) {
      uint32_t c, k, j, iv;
      decompose(r, v, c, k, j, iv);
      const int64_t off = (int64_t)c * r.sc + (int64_t)(k + r.s[2]) * r.sk +
                          (int64_t)(j + r.s[1]) * r.sj + (r.s[0] + (int32_t)iv * V);
      *reinterpret_cast<T *>(r.var + off) = val[u];
    }
  }
}

// ... through the neighbour tree's LogicalCoordinateTransformation: the box is enumerated in
// the sender's orientation, every element goes to InverseTransform({i, j, k}) times `fac`
// (boundary_communication.cpp:282-308, logical_coordinate_transformation.hpp:90-99).  Scalar
// accesses: a permuted box is not contiguous along i.
__device__ __forceinline__ void unpack_chunk_transformed(const DevRegion &r, uint32_t first,
                                                         const double *buf, bool has_data) {
  const double *in = buf + r.buf_off;
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const uint32_t v = first + u * kThreads + threadIdx.x;
    if (v < r.total_vec) {
      const double x = has_data ? r.fac * __ldcs(in + v) : r.value;
      uint32_t c, k, j, i;
      decompose(r, v, c, k, j, i);
      const int lin[3] = {r.s[0] + (int)i, r.s[1] + (int)j, r.s[2] + (int)k};
      int out[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const int q = lin[r.tr_dir[d]];
        out[d] = r.tr_flip[d] ? r.tr_ncell - 1 - q : q;
      }
      r.var[(int64_t)c * r.sc + (int64_t)out[2] * r.sk + (int64_t)out[1] * r.sj + out[0]] = x;
    }
  }
}

__global__ void __launch_bounds__(kThreads)
    unpack_kernel(const DevRegion *__restrict__ regions, const Chunk *__restrict__ chunks,
                  const double *__restrict__ buf, const int32_t *__restrict__ data_flags) {
  const Chunk ch = chunks[blockIdx.x];
  const DevRegion &r = regions[ch.region];
  if (r.status & PB2_REGION_SAME_TO_SAME) return; // :280
  if (!(r.status & PB2_REGION_ALLOCATED)) return; // :290,:311
  bool has_data = (r.status & PB2_REGION_BUF_ALLOCATED) != 0;
  if (data_flags != nullptr && r.flag_slot >= 0) has_data = data_flags[r.flag_slot] != 0;
  if (r.tr_on)
    unpack_chunk_transformed(r, ch.first_vec, buf, has_data);
  else if (r.vec == 2)
    unpack_chunk<2>(r, ch.first_vec, buf, has_data);
  else
    unpack_chunk<1>(r, ch.first_vec, buf, has_data);
}

// ---- fused copy: receiver box <- sender box ------------------------------------------------
// MODE 0: copy, and report "any |x| >= threshold" per region (dense fields, legacy)
// MODE 1: report only — the sender's half of a sparse exchange (SendBoundBufs decides between
//         Send and SendNull, boundary_communication.cpp:95-157); nothing is written
// MODE 2: deliver — the receiver's half (SetBounds :273-334): regions whose flag is set get
//         the data, the others (null message, or unallocated sender) the sparse default;
//         unallocated receivers are skipped
template <int V, int MODE>
__device__ __forceinline__ void copy_chunk(const DevRegion &r, uint32_t first,
                                           int32_t *flags) {
  using T = typename Vec<V>::type;
  T val[kUnroll];
  int64_t doff[kUnroll];
  bool ok[kUnroll];
  bool src_alloc = (r.status & PB2_REGION_ALLOCATED) != 0;
  if (MODE == 1 && !src_alloc) return; // flag stays 0: SendNull
  if (MODE == 2) {
    if (r.status & PB2_REGION_DST_UNALLOCATED) return;
    if (flags != nullptr && r.flag_slot >= 0) src_alloc = src_alloc && flags[r.flag_slot] != 0;
  }
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const uint32_t v = first + u * kThreads + threadIdx.x;
    ok[u] = v < r.total_vec;
    if (ok[u]) {
      uint32_t c, k, j, iv;
      decompose(r, v, c, k, j, iv);
      const int64_t soff = (int64_t)c * r.ssc + (int64_t)(k + r.ss[2]) * r.ssk +
                           (int64_t)(j + r.ss[1]) * r.ssj + (r.ss[0] + (int32_t)iv * V);
      doff[u] = (int64_t)c * r.sc + (int64_t)(k + r.s[2]) * r.sk +
                (int64_t)(j + r.s[1]) * r.sj + (r.s[0] + (int32_t)iv * V);
      if (src_alloc)
        val[u] = __ldg(reinterpret_cast<const T *>(r.src + soff));
      else
        splat(val[u], r.default_value);
    }
  }
  bool nz = false;
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    if (ok[u]) {
      if (MODE != 1) *reinterpret_cast<T *>(r.var + doff[u]) = val[u];
      nz = nz || above(val[u], r.value);
    }
  }
  if (MODE != 2 && flags != nullptr && r.flag_slot >= 0) {
    if (__syncthreads_or(nz) && threadIdx.x == 0) atomicOr(flags + r.flag_slot, 1);
  }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads)
    copy_kernel(const DevRegion *__restrict__ regions, const Chunk *__restrict__ chunks,
                int32_t *flags) {
  const Chunk ch = chunks[blockIdx.x];
  const DevRegion &r = regions[ch.region];
  if (r.status & PB2_REGION_SAME_TO_SAME) return;
  if (r.vec == 2)
    copy_chunk<2, MODE>(r, ch.first_vec, flags);
  else
    copy_chunk<1, MODE>(r, ch.first_vec, flags);
}


// ---- peer push: copy whose destinations lie in the memory of other GPUs, with arrival flags ---
// Same chunks as copy_kernel<0>.  Every thread fences its stores at system scope; the thread
// block that finishes last (device counter) writes the exchange number into the arrival flag of
// every peer, so a peer that sees its flag also sees the ghost cells (threadFenceReduction
// pattern across NVLink).
__global__ void __launch_bounds__(kThreads)
    copy_signal_kernel(const DevRegion *__restrict__ regions, const Chunk *__restrict__ chunks,
                       int32_t *counter, int32_t *const *__restrict__ peer_flags, int npeers,
                       int slot, int32_t seq) {
  __shared__ int s_last;
  const Chunk ch = chunks[blockIdx.x];
  const DevRegion &r = regions[ch.region];
  if (!(r.status & PB2_REGION_SAME_TO_SAME)) {
    if (r.vec == 2)
      copy_chunk<2, 0>(r, ch.first_vec, nullptr);
    else
      copy_chunk<1, 0>(r, ch.first_vec, nullptr);
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(counter, 1) == static_cast<int>(gridDim.x) - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  for (int p = threadIdx.x; p < npeers; p += kThreads)
    *reinterpret_cast<volatile int32_t *>(peer_flags[p] + slot) = seq;
  if (threadIdx.x == 0) *counter = 0;
}


// ---- uniform meshes: descriptor-free ghost fill ---------------------------------------------
// One CTA per (block, component, part of the ghost shell).  The shell is enumerated as
//   [low k planes][high k planes][per interior plane: low j rows, high j rows, x pieces]
// in units of V doubles; a thread turns its flat index into (k, j, i) with multiply-shift
// divisions, picks the owning neighbour from the 27-entry table of its block (shared memory)
// and copies V doubles.  All kHaloUnroll loads of a thread are issued before its stores.
#ifndef PB2_HALO_UNROLL
#define PB2_HALO_UNROLL 16
#endif
#ifndef PB2_HALO_MINB
#define PB2_HALO_MINB 2
#endif
constexpr int kHaloUnroll = PB2_HALO_UNROLL;

struct HaloGeom {
  int32_t nblocks, ncomp, parts;
  int32_t nx[3], ng[3], n[3]; // interior, ghost width (0 in symmetry directions), total
  int64_t sj, sk, sc, sb;
  uint32_t nivec;     // vectors per full row
  uint32_t xg;        // x-ghost vectors per row (both sides)
  uint32_t xg_half;   // ... per side
  uint32_t plane_full; // vectors in a full k plane
  uint32_t plane_in;   // ghost vectors in an interior k plane
  uint32_t rows_full;  // vectors in the 2 ng_j full rows of an interior plane
  uint32_t nA;         // vectors in the low (== high) k-ghost planes
  uint32_t total;      // ghost vectors per (block, component)
  uint32_t per_part;
  FastDiv d_nivec, d_plane_full, d_plane_in, d_xg;
};

template <int V>
__global__ void __launch_bounds__(kThreads, PB2_HALO_MINB)
    halo_uniform_kernel(const HaloGeom g, double *__restrict__ field,
                        const int32_t *__restrict__ nbr) {
  using T = typename Vec<V>::type;
  __shared__ int32_t s_nbr[27];
  const uint32_t part = blockIdx.x % g.parts;
  const uint32_t bc = blockIdx.x / g.parts;
  const uint32_t c = bc % g.ncomp, b = bc / g.ncomp;
  if (threadIdx.x < 27) s_nbr[threadIdx.x] = nbr[b * 27 + threadIdx.x];
  __syncthreads();
  const int64_t dbase = (int64_t)b * g.sb + (int64_t)c * g.sc;
  const uint32_t end = min(g.total, (part + 1) * g.per_part);
  for (uint32_t v0 = part * g.per_part + threadIdx.x; v0 < end; v0 += kThreads * kHaloUnroll) {
    T val[kHaloUnroll];
    int64_t doff[kHaloUnroll];
    bool ok[kHaloUnroll];
#pragma unroll
    for (int u = 0; u < kHaloUnroll; ++u) {
      const uint32_t v = v0 + u * kThreads;
      ok[u] = v < end;
      if (ok[u]) {
        uint32_t k, j, iv, r;
        if (v < 2 * g.nA) { // k-ghost planes
          const uint32_t w = v < g.nA ? v : v - g.nA;
          g.d_plane_full.divmod(w, k, r);
          if (v >= g.nA) k += g.ng[2] + g.nx[2];
          g.d_nivec.divmod(r, j, iv);
        } else {
          g.d_plane_in.divmod(v - 2 * g.nA, k, r);
          k += g.ng[2];
          if (r < g.rows_full) { // j-ghost rows
            g.d_nivec.divmod(r, j, iv);
            if (j >= (uint32_t)g.ng[1]) j += g.nx[1];
          } else { // x-ghost pieces of interior rows
            g.d_xg.divmod(r - g.rows_full, j, iv);
            j += g.ng[1];
            if (iv >= g.xg_half) iv += g.nivec - g.xg;
          }
        }
        const int i = (int)iv * V;
        const int ox = i < g.ng[0] ? -1 : (i >= g.ng[0] + g.nx[0] ? 1 : 0);
        const int oy = (int)j < g.ng[1] ? -1 : ((int)j >= g.ng[1] + g.nx[1] ? 1 : 0);
        const int oz = (int)k < g.ng[2] ? -1 : ((int)k >= g.ng[2] + g.nx[2] ? 1 : 0);
        const int sb = s_nbr[(ox + 1) + 3 * (oy + 1) + 9 * (oz + 1)];
        ok[u] = sb >= 0;
        if (ok[u]) {
          const int64_t cell = (int64_t)k * g.sk + (int64_t)j * g.sj + i;
          doff[u] = dbase + cell;
          const int64_t soff = (int64_t)sb * g.sb + (int64_t)c * g.sc + cell -
                               ((int64_t)oz * g.nx[2] * g.sk + (int64_t)oy * g.nx[1] * g.sj +
                                ox * g.nx[0]);
          val[u] = __ldg(reinterpret_cast<const T *>(field + soff));
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kHaloUnroll; ++u)
      if (ok[u]) *reinterpret_cast<T *>(field + doff[u]) = val[u];
  }
}

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Region tables are assembled in one pinned staging buffer (grown on demand, kept): an adaptive
// mesh re-creates tables of tens of thousands of regions (several MB) on every remesh, and a
// copy from pageable memory runs at a fraction of the PCIe rate.
struct Stage {
  std::mutex mu;
  void *ptr = nullptr;
  size_t cap = 0;
  std::vector<DevRegion> fallback;
  DevRegion *regions(size_t n) { // call with mu held
    const size_t bytes = n * sizeof(DevRegion);
    if (bytes > cap) {
      if (ptr) cudaFreeHost(ptr);
      ptr = nullptr;
      cap = 0;
      const size_t want = bytes + bytes / 2 + (size_t(1) << 20);
      if (cudaMallocHost(&ptr, want) == cudaSuccess) {
        cap = want;
      } else {
        cudaGetLastError();
        ptr = nullptr;
      }
    }
    if (ptr) return static_cast<DevRegion *>(ptr);
    fallback.resize(n);
    return fallback.data();
  }
};
static Stage g_stage;

static int build_table(pb2_bnd_table **out, const DevRegion *regs, size_t nregs, int kind) {
  std::vector<Chunk> chunks;
  chunks.reserve(nregs + nregs / 4);
  int64_t elements = 0;
  for (size_t r = 0; r < nregs; ++r) {
    const uint32_t per_chunk = kThreads * kUnroll;
    elements += (int64_t)regs[r].total_vec * regs[r].vec;
    for (uint32_t v = 0; v < regs[r].total_vec; v += per_chunk)
      chunks.push_back(Chunk{static_cast<int32_t>(r), v});
  }
  auto *t = new pb2_bnd_table();
  t->kind = kind;
  t->nregions = static_cast<int64_t>(nregs);
  t->nchunks = static_cast<int64_t>(chunks.size());
  t->elements = elements;
  t->d_regions = nullptr;
  t->d_chunks = nullptr;
  t->d_prores = nullptr;
  if (nregs > 0) {
    cudaError_t e = table_alloc(reinterpret_cast<void **>(&t->d_regions), nregs * sizeof(DevRegion));
    if (e == cudaSuccess) e = table_alloc(reinterpret_cast<void **>(&t->d_chunks), (chunks.size() + 1) * sizeof(Chunk));
    if (e == cudaSuccess)
      e = cudaMemcpy(t->d_regions, regs, nregs * sizeof(DevRegion), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !chunks.empty())
      e = cudaMemcpy(t->d_chunks, chunks.data(), chunks.size() * sizeof(Chunk),
                     cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      set_error("table upload failed: %s", cudaGetErrorString(e));
      table_free(t->d_regions);
      table_free(t->d_chunks);
      delete t;
      return PB2_ERR_CUDA;
    }
  }
  *out = t;
  return PB2_OK;
}

} // namespace pb2

using namespace pb2;

extern "C" {

int pb2_bnd_table_create(pb2_bnd_table **table, const pb2_bnd_region *regions, int64_t n) {
  PB2_REQUIRE(table && (regions || n == 0) && n >= 0, "bad arguments");
  if (int rc = require_device()) return rc;
  std::lock_guard<std::mutex> stage_lock(g_stage.mu);
  DevRegion *regs = g_stage.regions(static_cast<size_t>(n));
  for (int64_t i = 0; i < n; ++i) {
    const pb2_bnd_region &a = regions[i];
    DevRegion &d = regs[i];
    memset(&d, 0, sizeof(d));
    PB2_REQUIRE(a.n[0] >= 0 && a.n[1] >= 0 && a.n[2] >= 0 && a.ncomp >= 0, "negative extent");
    const int64_t total = (int64_t)a.ncomp * a.n[0] * a.n[1] * a.n[2];
    PB2_REQUIRE(total < (1ll << 31), "region too large");
    d.var = a.var;
    d.buf_off = a.buf_off;
    for (int q = 0; q < 3; ++q) d.s[q] = a.s[q];
    d.sj = a.stride_j;
    d.sk = a.stride_k;
    d.sc = a.stride_c;
    const bool v2 = (a.n[0] % 2 == 0) && (a.s[0] % 2 == 0) && (a.stride_j % 2 == 0) &&
                    (a.stride_k % 2 == 0) && (a.stride_c % 2 == 0) && (a.buf_off % 2 == 0) &&
                    aligned16(a.var) && !a.lcoord_on;
    if (a.lcoord_on) {
      bool used[3] = {false, false, false};
      for (int q = 0; q < 3; ++q) {
        PB2_REQUIRE(a.lcoord_dir[q] >= 0 && a.lcoord_dir[q] < 3 && !used[a.lcoord_dir[q]],
                    "lcoord_dir must be a permutation of 0, 1, 2");
        used[a.lcoord_dir[q]] = true;
        d.tr_dir[q] = a.lcoord_dir[q];
        d.tr_flip[q] = a.lcoord_flip[q] != 0;
      }
      d.tr_on = 1;
      d.tr_ncell = a.lcoord_ncell;
      d.fac = a.fac;
    }
    d.vec = v2 ? 2 : 1;
    d.dni.init(static_cast<uint32_t>(a.n[0] / (int)d.vec > 0 ? a.n[0] / (int)d.vec : 1));
    d.dnj.init(static_cast<uint32_t>(a.n[1] > 0 ? a.n[1] : 1));
    d.dnk.init(static_cast<uint32_t>(a.n[2] > 0 ? a.n[2] : 1));
    d.total_vec = static_cast<uint32_t>(total / d.vec);
    d.flag_slot = a.flag_slot;
    d.status = a.status;
    d.value = a.value;
  }
  return build_table(table, regs, static_cast<size_t>(n), kBnd);
}

int pb2_copy_table_create(pb2_bnd_table **table, const pb2_copy_region *regions, int64_t n) {
  PB2_REQUIRE(table && (regions || n == 0) && n >= 0, "bad arguments");
  if (int rc = require_device()) return rc;
  std::lock_guard<std::mutex> stage_lock(g_stage.mu);
  DevRegion *regs = g_stage.regions(static_cast<size_t>(n));
  for (int64_t i = 0; i < n; ++i) {
    const pb2_copy_region &a = regions[i];
    DevRegion &d = regs[i];
    memset(&d, 0, sizeof(d));
    PB2_REQUIRE(a.n[0] >= 0 && a.n[1] >= 0 && a.n[2] >= 0 && a.ncomp >= 0, "negative extent");
    const int64_t total = (int64_t)a.ncomp * a.n[0] * a.n[1] * a.n[2];
    PB2_REQUIRE(total < (1ll << 31), "region too large");
    d.var = a.dst;
    d.src = a.src;
    for (int q = 0; q < 3; ++q) {
      d.s[q] = a.ds[q];
      d.ss[q] = a.ss[q];
    }
    d.sj = a.dst_stride_j;
    d.sk = a.dst_stride_k;
    d.sc = a.dst_stride_c;
    d.ssj = a.src_stride_j;
    d.ssk = a.src_stride_k;
    d.ssc = a.src_stride_c;
    const bool v2 = (a.n[0] % 2 == 0) && (a.ss[0] % 2 == 0) && (a.ds[0] % 2 == 0) &&
                    (a.src_stride_j % 2 == 0) && (a.src_stride_k % 2 == 0) &&
                    (a.src_stride_c % 2 == 0) && (a.dst_stride_j % 2 == 0) &&
                    (a.dst_stride_k % 2 == 0) && (a.dst_stride_c % 2 == 0) &&
                    aligned16(a.src) && aligned16(a.dst);
    d.vec = v2 ? 2 : 1;
    d.dni.init(static_cast<uint32_t>(a.n[0] / (int)d.vec > 0 ? a.n[0] / (int)d.vec : 1));
    d.dnj.init(static_cast<uint32_t>(a.n[1] > 0 ? a.n[1] : 1));
    d.dnk.init(static_cast<uint32_t>(a.n[2] > 0 ? a.n[2] : 1));
    d.total_vec = static_cast<uint32_t>(total / d.vec);
    d.flag_slot = a.flag_slot;
    d.status = a.status;
    d.value = a.threshold;
    d.default_value = a.default_value;
  }
  return build_table(table, regs, static_cast<size_t>(n), kCopy);
}

int pb2_bnd_table_destroy(pb2_bnd_table *table) {
  if (!table) return PB2_OK;
  table_free(table->d_regions);
  table_free(table->d_chunks);
  table_free(table->d_prores);
  table_free(table->d_flxcor);
  table_free(table->d_bc);
  delete table;
  return PB2_OK;
}

int64_t pb2_bnd_table_elements(const pb2_bnd_table *table) {
  return table ? table->elements : 0;
}

int pb2_pack(const pb2_bnd_table *table, double *buf, int32_t *nonzero_flags,
             pb2_stream_t stream) {
  PB2_REQUIRE(table && table->kind == kBnd, "pack needs a boundary table");
  if (table->nchunks == 0) return PB2_OK;
  PB2_REQUIRE(buf, "null buffer");
  ProfScope prof(K_PACK, as_stream(stream), static_cast<double>(table->elements));
  pack_kernel<<<static_cast<unsigned>(table->nchunks), kThreads, 0, as_stream(stream)>>>(
      table->d_regions, table->d_chunks, buf, nonzero_flags);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

int pb2_unpack(const pb2_bnd_table *table, const double *buf, const int32_t *data_flags,
               pb2_stream_t stream) {
  PB2_REQUIRE(table && table->kind == kBnd, "unpack needs a boundary table");
  if (table->nchunks == 0) return PB2_OK;
  PB2_REQUIRE(buf, "null buffer");
  ProfScope prof(K_UNPACK, as_stream(stream), static_cast<double>(table->elements));
  unpack_kernel<<<static_cast<unsigned>(table->nchunks), kThreads, 0, as_stream(stream)>>>(
      table->d_regions, table->d_chunks, buf, data_flags);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

int pb2_copy(const pb2_bnd_table *table, int32_t *nonzero_flags, pb2_stream_t stream) {
  PB2_REQUIRE(table && table->kind == kCopy, "copy needs a copy table");
  if (table->nchunks == 0) return PB2_OK;
  ProfScope prof(K_COPY, as_stream(stream), static_cast<double>(table->elements));
  copy_kernel<0><<<static_cast<unsigned>(table->nchunks), kThreads, 0, as_stream(stream)>>>(
      table->d_regions, table->d_chunks, nonzero_flags);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

int pb2_copy_signal(const pb2_bnd_table *table, int32_t *counter, int32_t *const *peer_flags,
                    int npeers, int me, int nranks, int32_t seq, pb2_stream_t stream) {
  PB2_REQUIRE(table && table->kind == kCopy, "copy_signal needs a copy table");
  PB2_REQUIRE(counter && (peer_flags || npeers == 0) && npeers >= 0 && me >= 0 && me < nranks,
              "bad arguments");
  PB2_REQUIRE(table->nchunks > 0 || npeers == 0, "peers to signal but nothing to copy");
  if (table->nchunks == 0) return PB2_OK;
  ProfScope prof(K_COPY, as_stream(stream), static_cast<double>(table->elements));
  copy_signal_kernel<<<static_cast<unsigned>(table->nchunks), kThreads, 0, as_stream(stream)>>>(
      table->d_regions, table->d_chunks, counter, peer_flags, npeers, nranks + me, seq);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

int pb2_copy_flags(const pb2_bnd_table *table, int32_t *nonzero_flags, pb2_stream_t stream) {
  PB2_REQUIRE(table && table->kind == kCopy && nonzero_flags, "copy_flags needs a copy table");
  if (table->nchunks == 0) return PB2_OK;
  ProfScope prof(K_COPY, as_stream(stream), static_cast<double>(table->elements));
  copy_kernel<1><<<static_cast<unsigned>(table->nchunks), kThreads, 0, as_stream(stream)>>>(
      table->d_regions, table->d_chunks, nonzero_flags);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

int pb2_copy_select(const pb2_bnd_table *table, const int32_t *data_flags, pb2_stream_t stream) {
  PB2_REQUIRE(table && table->kind == kCopy, "copy_select needs a copy table");
  if (table->nchunks == 0) return PB2_OK;
  ProfScope prof(K_COPY, as_stream(stream), static_cast<double>(table->elements));
  copy_kernel<2><<<static_cast<unsigned>(table->nchunks), kThreads, 0, as_stream(stream)>>>(
      table->d_regions, table->d_chunks, const_cast<int32_t *>(data_flags));
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

int pb2_halo_copy_uniform(const pb2_pack_geom *pg, double *field, const int32_t *nbr,
                          pb2_stream_t stream) {
  PB2_REQUIRE(pg && field && nbr, "bad arguments");
  if (int rc = require_device()) return rc;
  if (pg->nblocks == 0 || pg->ncomp == 0) return PB2_OK;
  HaloGeom g;
  memset(&g, 0, sizeof(g));
  g.nblocks = pg->nblocks;
  g.ncomp = pg->ncomp;
  for (int d = 0; d < 3; ++d) {
    const bool sym = d >= pg->ndim;
    g.nx[d] = sym ? 1 : pg->nx[d];
    g.ng[d] = sym ? 0 : pg->ng;
    g.n[d] = g.nx[d] + 2 * g.ng[d];
  }
  g.sj = g.n[0];
  g.sk = (int64_t)g.n[0] * g.n[1];
  g.sc = g.sk * g.n[2];
  g.sb = pg->block_stride;
  const bool v2 = g.nx[0] % 2 == 0 && g.ng[0] % 2 == 0 && g.sb % 2 == 0 && aligned16(field);
  const uint32_t V = v2 ? 2 : 1;
  g.nivec = g.n[0] / V;
  g.xg_half = g.ng[0] / V;
  g.xg = 2 * g.xg_half;
  g.plane_full = g.n[1] * g.nivec;
  g.rows_full = 2 * g.ng[1] * g.nivec;
  g.plane_in = g.rows_full + g.nx[1] * g.xg;
  g.nA = g.ng[2] * g.plane_full;
  const int64_t total = 2 * (int64_t)g.nA + (int64_t)g.nx[2] * g.plane_in;
  PB2_REQUIRE(total < (1ll << 31), "block too large");
  g.total = static_cast<uint32_t>(total);
  if (g.total == 0) return PB2_OK;
  g.d_nivec.init(g.nivec);
  g.d_plane_full.init(g.plane_full > 0 ? g.plane_full : 1);
  g.d_plane_in.init(g.plane_in > 0 ? g.plane_in : 1);
  g.d_xg.init(g.xg > 0 ? g.xg : 1);
  // enough CTAs for ~8 waves of 8 resident CTAs on 148 SMs, each part a whole number of trips
  const uint32_t trip = kThreads * kHaloUnroll;
  int64_t parts = (148 * 8 * 8 + (int64_t)g.nblocks * g.ncomp - 1) / ((int64_t)g.nblocks * g.ncomp);
  const int64_t max_parts = (g.total + trip - 1) / trip;
  if (parts > max_parts) parts = max_parts;
  if (parts < 1) parts = 1;
  g.per_part = ((g.total + parts - 1) / parts + trip - 1) / trip * trip;
  g.parts = (g.total + g.per_part - 1) / g.per_part;
  const int64_t ctas = (int64_t)g.nblocks * g.ncomp * g.parts;
  PB2_REQUIRE(ctas < (1ll << 31), "grid too large");
  ProfScope prof(K_HALO_UNIFORM, as_stream(stream),
                 static_cast<double>(g.total) * (v2 ? 2 : 1) * g.nblocks * g.ncomp);
  if (v2)
    halo_uniform_kernel<2><<<static_cast<unsigned>(ctas), kThreads, 0, as_stream(stream)>>>(g, field, nbr);
  else
    halo_uniform_kernel<1><<<static_cast<unsigned>(ctas), kThreads, 0, as_stream(stream)>>>(g, field, nbr);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

} // extern "C"
