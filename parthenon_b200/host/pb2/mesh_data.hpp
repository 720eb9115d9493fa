// mesh_data.hpp — field storage in HBM and the MeshData batch (host only).
//
// Replaces Variable (src/interface/variable.hpp:57-160), MeshBlockData
// (meshblock_data.hpp:55), MeshData (mesh_data.hpp:191-530) and the device "views of views"
// packs (variable_pack.hpp:259-470, sparse_pack.hpp:53) by ONE contiguous slab per field
// and container:  data[block][component][k][j][i]  (i fastest, like LayoutRight).
// A pack is then nothing but {base pointer, block stride, component stride}: kernels index
// it arithmetically instead of dereferencing a 136-byte view handle per access, and every
// block of a 180 GB device sits in a handful of allocations.  Arrays are allocated lazily
// on first use, so containers that the fused stage never touches (dUdt, the six Ul*/Ur*
// reconstruction fields, flux arrays) cost no HBM.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "device.hpp"
#include "mesh.hpp"
#include "state.hpp"
#include "types.hpp"

namespace parthenon {

struct BvarsCache; // ghost-exchange tables of one MeshData (bvals.hpp)
struct SparsePackStorage; // device tables of one resolved SparsePack (sparse_pack.hpp)

// One field over every block of a MeshData batch.
class Variable {
 public:
  Variable(const std::string &label, const Metadata &m, int sparse_id, int nblocks,
           const IndexShape &cb, const IndexShape &ccb, bool multilevel, pb2_stream_t stream,
           int capacity = 0);
  const std::string &label() const { return label_; }
  const Metadata &metadata() const { return m_; }
  bool IsSet(MetadataFlag f) const { return m_.IsSet(f); }
  // slab components per block: topological elements x tensor components (element slowest, like
  // the reference's 7-D arrays (el, t, u, v, k, j, i))
  int NumComponents() const { return ncomp_; }
  // where the values live (Metadata::Cell / Face / Edge / Node) and how many topological
  // elements each cell carries: F1..F3 / E1..E3 = 3, cell and node = 1 (metadata.cpp:376-382)
  TopologicalType topological_type() const { return tt_; }
  int NumElements() const { return nel_; }
  int TensorComponents() const { return ncomp_ / nel_; }
  int GetDim(int i) const; // 1: ni, 2: nj, 3: nk, 4: ncomp (variable.hpp GetDim)
  int sparse_id() const { return sparse_id_; }

  Real *data();          // [nblocks][ncomp][nk][nj][ni]
  Real *coarse();        // [nblocks][ncomp][cnk][cnj][cni], multilevel + FillGhost only
  Real *flux(int dir);   // dir = 1..3 (X1DIR..), same extents as data (metadata.cpp:185)
  bool HasData() const { return static_cast<bool>(data_); }
  Real *block(int b) { return data() + b * block_stride; }

  int64_t block_stride = 0, comp_stride = 0;   // in Reals
  int64_t cblock_stride = 0, ccomp_stride = 0; // coarse buffer
  int ni = 1, nj = 1, nk = 1, cni = 1, cnj = 1, cnk = 1;

  // sparse allocation status per block (variable.hpp IsAllocated); dense fields: all true
  bool IsAllocated(int b) const { return allocated_[b] != 0; }
  void SetAllocated(int b, bool a) {
    allocated_[b] = a ? 1 : 0;
    mask_dirty_ = true;
  }
  // Variable::AllocateData (variable.cpp:112-127): the block's arrays start zeroed
  void AllocateBlock(int b);
  // device int32 [nblocks] allocation mask for the block-masked kernels; NULL for dense fields
  const int32_t *DeviceMask();
  // device int32 list of the allocated blocks (and its length) for the same kernels: they
  // launch over the list instead of over every block; NULL / 0 for dense fields
  const int32_t *DeviceList(int32_t *n);
  const std::vector<uint8_t> &AllocationStatus() const { return allocated_; }
  int dealloc_count(int b) const { return dealloc_count_[b]; }
  int &dealloc_count(int b) { return dealloc_count_[b]; }

 private:
  std::string label_;
  Metadata m_;
  int sparse_id_, ncomp_, nblocks_, capacity_;
  int data_shift_ = 0; // see Variable::data()
  TopologicalType tt_ = TopologicalType::Cell;
  int nel_ = 1;
  bool multilevel_;
  pb2_stream_t stream_;
  DeviceBuffer data_, coarse_, flux_[3];
  std::vector<uint8_t> allocated_;
  std::vector<int> dealloc_count_;
  DeviceBuffer mask_;
  bool mask_dirty_ = true;
  int32_t nlist_ = 0;
  void UploadAllocation();
};

// what md->PackVariables(names) returns: the selected fields of every block, addressable
// as pack(b, n) -> component pointer.  `ptrs` is [nblocks][nvar] on the host; DevicePtrs()
// uploads the same table for user kernels.
struct VariablePack {
  int nblocks = 0, nvar = 0;
  int dims[3] = {1, 1, 1}; // ni, nj, nk
  std::vector<Variable *> vars;             // in request order
  std::vector<std::pair<int, int>> ranges;  // component range of each requested field
  std::vector<Real *> ptrs;                 // [b * nvar + n]
  std::vector<Real *> flux_ptrs[3];
  DeviceBuffer dev_ptrs;
  int GetDim(int i) const { return i <= 3 ? dims[i - 1] : (i == 4 ? nvar : nblocks); }
  Real *operator()(int b, int n) const { return ptrs[static_cast<size_t>(b) * nvar + n]; }
  Real *const *DevicePtrs(pb2_stream_t stream);
};
using PackIndexMap = std::map<std::string, std::pair<int, int>>;

template <typename T>
class MeshData {
 public:
  MeshData(Mesh *pmesh, int partition_id, const std::string &label, MeshData<T> *base);
  ~MeshData();

  Mesh *GetMeshPointer() const { return pmesh_; }
  Mesh *GetParentPointer() const { return pmesh_; }
  int NumBlocks() const { return static_cast<int>(blocks_.size()); }
  int partition_id() const { return partition_; }
  const std::string &label() const { return label_; }
  MeshBlock *GetBlock(int b) const { return blocks_[b].get(); }
  const BlockList_t &GetBlockList() const { return blocks_; }
  pb2_stream_t stream() const { return pmesh_->stream; }

  IndexRange GetBoundsI(IndexDomain d) const { return blocks_[0]->cellbounds.GetBoundsI(d); }
  IndexRange GetBoundsJ(IndexDomain d) const { return blocks_[0]->cellbounds.GetBoundsJ(d); }
  IndexRange GetBoundsK(IndexDomain d) const { return blocks_[0]->cellbounds.GetBoundsK(d); }

  bool HasVariable(const std::string &name) const { return vars_.count(name) > 0; }
  Variable &Get(const std::string &name);
  std::vector<Variable *> GetVariablesByFlag(const std::vector<MetadataFlag> &flags);
  const std::vector<std::shared_ptr<Variable>> &GetVariableVector() const { return order_; }

  // mesh_data.hpp:349-440; cached per name list like the reference's pack cache
  VariablePack &PackVariables(const std::vector<std::string> &names, PackIndexMap *imap = nullptr);
  VariablePack &PackVariablesAndFluxes(const std::vector<std::string> &names,
                                       const std::vector<std::string> &flux_names,
                                       PackIndexMap *imap = nullptr);
  VariablePack &PackVariablesByFlag(const std::vector<MetadataFlag> &flags,
                                    PackIndexMap *imap = nullptr);

  // per-block geometry on the device: dx [nblocks][3], interior xmin [nblocks][3]
  const Real *DeviceDx();
  const Real *DeviceXmin();
  // C-ABI geometry descriptor of a field of this batch
  pb2_pack_geom Geometry(Variable &v);

  // one device Real owned by this batch: the min-dt cell its stage kernels reduce into (every
  // batch needs its own — task lists of different partitions interleave on the stream)
  Real *DtCell() {
    if (!dt_cell_) dt_cell_.Allocate(sizeof(Real), stream());
    return dt_cell_.get<Real>();
  }

  // a device scratch buffer of at least `bytes` that lives as long as the batch (per-cycle flag
  // reductions of sparse fields: no allocation / free inside the cycle)
  DeviceBuffer &SparseScratch(size_t bytes) {
    if (sparse_scratch_.bytes() < bytes) sparse_scratch_.Allocate(bytes, stream());
    return sparse_scratch_;
  }

  BvarsCache &bvars() { return *bvars_; }
  // resolved SparsePacks by descriptor identifier (MeshData::GetSparsePackCache)
  std::map<std::string, std::shared_ptr<SparsePackStorage>> &GetSparsePackCache() {
    return sparse_pack_cache_;
  }
  // bumped whenever an allocation status changes; the exchange tables are rebuilt when
  // their generation differs (replaces the per-call host walk of
  // CheckSendBufferCacheForRebuild, bvals_utils.hpp:140-202)
  uint64_t alloc_generation = 1;

 private:
  Mesh *pmesh_;
  int partition_;
  std::string label_;
  BlockList_t blocks_;
  std::map<std::string, std::shared_ptr<Variable>> vars_;
  std::vector<std::shared_ptr<Variable>> order_;
  std::map<std::string, std::unique_ptr<VariablePack>> pack_cache_;
  DeviceBuffer dx_, xmin_, dt_cell_, sparse_scratch_;
  std::map<std::string, std::shared_ptr<SparsePackStorage>> sparse_pack_cache_;
  std::unique_ptr<BvarsCache> bvars_;
};

// per-block view used by block-level hooks (ApplyBoundaryConditions, Refinement::Tag)
template <typename T>
class MeshBlockData {
 public:
  MeshBlockData(MeshData<T> *md, int b) : md_(md), b_(b) {}
  MeshBlock *GetBlockPointer() const { return md_->GetBlock(b_); }
  Real *Data(const std::string &name) { return md_->Get(name).block(b_); }
  MeshData<T> *GetMeshData() const { return md_; }
  int index() const { return b_; }

 private:
  MeshData<T> *md_;
  int b_;
};

} // namespace parthenon
