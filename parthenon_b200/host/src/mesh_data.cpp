// mesh_data.cpp — slab-per-field storage and MeshData batches (see pb2/mesh_data.hpp).
#include "pb2/mesh_data.hpp"

#include <algorithm>

#include "pb2/bvals.hpp"

namespace parthenon {

Variable::Variable(const std::string &label, const Metadata &m, int sparse_id, int nblocks,
                   const IndexShape &cb, const IndexShape &ccb, bool multilevel,
                   pb2_stream_t stream, int capacity)
    : label_(label), m_(m), sparse_id_(sparse_id), ncomp_(m.NumComponents()),
      nblocks_(nblocks), capacity_(std::max(capacity, nblocks)), multilevel_(multilevel),
      stream_(stream) {
  if (m.IsSet(Metadata::Face)) tt_ = TopologicalType::Face;
  if (m.IsSet(Metadata::Edge)) tt_ = TopologicalType::Edge;
  if (m.IsSet(Metadata::Node)) tt_ = TopologicalType::Node;
  nel_ = (tt_ == TopologicalType::Face || tt_ == TopologicalType::Edge) ? 3 : 1;
  ncomp_ *= nel_;
  ni = cb.ncellsi(IndexDomain::entire);
  nj = cb.ncellsj(IndexDomain::entire);
  nk = cb.ncellsk(IndexDomain::entire);
  if (tt_ != TopologicalType::Cell) {
    // face, edge and node arrays are one longer in every non-symmetry direction
    // (metadata.cpp:383-387)
    // (the flux of a face field is the separate edge field "bnd_flux::<name>",
    // StateDescriptor::AddField; edge and node fields with fluxes are not built)
    PARTHENON_REQUIRE(!m.IsSparse() &&
                          (!m.IsSet(Metadata::WithFluxes) || tt_ == TopologicalType::Face),
                      "non-cell-centred fields cannot be sparse, and only face fields carry "
                      "fluxes in this build (" + label + ")");
    ni++;
    if (nj > 1) nj++;
    if (nk > 1) nk++;
  }
  cni = ccb.ncellsi(IndexDomain::entire);
  cnj = ccb.ncellsj(IndexDomain::entire);
  cnk = ccb.ncellsk(IndexDomain::entire);
  if (tt_ != TopologicalType::Cell) { // coarse buffers are padded like the fine arrays
    cni++;
    if (cnj > 1) cnj++;
    if (cnk > 1) cnk++;
  }
  comp_stride = static_cast<int64_t>(ni) * nj * nk;
  block_stride = comp_stride * ncomp_;
  {
    const int lead = cb.is(IndexDomain::interior); // ghost cells in front of a row's interior
    const int shift = (8 - lead % 8) % 8;
    data_shift_ = (shift % 2 == 0) ? shift : 0; // keep 16-byte alignment for vector accesses
  }
  ccomp_stride = static_cast<int64_t>(cni) * cnj * cnk;
  cblock_stride = ccomp_stride * ncomp_;
  // sparse fields start unallocated (variable.cpp:112-160); dense ones are always there
  allocated_.assign(nblocks, m.IsSparse() ? 0 : 1);
  dealloc_count_.assign(nblocks, 0);
}

int Variable::GetDim(int i) const {
  switch (i) {
  case 1: return ni;
  case 2: return nj;
  case 3: return nk;
  case 4: return ncomp_;
  default: return 1;
  }
}

// slabs are allocated for `capacity_` blocks (>= the block count).  On adaptive meshes the mesh
// hands out a capacity that changes rarely (Mesh::SlabCapacity), so the slabs of successive
// remeshes have the same size and the allocator's exact-size cache (csrc/runtime.cu) recycles
// them instead of unmapping and mapping GBs
#define SlabBlocks(n) (static_cast<size_t>(std::max(capacity_, std::max((n), 1))))

// The slab starts `data_shift_` Reals into its allocation so that the first INTERIOR cell of a
// row — not the row's first ghost cell — sits on a 64-byte DRAM atom: a sweep that reads
// interior columns only (y / z direction, 256-byte rows of a 32^3 block) then touches 4 atoms
// per row instead of 5.  (Rows keep their alignment when the padded row length is a multiple of
// 8 Reals, e.g. 32 + 2 x 4.)
Real *Variable::data() {
  if (!data_)
    data_.Allocate(sizeof(Real) * (static_cast<size_t>(block_stride) * SlabBlocks(nblocks_) + 8),
                   stream_);
  return data_.get<Real>() + data_shift_;
}

Real *Variable::coarse() {
  PARTHENON_REQUIRE(multilevel_, "coarse buffers only exist on multilevel meshes");
  if (!coarse_)
    coarse_.Allocate(sizeof(Real) * static_cast<size_t>(cblock_stride) * SlabBlocks(nblocks_),
                     stream_);
  return coarse_.get<Real>();
}

Real *Variable::flux(int dir) {
  PARTHENON_REQUIRE(m_.IsSet(Metadata::WithFluxes), "field " + label_ + " has no fluxes");
  PARTHENON_REQUIRE(tt_ == TopologicalType::Cell,
                    "the flux of face field " + label_ + " is the edge field bnd_flux::" + label_);
  PARTHENON_REQUIRE(dir >= 1 && dir <= 3, "flux direction must be X1DIR..X3DIR");
  DeviceBuffer &f = flux_[dir - 1];
  if (!f)
    f.Allocate(sizeof(Real) * static_cast<size_t>(block_stride) * SlabBlocks(nblocks_), stream_);
  return f.get<Real>();
}

void Variable::AllocateBlock(int b) {
  SetAllocated(b, true);
  auto zero = [&](DeviceBuffer &buf, int64_t stride) {
    if (buf)
      PB2_CHECK(pb2_memset(buf.get<Real>() + b * stride, 0, sizeof(Real) * static_cast<size_t>(stride),
                           stream_));
  };
  if (data_)
    PB2_CHECK(pb2_memset(data() + b * block_stride, 0,
                         sizeof(Real) * static_cast<size_t>(block_stride), stream_));
  zero(coarse_, cblock_stride);
  for (auto &f : flux_) zero(f, block_stride);
}

void Variable::UploadAllocation() {
  if (!mask_dirty_) return;
  if (!mask_) mask_.Allocate(sizeof(int32_t) * 2 * std::max(nblocks_, 1), stream_);
  // [0, nblocks): the mask; [nblocks, 2 nblocks): the indices of the allocated blocks
  std::vector<int32_t> h(2 * static_cast<size_t>(std::max(nblocks_, 1)), 0);
  nlist_ = 0;
  for (int b = 0; b < nblocks_; ++b) {
    h[b] = allocated_[b];
    if (allocated_[b]) h[nblocks_ + nlist_++] = b;
  }
  PB2_CHECK(pb2_memcpy_h2d(mask_.get(), h.data(), sizeof(int32_t) * h.size(), stream_));
  PB2_CHECK(pb2_stream_sync(stream_)); // h goes out of scope
  mask_dirty_ = false;
}

const int32_t *Variable::DeviceMask() {
  if (!m_.IsSparse()) return nullptr;
  UploadAllocation();
  return mask_.get<int32_t>();
}

const int32_t *Variable::DeviceList(int32_t *n) {
  *n = 0;
  if (!m_.IsSparse()) return nullptr;
  UploadAllocation();
  *n = nlist_;
  return mask_.get<int32_t>() + nblocks_;
}

Real *const *VariablePack::DevicePtrs(pb2_stream_t stream) {
  if (!dev_ptrs && !ptrs.empty()) {
    dev_ptrs.Allocate(sizeof(Real *) * ptrs.size(), stream);
    PB2_CHECK(pb2_memcpy_h2d(dev_ptrs.get(), ptrs.data(), sizeof(Real *) * ptrs.size(), stream));
    PB2_CHECK(pb2_stream_sync(stream));
  }
  return dev_ptrs.get<Real *const>();
}

template <typename T>
MeshData<T>::MeshData(Mesh *pmesh, int partition_id, const std::string &label,
                      MeshData<T> *base)
    : pmesh_(pmesh), partition_(partition_id), label_(label) {
  for (auto &pmb : pmesh->block_list)
    if (pmb->partition == partition_id) blocks_.push_back(pmb);
  PARTHENON_REQUIRE(!blocks_.empty(), "MeshData partition without blocks");
  const IndexShape &cb = blocks_[0]->cellbounds, &ccb = blocks_[0]->c_cellbounds;
  for (auto &f : pmesh->resolved_fields) {
    std::shared_ptr<Variable> v;
    // OneCopy fields are shared with the base container (meshblock_data.cpp Add/Copy)
    if (base != nullptr && f.m.IsSet(Metadata::OneCopy)) {
      v = base->vars_.at(f.name);
    } else {
      v = std::make_shared<Variable>(f.name, f.m, f.sparse_id, NumBlocks(), cb, ccb,
                                     pmesh->multilevel, pmesh->stream,
                                     pmesh->SlabCapacity(NumBlocks()));
      if (base != nullptr) // a stage container starts with base's allocation status
        for (int b = 0; b < NumBlocks(); ++b) v->SetAllocated(b, base->vars_.at(f.name)->IsAllocated(b));
    }
    vars_[f.name] = v;
    order_.push_back(v);
  }
  bvars_ = std::make_unique<BvarsCache>();
}

template <typename T>
MeshData<T>::~MeshData() = default;

template <typename T>
Variable &MeshData<T>::Get(const std::string &name) {
  auto it = vars_.find(name);
  PARTHENON_REQUIRE(it != vars_.end(), "Couldn't find variable '" + name + "' in container '" +
                                           label_ + "'");
  return *it->second;
}

template <typename T>
std::vector<Variable *> MeshData<T>::GetVariablesByFlag(const std::vector<MetadataFlag> &flags) {
  std::vector<Variable *> out;
  for (auto &v : order_)
    if (v->metadata().AllFlagsSet(flags)) out.push_back(v.get());
  return out;
}

template <typename T>
VariablePack &MeshData<T>::PackVariablesAndFluxes(const std::vector<std::string> &names,
                                                  const std::vector<std::string> &flux_names,
                                                  PackIndexMap *imap) {
  std::string key;
  for (auto &n : names) key += n + "|";
  key += "#";
  for (auto &n : flux_names) key += n + "|";
  auto it = pack_cache_.find(key);
  if (it == pack_cache_.end()) {
    auto p = std::make_unique<VariablePack>();
    p->nblocks = NumBlocks();
    int n0 = 0;
    for (auto &n : names) {
      Variable &v = Get(n);
      p->vars.push_back(&v);
      p->ranges.emplace_back(n0, n0 + v.NumComponents() - 1);
      n0 += v.NumComponents();
      p->dims[0] = v.ni;
      p->dims[1] = v.nj;
      p->dims[2] = v.nk;
    }
    p->nvar = n0;
    p->ptrs.resize(static_cast<size_t>(p->nblocks) * n0);
    const bool with_flux = !flux_names.empty();
    if (with_flux)
      for (int d = 0; d < pmesh_->ndim; ++d) p->flux_ptrs[d].resize(p->ptrs.size());
    for (int b = 0; b < p->nblocks; ++b) {
      int n = 0;
      for (Variable *v : p->vars) {
        const bool fl = with_flux && std::find(flux_names.begin(), flux_names.end(),
                                               v->label()) != flux_names.end();
        for (int c = 0; c < v->NumComponents(); ++c, ++n) {
          const size_t idx = static_cast<size_t>(b) * n0 + n;
          p->ptrs[idx] = v->data() + b * v->block_stride + c * v->comp_stride;
          if (with_flux)
            for (int d = 0; d < pmesh_->ndim; ++d)
              p->flux_ptrs[d][idx] =
                  fl ? v->flux(d + 1) + b * v->block_stride + c * v->comp_stride : nullptr;
        }
      }
    }
    it = pack_cache_.emplace(key, std::move(p)).first;
  }
  if (imap) {
    imap->clear();
    for (size_t i = 0; i < names.size(); ++i) (*imap)[names[i]] = it->second->ranges[i];
  }
  return *it->second;
}

template <typename T>
VariablePack &MeshData<T>::PackVariables(const std::vector<std::string> &names,
                                         PackIndexMap *imap) {
  return PackVariablesAndFluxes(names, {}, imap);
}

template <typename T>
VariablePack &MeshData<T>::PackVariablesByFlag(const std::vector<MetadataFlag> &flags,
                                               PackIndexMap *imap) {
  std::vector<std::string> names;
  for (Variable *v : GetVariablesByFlag(flags)) names.push_back(v->label());
  return PackVariables(names, imap);
}

template <typename T>
const Real *MeshData<T>::DeviceDx() {
  if (!dx_) {
    std::vector<Real> h(static_cast<size_t>(NumBlocks()) * 3), x(h.size());
    for (int b = 0; b < NumBlocks(); ++b)
      for (int d = 0; d < 3; ++d) {
        h[3 * b + d] = blocks_[b]->coords.Dx()[d];
        x[3 * b + d] = blocks_[b]->block_size.xmin_[d];
      }
    dx_.Allocate(sizeof(Real) * h.size(), stream());
    xmin_.Allocate(sizeof(Real) * x.size(), stream());
    PB2_CHECK(pb2_memcpy_h2d(dx_.get(), h.data(), sizeof(Real) * h.size(), stream()));
    PB2_CHECK(pb2_memcpy_h2d(xmin_.get(), x.data(), sizeof(Real) * x.size(), stream()));
    PB2_CHECK(pb2_stream_sync(stream()));
  }
  return dx_.get<Real>();
}

template <typename T>
const Real *MeshData<T>::DeviceXmin() {
  DeviceDx();
  return xmin_.get<Real>();
}

template <typename T>
pb2_pack_geom MeshData<T>::Geometry(Variable &v) {
  pb2_pack_geom g{};
  g.nblocks = NumBlocks();
  g.ncomp = v.NumComponents();
  g.ndim = pmesh_->ndim;
  for (int d = 0; d < 3; ++d) g.nx[d] = pmesh_->base_block_size.nx_[d];
  g.ng = Globals::nghost;
  g.block_stride = v.block_stride;
  g.dx = DeviceDx();
  // sparse fields: the block-masked kernels launch over the allocated blocks only
  g.block_list = v.DeviceList(&g.nlist);
  return g;
}

template class MeshData<Real>;

std::shared_ptr<MeshData<Real>> &MeshDataCollection::GetOrAdd(const std::string &label,
                                                              int partition_id) {
  const std::string key = label + "_part-" + std::to_string(partition_id);
  auto it = map_.find(key);
  if (it != map_.end()) return it->second;
  MeshData<Real> *base = nullptr;
  if (label != "base") base = GetOrAdd("base", partition_id).get();
  return map_[key] = std::make_shared<MeshData<Real>>(pmesh_, partition_id, label, base);
}

void MeshDataCollection::PurgeNonBase() {
  for (auto it = map_.begin(); it != map_.end();)
    it = (it->first.compare(0, 5, "base_") == 0) ? std::next(it) : map_.erase(it);
}

} // namespace parthenon
