// update.cpp — dense per-stage update tasks (see pb2/update.hpp).
#include "pb2/update.hpp"

#include <algorithm>

namespace parthenon {
namespace Update {

template <>
TaskStatus FluxDivergence(MeshData<Real> *in, MeshData<Real> *dudt_cont) {
  for (Variable *v : in->GetVariablesByFlag({Metadata::Independent, Metadata::WithFluxes})) {
    Variable &d = dudt_cont->Get(v->label());
    pb2_pack_geom g = in->Geometry(*v);
    const double *flux[3] = {v->flux(1), g.ndim > 1 ? v->flux(2) : nullptr,
                             g.ndim > 2 ? v->flux(3) : nullptr};
    // update.cpp:78: only where both the field and dudt are allocated (sparse fields)
    PB2_CHECK(pb2_flux_divergence_blocks(&g, flux, d.data(), v->DeviceMask(), in->stream()));
  }
  return TaskStatus::complete;
}

template <>
TaskStatus WeightedSumData(const std::vector<MetadataFlag> &flags, MeshData<Real> *in1,
                           MeshData<Real> *in2, const Real w1, const Real w2,
                           MeshData<Real> *out) {
  for (Variable *x : in1->GetVariablesByFlag(flags)) {
    Variable &y = in2->Get(x->label());
    Variable &z = out->Get(x->label());
    if (x->metadata().IsSparse()) {
      // update.hpp:83-85: skipped where x, y or z is unallocated (they are allocated together)
      const pb2_pack_geom g = in1->Geometry(*x);
      PB2_CHECK(pb2_weighted_sum_blocks(&g, x->data(), y.data(), w1, w2, z.data(),
                                        x->DeviceMask(), in1->stream()));
      continue;
    }
    const int64_t n = x->block_stride * in1->NumBlocks();
    PB2_CHECK(pb2_weighted_sum(x->data(), y.data(), w1, w2, z.data(), n, in1->stream()));
  }
  return TaskStatus::complete;
}

template <>
TaskStatus EstimateTimestep(MeshData<Real> *rc) {
  Real dt_min = std::numeric_limits<Real>::max();
  for (const auto &pkg : rc->GetMeshPointer()->packages.AllPackages())
    if (pkg.second->EstimateTimestepMesh != nullptr)
      dt_min = std::min(dt_min, pkg.second->EstimateTimestepMesh(rc));
  // update.hpp:277-283: every block of the batch votes with the batch minimum
  for (auto &pmb : rc->GetBlockList()) pmb->SetAllowedDt(std::min(dt_min, pmb->NewDt()));
  return TaskStatus::complete;
}

template <>
TaskStatus PreCommFillDerived(MeshData<Real> *rc) {
  for (const auto &pkg : rc->GetMeshPointer()->packages.AllPackages())
    if (pkg.second->PreCommFillDerivedMesh != nullptr) pkg.second->PreCommFillDerivedMesh(rc);
  return TaskStatus::complete;
}

template <>
TaskStatus FillDerived(MeshData<Real> *rc) {
  auto &pkgs = rc->GetMeshPointer()->packages.AllPackages();
  for (const auto &pkg : pkgs)
    if (pkg.second->PreFillDerivedMesh != nullptr) pkg.second->PreFillDerivedMesh(rc);
  for (const auto &pkg : pkgs)
    if (pkg.second->FillDerivedMesh != nullptr) pkg.second->FillDerivedMesh(rc);
  for (const auto &pkg : pkgs)
    if (pkg.second->PostFillDerivedMesh != nullptr) pkg.second->PostFillDerivedMesh(rc);
  return TaskStatus::complete;
}

// update.cpp:143-217: a sparse field whose every value on a block (entire extents) stayed
// within the deallocation threshold for more than deallocation_count consecutive calls is
// deallocated on that block, in every container
TaskStatus SparseDealloc(MeshData<Real> *md) {
  Mesh *pm = md->GetMeshPointer();
  if (!pm->sparse_config.enabled || md->NumBlocks() == 0) return TaskStatus::complete;
  const int nb = md->NumBlocks();
  // one flag buffer, one read-back and one synchronisation for ALL sparse fields (the reference
  // reduces per (block, variable) team, update.cpp:161-186)
  const std::vector<Variable *> vars = md->GetVariablesByFlag({Metadata::Sparse});
  if (vars.empty()) return TaskStatus::complete;
  const size_t need = sizeof(int32_t) * static_cast<size_t>(nb) * vars.size();
  DeviceBuffer &flags = md->SparseScratch(need);
  std::vector<int32_t> quiet(static_cast<size_t>(nb) * vars.size());
  for (size_t iv = 0; iv < vars.size(); ++iv) {
    Variable *v = vars[iv];
    const pb2_pack_geom g = md->Geometry(*v);
    PB2_CHECK(pb2_block_quiet_flags(&g, v->data(), v->metadata().GetDeallocationThreshold(),
                                    v->DeviceMask(), flags.get<int32_t>() + iv * nb, md->stream()));
  }
  PB2_CHECK(pb2_memcpy_d2h(quiet.data(), flags.get(), need, md->stream()));
  PB2_CHECK(pb2_stream_sync(md->stream()));
  for (size_t iv = 0; iv < vars.size(); ++iv) {
    Variable *v = vars[iv];
    for (int b = 0; b < nb; ++b) {
      if (!v->IsAllocated(b)) continue;
      int &counter = v->dealloc_count(b);
      counter = quiet[iv * nb + b] ? counter + 1 : 0;
      if (counter > pm->sparse_config.deallocation_count) {
        counter = 0;
        pm->DeallocateSparse(v->label(), md->GetBlock(b)->lid);
      }
    }
  }
  return TaskStatus::complete;
}

} // namespace Update
} // namespace parthenon
