// update.hpp — per-stage dense update tasks (namespace parthenon::Update).
// Same task names and argument meaning as the reference's src/interface/update.hpp:43-383 /
// update.cpp:63-217; each is one launch through the C ABI over the whole MeshData slab.
#pragma once
#include <limits>
#include <vector>

#include "mesh_data.hpp"

namespace parthenon {
namespace Update {

// dudt_cont(Independent+WithFluxes) = -div(flux of `in`)   (update.cpp:63-86)
template <typename T>
TaskStatus FluxDivergence(T *in, T *dudt_cont);

// out = w1*in1 + w2*in2 over the ENTIRE extents of every field carrying `flags`
// (update.hpp:71-91)
template <typename F, typename T>
TaskStatus WeightedSumData(const F &flags, T *in1, T *in2, const Real w1, const Real w2, T *out);

template <>
TaskStatus FluxDivergence(MeshData<Real> *in, MeshData<Real> *dudt_cont);
template <>
TaskStatus WeightedSumData(const std::vector<MetadataFlag> &flags, MeshData<Real> *in1,
                           MeshData<Real> *in2, const Real w1, const Real w2,
                           MeshData<Real> *out);

template <typename F, typename T>
TaskStatus SumData(const F &flags, T *in1, T *in2, T *out) {
  return WeightedSumData(flags, in1, in2, 1.0, 1.0, out);
}
// out = in + dt*dudt (update.hpp:108-112)
template <typename F, typename T>
TaskStatus UpdateData(const F &flags, T *in, T *dudt, const Real dt, T *out) {
  return WeightedSumData(flags, in, dudt, 1.0, dt, out);
}
template <typename T>
TaskStatus UpdateIndependentData(T *in, T *dudt, const Real dt, T *out) {
  return WeightedSumData(std::vector<MetadataFlag>({Metadata::Independent}), in, dudt, 1.0, dt,
                         out);
}
// c1 = wgt1*c1 + (1-wgt1)*c2 (update.hpp:126-137)
template <typename F, typename T>
TaskStatus AverageData(const F &flags, T *c1, T *c2, const Real wgt1) {
  return WeightedSumData(flags, c1, c2, wgt1, 1.0 - wgt1, c1);
}
template <typename T>
TaskStatus AverageIndependentData(T *c1, T *c2, const Real wgt1) {
  return WeightedSumData(std::vector<MetadataFlag>({Metadata::Independent}), c1, c2, wgt1,
                         1.0 - wgt1, c1);
}

// package hooks (update.hpp:269-314)
template <typename T>
TaskStatus EstimateTimestep(T *rc);
template <typename T>
TaskStatus PreCommFillDerived(T *rc);
template <typename T>
TaskStatus FillDerived(T *rc);

template <>
TaskStatus EstimateTimestep(MeshData<Real> *rc);
template <>
TaskStatus PreCommFillDerived(MeshData<Real> *rc);
template <>
TaskStatus FillDerived(MeshData<Real> *rc);

// update.cpp:143-217
TaskStatus SparseDealloc(MeshData<Real> *md);

} // namespace Update

namespace Refinement {
// amr_criteria/refinement_package.cpp:150-172 for a whole MeshData batch
TaskStatus Tag(MeshData<Real> *rc);
} // namespace Refinement
} // namespace parthenon
