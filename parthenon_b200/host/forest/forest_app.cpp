// forest_app.cpp — see forest_app.hpp
#include "forest_app.hpp"

#include <unordered_map>
#include <vector>

namespace forest_example {
using namespace parthenon;

forest::ForestDefinition MakeForest(int variant) {
  PARTHENON_REQUIRE(variant >= 0 && variant <= 3, "forest/variant must be 0, 1, 2 or 3");
  //   6---7---8
  //   | 3 | 4 |      nodes and face ids as in example/boundary_exchange
  //   3---2---5
  //   | 0 | 1 |
  //   0---1---4
  std::unordered_map<uint64_t, std::shared_ptr<forest::Node>> n;
  const Real pos[9][2] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}, {2, 0}, {2, 1}, {0, 2}, {1, 2}, {2, 2}};
  for (uint64_t i = 0; i < 9; ++i) n[i] = forest::Node::create(i, {pos[i][0], pos[i][1]});
  forest::ForestDefinition def;
  using ar3_t = std::array<Real, 3>;
  using edge_t = forest::Edge;
  if (variant <= 1) {
    def.AddFace(0, {n[1], n[2], n[0], n[3]}, ar3_t{0.0, 0.0, 0.0}, ar3_t{1.0, 1.0, 1.0});
    def.AddFace(1, {n[1], n[4], n[2], n[5]}, ar3_t{2.0, 0.0, 0.0}, ar3_t{3.0, 1.0, 1.0});
    def.AddFace(3, {n[3], n[2], n[6], n[7]}, ar3_t{0.0, 2.0, 0.0}, ar3_t{1.0, 3.0, 1.0});
    def.AddFace(4, {n[2], n[5], n[7], n[8]}, ar3_t{2.0, 2.0, 0.0}, ar3_t{3.0, 3.0, 1.0});
  } else {
    // 0: rotated by 90 degrees; 1: by 180 degrees; 3: reflected about x1; 4: as laid out
    def.AddFace(0, {n[1], n[2], n[0], n[3]}, ar3_t{0.0, 0.0, 0.0}, ar3_t{1.0, 1.0, 1.0});
    def.AddFace(1, {n[5], n[2], n[4], n[1]}, ar3_t{2.0, 0.0, 0.0}, ar3_t{3.0, 1.0, 1.0});
    def.AddFace(3, {n[2], n[3], n[7], n[6]}, ar3_t{0.0, 2.0, 0.0}, ar3_t{1.0, 3.0, 1.0});
    def.AddFace(4, {n[2], n[5], n[7], n[8]}, ar3_t{2.0, 2.0, 0.0}, ar3_t{3.0, 3.0, 1.0});
  }
  const int outer[8][2] = {{0, 1}, {0, 3}, {1, 4}, {4, 5}, {6, 7}, {3, 6}, {5, 8}, {7, 8}};
  for (auto &e : outer) def.AddBC(edge_t({n[e[0]], n[e[1]]}));
  if (variant == 0) def.AddInitialRefinement(LogicalLocation(0, 1, 0, 0, 0));
  if (variant == 3) {
    def.AddInitialRefinement(LogicalLocation(3, 1, 1, 0, 0));
    def.AddInitialRefinement(LogicalLocation(4, 1, 0, 1, 0));
  }
  return def;
}

Packages_t ProcessPackages(std::unique_ptr<ParameterInput> &pin) {
  Packages_t packages;
  auto pkg = std::make_shared<StateDescriptor>("boundary_exchange");
  Metadata m({Metadata::Cell, Metadata::Independent, Metadata::FillGhost}, std::vector<int>{8});
  m.RegisterRefinementOps<refinement_ops::ProlongatePiecewiseConstant,
                          refinement_ops::RestrictAverage>();
  pkg->AddField("neighbor_info", m);
  packages.Add(pkg);
  return packages;
}

// value = (gid + 1) * 1e4 + component * 1e3 + flat (j, i) index, ghosts included
void MeshProblemGenerator(MeshData<Real> *md, ParameterInput *) {
  const int nb = md->NumBlocks();
  Variable &v = md->Get("neighbor_info");
  const int nc = v.TensorComponents();
  std::vector<Real> h(static_cast<size_t>(nb) * v.block_stride);
  for (int b = 0; b < nb; ++b) {
    const int gid = md->GetBlock(b)->gid;
    Real *vb = h.data() + static_cast<size_t>(b) * v.block_stride;
    for (int c = 0; c < nc; ++c)
      for (int64_t n = 0; n < v.comp_stride; ++n)
        vb[c * v.comp_stride + n] = (gid + 1) * 1.0e4 + c * 1.0e3 + static_cast<Real>(n);
  }
  PB2_CHECK(pb2_memcpy_h2d(v.data(), h.data(), sizeof(Real) * h.size(), md->stream()));
  PB2_CHECK(pb2_stream_sync(md->stream()));
}
} // namespace forest_example
