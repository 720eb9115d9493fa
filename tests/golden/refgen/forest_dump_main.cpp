// Fixture generator (test infrastructure, NOT product code).
//
// A minimal application on the UNMODIFIED reference library (libparthenon.a built out-of-tree,
// see make_fixtures.sh) that pins the ghost exchange across the trees of a 2-D FOREST whose
// trees meet with different orientations (LogicalCoordinateTransformation: axis permutation and
// flips in SetBounds, boundary_communication.cpp:282-308).  The forest is described through the
// reference's ForestDefinition API exactly as example/boundary_exchange does (nine nodes on a
// 3 x 3 lattice, four faces around the central node, user = outflow boundaries on the outer
// edges); $PB2_FOREST_VARIANT selects the node order of the faces (i.e. their orientations) and
// the initial refinement:
//   0  the example as shipped: face 0 listed as {n1, n2, n0, n3} (rotated), the other three
//      in lattice order; block (tree 0, level 1, 0, 0) refined
//   1  the same faces, no refinement
//   2  all four faces in different orientations (one of them a reflection), no refinement
//   3  as 2 with block (tree 3, level 1, 1, 0) and (tree 4, level 1, 0, 1) refined
// The package is the example's: one cell-centred field of 8 components with
// ProlongatePiecewiseConstant / RestrictAverage.  The problem generator writes
//     value = (gid + 1) * 1e4 + component * 1e3 + (j * ni + i)     (interior AND ghosts,
// so ghosts that no exchange, prolongation or boundary condition touches keep their own code).
// Only this file is ours.
//
// Dump layout: the one of burgers_dump_main.cpp with the per-block header
// {gid, level, lx1, lx2, tree} (2-D: lx3 is always 0, its slot carries the tree id).
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "mesh/forest/forest.hpp"
#include "parthenon_manager.hpp"
#include <parthenon/package.hpp>
#include <prolong_restrict/pr_ops.hpp>

namespace {
using namespace parthenon;
using namespace parthenon::package::prelude;
std::string g_prefix = "dump";

Packages_t ProcessPackages(std::unique_ptr<ParameterInput> &pin) {
  Packages_t packages;
  auto pkg = std::make_shared<StateDescriptor>("boundary_exchange");
  Metadata m({Metadata::Cell, Metadata::Independent, Metadata::FillGhost}, std::vector<int>{8});
  m.RegisterRefinementOps<parthenon::refinement_ops::ProlongatePiecewiseConstant,
                          parthenon::refinement_ops::RestrictAverage>();
  pkg->AddField("neighbor_info", m);
  packages.Add(pkg);
  return packages;
}

void ProblemGenerator(MeshBlock *pmb, ParameterInput *pin) {
  auto &v = pmb->meshblock_data.Get()->Get("neighbor_info");
  auto h = v.data.GetHostMirror();
  const int nc = v.data.GetDim(4), nk = v.data.GetDim(3), nj = v.data.GetDim(2),
            ni = v.data.GetDim(1);
  for (int c = 0; c < nc; ++c)
    for (int k = 0; k < nk; ++k)
      for (int j = 0; j < nj; ++j)
        for (int i = 0; i < ni; ++i)
          h(0, 0, 0, c, k, j, i) = (pmb->gid + 1) * 1.0e4 + c * 1.0e3 + ((k * nj + j) * ni + i);
  v.data.DeepCopy(h);
}

void Dump(Mesh *pmesh) {
  const std::string fname = g_prefix + ".0.bin";
  FILE *fp = std::fopen(fname.c_str(), "wb");
  if (!fp) std::abort();
  auto &v0 = pmesh->block_list[0]->meshblock_data.Get()->Get("neighbor_info");
  const int nc = v0.data.GetDim(4), nk = v0.data.GetDim(3), nj = v0.data.GetDim(2),
            ni = v0.data.GetDim(1);
  int hdr[7] = {0x50423230, static_cast<int>(pmesh->block_list.size()), nc, nk, nj, ni, 0};
  std::fwrite(hdr, sizeof(int), 7, fp);
  double td[2] = {0.0, 0.0};
  std::fwrite(td, sizeof(double), 2, fp);
  for (auto &pmb : pmesh->block_list) {
    int bh[5] = {pmb->gid, pmb->loc.level(), static_cast<int>(pmb->loc.lx1()),
                 static_cast<int>(pmb->loc.lx2()), static_cast<int>(pmb->loc.tree())};
    std::fwrite(bh, sizeof(int), 5, fp);
    double bb[6] = {pmb->block_size.xmin(X1DIR), pmb->block_size.xmin(X2DIR),
                    pmb->block_size.xmin(X3DIR), pmb->block_size.xmax(X1DIR),
                    pmb->block_size.xmax(X2DIR), pmb->block_size.xmax(X3DIR)};
    std::fwrite(bb, sizeof(double), 6, fp);
    auto &v = pmb->meshblock_data.Get()->Get("neighbor_info");
    auto h = v.data.GetHostMirrorAndCopy();
    for (int c = 0; c < nc; ++c)
      for (int k = 0; k < nk; ++k)
        for (int j = 0; j < nj; ++j)
          for (int i = 0; i < ni; ++i) {
            double x = h(0, 0, 0, c, k, j, i);
            std::fwrite(&x, sizeof(double), 1, fp);
          }
  }
  std::fclose(fp);
}
} // namespace

int main(int argc, char *argv[]) {
  ParthenonManager pman;
  if (const char *p = std::getenv("PB2_DUMP_PREFIX")) g_prefix = p;
  int variant = 0;
  if (const char *p = std::getenv("PB2_FOREST_VARIANT")) variant = std::atoi(p);
  pman.app_input->ProcessPackages = ProcessPackages;
  pman.app_input->ProblemGenerator = ProblemGenerator;
  auto manager_status = pman.ParthenonInitEnv(argc, argv);
  if (manager_status == ParthenonStatus::complete) {
    pman.ParthenonFinalize();
    return 0;
  }
  if (manager_status == ParthenonStatus::error) {
    pman.ParthenonFinalize();
    return 1;
  }

  // 3 x 3 lattice of nodes, numbered as in example/boundary_exchange
  //   6---7---8
  //   | 3 | 4 |
  //   3---2---5
  //   | 0 | 1 |
  //   0---1---4
  std::unordered_map<uint64_t, std::shared_ptr<forest::Node>> n;
  n[0] = forest::Node::create(0, {0.0, 0.0});
  n[1] = forest::Node::create(1, {1.0, 0.0});
  n[2] = forest::Node::create(2, {1.0, 1.0});
  n[3] = forest::Node::create(3, {0.0, 1.0});
  n[4] = forest::Node::create(4, {2.0, 0.0});
  n[5] = forest::Node::create(5, {2.0, 1.0});
  n[6] = forest::Node::create(6, {0.0, 2.0});
  n[7] = forest::Node::create(7, {1.0, 2.0});
  n[8] = forest::Node::create(8, {2.0, 2.0});

  forest::ForestDefinition forest_def;
  using edge_t = forest::Edge;
  using ar3_t = std::array<Real, 3>;
  if (variant <= 1) {
    forest_def.AddFace(0, {n[1], n[2], n[0], n[3]}, ar3_t{0.0, 0.0, 0.0}, ar3_t{1.0, 1.0, 1.0});
    forest_def.AddFace(1, {n[1], n[4], n[2], n[5]}, ar3_t{2.0, 0.0, 0.0}, ar3_t{3.0, 1.0, 1.0});
    forest_def.AddFace(3, {n[3], n[2], n[6], n[7]}, ar3_t{0.0, 2.0, 0.0}, ar3_t{1.0, 3.0, 1.0});
    forest_def.AddFace(4, {n[2], n[5], n[7], n[8]}, ar3_t{2.0, 2.0, 0.0}, ar3_t{3.0, 3.0, 1.0});
  } else {
    // 0: rotated by 90 degrees; 1: rotated by 180 degrees; 3: reflected about x1; 4: as laid out
    forest_def.AddFace(0, {n[1], n[2], n[0], n[3]}, ar3_t{0.0, 0.0, 0.0}, ar3_t{1.0, 1.0, 1.0});
    forest_def.AddFace(1, {n[5], n[2], n[4], n[1]}, ar3_t{2.0, 0.0, 0.0}, ar3_t{3.0, 1.0, 1.0});
    forest_def.AddFace(3, {n[2], n[3], n[7], n[6]}, ar3_t{0.0, 2.0, 0.0}, ar3_t{1.0, 3.0, 1.0});
    forest_def.AddFace(4, {n[2], n[5], n[7], n[8]}, ar3_t{2.0, 2.0, 0.0}, ar3_t{3.0, 3.0, 1.0});
  }
  forest_def.AddBC(edge_t({n[0], n[1]}));
  forest_def.AddBC(edge_t({n[0], n[3]}));
  forest_def.AddBC(edge_t({n[1], n[4]}));
  forest_def.AddBC(edge_t({n[4], n[5]}));
  forest_def.AddBC(edge_t({n[6], n[7]}));
  forest_def.AddBC(edge_t({n[3], n[6]}));
  forest_def.AddBC(edge_t({n[5], n[8]}));
  forest_def.AddBC(edge_t({n[7], n[8]}));
  if (variant == 0) forest_def.AddInitialRefinement(LogicalLocation(0, 1, 0, 0, 0));
  if (variant == 3) {
    forest_def.AddInitialRefinement(LogicalLocation(3, 1, 1, 0, 0));
    forest_def.AddInitialRefinement(LogicalLocation(4, 1, 0, 1, 0));
  }
  // Mesh::Initialize: problem generator on every block, then the boundary exchange
  pman.ParthenonInitPackagesAndMesh(forest_def);
  Dump(pman.pmesh.get());
  pman.ParthenonFinalize();
  return 0;
}
