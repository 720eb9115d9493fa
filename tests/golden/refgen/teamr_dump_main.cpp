// Fixture generator (test infrastructure, NOT product code).
//
// ADAPTIVE variant of tecomm_dump_main.cpp: the same three non-cell-centred fields on a mesh
// that is remeshed every "cycle" by a purely geometric criterion (refine the blocks whose centre
// lies within a radius of a point that moves with the cycle number, derefine the others); the
// fields never evolve, so the dumps after every remesh pin the refine / derefine data movement
// of face, edge and node fields (restriction into a new parent, shared + internal prolongation
// into new children, ownership with newly refined blocks) and the exchange that follows.
//
// A minimal application on the UNMODIFIED reference library (libparthenon.a built out-of-tree,
// see make_fixtures.sh) that pins the ghost exchange of NON-CELL-CENTRED fields: one package
// with a face field (2 components), an edge field and a node field, all Metadata::FillGhost.
// The problem generator writes a block-dependent integer code into EVERY entry of every array
// (ghosts and shared elements included):
//     value = (gid + 1) * 1e6 + element * 1e5 + component * 5e4 + flat (k, j, i) index
// so that after the boundary exchange Mesh::Initialize performs, every entry tells which block
// (and which entry of it) it came from — in particular which block OWNS each shared face, edge
// and node (mesh/forest/block_ownership.cpp).  Only this file is ours.
//
// With $PB2_TOTH_ROE set the face field registers ProlongateInternalTothAndRoe (the
// divergence-preserving internal prolongation, pr_ops.hpp:384-470) instead of the default
// ProlongateInternalAverage.
//
// Dump layout: the one of burgers_dump_main.cpp; the "cycle" slot of the file name and header
// enumerates the fields: 0 = face [3 elements x 2 components], 1 = edge [3 x 1], 2 = node
// [1 x 1]; ncomp in the header is elements x components, extents are the padded array extents.
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "parthenon_manager.hpp"
#include <parthenon/package.hpp>
#include <prolong_restrict/pr_ops.hpp>

#include <amr_criteria/refinement_package.hpp>

namespace {
using namespace parthenon;
using namespace parthenon::package::prelude;
std::string g_prefix = "dump";
int g_cycle = 0;
AmrTag TagByPosition(MeshBlockData<Real> *rc) {
  auto pmb = rc->GetBlockPointer();
  const Real xc = 0.5 * (pmb->block_size.xmin(X1DIR) + pmb->block_size.xmax(X1DIR));
  const Real yc = 0.5 * (pmb->block_size.xmin(X2DIR) + pmb->block_size.xmax(X2DIR));
  const Real zc = pmb->pmy_mesh->ndim > 2
                      ? 0.5 * (pmb->block_size.xmin(X3DIR) + pmb->block_size.xmax(X3DIR))
                      : 0.0;
  const Real px = -0.25 + 0.125 * g_cycle, py = -0.125 + 0.0625 * g_cycle,
             pz = pmb->pmy_mesh->ndim > 2 ? 0.125 : 0.0;
  const Real r2 = (xc - px) * (xc - px) + (yc - py) * (yc - py) + (zc - pz) * (zc - pz);
  return r2 < 0.2 * 0.2 ? AmrTag::refine : AmrTag::derefine;
}
const char *kFields[3] = {"face", "edge", "node"};

Packages_t ProcessPackages(std::unique_ptr<ParameterInput> &pin) {
  Packages_t packages;
  auto pkg = std::make_shared<StateDescriptor>("tecomm");
  Metadata mface({Metadata::Face, Metadata::Independent, Metadata::FillGhost},
                 std::vector<int>{2});
  if (std::getenv("PB2_TOTH_ROE"))
    mface.RegisterRefinementOps<parthenon::refinement_ops::ProlongateSharedMinMod,
                                parthenon::refinement_ops::RestrictAverage,
                                parthenon::refinement_ops::ProlongateInternalTothAndRoe>();
  Metadata medge({Metadata::Edge, Metadata::Independent, Metadata::FillGhost});
  Metadata mnode({Metadata::Node, Metadata::Independent, Metadata::FillGhost});
  // $PB2_SHARED_OP = linear | constant: ProlongateSharedLinear / ProlongatePiecewiseConstant
  // instead of the default ProlongateSharedMinMod, for all three fields
  if (const char *op = std::getenv("PB2_SHARED_OP")) {
    using namespace parthenon::refinement_ops;
    for (Metadata *m : {&mface, &medge, &mnode}) {
      if (std::string(op) == "linear")
        m->RegisterRefinementOps<ProlongateSharedLinear, RestrictAverage,
                                 ProlongateInternalAverage>();
      else
        m->RegisterRefinementOps<ProlongatePiecewiseConstant, RestrictAverage,
                                 ProlongateInternalAverage>();
    }
  }
  pkg->AddField("face", mface);
  pkg->AddField("edge", medge);
  pkg->AddField("node", mnode);
  pkg->CheckRefinementBlock = TagByPosition;
  packages.Add(pkg);
  return packages;
}

void ProblemGenerator(MeshBlock *pmb, ParameterInput *pin) {
  auto &rc = pmb->meshblock_data.Get();
  for (int f = 0; f < 3; ++f) {
    auto &v = rc->Get(kFields[f]);
    auto h = v.data.GetHostMirror();
    const int ne = v.data.GetDim(7), nc = v.data.GetDim(4), nk = v.data.GetDim(3),
              nj = v.data.GetDim(2), ni = v.data.GetDim(1);
    for (int e = 0; e < ne; ++e)
      for (int c = 0; c < nc; ++c)
        for (int k = 0; k < nk; ++k)
          for (int j = 0; j < nj; ++j)
            for (int i = 0; i < ni; ++i)
              h(e, 0, 0, c, k, j, i) =
                  (pmb->gid + 1) * 1.0e6 + e * 1.0e5 + c * 5.0e4 + ((k * nj + j) * ni + i);
    v.data.DeepCopy(h);
  }
}

void Dump(Mesh *pmesh, int cycle) {
  for (int f = 0; f < 3; ++f) {
    const std::string fname = g_prefix + "." + std::to_string(3 * cycle + f) + ".bin";
    FILE *fp = std::fopen(fname.c_str(), "wb");
    if (!fp) std::abort();
    auto &v0 = pmesh->block_list[0]->meshblock_data.Get()->Get(kFields[f]);
    const int ne = v0.data.GetDim(7), nc = v0.data.GetDim(4), nk = v0.data.GetDim(3),
              nj = v0.data.GetDim(2), ni = v0.data.GetDim(1);
    int hdr[7] = {0x50423230, static_cast<int>(pmesh->block_list.size()), ne * nc, nk, nj, ni,
                  3 * cycle + f};
    std::fwrite(hdr, sizeof(int), 7, fp);
    double td[2] = {0.0, 0.0};
    std::fwrite(td, sizeof(double), 2, fp);
    for (auto &pmb : pmesh->block_list) {
      int bh[5] = {pmb->gid, pmb->loc.level(), static_cast<int>(pmb->loc.lx1()),
                   static_cast<int>(pmb->loc.lx2()), static_cast<int>(pmb->loc.lx3())};
      std::fwrite(bh, sizeof(int), 5, fp);
      double bb[6] = {pmb->block_size.xmin(X1DIR), pmb->block_size.xmin(X2DIR),
                      pmb->block_size.xmin(X3DIR), pmb->block_size.xmax(X1DIR),
                      pmb->block_size.xmax(X2DIR), pmb->block_size.xmax(X3DIR)};
      std::fwrite(bb, sizeof(double), 6, fp);
      auto &v = pmb->meshblock_data.Get()->Get(kFields[f]);
      auto h = v.data.GetHostMirrorAndCopy();
      for (int e = 0; e < ne; ++e)
        for (int c = 0; c < nc; ++c)
          for (int k = 0; k < nk; ++k)
            for (int j = 0; j < nj; ++j)
              for (int i = 0; i < ni; ++i) {
                double x = h(e, 0, 0, c, k, j, i);
                std::fwrite(&x, sizeof(double), 1, fp);
              }
    }
    std::fclose(fp);
  }
}
} // namespace

int main(int argc, char *argv[]) {
  ParthenonManager pman;
  if (const char *p = std::getenv("PB2_DUMP_PREFIX")) g_prefix = p;
  pman.app_input->ProcessPackages = ProcessPackages;
  pman.app_input->ProblemGenerator = ProblemGenerator;
  pman.app_input->RegisterDefaultReflectingBoundaryConditions(); // "reflecting" in a deck
  auto manager_status = pman.ParthenonInitEnv(argc, argv);
  if (manager_status == ParthenonStatus::complete) {
    pman.ParthenonFinalize();
    return 0;
  }
  if (manager_status == ParthenonStatus::error) {
    pman.ParthenonFinalize();
    return 1;
  }
  // Mesh::Initialize: problem generator on every block, then CommunicateBoundaries
  pman.ParthenonInitPackagesAndMesh();
  Mesh *pm = pman.pmesh.get();
  Dump(pm, 0);
  const int ncycles = std::getenv("PB2_CYCLES") ? std::atoi(std::getenv("PB2_CYCLES")) : 6;
  for (int c = 1; c <= ncycles; ++c) {
    g_cycle = c;
    for (auto &pmb : pm->block_list)
      parthenon::Refinement::Tag(pmb->meshblock_data.Get().get());
    pm->LoadBalancingAndAdaptiveMeshRefinement(pman.pinput.get(), pman.app_input.get());
    Dump(pm, c);
  }
  pman.ParthenonFinalize();
  return 0;
}
