// tecomm_app.cpp — see tecomm_app.hpp
#include "tecomm_app.hpp"

#include <vector>

namespace tecomm_example {
using namespace parthenon;

namespace {
int g_cycle = 0;
void TagByPosition(MeshData<Real> *md, std::vector<AmrTag> &tags) {
  const int ndim = md->GetMeshPointer()->ndim;
  for (int b = 0; b < md->NumBlocks(); ++b) {
    const RegionSize &bs = md->GetBlock(b)->block_size;
    const Real xc = 0.5 * (bs.xmin_[0] + bs.xmax_[0]), yc = 0.5 * (bs.xmin_[1] + bs.xmax_[1]);
    const Real zc = ndim > 2 ? 0.5 * (bs.xmin_[2] + bs.xmax_[2]) : 0.0;
    const Real px = -0.25 + 0.125 * g_cycle, py = -0.125 + 0.0625 * g_cycle,
               pz = ndim > 2 ? 0.125 : 0.0;
    const Real r2 = (xc - px) * (xc - px) + (yc - py) * (yc - py) + (zc - pz) * (zc - pz);
    tags[b] = r2 < 0.2 * 0.2 ? AmrTag::refine : AmrTag::derefine;
  }
}
} // namespace

void SetCriterionCycle(int cycle) { g_cycle = cycle; }

Packages_t ProcessPackages(std::unique_ptr<ParameterInput> &pin) {
  Packages_t packages;
  auto pkg = std::make_shared<StateDescriptor>("tecomm");
  Metadata mface({Metadata::Face, Metadata::Independent, Metadata::FillGhost},
                 std::vector<int>{2});
  // tecomm/toth_roe = true: the divergence-preserving internal prolongation for the face field
  if (pin->GetOrAddBoolean("tecomm", "toth_roe", false))
    mface.RegisterRefinementOps<refinement_ops::ProlongateSharedMinMod,
                                refinement_ops::RestrictAverage,
                                refinement_ops::ProlongateInternalTothAndRoe>();
  Metadata medge({Metadata::Edge, Metadata::Independent, Metadata::FillGhost});
  Metadata mnode({Metadata::Node, Metadata::Independent, Metadata::FillGhost});
  // tecomm/shared_op = linear | constant: the other stock shared prolongations for all fields
  const std::string op = pin->GetOrAddString("tecomm", "shared_op", "minmod");
  PARTHENON_REQUIRE(op == "minmod" || op == "linear" || op == "constant",
                    "tecomm/shared_op must be minmod, linear or constant");
  if (op != "minmod")
    for (Metadata *m : {&mface, &medge, &mnode}) {
      if (op == "linear")
        m->RegisterRefinementOps<refinement_ops::ProlongateSharedLinear,
                                 refinement_ops::RestrictAverage>();
      else
        m->RegisterRefinementOps<refinement_ops::ProlongatePiecewiseConstant,
                                 refinement_ops::RestrictAverage>();
    }
  pkg->AddField("face", mface);
  pkg->AddField("edge", medge);
  pkg->AddField("node", mnode);
  // tecomm/flux_field = true: a face field B with fluxes — its flux is the edge field
  // "bnd_flux::B" — for the flux-correction tests (tests/golden/refgen/teflux_dump_main.cpp)
  if (pin->GetOrAddBoolean("tecomm", "flux_field", false))
    pkg->AddField("B", Metadata({Metadata::Face, Metadata::Independent, Metadata::WithFluxes,
                                 Metadata::FillGhost}));
  pkg->CheckRefinementMesh = TagByPosition;
  packages.Add(pkg);
  return packages;
}

// value = (gid + 1) * 1e6 + element * 1e5 + component * 5e4 + flat (k, j, i) index, in every
// entry of every array (ghosts and shared elements included)
void MeshProblemGenerator(MeshData<Real> *md, ParameterInput *) {
  const int nb = md->NumBlocks();
  for (const char *name : {"face", "edge", "node"}) {
    Variable &v = md->Get(name);
    const int nc = v.TensorComponents();
    std::vector<Real> h(static_cast<size_t>(nb) * v.block_stride);
    for (int b = 0; b < nb; ++b) {
      const int gid = md->GetBlock(b)->gid;
      Real *vb = h.data() + static_cast<size_t>(b) * v.block_stride;
      for (int e = 0; e < v.NumElements(); ++e)
        for (int c = 0; c < nc; ++c)
          for (int64_t n = 0; n < v.comp_stride; ++n)
            vb[(e * nc + c) * v.comp_stride + n] =
                (gid + 1) * 1.0e6 + e * 1.0e5 + c * 5.0e4 + static_cast<Real>(n);
    }
    PB2_CHECK(pb2_memcpy_h2d(v.data(), h.data(), sizeof(Real) * h.size(), md->stream()));
    PB2_CHECK(pb2_stream_sync(md->stream()));
  }
}

} // namespace tecomm_example
