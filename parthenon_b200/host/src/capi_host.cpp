// capi_host.cpp — C entry points of the host library (include/parthenon_b200_host.h).
// They let bench.py, the parity tests and foreign-language callers drive the C++ framework
// (ParthenonManager + BurgersDriver) without a C++ toolchain; exceptions never cross the
// boundary: every function returns 0 or a negative code and records the message.
#include <cstring>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../advection/advection_driver.hpp"
#include "../burgers/burgers_driver.hpp"
#include "../sparse_advection/sparse_advection_driver.hpp"
#include "../burgers/burgers_package.hpp"
#include "../tecomm/tecomm_app.hpp"
#include "../forest/forest_app.hpp"
#include "parthenon_b200_host.h"
#include "pb2/sparse_pack.hpp"
#include "pb2/parthenon.hpp"

using namespace parthenon;

namespace {
thread_local std::string g_error;

template <class F>
int Guard(F &&f) {
  try {
    f();
    return 0;
  } catch (const std::exception &e) {
    g_error = e.what();
    return -1;
  } catch (...) {
    g_error = "unknown exception";
    return -1;
  }
}

std::vector<std::string> SplitLines(const char *s) {
  std::vector<std::string> out;
  if (!s) return out;
  std::istringstream in(s);
  std::string line;
  while (std::getline(in, line))
    if (!line.empty()) out.push_back(line);
  return out;
}

// Two USER boundary conditions every application created through this interface can select in
// its deck ("pb2_user_outflow", "pb2_user_reflect"): written the way a downstream code would
// write its own — a function on a MeshData batch that builds C-ABI regions for the blocks on
// the face and launches them on the batch's stream.  They do what the stock outflow /
// reflecting conditions do, which is how the tests check the ApplicationInput::
// RegisterBoundaryCondition plumbing (call order relative to the stock directions, coarse
// buffers before prolongation, fine arrays after it).
void RegisterExampleUserBoundaries(ApplicationInput *app) {
  for (int face = 0; face < 6; ++face)
    for (int type : {PB2_BC_OUTFLOW, PB2_BC_REFLECT})
      app->RegisterBoundaryCondition(
          face, type == PB2_BC_OUTFLOW ? "pb2_user_outflow" : "pb2_user_reflect",
          [face, type](std::shared_ptr<MeshData<Real>> &md, bool coarse) {
            std::vector<pb2_bc_region> regs;
            for (auto &pmb : md->GetBlockList()) {
              if (pmb->boundary_flag[face] != BoundaryFlag::user) continue;
              for (Variable *v : md->GetVariablesByFlag({Metadata::FillGhost}))
                if (v->IsAllocated(pmb->pack_index))
                  regs.push_back(MakeBcRegion(*v, pmb.get(), face, type, coarse));
            }
            if (regs.empty()) return;
            pb2_bnd_table *table = nullptr;
            PB2_CHECK(pb2_bc_table_create(&table, regs.data(), static_cast<int64_t>(regs.size())));
            PB2_CHECK(pb2_apply_bcs(table, md->stream()));
            PB2_CHECK(pb2_stream_sync(md->stream())); // the table goes away
            PB2_CHECK(pb2_bnd_table_destroy(table));
          });
}

std::vector<LogicalLocation> Leaves(const int *leaves, int n) {
  std::vector<LogicalLocation> out;
  for (int i = 0; i < n; ++i) {
    LogicalLocation l;
    l.level = leaves[4 * i];
    for (int d = 0; d < 3; ++d) l.lx[d] = leaves[4 * i + 1 + d];
    out.push_back(l);
  }
  return out;
}
} // namespace

struct pb2h_sim {
  ParthenonManager pman;
  std::unique_ptr<MultiStageDriver> driver;
  bool topology_only = false;
  DeviceBuffer staging; // packed-interior staging for upload / download through host buffers
  // double-buffered host <-> device lanes (pb2h_sim_prefetch_interior ...): copies run on their
  // own streams so the H2D of batch n+1 and the D2H of batch n-1 overlap the cycle of batch n
  struct Lane {
    DeviceBuffer in, out;
    pb2_event_t h2d_done = nullptr, in_free = nullptr, gathered = nullptr, d2h_done = nullptr;
  };
  static constexpr int kLanes = 2;
  Lane lanes[kLanes];
  pb2_stream_t s_h2d = nullptr, s_d2h = nullptr;
  ~pb2h_sim() {
    for (auto &l : lanes)
      for (pb2_event_t e : {l.h2d_done, l.in_free, l.gathered, l.d2h_done})
        if (e) pb2_event_destroy(e);
    if (s_h2d) pb2_stream_destroy(s_h2d);
    if (s_d2h) pb2_stream_destroy(s_d2h);
  }
  // topology-only objects
  std::unique_ptr<ParameterInput> pin;
  std::unique_ptr<Mesh> mesh;
  Mesh *pm() { return topology_only ? mesh.get() : pman.pmesh.get(); }
};

extern "C" {

const char *pb2h_last_error(void) { return g_error.c_str(); }

int pb2h_sim_create(pb2h_sim **sim, const char *app, const char *deck, const char *overrides,
                    int rank, int nranks, const uint8_t *nccl_id, const int *leaves,
                    int nleaves) {
  return Guard([&] {
    PARTHENON_REQUIRE(sim && app && deck, "null argument");
    const std::string a(app);
    PARTHENON_REQUIRE(a == "burgers" || a == "advection" || a == "sparse_advection" ||
                          a == "tecomm" || a == "forest",
                      "unknown application (have: burgers, advection, sparse_advection, tecomm, "
                      "forest)");
    auto s = std::make_unique<pb2h_sim>();
    if (a == "burgers") {
      s->pman.app_input->ProcessPackages = burgers_benchmark::ProcessPackages;
      s->pman.app_input->MeshProblemGenerator = burgers_benchmark::MeshProblemGenerator;
    } else if (a == "advection") {
      s->pman.app_input->ProcessPackages = advection_example::ProcessPackages;
      s->pman.app_input->MeshProblemGenerator = advection_example::MeshProblemGenerator;
    } else if (a == "tecomm") {
      s->pman.app_input->ProcessPackages = tecomm_example::ProcessPackages;
      s->pman.app_input->MeshProblemGenerator = tecomm_example::MeshProblemGenerator;
    } else if (a == "forest") {
      s->pman.app_input->ProcessPackages = forest_example::ProcessPackages;
      s->pman.app_input->MeshProblemGenerator = forest_example::MeshProblemGenerator;
    } else {
      s->pman.app_input->ProcessPackages = sparse_advection_example::ProcessPackages;
      s->pman.app_input->MeshProblemGenerator = sparse_advection_example::MeshProblemGenerator;
    }
    RegisterExampleUserBoundaries(s->pman.app_input.get());
    tecomm_example::SetCriterionCycle(0);
    s->pman.ParthenonInitEnvFromString(deck, SplitLines(overrides));
    s->pman.SetRank(rank, nranks, nccl_id);
    if (a == "forest") // the mesh is a forest built in code, as example/boundary_exchange does
      s->pman.ParthenonInitPackagesAndMesh(forest_example::MakeForest(
          s->pman.pinput->GetOrAddInteger("forest", "variant", 0)));
    else
      s->pman.ParthenonInitPackagesAndMesh(Leaves(leaves, nleaves));
    std::unique_ptr<MultiStageDriver> drv;
    if (a == "burgers")
      drv = std::make_unique<burgers_benchmark::BurgersDriver>(
          s->pman.pinput.get(), s->pman.app_input.get(), s->pman.pmesh.get());
    else if (a == "advection")
      drv = std::make_unique<advection_example::AdvectionDriver>(
          s->pman.pinput.get(), s->pman.app_input.get(), s->pman.pmesh.get());
    else if (a == "sparse_advection")
      drv = std::make_unique<sparse_advection_example::SparseAdvectionDriver>(
          s->pman.pinput.get(), s->pman.app_input.get(), s->pman.pmesh.get());
    // tecomm has no time loop: fields, the exchange and the field accessors only
    if (drv) drv->quiet = true;
    s->driver = std::move(drv);
    *sim = s.release();
  });
}

int pb2h_topology_create(pb2h_sim **sim, const char *deck, const char *overrides, int rank,
                         int nranks, const int *leaves, int nleaves) {
  return Guard([&] {
    PARTHENON_REQUIRE(sim && deck, "null argument");
    auto s = std::make_unique<pb2h_sim>();
    s->topology_only = true;
    s->pin = std::make_unique<ParameterInput>();
    s->pin->LoadFromString(deck);
    for (auto &o : SplitLines(overrides)) s->pin->ModifyFromString(o);
    Packages_t none;
    s->mesh = std::make_unique<Mesh>(s->pin.get(), nullptr, none, rank, nranks,
                                     Leaves(leaves, nleaves));
    *sim = s.release();
  });
}

int pb2h_topology_create_forest(pb2h_sim **sim, const char *deck, const char *overrides,
                                int variant, int rank, int nranks) {
  return Guard([&] {
    PARTHENON_REQUIRE(sim && deck, "null argument");
    auto s = std::make_unique<pb2h_sim>();
    s->topology_only = true;
    s->pin = std::make_unique<ParameterInput>();
    s->pin->LoadFromString(deck);
    for (auto &o : SplitLines(overrides)) s->pin->ModifyFromString(o);
    Packages_t none;
    s->mesh = std::make_unique<Mesh>(s->pin.get(), nullptr, none,
                                     forest_example::MakeForest(variant), rank, nranks);
    *sim = s.release();
  });
}

int pb2h_sim_block_bcs(pb2h_sim *sim, int lid, int out[6]) {
  return Guard([&] {
    Mesh *pm = sim->pm();
    PARTHENON_REQUIRE(out && lid >= 0 && lid < pm->GetNumMeshBlocksThisRank(), "bad block index");
    for (int f = 0; f < 6; ++f) out[f] = static_cast<int>(pm->block_list[lid]->boundary_flag[f]);
  });
}

int pb2h_topology_regrid(pb2h_sim *sim, const int *tags, int nblocks, int *changed) {
  return Guard([&] {
    PARTHENON_REQUIRE(sim && tags && changed, "null argument");
    Mesh *pm = sim->pm();
    PARTHENON_REQUIRE(sim->topology_only, "pb2h_topology_regrid works on topology objects");
    PARTHENON_REQUIRE(nblocks == pm->GetNumMeshBlocksThisRank(), "one tag per block");
    // Refinement::Tag's bookkeeping (SetRefinement: level limits, derefinement counters), then
    // the tree update and the new block list
    for (int b = 0; b < nblocks; ++b) pm->SetRefinement(b, static_cast<AmrTag>(tags[b]));
    *changed = pm->RegridTopologyOnly() ? 1 : 0;
  });
}

int pb2h_topology_derefine_counts(pb2h_sim *sim, int *counts, int nblocks, int set) {
  return Guard([&] {
    PARTHENON_REQUIRE(sim && counts, "null argument");
    Mesh *pm = sim->pm();
    PARTHENON_REQUIRE(nblocks == pm->GetNumMeshBlocksThisRank(), "one counter per block");
    for (int b = 0; b < nblocks; ++b) {
      if (set)
        pm->block_list[b]->deref_count = counts[b];
      else
        counts[b] = pm->block_list[b]->deref_count;
    }
  });
}

// tecomm on an adaptive mesh: one "cycle" of the fixture generator — tag every block with the
// criterion of `cycle`, then LoadBalancingAndAdaptiveMeshRefinement
int pb2h_sim_tag_and_remesh(pb2h_sim *sim, int cycle, int *changed) {
  return Guard([&] {
    Mesh *pm = sim->pm();
    PARTHENON_REQUIRE(!sim->topology_only && pm->DefaultNumPartitions() == 1,
                      "tag_and_remesh needs a simulation with one MeshData per rank");
    tecomm_example::SetCriterionCycle(cycle);
    Refinement::Tag(pm->mesh_data.GetOrAdd("base", 0).get());
    pm->LoadBalancingAndAdaptiveMeshRefinement(sim->pman.pinput.get(), sim->pman.app_input.get());
    if (changed) *changed = pm->modified ? 1 : 0;
  });
}

int pb2h_sim_destroy(pb2h_sim *sim) {
  return Guard([&] {
    if (!sim) return;
    sim->driver.reset();
    delete sim;
  });
}

int pb2h_sim_pre_execute(pb2h_sim *sim) {
  return Guard([&] {
    PARTHENON_REQUIRE(sim->driver != nullptr, "this application has no time loop");
    sim->driver->PreExecute();
  });
}

int pb2h_sim_cycle(pb2h_sim *sim, int ncycles) {
  return Guard([&] {
    PARTHENON_REQUIRE(sim->driver != nullptr, "this application has no time loop");
    for (int c = 0; c < ncycles; ++c)
      PARTHENON_REQUIRE(sim->driver->DoCycle() == TaskListStatus::complete,
                        "Step failed to complete all tasks.");
  });
}

int pb2h_sim_cycle_phase(pb2h_sim *sim, int phase) {
  return Guard([&] {
    PARTHENON_REQUIRE(sim->driver != nullptr, "this application has no time loop");
    if (phase == 0) {
      PARTHENON_REQUIRE(sim->driver->StepAndAdvanceTime() == TaskListStatus::complete,
                        "Step failed to complete all tasks.");
    } else {
      sim->driver->AdaptMeshAndSetTimeStep();
    }
  });
}

int pb2h_sim_sync(pb2h_sim *sim) {
  return Guard([&] { PB2_CHECK(pb2_stream_sync(sim->pm()->stream)); });
}

void *pb2h_sim_stream(pb2h_sim *sim) { return sim->pm()->stream; }
double pb2h_sim_time(pb2h_sim *sim) { return sim->driver ? sim->driver->tm.time : 0.0; }
double pb2h_sim_dt(pb2h_sim *sim) { return sim->driver ? sim->driver->tm.dt : 0.0; }
int pb2h_sim_ncycle(pb2h_sim *sim) { return sim->driver ? sim->driver->tm.ncycle : 0; }
int pb2h_sim_set_dt(pb2h_sim *sim, double dt) {
  sim->driver->tm.dt = dt;
  return 0;
}

int pb2h_sim_info(pb2h_sim *sim, int out[12]) {
  return Guard([&] {
    Mesh *pm = sim->pm();
    const IndexShape &cb = pm->block_list[0]->cellbounds, &ccb = pm->block_list[0]->c_cellbounds;
    out[0] = pm->ndim;
    out[1] = pm->nbtotal;
    out[2] = pm->GetNumMeshBlocksThisRank();
    out[3] = cb.ncellsi(IndexDomain::entire);
    out[4] = cb.ncellsj(IndexDomain::entire);
    out[5] = cb.ncellsk(IndexDomain::entire);
    out[6] = ccb.ncellsi(IndexDomain::entire);
    out[7] = ccb.ncellsj(IndexDomain::entire);
    out[8] = ccb.ncellsk(IndexDomain::entire);
    out[9] = pm->multilevel ? 1 : 0;
    out[10] = pm->block_list[0]->gid;
    out[11] = Globals::nghost;
  });
}

int pb2h_sim_block(pb2h_sim *sim, int lid, int loc[4], double xmin[3], double xmax[3],
                   int *gid, int *nneighbors) {
  return Guard([&] {
    Mesh *pm = sim->pm();
    PARTHENON_REQUIRE(lid >= 0 && lid < pm->GetNumMeshBlocksThisRank(), "bad block index");
    const MeshBlock &mb = *pm->block_list[lid];
    loc[0] = mb.loc.level;
    for (int d = 0; d < 3; ++d) {
      loc[1 + d] = static_cast<int>(mb.loc.lx[d]);
      xmin[d] = mb.block_size.xmin_[d];
      xmax[d] = mb.block_size.xmax_[d];
    }
    if (pm->forest) loc[3] = static_cast<int>(mb.loc.tree); // 2-D: lx3 is 0, the slot names the tree
    *gid = mb.gid;
    *nneighbors = static_cast<int>(mb.neighbors.size());
  });
}

int pb2h_sim_neighbor(pb2h_sim *sim, int lid, int n, int out[6]) {
  return Guard([&] {
    Mesh *pm = sim->pm();
    const NeighborBlock &nb = pm->block_list.at(lid)->neighbors.at(n);
    out[0] = nb.gid;
    out[1] = nb.loc.level;
    out[2] = nb.offsets[0];
    out[3] = nb.offsets[1];
    out[4] = nb.offsets[2];
    out[5] = nb.rank;
  });
}

int pb2h_sim_calc_indices(pb2h_sim *sim, int lid, int n, int ir_type, int prores, int s[3],
                          int e[3]) {
  return Guard([&] {
    Mesh *pm = sim->pm();
    const MeshBlock &mb = *pm->block_list.at(lid);
    const IndexBox box = CalcIndices(mb.neighbors.at(n), &mb,
                                     ir_type == 0 ? IndexRangeType::BoundaryInteriorSend
                                                  : IndexRangeType::BoundaryExteriorRecv,
                                     prores != 0);
    for (int d = 0; d < 3; ++d) {
      s[d] = box.s[d];
      e[d] = box.e[d];
    }
  });
}

int pb2h_sim_ranklist(pb2h_sim *sim, int *ranks, int n) {
  return Guard([&] {
    Mesh *pm = sim->pm();
    PARTHENON_REQUIRE(n == pm->nbtotal, "ranklist length mismatch");
    for (int g = 0; g < n; ++g) ranks[g] = pm->ranklist[g];
  });
}

// kind: 0 local, 1 send, 2 recv.  rows: [sender_gid, receiver_gid, var, offset_index,
// slab_off, nelem, peer] as int64
int64_t pb2h_sim_plan(pb2h_sim *sim, int ncomp, int kind, int64_t *rows, int64_t max_rows,
                      int64_t *seg_off) {
  int64_t count = -1;
  Guard([&] {
    Mesh *pm = sim->pm();
    const ExchangePlan plan = BuildExchangePlan(pm, pm->block_list, {PlanVar{ncomp}});
    const std::vector<Channel> &chs = kind == 0 ? plan.local : (kind == 1 ? plan.send : plan.recv);
    count = static_cast<int64_t>(chs.size());
    if (rows) {
      for (int64_t i = 0; i < std::min<int64_t>(count, max_rows); ++i) {
        const Channel &c = chs[i];
        int64_t *r = rows + 7 * i;
        r[0] = c.sender_gid;
        r[1] = c.receiver_gid;
        r[2] = c.var;
        r[3] = c.offset_index;
        r[4] = c.slab_off;
        r[5] = (kind == 2 ? c.recv_box.size() : c.send_box.size()) * ncomp;
        r[6] = kind == 1 ? c.receiver_rank : c.sender_rank;
      }
    }
    if (seg_off && kind != 0) {
      const auto &off = kind == 1 ? plan.send_off : plan.recv_off;
      for (size_t p = 0; p < off.size(); ++p) seg_off[p] = off[p];
    }
  });
  return count;
}

// the same for a field of topological type tt (0 cell, 1 face, 2 edge, 3 node), with the index
// boxes: rows of 18 int64 [sender_gid, receiver_gid, offset_index, piece, comp0, ncomp,
// send_s(i,j,k), recv_s(i,j,k), n(i,j,k), slab_off, peer, coarse flags (1: the sender reads its
// coarse buffer, 2: the receiver writes its coarse buffer)]
int64_t pb2h_sim_plan_boxes(pb2h_sim *sim, int ncomp, int tt, int kind, int64_t *rows,
                            int64_t max_rows) {
  int64_t count = -1;
  Guard([&] {
    PARTHENON_REQUIRE(tt >= 0 && tt <= 3, "topological type: 0 cell, 1 face, 2 edge, 3 node");
    Mesh *pm = sim->pm();
    const ExchangePlan plan = BuildExchangePlan(
        pm, pm->block_list, {PlanVar{ncomp, static_cast<TopologicalType>(tt)}});
    const std::vector<Channel> &chs = kind == 0 ? plan.local : (kind == 1 ? plan.send : plan.recv);
    count = static_cast<int64_t>(chs.size());
    if (!rows) return;
    for (int64_t i = 0; i < std::min<int64_t>(count, max_rows); ++i) {
      const Channel &c = chs[i];
      int64_t *r = rows + 18 * i;
      r[0] = c.sender_gid;
      r[1] = c.receiver_gid;
      r[2] = c.offset_index;
      r[3] = c.piece;
      r[4] = c.comp0;
      r[5] = c.ncomp;
      for (int d = 0; d < 3; ++d) {
        r[6 + d] = c.send_box.s[d];
        r[9 + d] = c.recv_box.s[d];
        r[12 + d] = c.recv_box.n(d);
      }
      r[15] = c.slab_off;
      r[16] = kind == 1 ? c.receiver_rank : c.sender_rank;
      r[17] = (c.send_coarse ? 1 : 0) | (c.recv_coarse ? 2 : 0);
      if (c.transformed) {
        r[17] |= 4;
        for (int d = 0; d < 3; ++d) {
          r[17] |= static_cast<int64_t>(std::abs(c.lcoord_trans.dir_connection[d])) << (3 + 2 * d);
          r[17] |= static_cast<int64_t>(c.lcoord_trans.dir_flip[d] ? 1 : 0) << (9 + d);
        }
        r[17] |= static_cast<int64_t>(c.ncell) << 12;
      }
    }
  });
  return count;
}

static Variable &FindVar(pb2h_sim *sim, const char *container, const char *field) {
  PARTHENON_REQUIRE(sim->pm()->DefaultNumPartitions() == 1,
                    "field access through the C interface needs pack_size = -1");
  auto &md = sim->pm()->mesh_data.GetOrAdd(container, 0);
  EnsureLocalGhosts(md.get()); // whoever looks at a field sees current ghost cells
  return md->Get(field);
}

int pb2h_sim_field_ptr(pb2h_sim *sim, const char *container, const char *field, int which,
                       void **ptr, int64_t *nreal) {
  return Guard([&] {
    Variable &v = FindVar(sim, container, field);
    const int nb = sim->pm()->GetNumMeshBlocksThisRank();
    if (which == 0) {
      *ptr = v.data();
      *nreal = v.block_stride * nb;
    } else if (which == 4) {
      *ptr = v.coarse();
      *nreal = v.cblock_stride * nb;
    } else {
      *ptr = v.flux(which);
      *nreal = v.block_stride * nb;
    }
  });
}

int pb2h_sim_field_dims(pb2h_sim *sim, const char *container, const char *field, int out[6]) {
  return Guard([&] {
    Variable &v = FindVar(sim, container, field);
    out[0] = sim->pm()->GetNumMeshBlocksThisRank();
    out[1] = v.NumComponents();
    out[2] = v.nk;
    out[3] = v.nj;
    out[4] = v.ni;
    out[5] = v.NumElements();
  });
}

int pb2h_sim_get_field(pb2h_sim *sim, const char *container, const char *field, int which,
                       double *host, int64_t nreal) {
  return Guard([&] {
    void *p = nullptr;
    int64_t n = 0;
    PARTHENON_REQUIRE(pb2h_sim_field_ptr(sim, container, field, which, &p, &n) == 0, g_error);
    PARTHENON_REQUIRE(n == nreal, "field size mismatch");
    PB2_CHECK(pb2_memcpy_d2h(host, p, sizeof(double) * n, sim->pm()->stream));
    PB2_CHECK(pb2_stream_sync(sim->pm()->stream));
  });
}

int pb2h_sim_allocation(pb2h_sim *sim, const char *container, const char *field, int *out,
                        int nblocks) {
  return Guard([&] {
    PARTHENON_REQUIRE(sim && container && field && out, "null argument");
    Mesh *pm = sim->pm();
    int b = 0;
    for (int p = 0; p < pm->DefaultNumPartitions(); ++p) {
      auto &md = pm->mesh_data.GetOrAdd(container, p);
      Variable &v = md->Get(field);
      for (int i = 0; i < md->NumBlocks(); ++i, ++b) {
        PARTHENON_REQUIRE(b < nblocks, "allocation buffer too small");
        out[b] = v.IsAllocated(i) ? 1 : 0;
      }
    }
    PARTHENON_REQUIRE(b == nblocks, "allocation buffer size mismatch");
  });
}

int pb2h_sim_set_field(pb2h_sim *sim, const char *container, const char *field, int which,
                       const double *host, int64_t nreal) {
  return Guard([&] {
    void *p = nullptr;
    int64_t n = 0;
    PARTHENON_REQUIRE(pb2h_sim_field_ptr(sim, container, field, which, &p, &n) == 0, g_error);
    PARTHENON_REQUIRE(n == nreal, "field size mismatch");
    PB2_CHECK(pb2_memcpy_h2d(p, host, sizeof(double) * n, sim->pm()->stream));
    PB2_CHECK(pb2_stream_sync(sim->pm()->stream));
  });
}

// Upload / read back the INTERIOR cells of a field through host buffers laid out
// [block][comp][nx3][nx2][nx1]: H2D into a device staging buffer, one scatter launch, one ghost
// exchange (upload); one gather launch, D2H (download).  All asynchronous on the app stream.
static DeviceBuffer &Staging(pb2h_sim *sim, size_t bytes) {
  if (sim->staging.bytes() < bytes) sim->staging.Allocate(bytes, sim->pm()->stream);
  return sim->staging;
}

int pb2h_sim_upload_interior(pb2h_sim *sim, const char *container, const char *field,
                             const double *host, int64_t nreal) {
  return Guard([&] {
    Variable &v = FindVar(sim, container, field);
    auto &md = sim->pm()->mesh_data.GetOrAdd(container, 0);
    pb2_pack_geom g = md->Geometry(v);
    const int64_t n = static_cast<int64_t>(g.nblocks) * g.ncomp * sim->pm()->GetNumberOfMeshBlockCells();
    PARTHENON_REQUIRE(n == nreal, "interior size mismatch");
    DeviceBuffer &st = Staging(sim, sizeof(double) * n);
    PB2_CHECK(pb2_memcpy_h2d(st.get(), host, sizeof(double) * n, md->stream()));
    PB2_CHECK(pb2_interior_scatter(&g, st.get<double>(), v.data(), md->stream()));
    CommunicateBoundaries(md, true);
  });
}

int pb2h_sim_download_interior(pb2h_sim *sim, const char *container, const char *field,
                               double *host, int64_t nreal) {
  return Guard([&] {
    Variable &v = FindVar(sim, container, field);
    auto &md = sim->pm()->mesh_data.GetOrAdd(container, 0);
    pb2_pack_geom g = md->Geometry(v);
    const int64_t n = static_cast<int64_t>(g.nblocks) * g.ncomp * sim->pm()->GetNumberOfMeshBlockCells();
    PARTHENON_REQUIRE(n == nreal, "interior size mismatch");
    DeviceBuffer &st = Staging(sim, sizeof(double) * n);
    PB2_CHECK(pb2_interior_gather(&g, v.data(), st.get<double>(), md->stream()));
    PB2_CHECK(pb2_memcpy_d2h(host, st.get(), sizeof(double) * n, md->stream()));
  });
}

// ---- pipelined variant: independent batches of state streamed through two lanes ----------
namespace {
struct LaneCtx {
  pb2h_sim::Lane *lane;
  Variable *v;
  std::shared_ptr<MeshData<Real>> md;
  pb2_pack_geom g{};
  int64_t n;
};
LaneCtx GetLane(pb2h_sim *sim, const char *container, const char *field, int lane,
                int64_t nreal) {
  PARTHENON_REQUIRE(lane >= 0 && lane < pb2h_sim::kLanes, "lane out of range");
  LaneCtx c;
  c.v = &FindVar(sim, container, field);
  c.md = sim->pm()->mesh_data.GetOrAdd(container, 0);
  c.g = c.md->Geometry(*c.v);
  c.n = static_cast<int64_t>(c.g.nblocks) * c.g.ncomp * sim->pm()->GetNumberOfMeshBlockCells();
  PARTHENON_REQUIRE(nreal < 0 || c.n == nreal, "interior size mismatch");
  c.lane = &sim->lanes[lane];
  if (!sim->s_h2d) {
    PB2_CHECK(pb2_stream_create(&sim->s_h2d));
    PB2_CHECK(pb2_stream_create(&sim->s_d2h));
  }
  pb2h_sim::Lane &l = *c.lane;
  if (!l.h2d_done) {
    PB2_CHECK(pb2_event_create(&l.h2d_done));
    PB2_CHECK(pb2_event_create(&l.in_free));
    PB2_CHECK(pb2_event_create(&l.gathered));
    PB2_CHECK(pb2_event_create(&l.d2h_done));
  }
  const size_t bytes = sizeof(double) * c.n;
  if (l.in.bytes() < bytes) {
    l.in.Allocate(bytes, c.md->stream());
    l.out.Allocate(bytes, c.md->stream());
    PB2_CHECK(pb2_stream_sync(c.md->stream()));
  }
  return c;
}
} // namespace

int pb2h_sim_prefetch_interior(pb2h_sim *sim, const char *container, const char *field,
                               const double *host, int64_t nreal, int lane) {
  return Guard([&] {
    LaneCtx c = GetLane(sim, container, field, lane, nreal);
    // the lane's input staging is free once the scatter that last read it has run
    PB2_CHECK(pb2_stream_wait_event(sim->s_h2d, c.lane->in_free));
    PB2_CHECK(pb2_memcpy_h2d(c.lane->in.get(), host, sizeof(double) * c.n, sim->s_h2d));
    PB2_CHECK(pb2_event_record(c.lane->h2d_done, sim->s_h2d));
  });
}

int pb2h_sim_commit_interior(pb2h_sim *sim, const char *container, const char *field,
                             int lane) {
  return Guard([&] {
    LaneCtx c = GetLane(sim, container, field, lane, -1);
    PB2_CHECK(pb2_stream_wait_event(c.md->stream(), c.lane->h2d_done));
    PB2_CHECK(pb2_interior_scatter(&c.g, c.lane->in.get<double>(), c.v->data(), c.md->stream()));
    PB2_CHECK(pb2_event_record(c.lane->in_free, c.md->stream()));
    CommunicateBoundaries(c.md, true);
  });
}

int pb2h_sim_writeback_interior(pb2h_sim *sim, const char *container, const char *field,
                                double *host, int64_t nreal, int lane) {
  return Guard([&] {
    LaneCtx c = GetLane(sim, container, field, lane, nreal);
    // the lane's output staging is free once its previous D2H has drained
    PB2_CHECK(pb2_stream_wait_event(c.md->stream(), c.lane->d2h_done));
    PB2_CHECK(pb2_interior_gather(&c.g, c.v->data(), c.lane->out.get<double>(), c.md->stream()));
    PB2_CHECK(pb2_event_record(c.lane->gathered, c.md->stream()));
    PB2_CHECK(pb2_stream_wait_event(sim->s_d2h, c.lane->gathered));
    PB2_CHECK(pb2_memcpy_d2h(host, c.lane->out.get(), sizeof(double) * c.n, sim->s_d2h));
    PB2_CHECK(pb2_event_record(c.lane->d2h_done, sim->s_d2h));
  });
}

int pb2h_sim_lane_sync(pb2h_sim *sim, int lane) {
  return Guard([&] {
    PARTHENON_REQUIRE(lane >= 0 && lane < pb2h_sim::kLanes, "lane out of range");
    if (sim->lanes[lane].d2h_done) PB2_CHECK(pb2_event_sync(sim->lanes[lane].d2h_done));
  });
}

int pb2h_sim_exchange(pb2h_sim *sim, const char *container, int prolongate) {
  return Guard([&] {
    auto &md = sim->pm()->mesh_data.GetOrAdd(container, 0);
    CommunicateBoundaries(md, prolongate != 0);
  });
}

// the three phases separately (for timing / overlap tests): 0 send, 1 receive+set, 2 prolongate
int pb2h_sim_exchange_phase(pb2h_sim *sim, const char *container, int phase) {
  return Guard([&] {
    auto &md = sim->pm()->mesh_data.GetOrAdd(container, 0);
    if (phase == 0) SendBoundBufs<BoundaryType::any>(md);
    if (phase == 1) {
      ReceiveBoundBufs<BoundaryType::any>(md);
      SetBounds<BoundaryType::any>(md);
    }
    if (phase == 2) ProlongateBounds<BoundaryType::any>(md);
  });
}

int64_t pb2h_sim_exchange_elements(pb2h_sim *sim, const char *container, int64_t *local,
                                   int64_t *nonlocal) {
  int64_t total = -1;
  Guard([&] {
    auto &md = sim->pm()->mesh_data.GetOrAdd(container, 0);
    BuildBoundaryBuffers(md);
    const ExchangePlan &p = *md->bvars().plan;
    if (local) *local = p.local_elements;
    if (nonlocal) *nonlocal = p.recv_elements;
    total = p.local_elements + p.recv_elements;
  });
  return total;
}

int64_t pb2h_sim_edge_flux_plan(pb2h_sim *sim, int kind, int64_t *rows, int64_t max_rows) {
  int64_t count = -1;
  Guard([&] {
    PARTHENON_REQUIRE(kind >= 0 && kind <= 4, "kind: 0 restrict, 1 deliver, 2 send restrict, "
                                               "3 send, 4 receive");
    Mesh *pm = sim->pm();
    const EdgeFluxPlan plan = BuildEdgeFluxPlan(pm, pm->block_list);
    const bool is_restrict = kind == 0 || kind == 2;
    const std::vector<EdgeFluxRestrict> &rs = kind == 0 ? plan.restricts : plan.send_restricts;
    const std::vector<EdgeFluxPiece> &ps =
        kind == 1 ? plan.pieces : (kind == 3 ? plan.send : plan.recv);
    count = static_cast<int64_t>(is_restrict ? rs.size() : ps.size());
    if (!rows) return;
    for (int64_t i = 0; i < std::min<int64_t>(count, max_rows); ++i) {
      int64_t *r = rows + 16 * i;
      for (int q = 0; q < 16; ++q) r[q] = 0;
      if (is_restrict) {
        const EdgeFluxRestrict &x = rs[i];
        r[0] = x.gid;
        r[2] = x.el;
        for (int d = 0; d < 3; ++d) {
          r[4 + d] = x.box.s[d];
          r[10 + d] = x.box.n(d);
        }
      } else {
        const EdgeFluxPiece &x = ps[i];
        r[0] = x.sender_gid;
        r[1] = x.receiver_gid;
        r[2] = x.el;
        r[3] = x.pass;
        const IndexBox &nbox = kind == 3 ? x.send_box : x.recv_box;
        for (int d = 0; d < 3; ++d) {
          r[4 + d] = kind == 4 ? 0 : x.send_box.s[d];
          r[7 + d] = kind == 3 ? 0 : x.recv_box.s[d];
          r[10 + d] = nbox.n(d);
        }
        r[13] = x.seg;
        r[14] = x.slab_off + (x.seg >= 0 ? (kind == 3 ? plan.send_off : plan.recv_off)[x.seg] : 0);
        r[15] = x.offset_index * 64 + x.sub;
      }
    }
  });
  return count;
}

int pb2h_sim_flux_correction(pb2h_sim *sim, const char *container) {
  return Guard([&] {
    PARTHENON_REQUIRE(sim && container, "null argument");
    for (int p = 0; p < sim->pm()->DefaultNumPartitions(); ++p)
      FluxCorrection(sim->pm()->mesh_data.GetOrAdd(container, p).get());
    PB2_CHECK(pb2_stream_sync(sim->pm()->stream));
  });
}

int pb2h_sim_exchange_mode(pb2h_sim *sim, const char *container) {
  int mode = -1;
  Guard([&] {
    auto &md = sim->pm()->mesh_data.GetOrAdd(container, 0);
    BuildBoundaryBuffers(md);
    const BvarsCache &c = md->bvars();
    if (c.plan->send_elements + c.plan->recv_elements == 0)
      mode = 0;
    else if (!c.push_mode)
      mode = 1;
    else
      mode = c.push_direct ? 4 : (c.push_ce ? 2 : 3);
  });
  return mode;
}

namespace {
MetadataFlag FlagByName(const std::string &n) {
  static const std::map<std::string, MetadataFlag> m = {
      {"Cell", Metadata::Cell}, {"Face", Metadata::Face}, {"Edge", Metadata::Edge},
      {"Node", Metadata::Node}, {"Independent", Metadata::Independent},
      {"Derived", Metadata::Derived}, {"OneCopy", Metadata::OneCopy},
      {"FillGhost", Metadata::FillGhost}, {"WithFluxes", Metadata::WithFluxes},
      {"Sparse", Metadata::Sparse}, {"Vector", Metadata::Vector},
      {"Conserved", Metadata::Conserved}, {"Intensive", Metadata::Intensive}};
  auto it = m.find(n);
  PARTHENON_REQUIRE(it != m.end(), "unknown Metadata flag " + n);
  return it->second;
}
SparsePack MakePack(pb2h_sim *sim, const char *container, const char *names, const char *flags,
                    int options) {
  PARTHENON_REQUIRE(sim->pm()->DefaultNumPartitions() == 1,
                    "packs through the C interface need pack_size = -1");
  auto &md = sim->pm()->mesh_data.GetOrAdd(container, 0);
  EnsureLocalGhosts(md.get());
  std::vector<std::string> vars;
  std::vector<bool> regex;
  for (const std::string &n : SplitLines(names)) {
    const bool re = n.compare(0, 3, "re:") == 0;
    vars.push_back(re ? n.substr(3) : n);
    regex.push_back(re);
  }
  std::vector<MetadataFlag> fl;
  for (const std::string &f : SplitLines(flags)) fl.push_back(FlagByName(f));
  std::set<PDOpt> opt;
  if (options & 1) opt.insert(PDOpt::WithFluxes);
  if (options & 2) opt.insert(PDOpt::Coarse);
  if (options & 4) opt.insert(PDOpt::Flatten);
  std::vector<const StateDescriptor *> pk;
  const Packages_t &packages = sim->pm()->packages;
  for (const std::string &name : packages.Order()) pk.push_back(packages.Get(name).get());
  return MakePackDescriptor(pk, vars, regex, fl, opt).GetPack(md.get());
}
} // namespace

int pb2h_sim_sparse_pack(pb2h_sim *sim, const char *container, const char *names,
                         const char *flags, int options, pb2_sparse_pack *pack,
                         int32_t *host_bounds, int64_t host_bounds_len) {
  return Guard([&] {
    PARTHENON_REQUIRE(pack != nullptr, "null pack");
    SparsePack p = MakePack(sim, container, names, flags, options);
    *pack = p.pod();
    if (host_bounds) {
      const int64_t n = 2ll * pack->nblocks_md * (pack->nvar + 1);
      PARTHENON_REQUIRE(host_bounds_len >= n, "bounds buffer too small");
      for (int w = 0; w < 2; ++w)
        for (int b = 0; b < pack->nblocks_md; ++b) {
          for (int v = 0; v < pack->nvar; ++v)
            host_bounds[(static_cast<int64_t>(w) * pack->nblocks_md + b) * (pack->nvar + 1) + v] =
                w == 0 ? p.GetLowerBoundHost(b, PackIdx(v)) : p.GetUpperBoundHost(b, PackIdx(v));
          host_bounds[(static_cast<int64_t>(w) * pack->nblocks_md + b) * (pack->nvar + 1) +
                      pack->nvar] = w == 0 ? p.GetLowerBoundHost(b) : p.GetUpperBoundHost(b);
        }
    }
  });
}

const char *pb2h_sim_sparse_pack_label(pb2h_sim *sim, const char *container, const char *names,
                                       const char *flags, int options, int b, int idx) {
  static std::string label;
  label.clear();
  Guard([&] { label = MakePack(sim, container, names, flags, options).LabelHost(b, idx); });
  return label.c_str();
}

int pb2h_sim_set_sparse_allocation(pb2h_sim *sim, const char *field, int lid, int allocated) {
  return Guard([&] {
    if (allocated)
      sim->pm()->AllocateSparse(field, lid);
    else
      sim->pm()->DeallocateSparse(field, lid);
  });
}

int pb2h_sim_history(pb2h_sim *sim, double out[8]) {
  return Guard([&] {
    // one column per octant, summed over the MeshData batches of this rank (outputs/history.cpp:
    // 47-200), then over the ranks
    std::vector<Real> v(8, 0.0);
    for (int p = 0; p < sim->pm()->DefaultNumPartitions(); ++p) {
      const auto part = burgers_package::MassHistory(sim->pm()->mesh_data.GetOrAdd("base", p).get());
      for (int o = 0; o < 8; ++o) v[o] += part[o];
    }
    sim->pm()->ReduceHistory(v);
    for (int o = 0; o < 8; ++o) out[o] = v[o];
  });
}

double pb2h_sim_zone_cycles_per_second(pb2h_sim *sim) { return sim->driver->ZoneCyclesPerSecond(); }

int pb2h_sim_execute(pb2h_sim *sim) {
  return Guard([&] {
    PARTHENON_REQUIRE(sim->driver != nullptr, "this application has no time loop");
    PARTHENON_REQUIRE(sim->driver->Execute() != DriverStatus::failed, "driver failed");
  });
}

} // extern "C"
