// burgers_package.cpp — Parthenon-VIBE package: field registration and the tasks that hand
// a whole MeshData batch to the sm_100a kernels (see burgers_package.hpp).
#include "burgers_package.hpp"

#include <algorithm>
#include <cfloat>
#include <string>

namespace burgers_package {

namespace {
struct KernelConfig {
  int recon; // PB2_RECON_*
  int math;  // PB2_MATH_*
};

pb2_burgers_args MakeArgs(MeshData<Real> *md, Variable &u) {
  auto pkg = md->GetMeshPointer()->packages.Get("burgers_package");
  const auto &cfg = pkg->Param<KernelConfig>("kernel_config");
  pb2_burgers_args a{};
  a.geom = md->Geometry(u);
  a.recon = cfg.recon;
  a.math = cfg.math;
  a.u = u.data();
  return a;
}

// one device scalar per MeshData batch that the update kernel min-reduces into
Real *DtScratch(MeshData<Real> *md) { return md->DtCell(); }

void ResetDt(MeshData<Real> *md) {
  static const Real huge = DBL_MAX;
  PB2_CHECK(pb2_memcpy_h2d(DtScratch(md), &huge, sizeof(Real), md->stream()));
}
Real ReadDt(MeshData<Real> *md) {
  Real v = 0;
  PB2_CHECK(pb2_memcpy_d2h(&v, DtScratch(md), sizeof(Real), md->stream()));
  PB2_CHECK(pb2_stream_sync(md->stream()));
  return v;
}
} // namespace

std::shared_ptr<StateDescriptor> Initialize(ParameterInput *pin) {
  auto pkg = std::make_shared<StateDescriptor>("burgers_package");

  const Real cfl = pin->GetOrAddReal("burgers", "cfl", 0.8);
  pkg->AddParam("cfl", cfl);

  KernelConfig cfg{};
  const std::string recon = pin->GetOrAddString("burgers", "recon", "weno5");
  const int nghost = pin->GetInteger("parthenon/mesh", "nghost");
  if (recon == "weno5") {
    cfg.recon = PB2_RECON_WENO5;
    PARTHENON_REQUIRE_THROWS(nghost >= 4, "weno5 reconstruction requires 4 or more ghost "
                                          "cells.  Set <parthenon/mesh>/nghost = 4");
  } else if (recon == "linear") {
    cfg.recon = PB2_RECON_LINEAR;
    if (nghost > 2)
      PARTHENON_WARN("Using more ghost cells than required.  Consider setting "
                     "<parthenon/mesh>/nghost = 2");
  } else {
    PARTHENON_THROW(recon + " is an invalid option for <burgers>/recon.  Valid options are "
                            "weno5 and linear.");
  }
  // arithmetic mode of the stencil kernels: "strict" reproduces the reference's CPU build
  // bit for bit (no FMA contraction); "fast" allows contraction (<= 1e-12 relative)
  const std::string math = pin->GetOrAddString("pb2", "math", "fast", {"fast", "strict"});
  cfg.math = math == "strict" ? PB2_MATH_STRICT : PB2_MATH_FAST;
  pkg->AddParam("kernel_config", cfg);
  pkg->AddParam("fused_stage", pin->GetOrAddBoolean("pb2", "fused_stage", true));
  // fast fused stage on a uniform mesh: the sweeps read same-device neighbours directly
  // (pb2_burgers_args::nbr_direct) and the same-device ghost exchange leaves the cycle; ghost
  // cells are refreshed when something else reads them
  pkg->AddParam("lazy_ghosts", pin->GetOrAddBoolean("pb2", "lazy_ghosts", true));
  // multi-GPU: the blocks that feed other GPUs are listed first in a single launch and reported
  // from inside the last sweep (false: two launches per sweep, boundary blocks then the rest)
  pkg->AddParam("device_progress", pin->GetOrAddBoolean("pb2", "device_progress", true));

  const int num_scalars = pin->GetOrAddInteger("burgers", "num_scalars", 1);
  pkg->AddParam("num_scalars", num_scalars);
  PARTHENON_REQUIRE_THROWS(num_scalars > 0, "Burgers benchmark requires num_scalars >= 1");

  // always a three dimensional velocity + the passive scalars (burgers_package.cpp:75-80)
  const std::vector<int> vec_components(1, 3 + num_scalars);
  Metadata m({Metadata::Cell, Metadata::Independent, Metadata::Intensive, Metadata::Conserved,
              Metadata::FillGhost, Metadata::WithFluxes},
             vec_components);
  pkg->AddField("U", m);
  // The reference also registers six reconstruction scratch fields Ulx..Urz (:82-89).
  // They exist here too so packs by name resolve, but storage is lazy and the kernels keep
  // left/right states in registers, so they never occupy HBM.
  m = Metadata({Metadata::Cell, Metadata::Derived, Metadata::OneCopy}, vec_components);
  for (const char *n : {"Ulx", "Urx", "Uly", "Ury", "Ulz", "Urz"}) pkg->AddField(n, m);
  m = Metadata({Metadata::Cell, Metadata::Derived, Metadata::OneCopy});
  pkg->AddField("derived", m);

  Real mesh_vol = 1;
  for (int d = 1; d <= 3; ++d) {
    const std::string x = "x" + std::to_string(d);
    mesh_vol *= pin->GetReal("parthenon/mesh", x + "max") - pin->GetReal("parthenon/mesh", x + "min");
  }
  pkg->AddParam("mesh_volume", mesh_vol);

  // history: the eight octant masses, labels as in the reference (:110-135)
  HistoryOutputVec hv;
  hv.hst_op = UserHistoryOperation::sum;
  hv.hst_fun = MassHistory;
  for (int o = 0; o < 8; ++o) hv.labels.push_back("MS Mass " + std::to_string(o));
  pkg->AddParam(hist_vec_param_key, std::vector<HistoryOutputVec>{hv});

  pkg->EstimateTimestepMesh = EstimateTimestepMesh;
  pkg->FillDerivedMesh = CalculateDerived;
  return pkg;
}

TaskStatus CalculateFluxes(MeshData<Real> *md) {
  EnsureLocalGhosts(md);
  Variable &u = md->Get("U");
  pb2_burgers_args a = MakeArgs(md, u);
  for (int d = 0; d < a.geom.ndim; ++d) a.flux[d] = u.flux(d + 1);
  PB2_CHECK(pb2_burgers_calculate_fluxes(&a, md->stream()));
  return TaskStatus::complete;
}

void CalculateDerived(MeshData<Real> *md) {
  Variable &u = md->Get("U");
  pb2_pack_geom g = md->Geometry(u);
  PB2_CHECK(pb2_burgers_derived_dt(&g, u.data(), md->Get("derived").data(), nullptr, md->stream()));
}

Real EstimateTimestepMesh(MeshData<Real> *md) {
  auto pkg = md->GetMeshPointer()->packages.Get("burgers_package");
  Variable &u = md->Get("U");
  pb2_pack_geom g = md->Geometry(u);
  ResetDt(md);
  PB2_CHECK(pb2_burgers_derived_dt(&g, u.data(), nullptr, DtScratch(md), md->stream()));
  return pkg->Param<Real>("cfl") * ReadDt(md);
}

std::vector<Real> MassHistory(MeshData<Real> *md) {
  Mesh *pm = md->GetMeshPointer();
  Variable &u = md->Get("U");
  pb2_pack_geom g = md->Geometry(u);
  std::vector<Real> out(8, 0.0);
  PB2_CHECK(pb2_burgers_history(&g, u.data(), md->DeviceXmin(), pm->mesh_size.xmin_.data(),
                                pm->mesh_size.xmax_.data(), out.data(), md->stream()));
  return out;
}

TaskStatus FusedStage(MeshData<Real> *mc0, MeshData<Real> *mbase, MeshData<Real> *mc1,
                      Real beta, Real dt, bool last_stage) {
  auto pkg = mc0->GetMeshPointer()->packages.Get("burgers_package");
  Variable &u = mc0->Get("U");
  pb2_burgers_args a = MakeArgs(mc0, u);
  a.base = mbase->Get("U").data();
  a.out = mc1->Get("U").data();
  // only the bit-exact dataflow stores fluxes; the fast sweeps keep them in registers, so
  // the three flux arrays are never allocated (8.6 GB at 256^3, 69 GB at 512^3)
  // ... and meshes with fine-coarse faces, whose face fluxes must exist to be corrected
  Mesh *pm = mc0->GetMeshPointer();
  const bool flxcor = pm->HasFineCoarseFaces();
  if (a.math == PB2_MATH_STRICT || flxcor)
    for (int d = 0; d < a.geom.ndim; ++d) a.flux[d] = u.flux(d + 1);
  a.derived = mc1->Get("derived").data();
  a.beta = beta;
  a.dt = dt;
  if (last_stage) {
    ResetDt(mc1);
    a.dt_min = DtScratch(mc1);
  }
  // Blocks whose results travel to another GPU go first; their halo is then packed and
  // shipped on the communication stream while the interior blocks are advanced
  // (the local / nonlocal overlap of burgers_driver.cpp:106-119, without host polling).
  BvarsCache &bc = GetBvarsCache(mc1);
  const bool split = a.math == PB2_MATH_FAST && !pm->multilevel && bc.n_boundary > 0 &&
                     bc.n_interior > 0 && bc.plan->send_elements > 0;
  // Ghost cells across same-device faces: the fast sweeps read the neighbour's interior instead
  // (pb2_burgers_args::nbr_direct), so neither the input's local ghosts need to be current nor
  // do the output's have to be filled before the next stage — the local exchange of mc1 is
  // deferred until something else reads its ghost cells (EnsureLocalGhosts).
  BvarsCache &bc0 = GetBvarsCache(mc0);
  const bool direct = a.math == PB2_MATH_FAST && !flxcor && !pm->multilevel &&
                      pkg->Param<bool>("lazy_ghosts") && bc0.uniform_halo && bc.uniform_halo &&
                      bc0.vars.size() == 1 && bc0.vars[0] == &u && bc.vars.size() == 1 &&
                      bc.vars[0] == &mc1->Get("U");
  if (direct) {
    a.nbr_direct = bc0.halo_nbr.get<int32_t>();
    bc.defer_local = true;
    if (!(split && a.geom.ndim >= 2 && pkg->Param<bool>("device_progress"))) EnsureRemoteGhosts(mc0);
  } else {
    EnsureLocalGhosts(mc0);
    if (mbase != mc0) EnsureLocalGhosts(mbase);
  }
  if (flxcor) {
    // CalculateFluxes -> flux correction -> FluxDivergence + update (burgers_driver.cpp:92-104).
    // Only blocks with a face neighbour on another level exchange flux corrections; with
    // pb2/math = fast the others take the flux-free sweeps (same equations, <= 1e-12), with
    // strict every block keeps the reference's dataflow.
    BvarsCache &fc = GetBvarsCache(mc0);
    const bool mixed = a.math == PB2_MATH_FAST && fc.n_plain > 0;
    if (mixed) {
      a.block_ids = fc.ids_flxcor.get<int32_t>();
      a.num_block_ids = fc.n_flxcor;
    }
    PB2_CHECK(pb2_burgers_calculate_fluxes(&a, mc0->stream()));
    FluxCorrection(mc0);
    PB2_CHECK(pb2_burgers_update(&a, mc0->stream()));
    if (mixed) {
      a.block_ids = fc.ids_plain.get<int32_t>();
      a.num_block_ids = fc.n_plain;
      PB2_CHECK(pb2_burgers_stage(&a, mc0->stream()));
      a.block_ids = nullptr;
    }
    // the reference's Average/UpdateIndependentData run over the full extents; ghosts that
    // the exchange below does not refresh (fine ghosts facing a coarser block: the stage list
    // has no ProlongateBounds) carry beta*mc0 + (1-beta)*base into the next stencil
    if (fc.n_stale_ghosts > 0)
      PB2_CHECK(pb2_weighted_sum_ghosts_blocks(&a.geom, a.u, a.base, beta, 1.0 - beta, a.out,
                                               fc.ids_stale_ghosts.get<int32_t>(),
                                               fc.n_stale_ghosts, mc0->stream()));
  } else if (split && a.geom.ndim >= 2 && pkg->Param<bool>("device_progress")) {
    // ONE launch per sweep over [boundary blocks, interior blocks]; the last sweep reports the
    // boundary part from inside (pb2_burgers_args::progress) and the communication stream waits
    // for that count on the device: no second set of launches with its own tail
    const int32_t *ordered = bc.ids_ordered.get<int32_t>();
    const int nall = bc.n_boundary + bc.n_interior;
    if (direct && bc0.remote_pending) {
      // the input's inter-GPU ghosts are still being unpacked on the communication stream: the
      // first sweep starts on the blocks that have no remote face (they read neighbours'
      // interiors only), the compute stream waits for the unpack, then the other blocks follow
      a.sweeps = 1;
      a.block_ids = bc.ids_interior.get<int32_t>();
      a.num_block_ids = bc.n_interior;
      PB2_CHECK(pb2_burgers_stage(&a, mc0->stream()));
      EnsureRemoteGhosts(mc0);
      a.block_ids = bc.ids_boundary.get<int32_t>();
      a.num_block_ids = bc.n_boundary;
      PB2_CHECK(pb2_burgers_stage(&a, mc0->stream()));
      a.sweeps = 6;
    }
    a.block_ids = ordered;
    a.num_block_ids = nall;
    a.progress = bc.progress.get<int32_t>();
    a.progress_blocks = bc.n_boundary;
    PB2_CHECK(pb2_memset(a.progress, 0, sizeof(int32_t), mc0->stream()));
    PB2_CHECK(pb2_event_record(bc.early_ready, mc0->stream())); // the counter is reset
    bc.early_valid = true;
    bc.progress_target = pb2_burgers_progress_target(&a.geom, bc.n_boundary);
    PB2_CHECK(pb2_burgers_stage(&a, mc0->stream()));
    // the unpack of this stage's inter-GPU halo may run behind the exchange on the communication
    // stream: whoever reads mc1's ghost cells next waits for it (EnsureRemoteGhosts)
    if (direct) bc.defer_remote = true;
  } else if (split) {
    a.block_ids = bc.ids_boundary.get<int32_t>();
    a.num_block_ids = bc.n_boundary;
    PB2_CHECK(pb2_burgers_stage(&a, mc0->stream()));
    PB2_CHECK(pb2_event_record(bc.early_ready, mc0->stream()));
    bc.early_valid = true;
    a.block_ids = bc.ids_interior.get<int32_t>();
    a.num_block_ids = bc.n_interior;
    PB2_CHECK(pb2_burgers_stage(&a, mc0->stream()));
  } else {
    PB2_CHECK(pb2_burgers_stage(&a, mc0->stream()));
  }
  return TaskStatus::complete;
}

// EstimateTimestep task of the fused path: the stage kernel already reduced min dt
TaskStatus CollectFusedTimestep(MeshData<Real> *mc1) {
  auto pkg = mc1->GetMeshPointer()->packages.Get("burgers_package");
  const Real dt_min = pkg->Param<Real>("cfl") * ReadDt(mc1);
  for (auto &pmb : mc1->GetBlockList()) pmb->SetAllowedDt(std::min(dt_min, pmb->NewDt()));
  return TaskStatus::complete;
}

} // namespace burgers_package
