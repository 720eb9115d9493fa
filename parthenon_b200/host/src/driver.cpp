// driver.cpp — evolution loop and integrator tables (see pb2/driver.hpp).
#include "pb2/driver.hpp"

#include "pb2/bvals.hpp"

#include <algorithm>
#include <cstdio>
#include <iostream>

namespace parthenon {

StagedIntegrator::StagedIntegrator(ParameterInput *pin) {
  name_ = pin->GetOrAddString("parthenon/time", "integrator", "rk2");
  // two-register low-storage coefficient tables, low_storage_integrator.cpp:30-180
  auto set = [&](int n, std::vector<Real> d, std::vector<Real> b, std::vector<Real> g0,
                 std::vector<Real> g1, std::vector<Real> cc) {
    nstages = n;
    delta = std::move(d);
    beta = std::move(b);
    gam0 = std::move(g0);
    gam1 = std::move(g1);
    c = std::move(cc);
  };
  if (name_ == "rk1") {
    nbuffers = 1;
    set(1, {1.0}, {1.0}, {0.0}, {1.0}, {0.0});
  } else if (name_ == "rk2") {
    nbuffers = 2;
    set(2, {1.0, 0.0}, {1.0, 0.5}, {0.0, 0.5}, {1.0, 0.5}, {0.0, 1.0});
  } else if (name_ == "vl2") {
    nbuffers = 2;
    set(2, {1.0, 0.0}, {0.5, 1.0}, {0.0, 0.0}, {1.0, 1.0}, {0.0, 0.5});
  } else if (name_ == "rk3") {
    nbuffers = 2;
    set(3, {1.0, 0.0, 0.0}, {1.0, 0.25, 2.0 / 3.0}, {0.0, 0.25, 2.0 / 3.0},
        {1.0, 0.75, 1.0 / 3.0}, {0.0, 1.0, 0.5});
  } else {
    PARTHENON_THROW("Invalid selection for the time integrator: " + name_);
  }
  // staged_integrator.cpp:23-30: "base", "1", ..., back to "base"
  stage_name.resize(nstages + 1);
  stage_name[0] = "base";
  for (int i = 1; i < nstages; ++i) stage_name[i] = std::to_string(i);
  stage_name[nstages] = stage_name[0];
}

EvolutionDriver::EvolutionDriver(ParameterInput *pin, ApplicationInput *app_in, Mesh *pm)
    : Driver(pin, app_in, pm) {
  // driver.hpp:64-95
  const Real start_time = pin->GetOrAddReal("parthenon/time", "start_time", 0.0);
  tm.start_time = tm.time = start_time;
  tm.tlim = pin->GetOrAddReal("parthenon/time", "tlim", std::numeric_limits<Real>::infinity());
  tm.dt = pin->GetOrAddReal("parthenon/time", "dt", std::numeric_limits<Real>::max());
  tm.ncycle = pin->GetOrAddInteger("parthenon/time", "ncycle", 0);
  tm.nlim = pin->GetOrAddInteger("parthenon/time", "nlim", -1);
  tm.ncycle_out = pin->GetOrAddInteger("parthenon/time", "ncycle_out", 1);
  tm.ncycle_out_mesh = pin->GetOrAddInteger("parthenon/time", "ncycle_out_mesh", 0);
  dt_user = pin->GetOrAddReal("parthenon/time", "dt_user", dt_user);
  dt_force = pin->GetOrAddReal("parthenon/time", "dt_force", -1.0);
  dt_init = pin->GetOrAddReal("parthenon/time", "dt_init", dt_init);
  dt_init_force = pin->GetOrAddBoolean("parthenon/time", "dt_init_force", false);
  dt_factor = pin->GetOrAddReal("parthenon/time", "dt_factor", 2.0);
  dt_floor = pin->GetOrAddReal("parthenon/time", "dt_floor", dt_floor);
  dt_ceil = pin->GetOrAddReal("parthenon/time", "dt_ceil", dt_ceil);
  dt_min = pin->GetOrAddReal("parthenon/time", "dt_min", dt_min);
  dt_max = pin->GetOrAddReal("parthenon/time", "dt_max", dt_max);
  dt_min_count_max = pin->GetOrAddInteger("parthenon/time", "dt_min_cycle_limit", 10);
  dt_max_count_max = pin->GetOrAddInteger("parthenon/time", "dt_max_cycle_limit", 1);
  perf_cycle_offset = pin->GetOrAddInteger("parthenon/time", "perf_cycle_offset", 0);
  const std::string problem_id = pin->GetOrAddString("parthenon/job", "problem_id", "parthenon");
  for (auto &b : pin->BlockNames()) {
    if (b.compare(0, 16, "parthenon/output") != 0) continue;
    const std::string type = pin->GetOrAddString(b, "file_type", "none");
    const Real dt = pin->GetOrAddReal(b, "dt", -1.0);
    if (type != "hst" || dt <= 0.0) continue; // only history files are produced here
    HistoryOutput h;
    h.filename = problem_id + ".out" + b.substr(16) + ".hst";
    h.data_format = " " + pin->GetOrAddString(b, "data_format", "%12.5e");
    h.dt = dt;
    h.next_time = tm.start_time;
    hst_outputs_.push_back(h);
  }
}

void EvolutionDriver::InitializeBlockTimeSteps() {
  for (int p = 0; p < pmesh->DefaultNumPartitions(); ++p)
    Update::EstimateTimestep(pmesh->mesh_data.GetOrAdd("base", p).get());
}

void EvolutionDriver::SetGlobalTimeStep() {
  // driver.cpp:210-270
  if (dt_force > 0.0) {
    tm.dt = dt_force;
  } else if (tm.ncycle == 0 && dt_init_force && dt_init > 0.0) {
    tm.dt = dt_init;
  } else {
    if (tm.dt < 0.1 * std::numeric_limits<Real>::max()) tm.dt *= dt_factor;
    if (tm.ncycle == 0) tm.dt = std::min(tm.dt, dt_init);
    for (auto const &pmb : pmesh->block_list) {
      tm.dt = std::min(tm.dt, pmb->NewDt());
      pmb->SetAllowedDt(std::numeric_limits<Real>::max());
    }
    tm.dt = std::min(tm.dt, dt_user);
    tm.dt = std::max(dt_floor, std::min(tm.dt, dt_ceil));
    if (pmesh->nranks > 1) { // MPI_Allreduce(MIN), driver.cpp:237
      PARTHENON_REQUIRE(pmesh->comm != nullptr, "multi-rank mesh without a communicator");
      Real *d = pmesh->ScratchReal();
      PB2_CHECK(pb2_memcpy_h2d(d, &tm.dt, sizeof(Real), pmesh->stream));
      PB2_CHECK(pb2_comm_allreduce_min(pmesh->comm, d, pmesh->stream));
      PB2_CHECK(pb2_memcpy_d2h(&tm.dt, d, sizeof(Real), pmesh->stream));
      PB2_CHECK(pb2_stream_sync(pmesh->stream));
    }
  }
  // Check that we have not gone off the rails (driver.cpp:243-263)
  if (tm.dt <= dt_min) {
    PARTHENON_REQUIRE(++dt_min_count < dt_min_count_max,
                      "Timesetep has fallen bellow minimum (parthenon/time/dt_min=" +
                          std::to_string(dt_min) + ") for more than " +
                          std::to_string(dt_min_count_max) + " steps");
  } else {
    dt_min_count = 0;
  }
  if (tm.dt >= dt_max) {
    PARTHENON_REQUIRE(++dt_max_count < dt_max_count_max,
                      "Timesetep has risen above maximum (parthenon/time/dt_max=" +
                          std::to_string(dt_max) + ") for more than " +
                          std::to_string(dt_max_count_max) + " steps");
  } else {
    dt_max_count = 0;
  }
  // after the bounds check so that a step that lands epsilon before tlim does not fail
  if (tm.time < tm.tlim && (tm.tlim - tm.time) < tm.dt) tm.dt = tm.tlim - tm.time;
}

void EvolutionDriver::MakeHistoryOutput(bool force) {
  if (hst_outputs_.empty()) return;
  for (auto &h : hst_outputs_) {
    if (!(force || tm.time >= h.next_time || tm.tlim <= tm.time)) continue;
    // outputs/history.cpp:47-200: one column per enrolled reduction, summed over batches
    std::vector<Real> vals;
    std::vector<std::string> labels;
    for (const auto &pkg : pmesh->packages.AllPackages()) {
      const Params &params = pkg.second->AllParams();
      if (params.hasKey(hist_param_key)) {
        for (const auto &hv : params.Get<HstVar_list>(hist_param_key)) {
          Real acc = 0.0;
          for (int p = 0; p < pmesh->DefaultNumPartitions(); ++p) {
            const Real r = hv.hst_fun(pmesh->mesh_data.GetOrAdd("base", p).get());
            acc = p == 0 ? r
                         : (hv.hst_op == UserHistoryOperation::sum
                                ? acc + r
                                : (hv.hst_op == UserHistoryOperation::max ? std::max(acc, r)
                                                                          : std::min(acc, r)));
          }
          vals.push_back(acc);
          labels.push_back(hv.label);
        }
      }
      if (params.hasKey(hist_vec_param_key)) {
        for (const auto &hv : params.Get<std::vector<HistoryOutputVec>>(hist_vec_param_key)) {
          std::vector<Real> acc;
          for (int p = 0; p < pmesh->DefaultNumPartitions(); ++p) {
            const auto r = hv.hst_fun(pmesh->mesh_data.GetOrAdd("base", p).get());
            if (p == 0) {
              acc = r;
            } else {
              for (size_t i = 0; i < r.size(); ++i)
                acc[i] = hv.hst_op == UserHistoryOperation::sum
                             ? acc[i] + r[i]
                             : (hv.hst_op == UserHistoryOperation::max ? std::max(acc[i], r[i])
                                                                       : std::min(acc[i], r[i]));
            }
          }
          for (size_t i = 0; i < acc.size(); ++i) {
            vals.push_back(acc[i]);
            labels.push_back(hv.labels[i]);
          }
        }
      }
    }
    pmesh->ReduceHistory(vals); // sum over ranks (MPI_Reduce in outputs/history.cpp)
    if (pmesh->my_rank == 0) {
      std::FILE *f = std::fopen(h.filename.c_str(), h.header_written ? "a" : "w");
      PARTHENON_REQUIRE(f != nullptr, "cannot open history file " + h.filename);
      if (!h.header_written) {
        int col = 0;
        std::fprintf(f, "#  History data\n#");
        std::fprintf(f, " [%d]=time    ", ++col);
        std::fprintf(f, " [%d]=dt      ", ++col);
        std::fprintf(f, " [%d]=cycle   ", ++col);
        std::fprintf(f, " [%d]=nbtotal ", ++col);
        for (auto &l : labels) std::fprintf(f, " [%d]=%s", ++col, l.c_str());
        std::fprintf(f, "\n");
        h.header_written = true;
      }
      std::fprintf(f, h.data_format.c_str(), tm.time);
      std::fprintf(f, h.data_format.c_str(), tm.dt);
      std::fprintf(f, " %12d", tm.ncycle);
      std::fprintf(f, " %12d", pmesh->nbtotal);
      for (Real v : vals) std::fprintf(f, h.data_format.c_str(), v);
      std::fprintf(f, "\n");
      std::fclose(f);
    }
    if (!force) h.next_time += h.dt;
  }
}

void EvolutionDriver::OutputCycleDiagnostics() {
  if (quiet || pmesh->my_rank != 0 || tm.ncycle_out == 0 || tm.ncycle % tm.ncycle_out != 0) return;
  std::printf("cycle=%d time=%.16e dt=%.16e\n", tm.ncycle, tm.time, tm.dt);
}

void EvolutionDriver::PreExecute() {
  InitializeBlockTimeSteps();
  SetGlobalTimeStep();
  if (app_input && app_input->UserWorkBeforeLoop) app_input->UserWorkBeforeLoop(pmesh, pinput, tm);
  for (auto &pkg : pmesh->packages.AllPackages())
    if (pkg.second->UserWorkBeforeLoopMesh) pkg.second->UserWorkBeforeLoopMesh(pmesh, pinput, tm);
  MakeHistoryOutput(false);
  pmesh->mbcnt = 0;
  timer_main_ = std::chrono::steady_clock::now();
}

// driver.cpp:99-150 in the two halves either side of the PostStepUserWorkInLoop hook
TaskListStatus EvolutionDriver::StepAndAdvanceTime() {
  OutputCycleDiagnostics();
  const TaskListStatus status = Step();
  if (status != TaskListStatus::complete) return status;
  tm.ncycle++;
  tm.time += tm.dt;
  pmesh->mbcnt += pmesh->nbtotal;
  return status;
}

void EvolutionDriver::AdaptMeshAndSetTimeStep() {
  pmesh->LoadBalancingAndAdaptiveMeshRefinement(pinput, app_input);
  if (pmesh->modified) InitializeBlockTimeSteps();
  SetGlobalTimeStep();
}

TaskListStatus EvolutionDriver::DoCycle() {
  const TaskListStatus status = StepAndAdvanceTime();
  if (status != TaskListStatus::complete) return status;
  AdaptMeshAndSetTimeStep();
  if (tm.KeepGoing()) MakeHistoryOutput(false);
  if (tm.ncycle == perf_cycle_offset) {
    PB2_CHECK(pb2_stream_sync(pmesh->stream));
    pmesh->mbcnt = 0;
    timer_main_ = std::chrono::steady_clock::now();
  }
  return status;
}

double EvolutionDriver::ZoneCyclesPerSecond() const {
  const double wall =
      std::chrono::duration<double>(std::chrono::steady_clock::now() - timer_main_).count();
  return static_cast<double>(pmesh->mbcnt) * pmesh->GetNumberOfMeshBlockCells() / wall;
}

DriverStatus EvolutionDriver::Execute() {
  PreExecute();
  while (tm.KeepGoing()) {
    if (DoCycle() != TaskListStatus::complete) {
      std::cerr << "Step failed to complete all tasks." << std::endl;
      return DriverStatus::failed;
    }
  }
  EnsureLocalGhosts(pmesh); // deferred same-device ghost copies, before anyone looks
  PB2_CHECK(pb2_stream_sync(pmesh->stream));
  const double zcps = ZoneCyclesPerSecond();
  if (app_input && app_input->UserWorkAfterLoop) app_input->UserWorkAfterLoop(pmesh, pinput, tm);
  MakeHistoryOutput(true);
  if (!quiet && pmesh->my_rank == 0) {
    // driver.cpp:296-325 PostExecute summary
    std::printf("\ncycle=%d time=%.16e dt=%.16e\n", tm.ncycle, tm.time, tm.dt);
    std::printf("zone-cycles = %lld\n",
                static_cast<long long>(pmesh->mbcnt) * pmesh->GetNumberOfMeshBlockCells());
    std::printf("zone-cycles/wallsecond = %.6e\n", zcps);
  }
  return tm.KeepGoing() ? DriverStatus::timeout : DriverStatus::complete;
}

TaskListStatus MultiStageDriver::Step() {
  integrator->dt = tm.dt;
  TaskListStatus status = TaskListStatus::complete;
  for (int stage = 1; stage <= integrator->nstages; ++stage) {
    // driver.hpp:142 ConstructAndExecuteTaskLists: the graph is rebuilt per stage
    TaskCollection tc = MakeTaskCollection(pmesh->block_list, stage);
    status = tc.Execute();
    if (status != TaskListStatus::complete) break;
  }
  return status;
}

} // namespace parthenon
