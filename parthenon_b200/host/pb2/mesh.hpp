// mesh.hpp — block-structured mesh topology (host only).
//
// What the ghost-zone hot path needs from the reference's src/mesh + src/mesh/forest:
// Morton-ordered leaf list (logical_location.hpp:125-131, forest.cpp:145), contiguous gid
// ranges per rank (AssignBlocks, mesh-amr_loadbalance.cpp:362-386), the neighbour list of
// every block with same-level offsets (tree.cpp:139-226, mesh-gmg.cpp:48-101), index shapes
// (meshblock.cpp:204-216) and uniform Cartesian geometry (uniform_cartesian.hpp:27-190).
// One rank <-> one B200; the blocks of a rank are one (or `pack_size`-sized) MeshData batch.
// Supported: single-tree forests (2^n root blocks in every active direction), periodic
// boundaries, uniform or statically refined leaves with 2:1 nesting.
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "device.hpp"
#include "forest.hpp"
#include "parameter_input.hpp"
#include "state.hpp"
#include "types.hpp"

struct pb2_comm;

namespace parthenon {

namespace Globals {
extern int my_rank, nranks, nghost;
} // namespace Globals

// bvals/neighbor_block.hpp:48: what a block knows about one neighbour
struct NeighborBlock {
  int gid = -1, rank = 0, lid = -1; // lid: index on the owning rank
  LogicalLocation loc;              // wrapped location (as stored in the tree)
  LogicalLocation origin_loc;       // location in the frame of this block (may lie outside)
  int offsets[3] = {0, 0, 0};       // same-level offsets ox1, ox2, ox3
  // forests: how this block's logical coordinates read in the neighbour's tree
  // (neighbor_block.hpp:68); `transformed` = it is not the identity, i.e. what the neighbour
  // sends arrives in ITS orientation and is unpacked through the transformation
  forest::LogicalCoordinateTransformation lcoord_trans;
  bool transformed = false;
  // index of the offset in the 27-cube, the channel key element "loc idx"
  // (cell_center_offsets.hpp:98)
  int OffsetIndex() const { return (offsets[0] + 1) + 3 * (offsets[1] + 1) + 9 * (offsets[2] + 1); }
};

// coordinates/uniform_cartesian.hpp:27-190
class UniformCartesian {
 public:
  UniformCartesian() = default;
  UniformCartesian(const RegionSize &rs, int nghost) {
    for (int d = 0; d < 3; ++d) {
      dx_[d] = (rs.xmax_[d] - rs.xmin_[d]) / rs.nx_[d];
      istart_[d] = rs.symmetry_[d] ? 0 : nghost;
      xmin_[d] = rs.xmin_[d] - istart_[d] * dx_[d];
    }
  }
  UniformCartesian(const UniformCartesian &src, int coarsen) {
    istart_ = src.istart_;
    dx_ = src.dx_;
    xmin_ = src.xmin_;
    for (int d = 0; d < 3; ++d) xmin_[d] += istart_[d] * dx_[d] * (1 - coarsen);
    dx_[0] *= coarsen;
    dx_[1] *= (istart_[1] > 0 ? coarsen : 1);
    dx_[2] *= (istart_[2] > 0 ? coarsen : 1);
  }
  template <int dir>
  Real Dxc() const { return dx_[dir - 1]; }
  Real DxcFA(int dir) const { return dx_[dir - 1]; }
  template <int dir>
  Real Xc(int idx) const { return xmin_[dir - 1] + (idx + 0.5) * dx_[dir - 1]; }
  template <int dir>
  Real Xf(int idx) const { return xmin_[dir - 1] + idx * dx_[dir - 1]; }
  Real CellVolume() const { return dx_[0] * dx_[1] * dx_[2]; }
  Real FaceArea(int dir) const {
    return dir == 1 ? dx_[1] * dx_[2] : (dir == 2 ? dx_[0] * dx_[2] : dx_[0] * dx_[1]);
  }
  const std::array<Real, 3> &Dx() const { return dx_; }
  const std::array<Real, 3> &GetXmin() const { return xmin_; }
  const std::array<int, 3> &GetStartIndex() const { return istart_; }

 private:
  std::array<Real, 3> dx_{1, 1, 1}, xmin_{0, 0, 0};
  std::array<int, 3> istart_{0, 0, 0};
};
using Coordinates_t = UniformCartesian;

class Mesh;

class MeshBlock {
 public:
  int gid = 0, lid = 0; // global id (Morton order) and index on this rank
  LogicalLocation loc;
  // meshblock.hpp boundary_flag: the mesh's flag on faces that lie on the mesh boundary,
  // BoundaryFlag::block elsewhere (inner_x1, outer_x1, inner_x2, ...)
  BoundaryFlag boundary_flag[6] = {BoundaryFlag::block, BoundaryFlag::block, BoundaryFlag::block,
                                   BoundaryFlag::block, BoundaryFlag::block, BoundaryFlag::block};
  RegionSize block_size;
  IndexShape cellbounds, c_cellbounds;
  Coordinates_t coords;
  std::vector<NeighborBlock> neighbors;
  Mesh *pmy_mesh = nullptr;
  int partition = 0; // MeshData batch holding this block
  int pack_index = 0; // position inside that batch
  // MeshRefinement::refine_flag_ / deref_count_ (mesh/mesh_refinement.hpp): the tag of the last
  // Refinement::Tag and the number of consecutive "derefine" tags
  int refine_flag = 0, deref_count = 0;

  Real NewDt() const { return new_block_dt_; }
  void SetAllowedDt(Real dt) { new_block_dt_ = dt; }
  int GetNumberOfMeshBlockCells() const {
    return cellbounds.GetTotal(IndexDomain::interior);
  }

 private:
  Real new_block_dt_ = std::numeric_limits<Real>::max();
};
using BlockList_t = std::vector<std::shared_ptr<MeshBlock>>;

// interface/data_collection.hpp:46: named MeshData containers ("base", "1", "dUdt", ...) per
// block partition.  A new container shares OneCopy fields with "base" and gets its own
// arrays for everything else (meshblock_data.Add semantics, burgers_driver.cpp:64-73).
class MeshDataCollection {
 public:
  explicit MeshDataCollection(Mesh *pm) : pmesh_(pm) {}
  std::shared_ptr<MeshData<Real>> &GetOrAdd(const std::string &label, int partition_id);
  std::shared_ptr<MeshData<Real>> &Add(const std::string &label, int partition_id) {
    return GetOrAdd(label, partition_id);
  }
  std::shared_ptr<MeshData<Real>> &Get(const std::string &label = "base", int partition_id = 0) {
    return GetOrAdd(label, partition_id);
  }
  bool Contains(const std::string &label, int partition_id) const {
    return map_.count(label + "_part-" + std::to_string(partition_id)) > 0;
  }
  void PurgeNonBase();
  void Clear() { map_.clear(); }
  std::map<std::string, std::shared_ptr<MeshData<Real>>> &All() { return map_; }

 private:
  Mesh *pmesh_;
  std::map<std::string, std::shared_ptr<MeshData<Real>>> map_;
};

// application hooks (application_input.hpp)
// a user boundary condition (bvals/boundary_conditions.hpp:35 BValFunc) in the batched shape of
// this framework: it fills the ghost slabs of mesh face `face` for the blocks of the batch that
// lie on it (MeshBlock::boundary_flag[face] == BoundaryFlag::user), on the fine arrays or, if
// coarse, on the coarse buffers, by enqueueing its own kernels on md->stream()
using BValFuncMD = std::function<void(std::shared_ptr<MeshData<Real>> &, bool coarse)>;

struct ApplicationInput {
  // application_input.hpp:79-83: a deck selects it with <parthenon/mesh> ix1_bc = name
  void RegisterBoundaryCondition(int face, const std::string &name, BValFuncMD condition) {
    boundary_conditions_[face][name] = std::move(condition);
  }
  void RegisterBoundaryCondition(int face, BValFuncMD condition) {
    RegisterBoundaryCondition(face, "user", std::move(condition));
  }
  std::map<std::string, BValFuncMD> boundary_conditions_[6];
  std::function<Packages_t(std::unique_ptr<ParameterInput> &)> ProcessPackages = nullptr;
  // fills the interior of every block of a MeshData batch on the device
  std::function<void(MeshData<Real> *, ParameterInput *)> MeshProblemGenerator = nullptr;
  std::function<void(Mesh *, ParameterInput *, SimTime &)> UserWorkBeforeLoop = nullptr;
  std::function<void(Mesh *, ParameterInput *, SimTime &)> UserWorkAfterLoop = nullptr;
};

class Mesh {
 public:
  // leaves: explicit (level, lx1, lx2, lx3) list, or empty => root grid + the deck's
  // <parthenon/static_refinementN> regions
  Mesh(ParameterInput *pin, ApplicationInput *app_in, Packages_t &packages, int rank = 0,
       int nranks = 1, const std::vector<LogicalLocation> &leaves = {});
  // a 2-D forest of trees given face by face (mesh.cpp:189-218): static meshes of cell-centred
  // fields; trees may meet in any orientation
  Mesh(ParameterInput *pin, ApplicationInput *app_in, Packages_t &packages,
       const forest::ForestDefinition &forest_def, int rank = 0, int nranks = 1);
  std::shared_ptr<forest::Forest> forest; // null on hyper-rectangular meshes
  ~Mesh();

  int ndim = 3;
  RegionSize mesh_size, base_block_size;
  BoundaryFlag mesh_bcs[6];
  BValFuncMD user_bcs[6]; // set where mesh_bcs[f] == BoundaryFlag::user
  int root_level = 0, current_level = 0;
  int nrbx[3] = {1, 1, 1};
  bool multilevel = false, adaptive = false;
  int nbtotal = 0;
  int my_rank = 0, nranks = 1;
  int64_t mbcnt = 0; // block-cycles since the timer reset (driver.cpp:124)
  Packages_t packages;
  std::vector<FieldEntry> resolved_fields; // all packages, registration order

  std::vector<LogicalLocation> loclist; // every leaf, Morton order: index = gid
  std::vector<int> ranklist, nslist, nblist;
  BlockList_t block_list; // this rank's blocks, gid order
  MeshDataCollection mesh_data{this};
  void *stream = nullptr;      // compute stream of this rank (pb2_stream_t); NULL = default
  void *comm_stream = nullptr; // stream the NCCL halo exchange runs on
  Real *ScratchReal(); // 64 device doubles for tiny reductions
  // in-place sum over ranks of a host vector (MPI_Reduce of outputs/history.cpp)
  void ReduceHistory(std::vector<Real> &vals);
  // in-place sum over ranks of a host vector of any length (AMR: the MPI_Allgatherv of
  // refinement flags and counters, with every rank contributing its own gid range)
  void AllReduceSum(std::vector<Real> &vals);
  // true if some block of this rank has a neighbour on another level
  // adaptive mesh refinement (refinement = adaptive)
  int max_level = 63;          // numlevel + root_level - 1 (mesh.cpp:125)
  int derefine_count = 10;     // <parthenon/mesh>/derefine_count (mesh_refinement.cpp:56)
  bool modified = false;
  int nbnew = 0, nbdel = 0;
  // <parthenon/refinementN> blocks (AMRCriteria, amr_criteria/amr_criteria.cpp:25-101): the
  // stock derivative criteria on one component of a field
  struct AMRCriterion {
    int order = 1;     // method = derivative_order_1 | derivative_order_2
    std::string field;
    int comp = 0;      // vector_i (tensor indices are not supported)
    Real refine_criteria = 0.5, derefine_criteria = 0.05;
    int max_level = 1; // logical level: root level added (parthenon_manager.cpp:216-220)
  };
  std::vector<AMRCriterion> amr_criteria;
  // MeshRefinement::SetRefinement (mesh_refinement.cpp:81-118) for block `lid`
  void SetRefinement(int lid, AmrTag flag);
  // Mesh::LoadBalancingAndAdaptiveMeshRefinement (mesh-amr_loadbalance.cpp:324-351): update
  // the tree from the blocks' refine flags and, if it changed, move the data onto the new
  // block list (RedistributeAndRefineMeshBlocks :663-1010).  Sets `modified`.
  void LoadBalancingAndAdaptiveMeshRefinement(ParameterInput *pin, ApplicationInput *app_in);
  // tree update + new block list from the blocks' refine flags, no field data (tests)
  bool RegridTopologyOnly();

  // <parthenon/sparse> (globals.hpp:27-36, parthenon_manager.cpp:122-138)
  struct SparseConfig {
    bool enabled = true;
    Real allocation_threshold = 1.0e-12;
    Real deallocation_threshold = 1.0e-14;
    int deallocation_count = 5;
  } sparse_config;
  // MeshBlock::AllocateSparse / DeallocateSparse (meshblock.cpp:277-349) for block `lid` of
  // this rank: the field appears zero-filled in (or disappears from) EVERY container
  void AllocateSparse(const std::string &label, int lid);
  void DeallocateSparse(const std::string &label, int lid);
  // number of blocks field slabs are sized for: the block count itself on static meshes; on
  // adaptive meshes 25 % headroom that is only re-drawn when the block count leaves
  // [capacity / 2, capacity], so that successive remeshes allocate identical slab sizes
  int SlabCapacity(int nblocks);
  bool HasFineCoarseFaces() const;
  mutable int fine_coarse_faces_ = -1; // cached answer (static meshes)

  // exchange plans are pure topology: containers ("base", "1", ...) that exchange the same fields
  // over the same blocks share one (key: field signature + partition); a new block list drops them
  std::map<std::string, std::shared_ptr<const void>> plan_cache;

  pb2_comm *comm = nullptr; // NCCL communicator (nranks > 1), owned by the creator
  // test knob: split this rank's blocks over `virtual_ranks` pretend devices so the
  // slab (nonlocal) path runs on one GPU (the reference tests its MPI path the same way
  // with one-rank runs of BuffCommType::both)
  int virtual_ranks = 1;
  // test knob (pb2/table_halo): force the general region-table path for local channels even on
  // uniform meshes, where the descriptor-free pb2_halo_copy_uniform would be used
  bool table_halo = false;
  // pb2/peer_push (default: on for more than one rank): inter-GPU halos of uniform meshes with
  // dense fields are stored straight into the peers' ghost cells (BvarsCache::push_mode); false
  // keeps the slab + NCCL send / recv path.  With virtual ranks it must be asked for explicitly.
  bool peer_push = true;
  // pb2/peer_push_mode: how the halo gets into the peers' memory —
  //   ce      pack into the local send slab, then one device-to-device copy per peer into its
  //           receive slab (copy engines: no SMs, full-size NVLink packets); default
  //   sm      the pack kernel stores into the peers' receive slabs itself
  //   direct  no slabs at all: the copy kernel stores into the peers' ghost cells (cell-centred
  //           fields of uniform meshes; 32-byte x-face rows make poor NVLink packets)
  enum class PeerPush { ce, sm, direct } peer_push_mode = PeerPush::ce;
  // (pb2/unverified_sparse_multilevel: knob of round 1, when sparse fields on statically refined
  // meshes had not been run on a device yet; they are on by default now, the knob is ignored)
  bool unverified_sparse_multilevel = false;
  int VirtualRankOf(int gid) const;

  int GetNumMeshBlocksThisRank() const { return static_cast<int>(block_list.size()); }
  int DefaultPackSize() const; // mesh.hpp:151-153
  int DefaultPackSizeFor(int nblocks) const;
  int DefaultNumPartitions() const;
  int GetNumberOfMeshBlockCells() const {
    return base_block_size.nx_[0] * base_block_size.nx_[1] * base_block_size.nx_[2];
  }
  int64_t GetTotalCells() const { return static_cast<int64_t>(nbtotal) * GetNumberOfMeshBlockCells(); }
  int GetGidRank(int gid) const { return ranklist[gid]; }
  int GetLid(int gid) const { return gid - nslist[ranklist[gid]]; }
  RegionSize GetBlockSize(const LogicalLocation &loc) const;

  // ProblemGenerator + first ghost exchange + FillDerived (mesh.cpp:745)
  void Initialize(bool init_problem, ParameterInput *pin, ApplicationInput *app_in);

  // which of its 27 topological elements (index (o1+1) + 3 (o2+1) + 9 (o3+1)) leaf block `gid`
  // owns: DetermineOwnership, mesh/forest/block_ownership.cpp:42-83.  Any rank can ask about any
  // block (every rank holds the whole tree), so both sides of an inter-device channel derive the
  // same ownership mask without a message.
  const std::array<bool, 27> &Ownership(int gid) const;

  // pure topology helpers (also used by the CPU-only tests)
  static void AssignBlocks(const std::vector<double> &costlist, int nranks,
                           std::vector<int> &ranklist);

 private:
  int pack_size_ = -1;
  int slab_capacity_ = 0;
  DeviceBuffer scratch_;
  std::unordered_map<LogicalLocation, int, LogicalLocationHash> leaf_gid_;
  std::unordered_map<LogicalLocation, int, LogicalLocationHash> internal_;
  void BuildTree(ParameterInput *pin, const std::vector<LogicalLocation> &leaves);
  // what both constructors share: knobs of the deck, rank assignment, block list, fields
  void ReadKnobs(ParameterInput *pin);
  void FinishConstruction(ParameterInput *pin);
  // (re)create this rank's MeshBlocks from loclist / ranklist; blocks of `keep` that sit at an
  // unchanged location are reused (they carry their refinement counters and time step)
  void BuildBlockList(const BlockList_t *keep);
  bool UpdateMeshBlockTree(std::vector<LogicalLocation> &new_leaves, int &nnew, int &ndel);
  void RedistributeAndRefineMeshBlocks(const std::vector<LogicalLocation> &new_leaves);
  void RebuildFromLeaves(const std::vector<LogicalLocation> &new_leaves,
                         const BlockList_t &old_blocks);
  void FindNeighbors(MeshBlock &mb) const;
  std::vector<NeighborBlock> FindNeighbors(const LogicalLocation &loc) const;
  mutable std::unordered_map<int, std::array<bool, 27>> ownership_;
  // blocks created by refinement in the remesh in progress: they rank below older blocks of
  // their level when shared elements are owned (block_ownership.cpp:48-54); empty otherwise
  std::unordered_set<LogicalLocation, LogicalLocationHash> newly_refined_;
  bool WrapLocation(const LogicalLocation &in, LogicalLocation &out) const;
  int64_t BlocksAtLevel(int level, int d) const;
};

} // namespace parthenon
