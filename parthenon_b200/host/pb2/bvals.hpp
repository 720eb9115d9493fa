// bvals.hpp — ghost-zone exchange "in one" (host side).
//
// Same task functions as the reference's src/bvals/comms/bvals_in_one.hpp:41-92:
//   SendBoundBufs<bt>, StartReceiveBoundBufs<bt>, ReceiveBoundBufs<bt>, SetBounds<bt>,
//   ProlongateBounds<bt>, AddBoundaryExchangeTasks, and the flux-correction aliases.
// What changes underneath (B200 design):
//   * BndInfo / ProResInfo become flat POD region tables (include/parthenon_b200.h) built
//     once per MeshData and rebuilt only when its allocation generation changes — there is
//     no per-call host walk over 13 312 regions (bvals_utils.hpp:140-202).
//   * local channels (sender and receiver on the same GPU) are ONE fused launch that moves
//     sender box -> receiver ghost box with no intermediate buffer: half the HBM traffic of
//     pack + unpack.  SendBoundBufs<local> only publishes "sent"; SetBounds<local> copies.
//     On uniform meshes that launch does not even read a region table (see uniform_halo).
//   * nonlocal channels are packed straight into one contiguous slab per peer GPU, shipped
//     by one grouped ncclSend/ncclRecv per peer on a communication stream, and unpacked
//     after a stream-side event wait (no host polling; the CommBuffer state machine of
//     utils/communication_buffer.hpp collapses to two events).
//   * restriction runs before pack / after unpack, prolongation in ProlongateBounds, each
//     as one launch over all regions (boundary_communication.cpp:82-87, 338-346, 361-393).
#pragma once
#include <map>
#include <memory>
#include <vector>

#include "mesh_data.hpp"
#include "tasks.hpp"

namespace parthenon {

enum class IndexRangeType { BoundaryInteriorSend, BoundaryExteriorRecv, InteriorSend, InteriorRecv };

// index box of a boundary region, cell centred (bnd_info.cpp:105-252): s/e in (i,j,k) order
struct IndexBox {
  int s[3], e[3];
  int n(int d) const { return e[d] - s[d] + 1; }
  int64_t size() const { return static_cast<int64_t>(n(0)) * n(1) * n(2); }
};
IndexBox CalcIndices(const NeighborBlock &nb, const MeshBlock *pmb, IndexRangeType ir_type,
                     bool prores);
// the same for the face flux (Metadata::Flux, element F_dir with dir the direction of the face
// offset; GetFluxCorrectionElements bnd_info.cpp:71-83): one face thick along dir.  A block
// with a coarser neighbour gets the box in its coarse index space (bnd_info.cpp:124-125).
IndexBox CalcIndicesFlux(const NeighborBlock &nb, const MeshBlock *pmb);

// CalcIndices for one topological element (bnd_info.cpp:105-252 with TopologicalOffset and the
// element bounds of mesh/domain.hpp:162-251); el may be any of the ten elements whatever the
// field holds (the internal prolongation runs over boxes of container elements)
IndexBox CalcIndicesTE(const NeighborBlock &nb, const MeshBlock *pmb, TE el,
                       IndexRangeType ir_type, bool prores = false);
// the ownership mask a RECEIVING region applies to element el (bnd_info.cpp:232-248)
std::array<bool, 27> RecvMask(const Mesh *pm, const NeighborBlock &nb, const MeshBlock *pmb,
                              TE el);
// GetIndexRangeMaskFromOwnership (block_ownership.cpp:85-140): which entries of a receive box
// — first / inner / last index per direction, index (i+1) + 3 (j+1) + 9 (k+1) — the receiver
// takes from a sender with the given block ownership; sox = offsets of the receiver seen from
// the sender
std::array<bool, 27> IndexRangeMask(TE el, const std::array<bool, 27> &sender_ownership,
                                    const int sox[3]);
// the boxes (relative to the start of a box of extents n) that a mask lets through.  The unpack
// predicate of the reference (SpatiallyMaskedIndexer6D::IsActive, utils/indexer.hpp:163-175) is
// resolved on the host: masked-out entries are neither read nor sent, kernels stay predicate-free.
std::vector<IndexBox> ActivePieces(const int n[3], const std::array<bool, 27> &mask);

// the C-ABI region that fills the ghost slab of mesh face `face` of one block of a CELL-CENTRED
// field with a stock condition (type = PB2_BC_OUTFLOW / PB2_BC_REFLECT); user boundary
// conditions can build their own tables from these
pb2_bc_region MakeBcRegion(Variable &v, const MeshBlock *pmb, int face, int type, bool coarse);

// (o1, o2, o3) -> 0..26, and the entry of `sender`'s neighbour list that describes the channel
// towards `receiver_gid` seen from the receiver at offsets `roff` (channels are keyed by sender
// gid, receiver gid and the location index: bvals_utils.hpp:43-67)
int OffsetIndexOf(int o1, int o2, int o3);
const NeighborBlock *MatchingNeighbor(const MeshBlock *sender, int receiver_gid,
                                      const int roff[3]);

// one boundary channel as the host sees it (pure topology: testable without a device).  A
// channel of a face / edge / node field is split into pieces: one per topological element and
// active sub-box of the ownership mask.
struct Channel {
  int sender_gid, receiver_gid;
  int var;          // index into the MeshData's FillGhost variable list
  int offset_index; // sender-perspective offset index (0..26): the channel key
  int piece = 0;    // element * 32 + sub-box number (0 for cell-centred fields)
  int comp0 = 0;    // first slab component of the piece
  int ncomp = 0;    // slab components it moves
  IndexBox send_box, recv_box;
  bool send_coarse; // sender reads its coarse buffer (receiver is coarser)
  bool recv_coarse; // receiver writes its coarse buffer (sender is coarser)
  int sender_rank, receiver_rank;   // real ranks (GPUs)
  int sender_vrank, receiver_vrank; // virtual ranks inside one GPU (test knob)
  int64_t slab_off = -1;            // nonlocal: offset in Reals inside the peer segment
  // forests: the sender lives in a differently oriented tree.  recv_box is then the box in the
  // SENDER's logical coordinates (bnd_info.cpp:216-228) and the unpack writes each element at
  // lcoord_trans.InverseTransform (pb2_bnd_region::lcoord_*); ncell = the array extent along x1
  bool transformed = false;
  forest::LogicalCoordinateTransformation lcoord_trans;
  int ncell = 0;
};

// Everything one exchange of one MeshData needs, derived from topology alone.
struct ExchangePlan {
  std::vector<Channel> local;     // received by this MeshData from the same (virtual) rank
  std::vector<Channel> send;      // nonlocal, sent by this MeshData; sorted by peer, key
  std::vector<Channel> recv;      // nonlocal, received by this MeshData
  std::vector<int64_t> send_off;  // [npeers + 1] slab segment offsets in Reals
  std::vector<int64_t> recv_off;
  int npeers = 1;                 // real ranks, or virtual ranks in the test mode
  int64_t local_elements = 0, send_elements = 0, recv_elements = 0;
};
// what the plan needs to know about each FillGhost variable: tensor components per element and
// where the values live
struct PlanVar {
  int ncomp = 1;
  TopologicalType tt = TopologicalType::Cell;
};
// index boxes and elements of the flux correction of a face field, whose flux is an edge field
// (bnd_info.cpp:71-103, :207-218)
IndexBox CalcIndicesFluxTE(const NeighborBlock &nb, const MeshBlock *pmb, TE el,
                           IndexRangeType ir_type, bool prores);
std::vector<TE> FluxCorrectionEdgeElements(const int off[3]);
// The flux correction of a face field as pure topology: which coarse boxes of which element the
// fine blocks restrict, and which (sender coarse-buffer box -> receiver box) pieces deliver them
// — one piece per active sub-box of the sender's ownership mask; pass 0: across block edges,
// pass 1: across faces (delivered second).
struct EdgeFluxRestrict {
  int gid, el; // el: 0..2 = E1..E3
  IndexBox box;
};
struct EdgeFluxPiece {
  int sender_gid, receiver_gid, el, pass;
  IndexBox send_box, recv_box;
  // pieces between devices (or virtual ranks): the peer segment, the sender's offset index
  // towards the receiver and the sub-box number (the sort key both sides share), and the offset
  // in Reals inside the edge-flux part of the peer segment (for ncomp components per element)
  int seg = -1, offset_index = 0, sub = 0;
  int64_t slab_off = -1;
};
struct EdgeFluxPlan {
  // same-device neighbours, listed from the coarse receiver's side: what its fine neighbours
  // restrict, and the pieces that deliver it
  std::vector<EdgeFluxRestrict> restricts;
  std::vector<EdgeFluxPiece> pieces;
  // neighbours on another device: what this rank's fine blocks restrict and send (send_box
  // set), what its coarse blocks receive (recv_box set); both sorted by (segment, sender,
  // receiver, offset index, element, sub-box) so that the two sides agree on the offsets
  std::vector<EdgeFluxRestrict> send_restricts;
  std::vector<EdgeFluxPiece> send, recv;
  std::vector<int64_t> send_off, recv_off; // [npeers + 1], Reals, for `ncomp` components
  int npeers = 1;
};
// ncomp: tensor components per element of the flux field (sizes the slab offsets)
EdgeFluxPlan BuildEdgeFluxPlan(const Mesh *pm, const BlockList_t &blocks, int ncomp = 1);
ExchangePlan BuildExchangePlan(const Mesh *pm, const BlockList_t &blocks,
                               const std::vector<PlanVar> &vars);

struct BvarsCache {
  ~BvarsCache();
  void Clear();
  uint64_t built_generation = 0;
  // built once per block list and FillGhost signature, shared between the containers of a
  // partition through Mesh::plan_cache; never null
  std::shared_ptr<const ExchangePlan> plan = std::make_shared<const ExchangePlan>();
  bool plan_built = false;
  // the ownership of shared elements changed (end of a remesh): plan and tables are rebuilt
  void Invalidate() {
    plan_built = false;
    built_generation = 0;
  }
  std::vector<Variable *> vars;
  pb2_bnd_table *copy_local = nullptr;
  // uniform meshes, dense fields, one batch per device: local channels need no region table —
  // one descriptor-free launch per field pulls every ghost cell from the owning neighbour
  // (pb2_halo_copy_uniform); halo_nbr is [nblocks][27] on the device
  bool uniform_halo = false;
  // Lazy same-device ghosts.  A consumer that reads its neighbours' interiors directly (the fast
  // burgers stage, pb2_burgers_args::nbr_direct) does not need the same-device ghost cells of
  // its input: the producer of this container sets `defer_local`, the next SetBounds<local>
  // then skips its copy and leaves `local_ghosts_stale` set.  Anything else that reads ghost
  // cells calls EnsureLocalGhosts first (host field access, outputs, the bit-exact stage ...).
  bool defer_local = false, local_ghosts_stale = false;
  // Deferred inter-GPU unpack: a consumer that starts a stage on the blocks WITHOUT remote faces
  // (burgers FusedStage) lets SetBounds<nonlocal> unpack on the communication stream right
  // behind the NCCL exchange and waits for `unpacked` only before it touches the other blocks.
  // `defer_remote` is set by the producer; `remote_pending`: the compute stream has not waited yet
  // (EnsureLocalGhosts makes it wait for anyone else who reads ghost cells).
  bool defer_remote = false, remote_pending = false;
  DeviceBuffer halo_nbr;
  pb2_bnd_table *pack = nullptr, *unpack = nullptr;
  // [0]: regions whose neighbour is local, [1]: nonlocal
  pb2_bnd_table *restrict_send[2] = {nullptr, nullptr}, *restrict_set[2] = {nullptr, nullptr};
  pb2_bnd_table *prolongate[2][3] = {{nullptr, nullptr, nullptr},
                                     {nullptr, nullptr, nullptr}}; // per prolongation op
  // the same for face / edge / node fields: one region per element and active sub-box of the
  // ownership mask; te_internal: ProlongateInternalAverage over container-element boxes
  pb2_bnd_table *te_restrict_send[2] = {nullptr, nullptr}, *te_restrict_set[2] = {nullptr, nullptr};
  pb2_bnd_table *te_prolongate[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
  pb2_bnd_table *te_internal[2] = {nullptr, nullptr};
  pb2_bnd_table *te_toth_roe[2] = {nullptr, nullptr}; // ProlongateInternalTothAndRoe (faces)
  DeviceBuffer send_slab, recv_slab;
  // Peer push (include/parthenon_b200.h "peer push"): on uniform meshes with dense fields the
  // inter-GPU halo is ONE copy launch whose destinations are the ghost cells in the peers'
  // memory (field slabs mapped through CUDA IPC), with ready / arrival flags instead of NCCL
  // send / recv — no slabs, no unpack.  push_peers: the ranks this MeshData exchanges with.
  bool push_mode = false, push_direct = false, push_ce = false;
  // copy-engine form: per peer, where its segment of our send slab goes (the peer's receive slab
  // at its recv_off[this rank]), where it starts in the send slab and how many Reals it holds
  struct PushSegment {
    Real *dst;
    int64_t src_off, count;
  };
  std::vector<PushSegment> push_segments;
  pb2_bnd_table *push = nullptr;
  DeviceBuffer push_flags, push_counter, push_peer_flags, push_peer_ids;
  std::vector<pb2_ipc_handle> push_opened; // mappings this cache holds a reference to
  int push_npeers = 0;
  int32_t push_seq = 0;
  pb2_event_t packed = nullptr, received = nullptr, sent = nullptr;
  bool nonlocal_in_flight = false;
  // Overlap of inter-GPU halos with interior work (uniform meshes): blocks with a nonlocal
  // neighbour ("boundary") are advanced first and signal `early_ready`; SendBoundBufs<nonlocal>
  // then packs and ships on the communication stream while the remaining ("interior") blocks
  // are still being advanced on the compute stream.
  DeviceBuffer ids_boundary, ids_interior;
  int n_boundary = 0, n_interior = 0;
  // one launch over [boundary blocks..., interior blocks...]: the last sweep counts finished
  // thread blocks of the boundary part in `progress` (pb2_burgers_args::progress) and the
  // communication stream waits for the count on the device (pb2_stream_wait_value)
  DeviceBuffer ids_ordered, progress;
  int32_t progress_target = 0; // > 0: SendBoundBufs<nonlocal> waits for it before packing
  // multilevel meshes: blocks with a FACE neighbour on another level take part in flux
  // correction and need their face fluxes stored (ids_flxcor); all others (ids_plain) can run
  // the flux-free sweeps.  Blocks with a coarser neighbour own ghosts that a stage's exchange
  // does not refresh (ids_stale_ghosts: the ghost part of the full-extent WeightedSumData).
  DeviceBuffer ids_flxcor, ids_plain, ids_stale_ghosts;
  int n_flxcor = 0, n_plain = 0, n_stale_ghosts = 0;
  pb2_event_t early_ready = nullptr, unpacked = nullptr;
  bool early_valid = false, unpacked_valid = false;
  // local channels: SendBoundBufs<local> publishes a generation; receivers consume it
  uint64_t send_generation = 0;
  std::map<int, uint64_t> consumed_generation; // by sender partition
  // traffic accounting for bench.py: Reals moved by the last exchange
  int64_t elements_local = 0, elements_nonlocal = 0;

  // physical boundary conditions (outflow / reflect) of the blocks on a non-periodic mesh
  // face, one table per direction: [0] fine arrays, [1] coarse buffers
  pb2_bnd_table *bc[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
  bool has_bcs = false;

  // sparse fields (allocation-aware exchange): one "message is non-null" flag per local channel
  bool sparse = false;
  DeviceBuffer sparse_flags;
  std::vector<int32_t> sparse_flags_h;
  // ... and per inter-device channel: the pack kernel raises int32 flags, they travel as one
  // Real per channel in their own small slab (same per-peer order as the data slab, so the
  // offsets need no handshake either) and come back as the int32 data flags of the unpack
  DeviceBuffer send_flags, recv_flags;           // int32 per send / recv channel
  DeviceBuffer send_flag_slab, recv_flag_slab;   // Real per send / recv channel
  std::vector<int32_t> send_flags_h, recv_flags_h;
  std::vector<Real> flag_slab_h;
  std::vector<int64_t> send_flag_off, recv_flag_off; // [npeers + 1] channel counts per peer

  // flux correction at fine-coarse faces (flxcor_send / flxcor_recv), built on first use:
  // fused restrict+deliver for same-device channels, restrict-into-slab + unpack for the rest
  bool flxcor_built = false;
  pb2_bnd_table *flxcor_local = nullptr, *flxcor_pack = nullptr, *flxcor_unpack = nullptr;
  // ... of a FACE field: its flux is the edge field "bnd_flux::<name>".  Fine blocks restrict the
  // shared edge elements into the flux field's coarse buffer (teflx_restrict), then they are
  // copied into the coarser block's flux array under the sender's ownership mask, messages
  // across block edges first ([0]), across faces second ([1]) — see oracle/pb2_oracle.c on why
  // the order matters.
  pb2_bnd_table *teflx_restrict = nullptr, *teflx_copy[2] = {nullptr, nullptr};
  // ... and across devices: this rank's fine blocks restrict (teflx_restrict_send) and pack the
  // entries they own into the flux-correction slab behind the face fluxes of the peer segment;
  // the receiver unpacks block-edge messages ([0]) before face messages ([1])
  pb2_bnd_table *teflx_restrict_send = nullptr, *teflx_pack = nullptr,
                *teflx_unpack[2] = {nullptr, nullptr};
  int64_t teflx_elements = 0;
  std::vector<int64_t> flxcor_send_off, flxcor_recv_off; // [npeers + 1]
  int64_t flxcor_send_elements = 0, flxcor_recv_elements = 0, flxcor_local_elements = 0;
  DeviceBuffer flxcor_send_slab, flxcor_recv_slab;
  pb2_event_t flxcor_packed = nullptr, flxcor_received = nullptr;
  bool flxcor_in_flight = false;
};

void BuildBoundaryBuffers(std::shared_ptr<MeshData<Real>> &md);
// the (re)built exchange cache of a batch
BvarsCache &GetBvarsCache(MeshData<Real> *md);

template <BoundaryType bound_type>
TaskStatus StartReceiveBoundBufs(std::shared_ptr<MeshData<Real>> &md);
template <BoundaryType bound_type>
TaskStatus SendBoundBufs(std::shared_ptr<MeshData<Real>> &md);
template <BoundaryType bound_type>
TaskStatus ReceiveBoundBufs(std::shared_ptr<MeshData<Real>> &md);
template <BoundaryType bound_type>
TaskStatus SetBounds(std::shared_ptr<MeshData<Real>> &md);
template <BoundaryType bound_type>
TaskStatus ProlongateBounds(std::shared_ptr<MeshData<Real>> &md);

inline TaskStatus SendBoundaryBuffers(std::shared_ptr<MeshData<Real>> &md) {
  return SendBoundBufs<BoundaryType::any>(md);
}
inline TaskStatus StartReceiveBoundaryBuffers(std::shared_ptr<MeshData<Real>> &md) {
  return StartReceiveBoundBufs<BoundaryType::any>(md);
}
inline TaskStatus ReceiveBoundaryBuffers(std::shared_ptr<MeshData<Real>> &md) {
  return ReceiveBoundBufs<BoundaryType::any>(md);
}
inline TaskStatus SetBoundaries(std::shared_ptr<MeshData<Real>> &md) {
  return SetBounds<BoundaryType::any>(md);
}
// flux corrections only exist at fine-coarse faces; on meshes without them these complete
// immediately (boundary_communication.cpp:454-461).  FluxCorrection(md) is Send + Set in one
// call for fused stages.
TaskStatus StartReceiveFluxCorrections(std::shared_ptr<MeshData<Real>> &md);
TaskStatus LoadAndSendFluxCorrections(std::shared_ptr<MeshData<Real>> &md);
TaskStatus ReceiveFluxCorrections(std::shared_ptr<MeshData<Real>> &md);
TaskStatus SetFluxCorrections(std::shared_ptr<MeshData<Real>> &md);
void FluxCorrection(MeshData<Real> *md);

// physical boundaries (bvals/boundary_conditions.cpp:36-58, :85-95): generic outflow / reflect
// on the blocks that touch a non-periodic mesh face, whole MeshData batch in <= 3 launches.
// The per-block form of the reference is kept for source compatibility; batches should use
// the MD forms.
TaskStatus ApplyBoundaryConditions(std::shared_ptr<MeshBlockData<Real>> &rc);
TaskStatus ApplyBoundaryConditionsMD(std::shared_ptr<MeshData<Real>> &md);
// run the deferred same-device ghost copy of md (and the physical boundary fill that follows an
// exchange) if its ghost cells are stale; no-op otherwise
void EnsureLocalGhosts(MeshData<Real> *md);
// make the compute stream wait for a deferred inter-GPU unpack of md (see BvarsCache::defer_remote)
void EnsureRemoteGhosts(MeshData<Real> *md);
// ... of every container of the mesh
void EnsureLocalGhosts(Mesh *pm);
TaskStatus ApplyBoundaryConditionsOnCoarseOrFineMD(std::shared_ptr<MeshData<Real>> &md,
                                                   bool coarse);

// boundary_communication.cpp:406-452: Send -> Receive -> Set -> [Prolongate] in the
// local / nonlocal split
TaskID AddBoundaryExchangeTasks(TaskID dependency, TaskList &tl,
                                std::shared_ptr<MeshData<Real>> &md, bool multilevel);

// one blocking exchange (Mesh::CommunicateBoundaries, mesh.cpp:640-706)
void CommunicateBoundaries(std::shared_ptr<MeshData<Real>> &md, bool prolongate);

} // namespace parthenon
