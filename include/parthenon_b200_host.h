/* parthenon_b200_host.h — C entry points of the C++ host framework (libpb200_host.so).
 *
 * The host framework mirrors Parthenon's application API in C++ (namespace parthenon:
 * ParameterInput, Metadata, StateDescriptor, Mesh, MeshData, TaskList, MultiStageDriver,
 * SendBoundBufs/ReceiveBoundBufs/SetBounds/ProlongateBounds — headers under
 * parthenon_b200/host/pb2/) and reaches the GPU only through include/parthenon_b200.h.
 * These functions expose a running application (ParthenonManager + driver, reference
 * src/parthenon_manager.cpp, src/driver/driver.cpp:67-193) to non-C++ callers: bench.py,
 * the parity tests, language bindings.  All return 0 on success, <0 on error
 * (pb2h_last_error()); exceptions never cross this boundary.
 */
#ifndef PARTHENON_B200_HOST_H_
#define PARTHENON_B200_HOST_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pb2h_sim pb2h_sim;

const char *pb2h_last_error(void);

/* Build an application: parse `deck` (Athena++-style input text, reference
 * src/parameter_input.cpp) plus newline-separated "block/key=value" overrides, create the
 * packages, the mesh (rank `rank` of `nranks` GPUs; nccl_id = 128-byte id from
 * pb2_comm_unique_id when nranks > 1), run the problem generator, the first ghost exchange
 * and FillDerived (Mesh::Initialize, mesh.cpp:745).  app: "burgers".
 * leaves: optional explicit leaf list (level, lx1, lx2, lx3) x nleaves, else NULL/0. */
int pb2h_sim_create(pb2h_sim **sim, const char *app, const char *deck, const char *overrides,
                    int rank, int nranks, const uint8_t *nccl_id, const int *leaves,
                    int nleaves);
/* Mesh topology only (no device is touched): for CPU-side tests of partitioning,
 * neighbour lists, index boxes and slab layouts. */
int pb2h_topology_create(pb2h_sim **sim, const char *deck, const char *overrides, int rank,
                         int nranks, const int *leaves, int nleaves);
/* device-free AMR host logic on a topology object: apply one AmrTag per block (-1 derefine,
 * 0 same, 1 refine) through MeshRefinement::SetRefinement's rules, update the tree
 * (Mesh::UpdateMeshBlockTree: proper nesting, sibling-complete derefinement) and rebuild the
 * block list.  *changed = 1 if the mesh changed. */
int pb2h_topology_regrid(pb2h_sim *sim, const int *tags, int nblocks, int *changed);
/* the per-block derefinement counters (MeshRefinement::deref_count_) of a topology object:
 * set != 0 stores counts[] into the blocks, set == 0 reads them out */
int pb2h_topology_derefine_counts(pb2h_sim *sim, int *counts, int nblocks, int set);
/* application "tecomm" on an adaptive mesh: tag every block with the geometric criterion of
 * `cycle` (tests/golden/refgen/teamr_dump_main.cpp), then adapt the mesh
 * (Mesh::LoadBalancingAndAdaptiveMeshRefinement); *changed = 1 if blocks were (de)refined */
int pb2h_sim_tag_and_remesh(pb2h_sim *sim, int cycle, int *changed);
int pb2h_sim_destroy(pb2h_sim *sim);

/* EvolutionDriver pieces (driver.cpp:67-193): what Execute does before its loop, N cycles
 * of the loop body, or the whole thing */
int pb2h_sim_pre_execute(pb2h_sim *sim);
int pb2h_sim_cycle(pb2h_sim *sim, int ncycles);
/* one cycle in the two halves either side of the reference's PostStepUserWorkInLoop hook
 * (driver.cpp:112-129): phase 0 = Step + time advance, phase 1 =
 * LoadBalancingAndAdaptiveMeshRefinement + SetGlobalTimeStep.  pb2h_sim_cycle does both. */
int pb2h_sim_cycle_phase(pb2h_sim *sim, int phase);
int pb2h_sim_execute(pb2h_sim *sim);
int pb2h_sim_sync(pb2h_sim *sim);
void *pb2h_sim_stream(pb2h_sim *sim); /* cudaStream_t the application enqueues on */
double pb2h_sim_time(pb2h_sim *sim);
double pb2h_sim_dt(pb2h_sim *sim);
int pb2h_sim_ncycle(pb2h_sim *sim);
int pb2h_sim_set_dt(pb2h_sim *sim, double dt);
double pb2h_sim_zone_cycles_per_second(pb2h_sim *sim);

/* out: ndim, nbtotal, nblocks on this rank, ni, nj, nk (with ghosts), coarse ni, nj, nk,
 * multilevel, first gid of this rank, nghost */
int pb2h_sim_info(pb2h_sim *sim, int out[12]);
/* loc: level, lx1, lx2, lx3 — on a forest (2-D, lx3 = 0) the last slot holds the tree id */
int pb2h_sim_block(pb2h_sim *sim, int lid, int loc[4], double xmin[3], double xmax[3],
                   int *gid, int *nneighbors);
/* MeshBlock::boundary_flag per face (inner_x1, outer_x1, ...): -1 block, 1 reflect, 2 outflow,
 * 3 periodic, 4 user */
int pb2h_sim_block_bcs(pb2h_sim *sim, int lid, int out[6]);
/* topology of the forest application (host/forest/forest_app.hpp: the 2 x 2 forests of
 * example/boundary_exchange, `variant` 0..3), no device touched */
int pb2h_topology_create_forest(pb2h_sim **sim, const char *deck, const char *overrides,
                                int variant, int rank, int nranks);
/* out: gid, level, ox1, ox2, ox3, rank */
int pb2h_sim_neighbor(pb2h_sim *sim, int lid, int n, int out[6]);
/* ir_type 0 = BoundaryInteriorSend, 1 = BoundaryExteriorRecv (bnd_info.cpp:105-252) */
int pb2h_sim_calc_indices(pb2h_sim *sim, int lid, int n, int ir_type, int prores, int s[3],
                          int e[3]);
int pb2h_sim_ranklist(pb2h_sim *sim, int *ranks, int n);
/* channel plan of this rank for one ncomp-component field.  kind 0 local, 1 send, 2 recv;
 * rows (7 int64 each): sender gid, receiver gid, var, offset index, slab offset, Reals,
 * peer rank.  seg_off: [npeers+1] slab segment offsets (kinds 1,2).  Returns the count. */
int64_t pb2h_sim_plan(pb2h_sim *sim, int ncomp, int kind, int64_t *rows, int64_t max_rows,
                      int64_t *seg_off);
/* the same for one field of topological type tt (0 cell, 1 face, 2 edge, 3 node) with `ncomp`
 * tensor components, index boxes included: rows of 18 int64 [sender_gid, receiver_gid,
 * offset_index, piece, comp0, ncomp, send_s(i,j,k), recv_s(i,j,k), n(i,j,k), slab_off, peer,
 * coarse flags (bit 0: the sender reads its coarse buffer, bit 1: the receiver writes its;
 * forests: bit 2 = the sender's tree is oriented differently — the receive box is then given in
 * the SENDER's logical coordinates and element (i, j, k) of it lands at
 * out[d] = flip[d] ? ncell - 1 - in[dir[d]] : in[dir[d]], with dir[d] in bits 3+2d..4+2d,
 * flip[d] in bit 9+d and ncell in bits 12 and up)].
 * Channels of non-cell-centred fields come in pieces: one per topological element and active
 * sub-box of the sender's ownership mask (block_ownership.cpp:85-140). */
int64_t pb2h_sim_plan_boxes(pb2h_sim *sim, int ncomp, int tt, int kind, int64_t *rows,
                            int64_t max_rows);

/* field access: which = 0 data, 1..3 flux X1..X3, 4 coarse buffer.  Layout
 * [block][component][k][j][i] over this rank's blocks. */
int pb2h_sim_field_ptr(pb2h_sim *sim, const char *container, const char *field, int which,
                       void **dev_ptr, int64_t *nreal);
/* extents of a field's fine array: out = [blocks on this rank, slab components (topological
 * elements x tensor components), nk, nj, ni, topological elements].  Face / edge / node fields
 * are one entry longer than cell-centred ones in every non-symmetry direction. */
int pb2h_sim_field_dims(pb2h_sim *sim, const char *container, const char *field, int out[6]);
int pb2h_sim_get_field(pb2h_sim *sim, const char *container, const char *field, int which,
                       double *host, int64_t nreal);
int pb2h_sim_set_field(pb2h_sim *sim, const char *container, const char *field, int which,
                       const double *host, int64_t nreal);
/* sparse fields: out[b] = 1 if `field` is allocated on this rank's block b (Variable::IsAllocated,
 * reference src/interface/variable.hpp), else 0; dense fields report 1 everywhere */
int pb2h_sim_allocation(pb2h_sim *sim, const char *container, const char *field, int *out,
                        int nblocks);

/* state through HOST buffers holding interior cells only, [block][comp][nx3][nx2][nx1] over
 * this rank's blocks (pinned memory recommended).  upload = H2D + scatter + ghost exchange;
 * download = gather + D2H.  Asynchronous on the application's stream: call pb2h_sim_sync
 * before reading `host` after a download. */
int pb2h_sim_upload_interior(pb2h_sim *sim, const char *container, const char *field,
                             const double *host, int64_t nreal);
int pb2h_sim_download_interior(pb2h_sim *sim, const char *container, const char *field,
                               double *host, int64_t nreal);
/* Pipelined variant for streams of INDEPENDENT batches of state (ensemble members, frames of a
 * parameter scan): two lanes, each with its own device staging; the H2D copy runs on a copy
 * stream as soon as the lane's staging is free (prefetch), the scatter + ghost exchange joins
 * the application stream (commit), and the gather + D2H drains on a second copy stream
 * (writeback), so the copies of neighbouring batches overlap the cycle of the current one.
 *   prefetch(lane, batch n+1); commit(lane'); cycle; writeback(lane'); ... lane_sync(lane)
 * `host` buffers must stay valid until pb2h_sim_lane_sync(lane) (writeback) or the matching
 * commit has run (prefetch). */
int pb2h_sim_prefetch_interior(pb2h_sim *sim, const char *container, const char *field,
                               const double *host, int64_t nreal, int lane);
int pb2h_sim_commit_interior(pb2h_sim *sim, const char *container, const char *field, int lane);
int pb2h_sim_writeback_interior(pb2h_sim *sim, const char *container, const char *field,
                                double *host, int64_t nreal, int lane);
int pb2h_sim_lane_sync(pb2h_sim *sim, int lane);
/* one full ghost exchange of a container (Send -> Receive -> Set [-> Prolongate]),
 * Mesh::CommunicateBoundaries mesh.cpp:640-706 */
int pb2h_sim_exchange(pb2h_sim *sim, const char *container, int prolongate);
int pb2h_sim_exchange_phase(pb2h_sim *sim, const char *container, int phase);
/* Reals one exchange moves on this rank (ghost cells filled x components) */
int64_t pb2h_sim_exchange_elements(pb2h_sim *sim, const char *container, int64_t *local,
                                   int64_t *nonlocal);
/* flux correction of a face field (edge-centred flux) as pure topology, rows of 16 int64.
 * kind 0, restrictions on fine blocks whose coarser neighbour is on the same device: [gid, 0,
 * element 0..2 (E1..E3), 0, s(i,j,k) in the coarse index space, 0, 0, 0, n(i,j,k), 0, 0, 0];
 * kind 1, same-device deliveries: [sender gid, receiver gid, element, pass (0: across a block
 * edge, 1: across a face, delivered second), send_s(i,j,k) in the sender's coarse buffer,
 * recv_s(i,j,k) in the receiver's array, n(i,j,k), -1, -1, key];
 * kind 2: like 0 for this rank's fine blocks whose coarser neighbour is on another device;
 * kind 3 / 4: pieces this rank sends / receives across devices, like kind 1 with only the send /
 * receive start set, then [peer, offset in Reals inside the edge-flux part of the slab for one
 * component per element, key] - both sides list a peer's pieces in the same order at the same
 * offsets.  Returns the count. */
int64_t pb2h_sim_edge_flux_plan(pb2h_sim *sim, int kind, int64_t *rows, int64_t max_rows);
/* AddFluxCorrectionTasks (boundary_communication.cpp:454-461) on every partition of `container`:
 * fine blocks restrict the fluxes they share with coarser neighbours, the coarser blocks take
 * them — face fluxes of cell-centred fields and the edge-centred fluxes of face fields */
int pb2h_sim_flux_correction(pb2h_sim *sim, const char *container);
/* how the inter-device halo of `container` travels: 0 no inter-device channels, 1 per-peer slabs
 * + grouped ncclSend / ncclRecv, 2 peer push through the copy engines (pack, one device-to-device
 * copy per peer into its receive slab, arrival flags), 3 peer push by the pack kernel (stores
 * into the peers' receive slabs), 4 direct peer push (stores into the peers' ghost cells);
 * -1 on error */
int pb2h_sim_exchange_mode(pb2h_sim *sim, const char *container);
/* the history columns "MS Mass 0..7" of the burgers benchmark, reduced over ranks */
int pb2h_sim_history(pb2h_sim *sim, double out[8]);

/* A SparsePack over one container of the running application (parthenon::MakePackDescriptor +
 * Descriptor::GetPack, host/pb2/sparse_pack.hpp), handed out as the POD a user kernel takes by
 * value (include/parthenon_b200_pack.h).  names: '\n'-separated variable (or sparse pool) names,
 * each optionally prefixed "re:" for a regular expression; flags: '\n'-separated Metadata flag
 * names every selected field must carry ("WithFluxes", "FillGhost", "Independent", "Sparse" ...);
 * options: bit 0 WithFluxes, bit 1 Coarse, bit 2 Flatten.  The tables stay valid until the
 * container's allocation status changes (ask again then: the pack is cached otherwise).
 * host_bounds (optional): the bounds table [2][nblocks][nvar+1] copied to the host. */
#include "parthenon_b200_pack.h"
int pb2h_sim_sparse_pack(pb2h_sim *sim, const char *container, const char *names,
                         const char *flags, int options, pb2_sparse_pack *pack,
                         int32_t *host_bounds, int64_t host_bounds_len);
/* label of pack entry (b, idx), as SparsePack::LabelHost */
const char *pb2h_sim_sparse_pack_label(pb2h_sim *sim, const char *container, const char *names,
                                       const char *flags, int options, int b, int idx);
/* (de)allocate a sparse field on one block of every container (MeshBlock::AllocateSparse /
 * DeallocateSparse) — lets a test reproduce tst/unit/test_sparse_pack.cpp */
int pb2h_sim_set_sparse_allocation(pb2h_sim *sim, const char *field, int lid, int allocated);

#ifdef __cplusplus
}
#endif
#endif /* PARTHENON_B200_HOST_H_ */
