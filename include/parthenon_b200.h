/* parthenon_b200.h — C ABI of the B200-native ghost-zone hot path.
 *
 * This is the drop-in boundary.  The reference (parthenon-hpc-lab/parthenon @79d5d30) has no
 * FFI: its hot path is a set of C++ task functions that launch Kokkos lambdas.  Each entry
 * point below replaces the DEVICE part of one of those functions; the host-side shims in
 * parthenon_b200/host/ keep the reference's C++ signatures and call only this header (see
 * INTEGRATION.md for the binding a Parthenon maintainer would add).
 *
 * Conventions: plain pointers and sizes, POD structs, no C++/torch types; every function
 * returns 0 (PB2_OK) or a negative error code and never throws; the caller owns all field
 * memory, the library owns only tables/slabs created through it; all launches are
 * asynchronous on the given stream (a cudaStream_t passed as void*; NULL = default stream).
 * All field arithmetic is FP64 (`Real` = double, reference basic_types.hpp:30-39); arrays
 * are row-major with i fastest (reference kokkos_abstraction.hpp:52).
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with
 * PB2_ERR_NO_DEVICE.
 */
#ifndef PARTHENON_B200_H_
#define PARTHENON_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB2_OK 0
#define PB2_ERR_INVALID (-1)
#define PB2_ERR_CUDA (-2)
#define PB2_ERR_NO_DEVICE (-3)
#define PB2_ERR_NCCL (-4)
#define PB2_ERR_UNSUPPORTED (-5)

typedef void *pb2_stream_t; /* cudaStream_t */
typedef void *pb2_event_t;  /* cudaEvent_t */

/* ---------------------------------------------------------------------------------------
 * plumbing: device, memory, streams, events (so host C++ never links the CUDA runtime)
 * ------------------------------------------------------------------------------------- */
int pb2_version(void);
const char *pb2_last_error(void);
int pb2_device_count(int *count);
int pb2_set_device(int device);
int pb2_device_sm_count(int *count);
int pb2_malloc(void **ptr, size_t bytes);
int pb2_free(void *ptr);
/* pb2_malloc recycles large allocations through an exact-size cache per device (adaptive runs
 * re-create multi-GB slabs on every remesh); this returns everything it holds to the driver.
 * PB2_ALLOC_CACHE_GB (default 48) bounds what the cache may keep. */
int pb2_cache_trim(void);
int pb2_host_alloc(void **ptr, size_t bytes); /* pinned */
int pb2_host_free(void *ptr);
int pb2_memset(void *ptr, int value, size_t bytes, pb2_stream_t stream);
int pb2_memcpy_h2d(void *dst, const void *src, size_t bytes, pb2_stream_t stream);
int pb2_memcpy_d2h(void *dst, const void *src, size_t bytes, pb2_stream_t stream);
int pb2_memcpy_d2d(void *dst, const void *src, size_t bytes, pb2_stream_t stream);
int pb2_stream_create(pb2_stream_t *stream);
/* a stream of the device's highest (high != 0) or lowest priority: kernels of a high-priority
 * stream get free SM slots before pending work of other streams (the halo pack / NCCL stream) */
int pb2_stream_create_priority(pb2_stream_t *stream, int high);
/* Device-side wait: work enqueued on `stream` after this call starts once *counter >= target
 * (a one-thread kernel polls the counter).  The producer is a kernel on another stream that
 * advances the counter from inside (pb2_burgers_args::progress): finer than an event, which only
 * fires at a kernel boundary.  Do not use under a profiler that serialises kernels. */
int pb2_stream_wait_value(pb2_stream_t stream, const int32_t *counter, int32_t target);
int pb2_stream_destroy(pb2_stream_t stream);
int pb2_stream_sync(pb2_stream_t stream);
int pb2_device_sync(void);
int pb2_event_create(pb2_event_t *ev);
int pb2_event_destroy(pb2_event_t ev);
int pb2_event_record(pb2_event_t ev, pb2_stream_t stream);
int pb2_event_sync(pb2_event_t ev);
int pb2_event_query(pb2_event_t ev); /* 0 done, 1 not yet */
int pb2_stream_wait_event(pb2_stream_t stream, pb2_event_t ev);
int pb2_event_elapsed_ms(pb2_event_t start, pb2_event_t stop, float *ms);
/* optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline
 * figures).  enable(1) brackets every launch with an event pair; get() synchronises the
 * device and returns the accumulated device time and launch count of kernel class `id`
 * (0 <= id < pb2_profile_kernels()). */
int pb2_profile_enable(int on);
int pb2_profile_reset(void);
int pb2_profile_kernels(void);
int pb2_profile_get(int id, const char **name, double *total_ms, int64_t *launches);
/* work the recorded launches of kernel `id` processed, in the kernel's own unit: zones (cells of
 * the blocks actually launched) for the stencil kernels, values moved for the copy / pack /
 * unpack / ghost-fill kernels, 0 where a kernel does not report it.  Bytes per launch in
 * bench.py's roofline = work / launches x the kernel's algorithmic bytes per unit. */
int pb2_profile_get_work(int id, double *work);
/* FP64 FMA throughput of the current device in TFLOP/s (2 flops per FMA), measured with a
 * register-resident microbenchmark: the FP64-pipe roofline of the burgers stencil */
int pb2_measure_fp64_peak(double *tflops);
/* number of kernels this library has launched in this process (bench.py gpu_launches) */
int64_t pb2_launch_count(void);

/* ---------------------------------------------------------------------------------------
 * ghost exchange: boundary-region tables
 * replaces BndInfo (src/bvals/comms/bnd_info.hpp:51-83) and the pack / unpack kernels of
 * SendBoundBufs (src/bvals/comms/boundary_communication.cpp:95-140) and SetBounds (:273-334)
 * ------------------------------------------------------------------------------------- */
#define PB2_REGION_ALLOCATED 1u     /* BndInfo::allocated */
#define PB2_REGION_BUF_ALLOCATED 2u /* BndInfo::buf_allocated (receive side) */
#define PB2_REGION_SAME_TO_SAME 4u  /* BndInfo::same_to_same */
#define PB2_REGION_DST_UNALLOCATED 8u /* copy regions: the RECEIVING field is not allocated */

/* One side of a boundary channel: an index box of one block's array <-> a contiguous run
 * of the buffer slab.  Buffer order is [comp][k][j][i] (Indexer6D, src/utils/indexer.hpp). */
typedef struct pb2_bnd_region {
  double *var;       /* component 0 of the array the box indexes (fine data, or the coarse
                        buffer when the neighbour is coarser: bnd_info.cpp:285-289) */
  int64_t buf_off;   /* offset in Reals into the buffer slab given at launch */
  int32_t s[3];      /* box start (i, j, k) — CalcIndices, bnd_info.cpp:105-252 */
  int32_t n[3];      /* box extent (i, j, k) */
  int32_t ncomp;     /* flattened tensor components (t,u,v) */
  int32_t stride_j;  /* array strides in Reals */
  int32_t stride_k;
  int32_t stride_c;
  int32_t flag_slot; /* send: slot of sending_nonzero_flags; recv: slot of the flag that
                        says whether the buffer holds data (<0: always data) */
  uint32_t status;   /* PB2_REGION_* */
  double value;      /* send: allocation threshold; recv: sparse default value */
  /* RECEIVE regions only: the neighbour's LogicalCoordinateTransformation (trees of a forest
   * that meet with different orientations, mesh/forest/logical_coordinate_transformation.hpp:
   * 36-112).  If lcoord_on != 0 the buffer element of box cell (i, j, k) — start s, extent n, in
   * the SENDER's orientation — lands at InverseTransform({i, j, k}) (:90-99):
   *   out[d] = lcoord_flip[d] ? lcoord_ncell - 1 - in[lcoord_dir[d]] : in[lcoord_dir[d]]
   * and is multiplied by `fac` (the sign a vector component picks up, :66-77), exactly what
   * SetBounds does per element (boundary_communication.cpp:282-308).  lcoord_dir = |dir_connection|,
   * lcoord_ncell = the array's extent GetDim(1) (bnd_info.cpp:297).  All zero = identity. */
  int32_t lcoord_on;
  int32_t lcoord_dir[3];
  int32_t lcoord_flip[3];
  int32_t lcoord_ncell;
  double fac;
} pb2_bnd_region;

/* Fused same-device channel: sender box -> receiver box with no intermediate buffer
 * (BuffCommType::both channels, src/utils/communication_buffer.hpp:145-147, 390-400). */
typedef struct pb2_copy_region {
  const double *src;
  double *dst;
  int32_t ss[3]; /* source box start (i,j,k) */
  int32_t ds[3]; /* destination box start */
  int32_t n[3];  /* common extent */
  int32_t ncomp;
  int32_t src_stride_j, src_stride_k, src_stride_c;
  int32_t dst_stride_j, dst_stride_k, dst_stride_c;
  int32_t flag_slot; /* sparse: slot receiving "any |x| >= threshold" (or <0) */
  uint32_t status;   /* PB2_REGION_ALLOCATED refers to the source */
  double threshold;
  double default_value; /* written when the source is unallocated / all below threshold */
} pb2_copy_region;

typedef struct pb2_bnd_table pb2_bnd_table; /* opaque, device resident */

/* Upload `n` regions (host array) and build the work decomposition.  Called on cache
 * rebuild only (reference: RebuildBufferCache, src/bvals/comms/bvals_utils.hpp:212-260). */
int pb2_bnd_table_create(pb2_bnd_table **table, const pb2_bnd_region *regions, int64_t n);
int pb2_copy_table_create(pb2_bnd_table **table, const pb2_copy_region *regions, int64_t n);
/* Destroying a table never waits for the device: launches that still read it stay valid, its
 * device memory goes back to a pool (size classes 4 KB x 4^k per device) and is handed out again
 * only after a device synchronisation that followed the destroy (csrc/runtime.cu: table_alloc).
 * Tables of every kind (boundary, copy, prolongation / restriction, flux correction, boundary
 * condition) are destroyed with this one call. */
int pb2_bnd_table_destroy(pb2_bnd_table *table);
/* total Reals covered by the table's boxes */
int64_t pb2_bnd_table_elements(const pb2_bnd_table *table);

/* One launch packs every region of the table: buf[buf_off + m] = var(box)[m];
 * nonzero_flags[flag_slot] |= any(|x| >= value) when nonzero_flags != NULL
 * (boundary_communication.cpp:95-140).  nonzero_flags is int32 per slot, zeroed by caller. */
int pb2_pack(const pb2_bnd_table *table, double *buf, int32_t *nonzero_flags,
             pb2_stream_t stream);
/* One launch unpacks every region: var(box)[m] = buf[buf_off + m]; regions whose
 * data_flags[flag_slot] == 0 are filled with `value` (boundary_communication.cpp:273-334). */
int pb2_unpack(const pb2_bnd_table *table, const double *buf, const int32_t *data_flags,
               pb2_stream_t stream);
/* One launch moves every same-device channel straight from sender box to receiver box. */
int pb2_copy(const pb2_bnd_table *table, int32_t *nonzero_flags, pb2_stream_t stream);
/* The same exchange in the two halves a SPARSE field needs, because the receiver must know
 * whether a message is null before anything is written (and may have to allocate first):
 *  pb2_copy_flags   sender half (boundary_communication.cpp:95-157): nonzero_flags[flag_slot]
 *                   |= allocated source && any |x| >= threshold in the send box; writes no field
 *  pb2_copy_select  receiver half (:273-334): regions with data_flags[flag_slot] != 0 (or
 *                   flag_slot < 0) receive the data, all others default_value; regions marked
 *                   PB2_REGION_DST_UNALLOCATED are skipped */
int pb2_copy_flags(const pb2_bnd_table *table, int32_t *nonzero_flags, pb2_stream_t stream);
int pb2_copy_select(const pb2_bnd_table *table, const int32_t *data_flags, pb2_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * prolongation / restriction at fine-coarse boundaries
 * replaces ProResInfo (bnd_info.hpp:85-125) + refinement::Restrict / ProlongateShared
 * (src/prolong_restrict/prolong_restrict.cpp:37-79, pr_loops.hpp:113-155, pr_ops.hpp)
 * ------------------------------------------------------------------------------------- */
typedef struct pb2_prores_region {
  double *fine;   /* component 0 of the block's fine array */
  double *coarse; /* component 0 of the block's coarse buffer */
  int32_t s[3];   /* box in COARSE index space (i,j,k) */
  int32_t n[3];
  int32_t ncomp;
  int32_t fine_stride_j, fine_stride_k, fine_stride_c;
  int32_t coarse_stride_j, coarse_stride_k, coarse_stride_c;
  int32_t fine_is[3];   /* interior start of the fine index space  (ib.s, jb.s, kb.s) */
  int32_t coarse_is[3]; /* interior start of the coarse index space (cib.s, ...) */
  int32_t ndim;
  uint32_t status;
  double fine_xmin[3], fine_dx[3];     /* UniformCartesian xmin_, dx_ of the block */
  double coarse_xmin[3], coarse_dx[3]; /* and of its coarse coordinates */
  /* face / edge / node elements (the *_te entry points; zero for cell-centred regions): in
   * which directions (i, j, k) the region's element is displaced by half a cell
   * (TopologicalOffsetI/J/K, basic_types.hpp:195-203), and the same for the coarse CONTAINER
   * element of an internal prolongation.  `fine` / `coarse` point at the element's first slab
   * component; the box counts entries of the (container) element. */
  int32_t ftop[3];
  int32_t ctop[3];
} pb2_prores_region;

#define PB2_PROLONG_MINMOD 0             /* ProlongateSharedMinMod (default, metadata.hpp:337) */
#define PB2_PROLONG_LINEAR 1             /* ProlongateSharedLinear */
#define PB2_PROLONG_PIECEWISE_CONSTANT 2 /* ProlongatePiecewiseConstant */

int pb2_prores_table_create(pb2_bnd_table **table, const pb2_prores_region *regions,
                            int64_t n);
/* RestrictAverage over every region of the table (pr_ops.hpp:105-165) */
int pb2_restrict(const pb2_bnd_table *table, pb2_stream_t stream);
/* ProlongateSharedGeneral over every region of the table (pr_ops.hpp:167-280) */
int pb2_prolongate(const pb2_bnd_table *table, int op, pb2_stream_t stream);
/* The element forms for face / edge / node fields (regions carry ftop): RestrictAverage::Do<el>
 * and ProlongateSharedGeneral::Do<el> average / interpolate only along the directions the
 * element is centred in; pb2_prolongate_internal is ProlongateInternalAverage::Do<fel, cel>
 * (pr_ops.hpp:291-382) over boxes of coarse container elements (ctop) and must follow the
 * pb2_prolongate_te of the same exchange on the same stream.  The ownership mask of the
 * reference's loops (pr_loops.hpp:69-77 IsActive) is resolved by the caller into boxes. */
int pb2_restrict_te(const pb2_bnd_table *table, pb2_stream_t stream);
int pb2_prolongate_te(const pb2_bnd_table *table, int op, pb2_stream_t stream);
int pb2_prolongate_internal(const pb2_bnd_table *table, pb2_stream_t stream);
/* ProlongateInternalTothAndRoe::Do<fel, CC> (pr_ops.hpp:384-470) for face fields: regions are
 * boxes of coarse CELLS, `fine` points at element F1 of the field (all three are read), ncomp =
 * tensor components per element, ftop names the element whose internal faces are written.
 * Same ordering rule as pb2_prolongate_internal. */
int pb2_prolongate_toth_roe(const pb2_bnd_table *table, pb2_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * physical boundary conditions
 * replaces ApplyBoundaryConditionsOnCoarseOrFine (src/bvals/boundary_conditions.cpp:36-58) with
 * the generic outflow / reflect functions (boundary_conditions_generic.hpp:174-268): one launch fills the ghost slabs of every listed (block, face) — the slab
 * spans the ENTIRE extents of the other directions (mesh/domain.hpp:183-251).  Faces of
 * different directions must be applied in order x1, x2, x3 (one table per direction), because
 * the x2 slab reads x1 ghosts etc.
 * ------------------------------------------------------------------------------------- */
#define PB2_BC_OUTFLOW 0 /* ghost = value of the last interior cell along the normal */
#define PB2_BC_REFLECT 1 /* ghost = mirror image; sign flipped for the normal vector component */
typedef struct pb2_bc_region {
  double *var;        /* component 0 of the block's array (fine data, or the coarse buffer) */
  int32_t face;       /* BoundaryFace: 0 inner_x1, 1 outer_x1, 2 inner_x2, ... 5 outer_x3 */
  int32_t type;       /* PB2_BC_* */
  int32_t ncomp;
  int32_t n[3];       /* entire extents (i,j,k) of the array */
  int32_t is, ie;     /* interior bounds along the face normal */
  int32_t stride_c;   /* component stride in Reals (row stride n[0], plane stride n[0]*n[1]) */
  uint32_t flip_mask; /* reflect: bit c set => component c is the vector component along the
                         normal (Metadata::Vector, vector_component == DIR) and changes sign */
  /* face / edge / node elements: one region per topological element, `var` at the element's
   * first slab component, n = the index range the element USES (cells + 1 where it is
   * displaced), is / ie = its first / last interior entry along the normal (the boundary face
   * itself for a displaced element), and the strides of the (padded) array: */
  int32_t stride_j, stride_k; /* 0: n[0] and n[0] * n[1] */
} pb2_bc_region;
int pb2_bc_table_create(pb2_bnd_table **table, const pb2_bc_region *regions, int64_t n);
int pb2_apply_bcs(const pb2_bnd_table *table, pb2_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * flux correction at fine-coarse faces
 * replaces SendBoundBufs<flxcor_send> / SetBounds<flxcor_recv> (boundary_communication.cpp:
 * 454-461) for the face fluxes of cell-centred fields: region selection loop_utils.hpp:145-160,
 * index boxes bnd_info.cpp:105-252 with flux = true, RestrictAverage on faces pr_ops.hpp:105-165.
 * The reference restricts the fine flux into the flux field's coarse buffer, packs it, and the
 * coarser block unpacks it over its own face flux; here the restriction writes straight into
 * the coarser block's flux array (same device) or into the peer slab (other device, unpacked
 * there with pb2_unpack).
 * ------------------------------------------------------------------------------------- */
typedef struct pb2_flxcor_region {
  const double *fine; /* component 0 of the FINER block's flux array of direction `dir` */
  double *coarse;     /* component 0 of the COARSER block's flux array of the same direction,
                         or NULL: write to the slab given at launch at buf_off */
  int64_t buf_off;    /* slab offset in Reals; slab order is [comp][k][j][i] of the box */
  int32_t dir;        /* 0..2: normal of the shared face (the neighbour offset direction) */
  int32_t ndim;
  int32_t fs[3];      /* fine-array index (i,j,k) of the first fine face of the box */
  int32_t ds[3];      /* destination start (i,j,k) in the coarser block's array */
  int32_t n[3];       /* extent in COARSE faces; n[dir] == 1 */
  int32_t ncomp;
  int32_t fine_stride_j, fine_stride_k, fine_stride_c;
  int32_t coarse_stride_j, coarse_stride_k, coarse_stride_c;
  uint32_t status;    /* PB2_REGION_ALLOCATED */
  double area;        /* coords.Volume<F_dir>: face area of the finer block */
} pb2_flxcor_region;

int pb2_flxcor_table_create(pb2_bnd_table **table, const pb2_flxcor_region *regions, int64_t n);
/* One launch restricts and delivers every region of the table.  `slab` may be NULL when no
 * region has coarse == NULL. */
int pb2_flux_correct(const pb2_bnd_table *table, double *slab, pb2_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * dense per-stage updates: src/interface/update.hpp:43-91, update.cpp:63-86
 * ------------------------------------------------------------------------------------- */
/* z = w1*x + w2*y over n Reals (WeightedSumData, update.hpp:71-91) */
int pb2_weighted_sum(const double *x, const double *y, double w1, double w2, double *z,
                     int64_t n, pb2_stream_t stream);

/* Block-batched field geometry shared by the stencil entry points. */
typedef struct pb2_pack_geom {
  int32_t nblocks, ncomp, ndim;
  int32_t nx[3];   /* interior cells (i,j,k) */
  int32_t ng;      /* ghost width (0 in symmetry directions) */
  int64_t block_stride; /* Reals between blocks; component stride is ni*nj*nk */
  const double *dx;     /* device [nblocks][3] cell widths */
  /* optional compact list for the block-masked (`_blocks`) entry points: if block_list != NULL
   * they launch over its nlist entries (device int32 block indices, e.g. the blocks on which a
   * sparse field is allocated) instead of over all nblocks blocks; a mask, if given as well, is
   * still honoured.  Zero / NULL elsewhere. */
  const int32_t *block_list;
  int32_t nlist;
} pb2_pack_geom;

/* Block-masked forms for SPARSE fields: block b of the batch is processed only if
 * block_mask[b] != 0 (device int32 [nblocks]; NULL = every block), which is how the
 * reference's kernels skip unallocated variables (IsAllocated guards in update.hpp:83-85,
 * update.cpp:78, dc_inline.hpp:39).  x, y, z are whole slabs [nblocks][ncomp][nk][nj][ni]. */
int pb2_weighted_sum_blocks(const pb2_pack_geom *g, const double *x, const double *y, double w1,
                            double w2, double *z, const int32_t *block_mask,
                            pb2_stream_t stream);
int pb2_flux_divergence_blocks(const pb2_pack_geom *g, const double *const flux[3],
                               double *dudt, const int32_t *block_mask, pb2_stream_t stream);
int pb2_advection_fluxes_blocks(const pb2_pack_geom *g, const double *u, double *const flux[3],
                                const double v[3], const int32_t *block_mask,
                                pb2_stream_t stream);
/* Update::SparseDealloc's device pass (update.cpp:161-186): quiet[b] = 1 if every |x| of
 * block b (all components, ENTIRE extents) is <= threshold, else 0; blocks with
 * block_mask[b] == 0 are left untouched.  quiet: device int32 [nblocks]. */
int pb2_block_quiet_flags(const pb2_pack_geom *g, const double *u, double threshold,
                          const int32_t *block_mask, int32_t *quiet, pb2_stream_t stream);

/* Refinement tagging reduction (example/advection CheckRefinement, advection_package.cpp:
 * 239-273): minmax[2b] / minmax[2b+1] = minimum / maximum of block b over all components and
 * the ENTIRE extents; blocks with block_mask[b] == 0 are left untouched.  minmax: device. */
int pb2_block_minmax(const pb2_pack_geom *g, const double *u, const int32_t *block_mask,
                     double *minmax, pb2_stream_t stream);

/* The stock refinement criteria <parthenon/refinementN> method = derivative_order_1 / _2
 * (Refinement::FirstDerivative / SecondDerivative, amr_criteria/refinement_package.cpp:92-150):
 * maxd[b] = largest normalised first (order 1) or second (order 2) difference of component
 * `comp` over the interior cells of block b.  maxd: device [nblocks]. */
int pb2_block_derivative(const pb2_pack_geom *g, const double *u, int comp, int order,
                         double *maxd, pb2_stream_t stream);

/* z = w1*x + w2*y over the GHOST cells of every block only: what the reference's full-extent
 * WeightedSumData passes (AverageIndependentData / UpdateIndependentData, update.hpp:122-137)
 * do outside the interior.  A fused interior update plus this call equals the reference on
 * multilevel meshes, where fine ghosts facing a coarser block are not refreshed by the
 * stage's exchange.  x, y, z may alias. */
int pb2_weighted_sum_ghosts(const pb2_pack_geom *g, const double *x, const double *y, double w1,
                            double w2, double *z, pb2_stream_t stream);
/* the same for a subset of the batch: block_ids = device array of num_block_ids block indices
 * (only blocks with a coarser neighbour own ghosts that the exchange does not refresh) */
int pb2_weighted_sum_ghosts_blocks(const pb2_pack_geom *g, const double *x, const double *y,
                                   double w1, double w2, double *z, const int32_t *block_ids,
                                   int32_t num_block_ids, pb2_stream_t stream);

/* interior cells of a field <-> a packed buffer [block][comp][nx3][nx2][nx1] without ghosts
 * (the layout of an application's host arrays): the device side of uploading / reading back a
 * state through host buffers.  Ghosts are then filled by one exchange instead of crossing
 * PCIe (40^3 vs 32^3: half the bytes). */
int pb2_interior_scatter(const pb2_pack_geom *g, const double *packed, double *field,
                         pb2_stream_t stream);
int pb2_interior_gather(const pb2_pack_geom *g, const double *field, double *packed,
                        pb2_stream_t stream);
/* Uniform-mesh fast path of SetBounds<local> (boundary_communication.cpp:251-348 with the
 * same-level index boxes of bnd_info.cpp:205-212): every ghost cell of every block of a dense
 * field is pulled straight from the interior cell of the same-level neighbour that owns it.
 * No region table: the work decomposition is arithmetic over (block, component, ghost-shell
 * vector), the only indirection is `nbr` (device, [nblocks][27] block indices of the batch by
 * receiver-side offset index (ox1+1) + 3 (ox2+1) + 9 (ox3+1); < 0: no same-device neighbour in
 * that direction, those ghosts are left to pb2_unpack).  Periodic wrap onto the block itself
 * is allowed. */
int pb2_halo_copy_uniform(const pb2_pack_geom *g, double *field, const int32_t *nbr,
                          pb2_stream_t stream);
/* dudt = -div(F) (FluxDivergence + FluxDivHelper), interior cells */
int pb2_flux_divergence(const pb2_pack_geom *g, const double *const flux[3], double *dudt,
                        pb2_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * example/advection stencil: CalculateFluxes with a constant velocity
 * (example/advection/advection_package.cpp:540-646, DonorCellX1/2/3
 * src/reconstruct/dc_inline.hpp:31-71): flux[d](k,j,i) = v[d] * (v[d] > 0 ? u(cell - e_d) : u(cell))
 * on the d-faces of the interior, all blocks and components in one launch.  v is a HOST array.
 * ------------------------------------------------------------------------------------- */
int pb2_advection_fluxes(const pb2_pack_geom *g, const double *u, double *const flux[3],
                         const double v[3], pb2_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * benchmarks/burgers stencil: CalculateFluxes (burgers_package.cpp:202-404), fused with
 * FluxDivergence / AverageIndependentData / UpdateIndependentData / CalculateDerived /
 * EstimateTimestepMesh (burgers_driver.cpp:92-127)
 * ------------------------------------------------------------------------------------- */
#define PB2_RECON_WENO5 0
#define PB2_RECON_LINEAR 1
/* arithmetic mode: STRICT keeps the reference's expression order with FMA contraction off
 * (bit-identical to the reference's CPU build); FAST allows FMA contraction (<=1e-12 rel) */
#define PB2_MATH_STRICT 0
#define PB2_MATH_FAST 1

typedef struct pb2_burgers_args {
  pb2_pack_geom geom;
  int32_t recon; /* PB2_RECON_* */
  int32_t math;  /* PB2_MATH_* */
  const double *u;    /* stage input  mc0  [nblocks][ncomp][nk][nj][ni] */
  const double *base; /* base container (== u in stage 1) */
  double *out;        /* stage output mc1 */
  double *flux[3];    /* face fluxes, same extents as u (WithFluxes, metadata.cpp:185) */
  double *derived;    /* [nblocks][nk][nj][ni] or NULL */
  double *dt_min;     /* device scalar: min over cells of 1/sum(|u_d|/dx_d), or NULL.
                         Must be initialised to +huge by the caller. */
  double beta;        /* integrator->beta[stage-1] */
  double dt;
  /* restrict the launch to these blocks of the batch (device array of num_block_ids indices), or
   * NULL for all geom.nblocks blocks.  Lets a caller run the blocks that feed inter-GPU halos
   * first and overlap their exchange with the rest, and lets a multilevel mesh keep the
   * stored-flux path for the blocks that take part in flux correction only. */
  const int32_t *block_ids;
  int32_t num_block_ids;
  /* PB2_MATH_FAST, pb2_burgers_stage only.  Device table [geom.nblocks][27] of same-device
   * neighbour blocks (index (ox+1) + 3(oy+1) + 9(oz+1), -1 = none: physical boundary, another
   * device, another level) as pb2_halo_copy_uniform takes it, or NULL.  If given, stencil values
   * beyond a block face that has such a neighbour are read from the neighbour's INTERIOR cells
   * instead of the block's own ghost cells — the values SendBoundBufs<local> + SetBounds<local>
   * (boundary_communication.cpp:95-140, :273-334) would have copied there — so the ghost cells
   * of `u` across those faces need not be current: the caller may leave the same-device ghost
   * exchange out of the cycle and run it only when something else reads ghost cells.  Only the
   * six face entries are used (a direction sweep never reads edge or corner ghosts). */
  const int32_t *nbr_direct;
  /* PB2_MATH_FAST, pb2_burgers_stage only.  If progress != NULL, every thread block of the LAST
   * direction sweep that works on one of the first progress_blocks launched blocks (order of
   * block_ids) adds 1 to *progress when its results are in memory.  A consumer on another stream
   * (pb2_stream_wait_value) can then start on those blocks — the ones that feed inter-GPU halos,
   * listed first — while the same launch is still busy with the rest.  The counter reaches
   * pb2_burgers_progress_target(); the caller zeroes it before the stage.  2-D / 3-D only. */
  int32_t *progress;
  int32_t progress_blocks;
  /* PB2_MATH_FAST, pb2_burgers_stage only.  Which direction sweeps this call runs: bit 0 = x1,
   * bit 1 = x2, bit 2 = x3; 0 = all of them.  Lets a caller run the first sweep of a stage on the
   * blocks whose ghost cells are complete while the inter-GPU halo of the others is still in
   * flight, then the rest.  A stage must run its sweeps in the order x1, x2, x3 over every block. */
  int32_t sweeps;
} pb2_burgers_args;

/* fluxes only: writes args->flux[0..ndim-1] from args->u */
int pb2_burgers_calculate_fluxes(const pb2_burgers_args *args, pb2_stream_t stream);
/* out = (beta*u + (1-beta)*base) + beta*dt*(-div flux); derived; dt_min — interior cells */
int pb2_burgers_update(const pb2_burgers_args *args, pb2_stream_t stream);
/* One call per stage.  PB2_MATH_STRICT: the two calls above (needs args->flux).
 * PB2_MATH_FAST: three direction sweeps that keep every flux in registers and accumulate its
 * divergence straight into `out` (args->flux is ignored and may be NULL; out may alias base
 * but not u). */
int pb2_burgers_stage(const pb2_burgers_args *args, pb2_stream_t stream);
/* value *progress reaches once the first nblocks launched blocks of a stage are done */
int32_t pb2_burgers_progress_target(const pb2_pack_geom *g, int32_t nblocks);
/* stand-alone CalculateDerived (burgers_package.cpp:143-167) and/or EstimateTimestepMesh
 * (:170-200) over interior cells: derived [nblocks][nk][nj][ni] or NULL; dt_min device scalar
 * (initialised to +huge by the caller) or NULL */
int pb2_burgers_derived_dt(const pb2_pack_geom *g, const double *u, double *derived,
                           double *dt_min, pb2_stream_t stream);
/* 8 octant mass histories (MassHistory, burgers_package.cpp:406-439) -> host out[8] */
int pb2_burgers_history(const pb2_pack_geom *g, const double *u, const double *block_xmin,
                        const double mesh_xmin[3], const double mesh_xmax[3],
                        double out[8], pb2_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * inter-GPU halo slabs: replaces CommBuffer's MPI_Isend/Irecv
 * (src/utils/communication_buffer.hpp:209-406) with one grouped NCCL send/recv per peer
 * ------------------------------------------------------------------------------------- */
typedef struct pb2_comm pb2_comm;
#define PB2_NCCL_UNIQUE_ID_BYTES 128
int pb2_comm_unique_id(uint8_t id[PB2_NCCL_UNIQUE_ID_BYTES]);
int pb2_comm_create(pb2_comm **comm, int rank, int nranks,
                    const uint8_t id[PB2_NCCL_UNIQUE_ID_BYTES]);
int pb2_comm_destroy(pb2_comm *comm);
/* send_off/recv_off: [nranks+1] offsets in Reals into the slabs (peer p owns
 * [off[p], off[p+1])); empty ranges are skipped. */
int pb2_comm_exchange(pb2_comm *comm, const double *send_slab, const int64_t *send_off,
                      double *recv_slab, const int64_t *recv_off, pb2_stream_t stream);
/* in-place min all-reduce of one double (replaces MPI_Allreduce, driver.cpp:237) */
int pb2_comm_allreduce_min(pb2_comm *comm, double *dev_value, pb2_stream_t stream);
/* in-place sum all-reduce of n doubles (history reductions, MPI_Reduce in outputs/history.cpp) */
int pb2_comm_allreduce_sum(pb2_comm *comm, double *dev_values, int64_t n, pb2_stream_t stream);
/* all ranks: returns once everything every rank enqueued on `stream` before the call has
 * completed (an all-reduce followed by a stream synchronisation) */
int pb2_comm_barrier(pb2_comm *comm, pb2_stream_t stream);

/* ---- peer push: the inter-GPU halo without slabs, NCCL or an unpack ---------------------------
 * CommBuffer's Send / TryReceive (utils/communication_buffer.hpp:209-406) for GPUs of one
 * NVLink / NVSwitch domain: the SENDER's copy kernel stores each region straight into the ghost
 * cells of the receiving block in the peer's memory (the peer's field slab is mapped into this
 * process through CUDA IPC) and the last thread block of the launch raises an arrival flag in
 * the peer's memory; the receiver only waits for its flags.  One launch replaces pack ->
 * ncclSend / ncclRecv -> unpack, and the transfer overlaps whatever else runs.
 *
 * Flags: every rank owns int32 flags[2][nranks] in device memory (zeroed), mapped by its peers:
 *   flags[0][p] = s : rank p is ready for exchange number s — nothing on p reads the ghost cells
 *                     of exchange s - 1 any more, they may be overwritten (receiver -> sender);
 *   flags[1][p] = s : the halo of exchange s from rank p has landed (sender -> receiver).
 * Exchange numbers grow by one per exchange of a container, the same on every rank. */
typedef struct pb2_ipc_handle {
  uint8_t bytes[64]; /* cudaIpcMemHandle_t of the allocation that holds the pointer */
  int64_t offset;    /* of the pointer inside that allocation, in bytes */
} pb2_ipc_handle;
/* handle of device memory of THIS process (memory from pb2_malloc), to be sent to a peer */
int pb2_ipc_export(const void *ptr, pb2_ipc_handle *handle);
/* map a peer's memory into this process (peer access is enabled on demand); mappings are cached
 * per allocation and reference counted, pb2_ipc_close drops one reference */
int pb2_ipc_open(const pb2_ipc_handle *handle, void **ptr);
int pb2_ipc_close(const pb2_ipc_handle *handle);
/* Step 1 on `stream`: tell every peer "ready for exchange seq" (peer_flags[i][0 * nranks + me] =
 * seq) and wait until every peer said the same to us (my_flags[0 * nranks + peers[i]] >= seq).
 * peer_flags: DEVICE array of npeers pointers to the peers' flag arrays; peers: DEVICE array of
 * their ranks.  One small kernel. */
int pb2_peer_handshake(int32_t *const *peer_flags, const int32_t *my_flags, const int32_t *peers,
                       int npeers, int me, int nranks, int32_t seq, pb2_stream_t stream);
/* Step 2: pb2_copy whose destinations may lie in peer memory; when the last thread block of the
 * launch has stored its part (system-scope fence), peer_flags[i][1 * nranks + me] = seq for every
 * peer.  `counter`: one zeroed device int32 owned by the caller (left at zero). */
int pb2_copy_signal(const pb2_bnd_table *table, int32_t *counter, int32_t *const *peer_flags,
                    int npeers, int me, int nranks, int32_t seq, pb2_stream_t stream);
/* Step 2, copy-engine form: after the caller has enqueued its own transfers on `stream` (e.g.
 * pb2_memcpy_d2d from a packed send slab into the peers' receive slabs: DMA engines, no SMs),
 * one small kernel raises the arrival flags peer_flags[i][1 * nranks + me] = seq. */
int pb2_peer_signal(int32_t *const *peer_flags, int npeers, int me, int nranks, int32_t seq,
                    pb2_stream_t stream);
/* Step 3 (receiver) on `stream`: wait until my_flags[1 * nranks + peers[i]] >= seq for all i */
int pb2_peer_wait(const int32_t *my_flags, const int32_t *peers, int npeers, int nranks,
                  int32_t seq, pb2_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PARTHENON_B200_H_ */
