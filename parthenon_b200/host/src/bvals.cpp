// bvals.cpp — ghost-zone exchange: device tables and task functions (the index boxes and the
// channel plan they are built from live in bvals_plan.cpp).
// See pb2/bvals.hpp for the design and the reference files each piece replaces.
#include "pb2/bvals.hpp"

#ifdef _OPENMP
#include <omp.h>
#endif
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include <algorithm>
#include <tuple>

namespace parthenon {
// ---------------------------------------------------------------------------------------
// device tables
// ---------------------------------------------------------------------------------------
BvarsCache::~BvarsCache() { Clear(); }

void BvarsCache::Clear() {
  for (int c = 0; c < 2; ++c) {
    pb2_bnd_table_destroy(restrict_send[c]);
    pb2_bnd_table_destroy(restrict_set[c]);
    restrict_send[c] = restrict_set[c] = nullptr;
    pb2_bnd_table_destroy(te_restrict_send[c]);
    pb2_bnd_table_destroy(te_restrict_set[c]);
    pb2_bnd_table_destroy(te_internal[c]);
    pb2_bnd_table_destroy(te_toth_roe[c]);
    te_restrict_send[c] = te_restrict_set[c] = te_internal[c] = te_toth_roe[c] = nullptr;
    for (int o = 0; o < 3; ++o) {
      pb2_bnd_table_destroy(te_prolongate[c][o]);
      te_prolongate[c][o] = nullptr;
    }
    for (int o = 0; o < 3; ++o) {
      pb2_bnd_table_destroy(prolongate[c][o]);
      prolongate[c][o] = nullptr;
    }
  }
  pb2_bnd_table_destroy(push);
  push = nullptr;
  for (const pb2_ipc_handle &h : push_opened) pb2_ipc_close(&h);
  push_opened.clear();
  push_mode = false;
  pb2_bnd_table_destroy(copy_local);
  pb2_bnd_table_destroy(pack);
  pb2_bnd_table_destroy(unpack);
  copy_local = pack = unpack = nullptr;
  for (int cf = 0; cf < 2; ++cf)
    for (int d = 0; d < 3; ++d) {
      pb2_bnd_table_destroy(bc[cf][d]);
      bc[cf][d] = nullptr;
    }
  pb2_bnd_table_destroy(flxcor_local);
  pb2_bnd_table_destroy(flxcor_pack);
  pb2_bnd_table_destroy(flxcor_unpack);
  flxcor_local = flxcor_pack = flxcor_unpack = nullptr;
  for (pb2_bnd_table **t : {&teflx_restrict, &teflx_copy[0], &teflx_copy[1], &teflx_restrict_send,
                            &teflx_pack, &teflx_unpack[0], &teflx_unpack[1]}) {
    pb2_bnd_table_destroy(*t);
    *t = nullptr;
  }
  flxcor_built = false;
  built_generation = 0;
}

namespace {

MeshData<Real> *ContainerOf(MeshData<Real> *md, const MeshBlock *pmb) {
  if (pmb->partition == md->partition_id()) return md;
  return md->GetMeshPointer()->mesh_data.GetOrAdd(md->label(), pmb->partition).get();
}

pb2_prores_region MakeProRes(Variable &v, const MeshBlock *pmb, const IndexBox &box, int ndim) {
  pb2_prores_region r{};
  r.fine = v.data() + pmb->pack_index * v.block_stride;
  r.coarse = v.coarse() + pmb->pack_index * v.cblock_stride;
  for (int d = 0; d < 3; ++d) {
    r.s[d] = box.s[d];
    r.n[d] = box.n(d);
    r.fine_is[d] = pmb->cellbounds.Bounds(d, IndexDomain::interior).s;
    r.coarse_is[d] = pmb->c_cellbounds.Bounds(d, IndexDomain::interior).s;
  }
  r.ncomp = v.NumComponents();
  r.fine_stride_j = v.ni;
  r.fine_stride_k = v.ni * v.nj;
  r.fine_stride_c = static_cast<int32_t>(v.comp_stride);
  r.coarse_stride_j = v.cni;
  r.coarse_stride_k = v.cni * v.cnj;
  r.coarse_stride_c = static_cast<int32_t>(v.ccomp_stride);
  r.ndim = ndim;
  // ProResInfo::allocated (bnd_info.cpp:344): a sparse field that is not allocated on the
  // block is neither restricted nor prolongated (pr_loops.hpp:43-46)
  r.status = v.IsAllocated(pmb->pack_index) ? PB2_REGION_ALLOCATED : 0u;
  const UniformCartesian cc(pmb->coords, 2); // MeshRefinement::GetCoarseCoords
  for (int d = 0; d < 3; ++d) {
    r.fine_xmin[d] = pmb->coords.GetXmin()[d];
    r.fine_dx[d] = pmb->coords.Dx()[d];
    r.coarse_xmin[d] = cc.GetXmin()[d];
    r.coarse_dx[d] = cc.Dx()[d];
  }
  return r;
}

// one region per active sub-box of `mask` over `box`, for element number e (storage order) of
// a face / edge / node field; ctop != nullptr: boxes of a container element (internal
// prolongation)
void AddTeRegions(std::vector<pb2_prores_region> &out, Variable &v, const MeshBlock *pmb, int e,
                  TE fel, const TE *cel, const IndexBox &box, const std::array<bool, 27> &mask,
                  int ndim) {
  int n[3] = {box.n(0), box.n(1), box.n(2)};
  for (const IndexBox &rel : ActivePieces(n, mask)) {
    IndexBox sub;
    for (int d = 0; d < 3; ++d) {
      sub.s[d] = box.s[d] + rel.s[d];
      sub.e[d] = box.s[d] + rel.e[d];
    }
    pb2_prores_region r = MakeProRes(v, pmb, sub, ndim);
    const int nc = v.TensorComponents();
    r.fine += static_cast<int64_t>(e) * nc * v.comp_stride;
    r.coarse += static_cast<int64_t>(e) * nc * v.ccomp_stride;
    r.ncomp = nc;
    for (int d = 0; d < 3; ++d) {
      r.ftop[d] = TopologicalOffset(fel, d);
      r.ctop[d] = cel ? TopologicalOffset(*cel, d) : 0;
    }
    out.push_back(r);
  }
}

} // namespace

pb2_bc_region MakeBcRegion(Variable &v, const MeshBlock *pmb, int face, int type, bool coarse) {
  PARTHENON_REQUIRE(v.topological_type() == TopologicalType::Cell,
                    "MakeBcRegion describes cell-centred fields");
  const IndexShape &shape = coarse ? pmb->c_cellbounds : pmb->cellbounds;
  const int d = face / 2;
  pb2_bc_region r{};
  r.var = coarse ? v.coarse() + pmb->pack_index * v.cblock_stride
                 : v.data() + pmb->pack_index * v.block_stride;
  r.face = face;
  r.type = type;
  r.ncomp = v.NumComponents();
  r.n[0] = coarse ? v.cni : v.ni;
  r.n[1] = coarse ? v.cnj : v.nj;
  r.n[2] = coarse ? v.cnk : v.nk;
  const IndexRange b = shape.Bounds(d, IndexDomain::interior);
  r.is = b.s;
  r.ie = b.e;
  r.stride_c = static_cast<int32_t>(coarse ? v.ccomp_stride : v.comp_stride);
  // Metadata::Vector fields flip the component along the normal
  // (boundary_conditions_generic.hpp:225-228)
  r.flip_mask = v.IsSet(Metadata::Vector) && d < r.ncomp ? (1u << d) : 0u;
  return r;
}

namespace {
// Peer push tables of a MeshData (BvarsCache::push_mode): every rank publishes CUDA IPC handles of
// its FillGhost field slabs and of its flag array, maps those of the ranks it exchanges with and
// turns its send channels into copy regions whose destination is the receiving block's ghost box
// in the peer's slab.  Collective over the ranks (every rank rebuilds a container's cache at the
// same point of the program); if any rank cannot map a peer, all fall back to slabs + NCCL.
void BuildPeerPush(MeshData<Real> *md, BvarsCache &c, bool direct) {
  Mesh *pm = md->GetMeshPointer();
  pb2_stream_t st = md->stream();
  const int R = pm->nranks, me = pm->my_rank;
  const bool real = R > 1;
  // direct: destinations are the ghost boxes in the peers' field slabs; else the peers' receive
  // slabs (contiguous stores: full NVLink packets), which they unpack themselves
  const int nv = direct ? static_cast<int>(c.vars.size()) : 1;
  auto exported = [&](int iv) -> void * {
    return direct ? static_cast<void *>(c.vars[iv]->data()) : c.recv_slab.get();
  };
  if (!c.push_flags) {
    c.push_flags.Allocate(sizeof(int32_t) * 2 * R, st);
    c.push_counter.Allocate(sizeof(int32_t), st);
    c.push_seq = 0;
    PB2_CHECK(pb2_stream_sync(st));
  }
  // ranks this MeshData exchanges with (virtual ranks: the device itself)
  std::vector<int32_t> peers;
  if (real) {
    std::vector<char> is_peer(R, 0);
    for (const Channel &ch : c.plan->send) is_peer[ch.receiver_rank] = 1;
    for (const Channel &ch : c.plan->recv) is_peer[ch.sender_rank] = 1;
    for (int p = 0; p < R; ++p)
      if (is_peer[p] && p != me) peers.push_back(p);
  } else {
    peers.push_back(0);
  }
  // all-gather of (nv + 1) handles per rank, one byte per Real through the sum all-reduce
  constexpr int HB = static_cast<int>(sizeof(pb2_ipc_handle));
  std::vector<pb2_ipc_handle> mine(nv + 1), all(static_cast<size_t>(R) * (nv + 1));
  std::vector<std::vector<Real *>> peer_var(R, std::vector<Real *>(nv, nullptr));
  std::vector<int32_t *> peer_flags(R, nullptr);
  int failed = 0;
  if (real) {
    for (int iv = 0; iv < nv && !failed; ++iv)
      failed = pb2_ipc_export(exported(iv), &mine[iv]) != PB2_OK;
    if (!failed) failed = pb2_ipc_export(c.push_flags.get(), &mine[nv]) != PB2_OK;
    std::vector<Real> wire(static_cast<size_t>(R) * (nv + 1) * HB, 0.0);
    if (!failed) {
      const unsigned char *b = reinterpret_cast<const unsigned char *>(mine.data());
      for (size_t i = 0; i < static_cast<size_t>(nv + 1) * HB; ++i)
        wire[static_cast<size_t>(me) * (nv + 1) * HB + i] = b[i];
    }
    pm->AllReduceSum(wire);
    unsigned char *a = reinterpret_cast<unsigned char *>(all.data());
    for (size_t i = 0; i < wire.size(); ++i) a[i] = static_cast<unsigned char>(wire[i]);
    for (int p : peers) {
      for (int iv = 0; iv <= nv && !failed; ++iv) {
        const pb2_ipc_handle &h = all[static_cast<size_t>(p) * (nv + 1) + iv];
        void *ptr = nullptr;
        if (pb2_ipc_open(&h, &ptr) != PB2_OK) {
          failed = 1;
          break;
        }
        c.push_opened.push_back(h);
        if (iv < nv)
          peer_var[p][iv] = static_cast<Real *>(ptr);
        else
          peer_flags[p] = static_cast<int32_t *>(ptr);
      }
    }
    std::vector<Real> nfail(1, static_cast<Real>(failed));
    pm->AllReduceSum(nfail);
    if (nfail[0] > 0.0) {
      for (const pb2_ipc_handle &h : c.push_opened) pb2_ipc_close(&h);
      c.push_opened.clear();
      if (me == 0)
        std::fprintf(stderr, "[pb2] peer push unavailable (%s); inter-GPU halos use slabs + NCCL\n",
                     pb2_last_error());
      return;
    }
  } else {
    for (int iv = 0; iv < nv; ++iv) peer_var[0][iv] = static_cast<Real *>(exported(iv));
    peer_flags[0] = c.push_flags.get<int32_t>();
  }
  // slab variant: where our segment starts inside each peer's receive slab (its recv_off[me])
  std::vector<Real> peer_recv_off(static_cast<size_t>(R) * R, 0.0);
  if (!direct && real) {
    for (int p = 0; p < R; ++p) peer_recv_off[static_cast<size_t>(me) * R + p] = static_cast<Real>(c.plan->recv_off[p]);
    pm->AllReduceSum(peer_recv_off);
  }
  std::vector<pb2_copy_region> regs;
  regs.reserve(c.plan->send.size());
  for (const Channel &ch : c.plan->send) {
    PARTHENON_REQUIRE(!direct || (!ch.send_coarse && !ch.recv_coarse),
                      "direct peer push is for uniform meshes");
    const MeshBlock *sb = pm->block_list[pm->GetLid(ch.sender_gid)].get();
    Variable &v = *c.vars[ch.var];
    const int rrank = real ? ch.receiver_rank : 0;
    const int64_t rindex = real ? ch.receiver_gid - pm->nslist[ch.receiver_rank]
                                : pm->block_list[pm->GetLid(ch.receiver_gid)]->pack_index;
    pb2_copy_region r{};
    if (ch.send_coarse) {
      r.src = v.coarse() + sb->pack_index * v.cblock_stride + ch.comp0 * v.ccomp_stride;
      r.src_stride_j = v.cni;
      r.src_stride_k = v.cni * v.cnj;
      r.src_stride_c = static_cast<int32_t>(v.ccomp_stride);
    } else {
      r.src = v.data() + sb->pack_index * v.block_stride + ch.comp0 * v.comp_stride;
      r.src_stride_j = v.ni;
      r.src_stride_k = v.ni * v.nj;
      r.src_stride_c = static_cast<int32_t>(v.comp_stride);
    }
    if (!direct) {
      // the channel's place in the receiver's slab: same offset inside the peer segment on both
      // sides (BuildExchangePlan), the segment at the receiver's recv_off[this rank]; with
      // virtual ranks the send and the receive slab of the device have the same layout
      const int64_t seg0 = real ? c.plan->send_off[ch.receiver_rank] : 0;
      const int64_t base = real ? static_cast<int64_t>(peer_recv_off[static_cast<size_t>(ch.receiver_rank) * R + me]) : 0;
      r.dst = peer_var[rrank][0] + base + (ch.slab_off - seg0);
      for (int d = 0; d < 3; ++d) {
        r.ss[d] = ch.send_box.s[d];
        r.ds[d] = 0;
        r.n[d] = ch.send_box.n(d);
      }
      r.dst_stride_j = r.n[0];
      r.dst_stride_k = r.n[0] * r.n[1];
      r.dst_stride_c = r.n[0] * r.n[1] * r.n[2];
      r.ncomp = ch.ncomp;
      r.flag_slot = -1;
      r.status = PB2_REGION_ALLOCATED;
      r.default_value = v.metadata().GetDefaultValue();
      regs.push_back(r);
      continue;
    }
    r.dst = peer_var[rrank][ch.var] + rindex * v.block_stride + ch.comp0 * v.comp_stride;
    r.dst_stride_j = v.ni;
    r.dst_stride_k = v.ni * v.nj;
    r.dst_stride_c = static_cast<int32_t>(v.comp_stride);
    // the receiver's box (a send channel of the plan carries only the sender's): on a uniform mesh
    // it is CalcIndices of a block of the same shape that sees the sender at the mirrored offsets
    NeighborBlock rev;
    rev.loc = rev.origin_loc = sb->loc;
    for (int d = 0; d < 3; ++d) rev.offsets[d] = -((ch.offset_index / (d == 0 ? 1 : d == 1 ? 3 : 9)) % 3 - 1);
    const IndexBox rbox = CalcIndices(rev, sb, IndexRangeType::BoundaryExteriorRecv, false);
    for (int d = 0; d < 3; ++d) {
      PARTHENON_REQUIRE(rbox.n(d) == ch.send_box.n(d), "send/receive extents of a channel differ");
      r.ss[d] = ch.send_box.s[d];
      r.ds[d] = rbox.s[d];
      r.n[d] = rbox.n(d);
    }
    r.ncomp = ch.ncomp;
    r.flag_slot = -1;
    r.status = PB2_REGION_ALLOCATED;
    r.default_value = v.metadata().GetDefaultValue();
    regs.push_back(r);
  }
  pb2_bnd_table_destroy(c.push);
  c.push = nullptr;
  PB2_CHECK(pb2_copy_table_create(&c.push, regs.data(), static_cast<int64_t>(regs.size())));
  c.push_segments.clear();
  c.push_ce = !direct && pm->peer_push_mode == Mesh::PeerPush::ce;
  if (c.push_ce) {
    if (real) {
      for (int p : peers) {
        const int64_t n = c.plan->send_off[p + 1] - c.plan->send_off[p];
        if (n == 0) continue;
        c.push_segments.push_back(
            {peer_var[p][0] + static_cast<int64_t>(peer_recv_off[static_cast<size_t>(p) * R + me]),
             c.plan->send_off[p], n});
      }
    } else {
      c.push_segments.push_back({c.recv_slab.get<Real>(), 0, c.plan->send_elements});
    }
  }
  std::vector<int32_t *> pf;
  for (int p : peers) pf.push_back(peer_flags[p]);
  c.push_npeers = static_cast<int>(peers.size());
  c.push_peer_flags.Allocate(sizeof(int32_t *) * std::max<size_t>(pf.size(), 1), st);
  c.push_peer_ids.Allocate(sizeof(int32_t) * std::max<size_t>(peers.size(), 1), st);
  PB2_CHECK(pb2_memcpy_h2d(c.push_peer_flags.get(), pf.data(), sizeof(int32_t *) * pf.size(), st));
  PB2_CHECK(pb2_memcpy_h2d(c.push_peer_ids.get(), peers.data(), sizeof(int32_t) * peers.size(), st));
  PB2_CHECK(pb2_stream_sync(st));
  c.push_mode = true;
  c.push_direct = direct;
}
} // namespace

namespace {
void Rebuild(MeshData<Real> *md) {
  BvarsCache &c = md->bvars();
  c.Clear();
  Mesh *pm = md->GetMeshPointer();
  c.vars = md->GetVariablesByFlag({Metadata::FillGhost});
  std::vector<PlanVar> pvars;
  bool all_cell = true;
  for (Variable *v : c.vars) {
    pvars.push_back(PlanVar{v->TensorComponents(), v->topological_type()});
    all_cell = all_cell && v->topological_type() == TopologicalType::Cell;
  }
  // the plan is pure topology: the block list and the FillGhost fields of a MeshData never
  // change during its life (a remesh builds new MeshData), so a rebuild after a sparse
  // (de)allocation reuses it — only region statuses and tables change
  static const bool timing = std::getenv("PB2_TIME_REMESH") != nullptr;
  const auto tr0 = std::chrono::steady_clock::now();
  if (!c.plan_built) {
    std::string key = "p" + std::to_string(md->partition_id());
    for (const PlanVar &pv : pvars)
      key += "|" + std::to_string(pv.ncomp) + ":" + std::to_string(static_cast<int>(pv.tt));
    auto it = pm->plan_cache.find(key);
    if (it == pm->plan_cache.end()) {
      auto sp = std::make_shared<ExchangePlan>(BuildExchangePlan(pm, md->GetBlockList(), pvars));
      it = pm->plan_cache.emplace(key, std::static_pointer_cast<const void>(sp)).first;
    }
    // shared, not copied: a plan of a few thousand blocks is tens of MB
    c.plan = std::static_pointer_cast<const ExchangePlan>(it->second);
    c.plan_built = true;
  }
  const auto tr1 = std::chrono::steady_clock::now();
  const bool slabs = c.plan->send_elements > 0 || c.plan->recv_elements > 0;
  PARTHENON_REQUIRE(!slabs || pm->DefaultNumPartitions() == 1,
                    "inter-device halos need one MeshData per rank (parthenon/mesh/pack_size=-1)");
  // sparse fields: allocation-aware exchange (null messages, allocate-on-receive), built for
  // same-device channels of a uniform mesh with one MeshData per device
  c.sparse = false;
  for (Variable *v : c.vars) c.sparse = c.sparse || (v->metadata().IsSparse() && pm->sparse_config.enabled);
  if (c.sparse) {
    // statically refined meshes: allocation-aware restriction / prolongation / flux correction
    // on same-device channels (tests/test_late_round1_gpu.py, bit-exact against the reference's
    // dumps); remeshing sparse fields and shipping them between devices across levels are not built
    PARTHENON_REQUIRE(!pm->multilevel || !slabs,
                      "sparse fields on multilevel meshes: one device only in this build");
    PARTHENON_REQUIRE(pm->DefaultNumPartitions() == 1,
                      "sparse fields need one MeshData per rank (parthenon/mesh/pack_size=-1)");
  }

  // fused local channels: receiver ghost box <- sender interior box.  One region per channel,
  // built on all host threads: the variables of every partition a sender lives in are looked up
  // (and their lazily allocated arrays touched) once, before the loop
  const auto tr2 = std::chrono::steady_clock::now();
  const std::vector<Channel> &local_chs = c.plan->local;
  const int64_t nlocal = static_cast<int64_t>(local_chs.size());
  const int nvars = static_cast<int>(c.vars.size());
  std::vector<std::vector<Variable *>> part_vars(static_cast<size_t>(pm->DefaultNumPartitions()));
  auto touch = [&](Variable *v) {
    v->data();
    if (pm->multilevel) v->coarse();
  };
  auto fill_partition = [&](int p) {
    if (!part_vars[p].empty() || nvars == 0) return;
    MeshData<Real> *pmd =
        p == md->partition_id() ? md : pm->mesh_data.GetOrAdd(md->label(), p).get();
    for (Variable *v : c.vars) {
      part_vars[p].push_back(p == md->partition_id() ? v : &pmd->Get(v->label()));
      touch(part_vars[p].back());
    }
  };
  fill_partition(md->partition_id());
  if (part_vars.size() > 1)
    for (const Channel &ch : local_chs)
      fill_partition(pm->block_list[pm->GetLid(ch.sender_gid)]->partition);
  std::vector<pb2_copy_region> copies(static_cast<size_t>(nlocal));
  std::vector<char> keep(static_cast<size_t>(nlocal), 1);
#pragma omp parallel for schedule(static) if (nlocal > 4096)
  for (int64_t i = 0; i < nlocal; ++i) {
    const Channel &ch = local_chs[i];
    const MeshBlock *rb = pm->block_list[pm->GetLid(ch.receiver_gid)].get();
    const MeshBlock *sb = pm->block_list[pm->GetLid(ch.sender_gid)].get();
    Variable &rv = *c.vars[ch.var];
    Variable &sv = *part_vars[sb->partition][ch.var];
    pb2_copy_region r{};
    if (ch.send_coarse) {
      r.src = sv.coarse() + sb->pack_index * sv.cblock_stride + ch.comp0 * sv.ccomp_stride;
      r.src_stride_j = sv.cni;
      r.src_stride_k = sv.cni * sv.cnj;
      r.src_stride_c = static_cast<int32_t>(sv.ccomp_stride);
    } else {
      r.src = sv.data() + sb->pack_index * sv.block_stride + ch.comp0 * sv.comp_stride;
      r.src_stride_j = sv.ni;
      r.src_stride_k = sv.ni * sv.nj;
      r.src_stride_c = static_cast<int32_t>(sv.comp_stride);
    }
    if (ch.recv_coarse) {
      r.dst = rv.coarse() + rb->pack_index * rv.cblock_stride + ch.comp0 * rv.ccomp_stride;
      r.dst_stride_j = rv.cni;
      r.dst_stride_k = rv.cni * rv.cnj;
      r.dst_stride_c = static_cast<int32_t>(rv.ccomp_stride);
    } else {
      r.dst = rv.data() + rb->pack_index * rv.block_stride + ch.comp0 * rv.comp_stride;
      r.dst_stride_j = rv.ni;
      r.dst_stride_k = rv.ni * rv.nj;
      r.dst_stride_c = static_cast<int32_t>(rv.comp_stride);
    }
    for (int d = 0; d < 3; ++d) {
      r.ss[d] = ch.send_box.s[d];
      r.ds[d] = ch.recv_box.s[d];
      r.n[d] = ch.recv_box.n(d);
    }
    r.ncomp = ch.ncomp;
    r.flag_slot = -1;
    r.status = PB2_REGION_ALLOCATED;
    r.threshold = 0.0;
    r.default_value = rv.metadata().GetDefaultValue();
    if (c.sparse && rv.metadata().IsSparse()) {
      // BndInfo::allocated of the sender / the receiver (bnd_info.cpp:277, :290-292).  Flags are
      // indexed by CHANNEL; a channel with neither side allocated can neither carry data nor
      // need a default fill, so it gets no region at all (its flag stays 0): with a few percent
      // of the fields allocated the launches shrink by the same factor
      const bool src_alloc = sv.IsAllocated(sb->pack_index), dst_alloc = rv.IsAllocated(rb->pack_index);
      if (!src_alloc && !dst_alloc) keep[i] = 0;
      r.flag_slot = static_cast<int32_t>(i);
      r.status = (src_alloc ? PB2_REGION_ALLOCATED : 0u) |
                 (dst_alloc ? 0u : PB2_REGION_DST_UNALLOCATED);
      r.threshold = rv.metadata().GetAllocationThreshold();
    }
    copies[i] = r;
  }
  if (c.sparse) { // drop the channels without a region, order kept
    size_t w = 0;
    for (int64_t i = 0; i < nlocal; ++i)
      if (keep[i]) copies[w++] = copies[i];
    copies.resize(w);
  }
  const auto tr3 = std::chrono::steady_clock::now();
  if (c.sparse && !c.sparse_flags) {
    c.sparse_flags.Allocate(sizeof(int32_t) * std::max<size_t>(c.plan->local.size(), 1), md->stream());
    c.sparse_flags_h.assign(c.plan->local.size(), 0);
  }
  PB2_CHECK(pb2_copy_table_create(&c.copy_local, copies.data(), static_cast<int64_t>(copies.size())));
  const auto tr4 = std::chrono::steady_clock::now();

  // uniform fast path: all local channels are same-level boxes between blocks of this batch
  c.uniform_halo = !pm->multilevel && pm->DefaultNumPartitions() == 1 && !c.plan->local.empty() &&
                   !pm->table_halo && all_cell;
  for (Variable *v : c.vars) c.uniform_halo = c.uniform_halo && !v->metadata().IsSparse();
  if (c.uniform_halo) {
    std::vector<int32_t> nbr(static_cast<size_t>(md->NumBlocks()) * 27, -1);
    for (const Channel &ch : c.plan->local) {
      if (ch.var != 0) continue; // the topology is the same for every field
      const MeshBlock *rb = pm->block_list[pm->GetLid(ch.receiver_gid)].get();
      const MeshBlock *sb = pm->block_list[pm->GetLid(ch.sender_gid)].get();
      // offset_index is the sender's view of the receiver; the receiver sees the mirror image
      nbr[static_cast<size_t>(rb->pack_index) * 27 + (26 - ch.offset_index)] = sb->pack_index;
    }
    c.halo_nbr.Allocate(sizeof(int32_t) * nbr.size(), md->stream());
    PB2_CHECK(pb2_memcpy_h2d(c.halo_nbr.get(), nbr.data(), sizeof(int32_t) * nbr.size(), md->stream()));
    PB2_CHECK(pb2_stream_sync(md->stream()));
  }

  // slab channels
  auto bnd = [&](const Channel &ch, bool send) {
    const int gid = send ? ch.sender_gid : ch.receiver_gid;
    const MeshBlock *pmb = pm->block_list[pm->GetLid(gid)].get();
    Variable &v = *c.vars[ch.var];
    const bool coarse = send ? ch.send_coarse : ch.recv_coarse;
    const IndexBox &box = send ? ch.send_box : ch.recv_box;
    pb2_bnd_region r{};
    if (coarse) {
      r.var = v.coarse() + pmb->pack_index * v.cblock_stride + ch.comp0 * v.ccomp_stride;
      r.stride_j = v.cni;
      r.stride_k = v.cni * v.cnj;
      r.stride_c = static_cast<int32_t>(v.ccomp_stride);
    } else {
      r.var = v.data() + pmb->pack_index * v.block_stride + ch.comp0 * v.comp_stride;
      r.stride_j = v.ni;
      r.stride_k = v.ni * v.nj;
      r.stride_c = static_cast<int32_t>(v.comp_stride);
    }
    r.buf_off = ch.slab_off;
    for (int d = 0; d < 3; ++d) {
      r.s[d] = box.s[d];
      r.n[d] = box.n(d);
    }
    r.ncomp = ch.ncomp;
    r.flag_slot = -1;
    r.status = PB2_REGION_ALLOCATED | (send ? 0u : PB2_REGION_BUF_ALLOCATED);
    r.value = send ? v.metadata().GetAllocationThreshold() : v.metadata().GetDefaultValue();
    if (!send && ch.transformed) {
      r.lcoord_on = 1;
      for (int d = 0; d < 3; ++d) {
        r.lcoord_dir[d] = std::abs(ch.lcoord_trans.dir_connection[d]);
        r.lcoord_flip[d] = ch.lcoord_trans.dir_flip[d] ? 1 : 0;
      }
      r.lcoord_ncell = ch.ncell;
      r.fac = 1.0; // cell-centred fields carry no sign (logical_coordinate_transformation.hpp:66-77)
    }
    return r;
  };
  std::vector<pb2_bnd_region> packs, unpacks;
  for (const Channel &ch : c.plan->send) packs.push_back(bnd(ch, true));
  for (const Channel &ch : c.plan->recv) unpacks.push_back(bnd(ch, false));
  if (c.sparse) {
    // every inter-device channel gets a flag slot; BndInfo::allocated is this side's own field
    // (bnd_info.cpp:277): an unallocated sender packs nothing and its flag stays 0 (a null
    // message), an unallocated receiver is skipped by the unpack
    for (size_t i = 0; i < packs.size(); ++i) {
      const Channel &ch = c.plan->send[i];
      const MeshBlock *pmb = pm->block_list[pm->GetLid(ch.sender_gid)].get();
      packs[i].flag_slot = static_cast<int32_t>(i);
      packs[i].status = c.vars[ch.var]->IsAllocated(pmb->pack_index) ? PB2_REGION_ALLOCATED : 0u;
    }
    for (size_t i = 0; i < unpacks.size(); ++i) {
      const Channel &ch = c.plan->recv[i];
      const MeshBlock *pmb = pm->block_list[pm->GetLid(ch.receiver_gid)].get();
      unpacks[i].flag_slot = static_cast<int32_t>(i);
      unpacks[i].status = c.vars[ch.var]->IsAllocated(pmb->pack_index) ? PB2_REGION_ALLOCATED : 0u;
    }
    auto counts = [&](const std::vector<Channel> &chs, bool send, std::vector<int64_t> &off) {
      const int V = pm->virtual_ranks > 1 ? pm->virtual_ranks : 1;
      off.assign(c.plan->npeers + 1, 0);
      for (const Channel &ch : chs) {
        const int seg = V > 1 ? ch.sender_vrank * V + ch.receiver_vrank
                              : (send ? ch.receiver_rank : ch.sender_rank);
        off[seg + 1]++;
      }
      for (int p = 0; p < c.plan->npeers; ++p) off[p + 1] += off[p];
    };
    counts(c.plan->send, true, c.send_flag_off);
    counts(c.plan->recv, false, c.recv_flag_off);
    auto ensure = [&](DeviceBuffer &b, size_t bytes) {
      if (b.bytes() != std::max<size_t>(bytes, 8)) b.Allocate(std::max<size_t>(bytes, 8), md->stream());
    };
    ensure(c.send_flags, sizeof(int32_t) * packs.size());
    ensure(c.recv_flags, sizeof(int32_t) * unpacks.size());
    ensure(c.send_flag_slab, sizeof(Real) * packs.size());
    ensure(c.recv_flag_slab, sizeof(Real) * unpacks.size());
    c.send_flags_h.assign(packs.size(), 0);
    c.recv_flags_h.assign(unpacks.size(), 0);
  }
  PB2_CHECK(pb2_bnd_table_create(&c.pack, packs.data(), static_cast<int64_t>(packs.size())));
  PB2_CHECK(pb2_bnd_table_create(&c.unpack, unpacks.data(), static_cast<int64_t>(unpacks.size())));
  // slabs of an unchanged size survive a rebuild: allocate-on-receive rebuilds the tables
  // between the arrival of a slab and its unpack
  const bool want_push = pm->peer_push && slabs && !pm->adaptive && !c.sparse &&
                         (pm->nranks > 1 || pm->virtual_ranks > 1);
  // direct variant — cell-centred fields of uniform meshes only: send boxes are interior cells
  // and receive boxes ghost cells, so stores of one channel never touch what another channel
  // reads (shared faces / edges / nodes are both and need every pack to precede every unpack)
  const bool want_direct = want_push && pm->peer_push_mode == Mesh::PeerPush::direct && all_cell && !pm->multilevel;
  if (want_direct) BuildPeerPush(md, c, true);
  if (!c.push_mode && c.plan->send_elements > 0 &&
      c.send_slab.bytes() != sizeof(Real) * static_cast<size_t>(c.plan->send_elements))
    c.send_slab.Allocate(sizeof(Real) * static_cast<size_t>(c.plan->send_elements), md->stream());
  if (!c.push_mode && c.plan->recv_elements > 0 &&
      c.recv_slab.bytes() != sizeof(Real) * static_cast<size_t>(c.plan->recv_elements))
    c.recv_slab.Allocate(sizeof(Real) * static_cast<size_t>(c.plan->recv_elements), md->stream());
  // slab variant: the pack writes straight into the peers' receive slabs (no send slab needed,
  // it is kept for the fallback), the unpack stays with the receiver
  if (want_push && !c.push_mode && c.plan->recv_elements > 0) BuildPeerPush(md, c, false);

  const auto tr5 = std::chrono::steady_clock::now();
  // restriction / prolongation regions (ProResInfo::GetSend / GetSet, bnd_info.cpp:387-448),
  // split by whether the neighbour is local so the local / nonlocal task split still works
  if (pm->multilevel) {
    struct ProResLists {
      std::vector<pb2_prores_region> rsend[2], rset[2], pro[2][3];
      std::vector<pb2_prores_region> te_rsend[2], te_rset[2], te_pro[2][3], te_int[2], te_tr[2];
    };
    std::array<bool, 27> all_true;
    all_true.fill(true);
    // containers of the internal prolongation in the order the reference visits them
    // (pr_loops.hpp:82-108)
    const TE containers[8] = {TE::NN, TE::E3, TE::E2, TE::E1, TE::F1, TE::F2, TE::F3, TE::CC};
    auto is_submanifold = [](TE f, TE cnt) { // basic_types.hpp:207-232
      int more = 0;
      for (int d = 0; d < 3; ++d) {
        if (TopologicalOffset(cnt, d) && !TopologicalOffset(f, d)) return false;
        more += TopologicalOffset(f, d) && !TopologicalOffset(cnt, d);
      }
      return more > 0;
    };
    auto collect = [&](const std::shared_ptr<MeshBlock> &pmb, ProResLists &L) {
      const int my_vr = pm->VirtualRankOf(pmb->gid);
      bool restricted = false;
      if (pmb->loc.level > 0)
        for (auto &nb : pmb->neighbors)
          restricted = restricted || nb.origin_loc.level == pmb->loc.level - 1;
      for (auto &nb : pmb->neighbors) {
        // (same rule as BuildExchangePlan: neighbours in differently oriented trees are served
        // by the slab path, their restriction / prolongation regions follow its unpack)
        const bool local =
            nb.rank == pm->my_rank && pm->VirtualRankOf(nb.gid) == my_vr && !nb.transformed;
        const int cls = local ? 0 : 1;
        for (Variable *v : c.vars) {
          if (v->topological_type() != TopologicalType::Cell) {
            const std::vector<TE> els = GetTopologicalElements(v->topological_type());
            const int op = v->metadata().ProlongationOp();
            for (size_t e = 0; e < els.size(); ++e) {
              const int ei = static_cast<int>(e);
              if (nb.origin_loc.level < pmb->loc.level) {
                AddTeRegions(L.te_rsend[cls], *v, pmb.get(), ei, els[e], nullptr,
                             CalcIndicesTE(nb, pmb.get(), els[e],
                                           IndexRangeType::BoundaryInteriorSend, true),
                             all_true, pm->ndim);
                AddTeRegions(L.te_pro[cls][op], *v, pmb.get(), ei, els[e], nullptr,
                             CalcIndicesTE(nb, pmb.get(), els[e],
                                           IndexRangeType::BoundaryExteriorRecv, true),
                             RecvMask(pm, nb, pmb.get(), els[e]), pm->ndim);
                if (v->metadata().InternalProlongationOp() == 1) {
                  // Toth & Roe: coarse cells only, all three face elements are read — the
                  // regions point at element F1 (pr_ops.hpp:390-393)
                  PARTHENON_REQUIRE(v->topological_type() == TopologicalType::Face,
                                    "ProlongateInternalTothAndRoe is defined for face fields");
                  const TE cc = TE::CC;
                  const size_t first = L.te_tr[cls].size();
                  AddTeRegions(L.te_tr[cls], *v, pmb.get(), ei, els[e], &cc,
                               CalcIndicesTE(nb, pmb.get(), cc,
                                             IndexRangeType::BoundaryExteriorRecv, true),
                               RecvMask(pm, nb, pmb.get(), cc), pm->ndim);
                  for (size_t q = first; q < L.te_tr[cls].size(); ++q)
                    L.te_tr[cls][q].fine -=
                        static_cast<int64_t>(ei) * v->TensorComponents() * v->comp_stride;
                } else {
                  for (const TE &cel : containers)
                    if (is_submanifold(els[e], cel))
                      AddTeRegions(L.te_int[cls], *v, pmb.get(), ei, els[e], &cel,
                                   CalcIndicesTE(nb, pmb.get(), cel,
                                                 IndexRangeType::BoundaryExteriorRecv, true),
                                   RecvMask(pm, nb, pmb.get(), cel), pm->ndim);
                }
              } else if (restricted) {
                AddTeRegions(L.te_rset[cls], *v, pmb.get(), ei, els[e], nullptr,
                             CalcIndicesTE(nb, pmb.get(), els[e],
                                           IndexRangeType::BoundaryExteriorRecv, true),
                             RecvMask(pm, nb, pmb.get(), els[e]), pm->ndim);
              }
            }
            continue;
          }
          if (nb.origin_loc.level < pmb->loc.level) {
            L.rsend[cls].push_back(MakeProRes(
                *v, pmb.get(),
                CalcIndices(nb, pmb.get(), IndexRangeType::BoundaryInteriorSend, true), pm->ndim));
            L.pro[cls][v->metadata().ProlongationOp()].push_back(MakeProRes(
                *v, pmb.get(),
                CalcIndices(nb, pmb.get(), IndexRangeType::BoundaryExteriorRecv, true), pm->ndim));
          } else if (restricted) {
            L.rset[cls].push_back(MakeProRes(
                *v, pmb.get(),
                CalcIndices(nb, pmb.get(), IndexRangeType::BoundaryExteriorRecv, true), pm->ndim));
          }
        }
      }
    };
    // blocks are independent: cell-centred fields are collected on all host threads, each thread
    // a contiguous range of blocks, and the lists are joined in thread order (the serial order).
    // Face / edge / node fields consult the lazily filled ownership cache and stay serial.
    const BlockList_t &blist = md->GetBlockList();
    const int nblk_pr = static_cast<int>(blist.size());
    int nthreads_pr = 1;
#ifdef _OPENMP
    if (all_cell && nblk_pr > 256) nthreads_pr = omp_get_max_threads();
#endif
    std::vector<ProResLists> lists(static_cast<size_t>(nthreads_pr));
    std::string pr_failure;
#pragma omp parallel num_threads(nthreads_pr) if (nthreads_pr > 1)
    {
#ifdef _OPENMP
      const int tid = nthreads_pr > 1 ? omp_get_thread_num() : 0;
      const int nth = nthreads_pr > 1 ? omp_get_num_threads() : 1;
#else
      const int tid = 0, nth = 1;
#endif
      const int lo = static_cast<int>(static_cast<int64_t>(nblk_pr) * tid / nth);
      const int hi = static_cast<int>(static_cast<int64_t>(nblk_pr) * (tid + 1) / nth);
      try {
        for (int ib = lo; ib < hi; ++ib) collect(blist[ib], lists[tid]);
      } catch (const std::exception &e) {
#pragma omp critical
        pr_failure = e.what();
      }
    }
    PARTHENON_REQUIRE(pr_failure.empty(), pr_failure);
    ProResLists &L0 = lists[0];
    auto join = [&](auto member) {
      for (int t = 1; t < nthreads_pr; ++t) {
        auto &src = member(lists[t]);
        auto &dst = member(L0);
        dst.insert(dst.end(), src.begin(), src.end());
      }
    };
    for (int cls = 0; cls < 2; ++cls) {
      join([cls](ProResLists &l) -> std::vector<pb2_prores_region> & { return l.rsend[cls]; });
      join([cls](ProResLists &l) -> std::vector<pb2_prores_region> & { return l.rset[cls]; });
      for (int o = 0; o < 3; ++o)
        join([cls, o](ProResLists &l) -> std::vector<pb2_prores_region> & { return l.pro[cls][o]; });
    }
    auto &rsend = L0.rsend;
    auto &rset = L0.rset;
    auto &pro = L0.pro;
    auto &te_rsend = L0.te_rsend;
    auto &te_rset = L0.te_rset;
    auto &te_pro = L0.te_pro;
    auto &te_int = L0.te_int;
    auto &te_tr = L0.te_tr;
    for (int cls = 0; cls < 2; ++cls) {
      PB2_CHECK(pb2_prores_table_create(&c.restrict_send[cls], rsend[cls].data(),
                                        static_cast<int64_t>(rsend[cls].size())));
      PB2_CHECK(pb2_prores_table_create(&c.restrict_set[cls], rset[cls].data(),
                                        static_cast<int64_t>(rset[cls].size())));
      for (int o = 0; o < 3; ++o)
        PB2_CHECK(pb2_prores_table_create(&c.prolongate[cls][o], pro[cls][o].data(),
                                          static_cast<int64_t>(pro[cls][o].size())));
      PB2_CHECK(pb2_prores_table_create(&c.te_restrict_send[cls], te_rsend[cls].data(),
                                        static_cast<int64_t>(te_rsend[cls].size())));
      PB2_CHECK(pb2_prores_table_create(&c.te_restrict_set[cls], te_rset[cls].data(),
                                        static_cast<int64_t>(te_rset[cls].size())));
      PB2_CHECK(pb2_prores_table_create(&c.te_internal[cls], te_int[cls].data(),
                                        static_cast<int64_t>(te_int[cls].size())));
      PB2_CHECK(pb2_prores_table_create(&c.te_toth_roe[cls], te_tr[cls].data(),
                                        static_cast<int64_t>(te_tr[cls].size())));
      for (int o = 0; o < 3; ++o)
        PB2_CHECK(pb2_prores_table_create(&c.te_prolongate[cls][o], te_pro[cls][o].data(),
                                          static_cast<int64_t>(te_pro[cls][o].size())));
    }
  }
  const auto tr6 = std::chrono::steady_clock::now();
  // physical boundary conditions: blocks on a non-periodic mesh face
  {
    std::vector<pb2_bc_region> regs[2][3];
    for (auto &pmb : md->GetBlockList())
      for (int face = 0; face < 2 * pm->ndim; ++face) {
        const BoundaryFlag flag = pmb->boundary_flag[face];
        if (flag != BoundaryFlag::outflow && flag != BoundaryFlag::reflect) continue;
        const int d = face / 2;
        for (Variable *v : c.vars) {
          if (!v->IsAllocated(pmb->pack_index)) continue;
          if (v->topological_type() != TopologicalType::Cell) {
            // GenericBC per topological element (boundary_conditions_generic.hpp:268-273)
            PARTHENON_REQUIRE(!v->IsSet(Metadata::Vector),
                              "Metadata::Vector on non-cell-centred fields is not supported");
            const std::vector<TE> els = GetTopologicalElements(v->topological_type());
            const int nc = v->TensorComponents();
            for (int cf = 0; cf < (pm->multilevel ? 2 : 1); ++cf) {
              const IndexShape &shape = cf ? pmb->c_cellbounds : pmb->cellbounds;
              for (size_t e = 0; e < els.size(); ++e) {
                pb2_bc_region r{};
                const int64_t cs = cf ? v->ccomp_stride : v->comp_stride;
                r.var = (cf ? v->coarse() + pmb->pack_index * v->cblock_stride
                            : v->data() + pmb->pack_index * v->block_stride) +
                        static_cast<int64_t>(e) * nc * cs;
                r.face = face;
                r.type = flag == BoundaryFlag::outflow ? PB2_BC_OUTFLOW : PB2_BC_REFLECT;
                r.ncomp = nc;
                for (int q = 0; q < 3; ++q) r.n[q] = shape.Bounds(q, IndexDomain::entire, els[e]).e + 1;
                const IndexRange b = shape.Bounds(d, IndexDomain::interior, els[e]);
                r.is = b.s;
                r.ie = b.e;
                r.stride_c = static_cast<int32_t>(cs);
                r.stride_j = cf ? v->cni : v->ni;
                r.stride_k = cf ? v->cni * v->cnj : v->ni * v->nj;
                regs[cf][d].push_back(r);
                c.has_bcs = true;
              }
            }
            continue;
          }
          for (int cf = 0; cf < (pm->multilevel ? 2 : 1); ++cf) {
            const pb2_bc_region r = MakeBcRegion(
                *v, pmb.get(), face,
                flag == BoundaryFlag::outflow ? PB2_BC_OUTFLOW : PB2_BC_REFLECT, cf != 0);
            regs[cf][d].push_back(r);
            c.has_bcs = true;
          }
        }
      }
    for (int cf = 0; cf < 2; ++cf)
      for (int d = 0; d < 3; ++d)
        PB2_CHECK(pb2_bc_table_create(&c.bc[cf][d], regs[cf][d].data(),
                                      static_cast<int64_t>(regs[cf][d].size())));
  }
  const auto tr7 = std::chrono::steady_clock::now();
  // boundary / interior split of the batch for comm-compute overlap
  {
    std::vector<int32_t> bnd, inr;
    for (auto &pmb : md->GetBlockList()) {
      const int my_vr = pm->VirtualRankOf(pmb->gid);
      bool nonlocal = false;
      for (auto &nb : pmb->neighbors)
        nonlocal = nonlocal || nb.rank != pm->my_rank || pm->VirtualRankOf(nb.gid) != my_vr ||
                   nb.transformed;
      (nonlocal ? bnd : inr).push_back(pmb->pack_index);
    }
    c.n_boundary = static_cast<int>(bnd.size());
    c.n_interior = static_cast<int>(inr.size());
    auto upload = [&](DeviceBuffer &buf, const std::vector<int32_t> &v) {
      if (v.empty()) return;
      buf.Allocate(sizeof(int32_t) * v.size(), md->stream());
      PB2_CHECK(pb2_memcpy_h2d(buf.get(), v.data(), sizeof(int32_t) * v.size(), md->stream()));
      PB2_CHECK(pb2_stream_sync(md->stream()));
    };
    upload(c.ids_boundary, bnd);
    upload(c.ids_interior, inr);
    std::vector<int32_t> ordered(bnd);
    ordered.insert(ordered.end(), inr.begin(), inr.end());
    upload(c.ids_ordered, ordered);
    if (!c.progress) c.progress.Allocate(sizeof(int32_t), md->stream());
    // classes for the multilevel stage
    std::vector<int32_t> flx, plain, stale;
    for (auto &pmb : md->GetBlockList()) {
      bool face_other_level = false, coarser = false;
      for (auto &nb : pmb->neighbors) {
        const int nz = (nb.offsets[0] != 0) + (nb.offsets[1] != 0) + (nb.offsets[2] != 0);
        if (nz == 1 && nb.loc.level != pmb->loc.level) face_other_level = true;
        if (nb.loc.level < pmb->loc.level) coarser = true;
      }
      (face_other_level ? flx : plain).push_back(pmb->pack_index);
      if (coarser) stale.push_back(pmb->pack_index);
    }
    c.n_flxcor = static_cast<int>(flx.size());
    c.n_plain = static_cast<int>(plain.size());
    c.n_stale_ghosts = static_cast<int>(stale.size());
    upload(c.ids_flxcor, flx);
    upload(c.ids_plain, plain);
    upload(c.ids_stale_ghosts, stale);
  }
  if (!c.packed) {
    PB2_CHECK(pb2_event_create(&c.early_ready));
    PB2_CHECK(pb2_event_create(&c.unpacked));
    PB2_CHECK(pb2_event_create(&c.packed));
    PB2_CHECK(pb2_event_create(&c.received));
    PB2_CHECK(pb2_event_create(&c.sent));
  }
  c.built_generation = md->alloc_generation;
  if (timing) {
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    const auto tr8 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "rebuild %s: plan %.2f ms, tables %.2f ms (%zu local, %zu send channels): "
                         "copy regions %.2f, copy table %.2f, slab tables %.2f, prores %.2f, bcs %.2f, "
                         "block classes %.2f\n",
                 md->label().c_str(), ms(tr0, tr1), ms(tr1, tr8), c.plan->local.size(),
                 c.plan->send.size(), ms(tr2, tr3), ms(tr3, tr4), ms(tr4, tr5), ms(tr5, tr6),
                 ms(tr6, tr7), ms(tr7, tr8));
  }
}

inline BvarsCache &Cache(std::shared_ptr<MeshData<Real>> &md) {
  if (md->bvars().built_generation != md->alloc_generation) Rebuild(md.get());
  return md->bvars();
}

constexpr bool DoesLocal(BoundaryType bt) {
  return bt == BoundaryType::local || bt == BoundaryType::any;
}
constexpr bool DoesNonlocal(BoundaryType bt) {
  return bt == BoundaryType::nonlocal || bt == BoundaryType::any;
}

} // namespace

void BuildBoundaryBuffers(std::shared_ptr<MeshData<Real>> &md) { Cache(md); }

BvarsCache &GetBvarsCache(MeshData<Real> *md) {
  if (md->bvars().built_generation != md->alloc_generation) Rebuild(md);
  return md->bvars();
}

template <BoundaryType bt>
TaskStatus StartReceiveBoundBufs(std::shared_ptr<MeshData<Real>> &md) {
  // receives are posted together with the sends inside one NCCL group; nothing to pre-post
  Cache(md);
  return TaskStatus::complete;
}

template <BoundaryType bt>
TaskStatus SendBoundBufs(std::shared_ptr<MeshData<Real>> &md) {
  BvarsCache &c = Cache(md);
  Mesh *pm = md->GetMeshPointer();
  pb2_stream_t st = md->stream();
  if (DoesLocal(bt)) {
    // boundary_communication.cpp:82-87: restrict before anything reads the coarse buffers
    if (pm->multilevel) {
      PB2_CHECK(pb2_restrict(c.restrict_send[0], st));
      PB2_CHECK(pb2_restrict_te(c.te_restrict_send[0], st));
    }
    if (c.sparse) {
      // sender half of a sparse exchange: which messages are null (:95-157)
      PB2_CHECK(pb2_memset(c.sparse_flags.get(), 0, sizeof(int32_t) * c.sparse_flags_h.size(), st));
      PB2_CHECK(pb2_copy_flags(c.copy_local, c.sparse_flags.get<int32_t>(), st));
    }
    c.send_generation++; // the copy itself happens in SetBounds<local> of the receiver
  }
  if (DoesNonlocal(bt) && c.plan->send_elements + c.plan->recv_elements > 0) {
    if (pm->multilevel) {
      PB2_CHECK(pb2_restrict(c.restrict_send[1], st));
      PB2_CHECK(pb2_restrict_te(c.te_restrict_send[1], st));
    }
    pb2_stream_t cs = pm->comm_stream;
    // Where the pack runs: normally on the compute stream, after everything enqueued so far.
    // If the producer signalled `early_ready` (all blocks that feed nonlocal channels are
    // final), pack on the communication stream right behind that event so that packing and
    // the transfer overlap whatever the compute stream still has to do.
    pb2_stream_t ps = st;
    if (c.early_valid && !pm->multilevel) {
      ps = cs;
      PB2_CHECK(pb2_stream_wait_event(cs, c.early_ready));
      // ... and, if the producer reports from inside its last kernel, for the boundary blocks
      if (c.progress_target > 0)
        PB2_CHECK(pb2_stream_wait_value(cs, c.progress.get<int32_t>(), c.progress_target));
    }
    c.early_valid = false;
    c.progress_target = 0;
    // the previous exchange must have left the slabs: its sends (same stream order on cs, or
    // the `sent` event) and its unpack (`unpacked`, recorded on the compute stream)
    if (c.nonlocal_in_flight && ps == st) PB2_CHECK(pb2_stream_wait_event(st, c.sent));
    if (c.push_mode) {
      // peer push: tell the peers this rank's ghost cells may be overwritten, wait for the same
      // from them, store the halo into their ghost cells and raise the arrival flags — one
      // handshake kernel and one copy launch on `ps`; nothing to receive or unpack afterwards
      c.push_seq++;
      // "ready" below promises that our receive slab / ghost cells of the previous exchange are
      // not in use any more: its unpack may have run on the other stream
      if (c.unpacked_valid) PB2_CHECK(pb2_stream_wait_event(ps, c.unpacked));
      const int me = pm->nranks > 1 ? pm->my_rank : 0;
      if (c.push_ce) {
        // the previous exchange's copies have left the send slab (they may have run on the other
        // stream); pack locally, then let the copy engines carry each peer's segment
        if (c.nonlocal_in_flight) PB2_CHECK(pb2_stream_wait_event(ps, c.sent));
        PB2_CHECK(pb2_pack(c.pack, c.send_slab.get<Real>(), nullptr, ps));
      }
      PB2_CHECK(pb2_peer_handshake(c.push_peer_flags.get<int32_t *>(), c.push_flags.get<int32_t>(),
                                   c.push_peer_ids.get<int32_t>(), c.push_npeers, me, pm->nranks,
                                   c.push_seq, ps));
      if (c.push_ce) {
        for (const BvarsCache::PushSegment &sg : c.push_segments)
          PB2_CHECK(pb2_memcpy_d2d(sg.dst, c.send_slab.get<Real>() + sg.src_off,
                                   sizeof(Real) * static_cast<size_t>(sg.count), ps));
        PB2_CHECK(pb2_peer_signal(c.push_peer_flags.get<int32_t *>(), c.push_npeers, me,
                                  pm->nranks, c.push_seq, ps));
      } else {
        PB2_CHECK(pb2_copy_signal(c.push, c.push_counter.get<int32_t>(),
                                  c.push_peer_flags.get<int32_t *>(), c.push_npeers, me,
                                  pm->nranks, c.push_seq, ps));
      }
      PB2_CHECK(pb2_event_record(c.sent, ps));
      c.nonlocal_in_flight = true;
      return TaskStatus::complete;
    }
    if (c.sparse) {
      // which messages are null (:95-157): flags out of the pack, as Reals into the flag slab
      PB2_CHECK(pb2_memset(c.send_flags.get(), 0, sizeof(int32_t) * c.send_flags_h.size(), ps));
      PB2_CHECK(pb2_pack(c.pack, c.send_slab.get<Real>(), c.send_flags.get<int32_t>(), ps));
      PB2_CHECK(pb2_memcpy_d2h(c.send_flags_h.data(), c.send_flags.get(),
                               sizeof(int32_t) * c.send_flags_h.size(), ps));
      PB2_CHECK(pb2_stream_sync(ps));
      c.flag_slab_h.assign(c.send_flags_h.begin(), c.send_flags_h.end());
      PB2_CHECK(pb2_memcpy_h2d(c.send_flag_slab.get(), c.flag_slab_h.data(),
                               sizeof(Real) * c.flag_slab_h.size(), ps));
      PB2_CHECK(pb2_stream_sync(ps)); // flag_slab_h is reused by the receive side
    } else {
      PB2_CHECK(pb2_pack(c.pack, c.send_slab.get<Real>(), nullptr, ps));
    }
    if (ps == st) {
      PB2_CHECK(pb2_event_record(c.packed, st));
      PB2_CHECK(pb2_stream_wait_event(cs, c.packed));
    }
    if (c.unpacked_valid) PB2_CHECK(pb2_stream_wait_event(cs, c.unpacked));
    if (pm->nranks > 1) {
      PARTHENON_REQUIRE(pm->comm != nullptr, "multi-rank mesh without a communicator");
      PB2_CHECK(pb2_comm_exchange(pm->comm, c.send_slab.get<Real>(), c.plan->send_off.data(),
                                  c.recv_slab.get<Real>(), c.plan->recv_off.data(), cs));
      if (c.sparse)
        PB2_CHECK(pb2_comm_exchange(pm->comm, c.send_flag_slab.get<Real>(),
                                    c.send_flag_off.data(), c.recv_flag_slab.get<Real>(),
                                    c.recv_flag_off.data(), cs));
    } else {
      // virtual ranks on one device: the "wire" is a device-to-device copy of the slab
      PB2_CHECK(pb2_memcpy_d2d(c.recv_slab.get(), c.send_slab.get(),
                               sizeof(Real) * static_cast<size_t>(c.plan->send_elements), cs));
      if (c.sparse)
        PB2_CHECK(pb2_memcpy_d2d(c.recv_flag_slab.get(), c.send_flag_slab.get(),
                                 sizeof(Real) * c.send_flags_h.size(), cs));
    }
    PB2_CHECK(pb2_event_record(c.received, cs));
    PB2_CHECK(pb2_event_record(c.sent, cs));
    c.nonlocal_in_flight = true;
  }
  return TaskStatus::complete;
}

template <BoundaryType bt>
TaskStatus ReceiveBoundBufs(std::shared_ptr<MeshData<Real>> &md) {
  BvarsCache &c = Cache(md);
  Mesh *pm = md->GetMeshPointer();
  if (DoesLocal(bt)) {
    // every partition that sends to us must have published this exchange
    // (CommBuffer::TryReceive for same-rank buffers, communication_buffer.hpp:390-400)
    for (const Channel &ch : c.plan->local) {
      const MeshBlock *sb = pm->block_list[pm->GetLid(ch.sender_gid)].get();
      if (sb->partition == md->partition_id()) continue;
      MeshData<Real> *smd = ContainerOf(md.get(), sb);
      if (smd->bvars().send_generation <= c.consumed_generation[sb->partition])
        return TaskStatus::incomplete;
    }
    if (c.send_generation <= c.consumed_generation[md->partition_id()])
      return TaskStatus::incomplete;
    if (c.sparse) {
      // :216-232: a field that receives actual data in any boundary is allocated there
      PB2_CHECK(pb2_memcpy_d2h(c.sparse_flags_h.data(), c.sparse_flags.get(),
                               sizeof(int32_t) * c.sparse_flags_h.size(), md->stream()));
      PB2_CHECK(pb2_stream_sync(md->stream()));
      for (size_t i = 0; i < c.plan->local.size(); ++i) {
        const Channel &ch = c.plan->local[i];
        Variable &rv = *c.vars[ch.var];
        if (!rv.metadata().IsSparse() || !c.sparse_flags_h[i]) continue;
        const MeshBlock *rb = pm->block_list[pm->GetLid(ch.receiver_gid)].get();
        if (!rv.IsAllocated(rb->pack_index)) pm->AllocateSparse(rv.label(), rb->lid);
      }
    }
  }
  // nonlocal: completion is a stream-side event wait in SetBounds — no host polling; only
  // sparse fields need the host: which messages were null decides what gets allocated
  if (DoesNonlocal(bt) && c.sparse && !c.recv_flags_h.empty()) {
    PB2_CHECK(pb2_event_sync(c.received));
    c.flag_slab_h.resize(c.recv_flags_h.size());
    PB2_CHECK(pb2_memcpy_d2h(c.flag_slab_h.data(), c.recv_flag_slab.get(),
                             sizeof(Real) * c.flag_slab_h.size(), md->stream()));
    PB2_CHECK(pb2_stream_sync(md->stream()));
    for (size_t i = 0; i < c.plan->recv.size(); ++i) {
      const Channel &ch = c.plan->recv[i];
      Variable &rv = *c.vars[ch.var];
      c.recv_flags_h[i] = c.flag_slab_h[i] != 0.0 ? 1 : 0;
      if (!rv.metadata().IsSparse() || !c.recv_flags_h[i]) continue;
      const MeshBlock *rb = pm->block_list[pm->GetLid(ch.receiver_gid)].get();
      if (!rv.IsAllocated(rb->pack_index)) pm->AllocateSparse(rv.label(), rb->lid);
    }
    // (an allocation bumps alloc_generation: the next Cache() rebuilds the tables; the flag
    // buffers and slabs keep their size and survive)
    PB2_CHECK(pb2_memcpy_h2d(c.recv_flags.get(), c.recv_flags_h.data(),
                             sizeof(int32_t) * c.recv_flags_h.size(), md->stream()));
    PB2_CHECK(pb2_stream_sync(md->stream()));
  }
  return TaskStatus::complete;
}

template <BoundaryType bt>
TaskStatus SetBounds(std::shared_ptr<MeshData<Real>> &md) {
  BvarsCache &c = Cache(md);
  Mesh *pm = md->GetMeshPointer();
  pb2_stream_t st = md->stream();
  if (DoesLocal(bt)) {
    if (c.uniform_halo && c.defer_local) {
      // the consumer reads its neighbours' interiors: the copy happens only if someone else
      // asks for these ghost cells (EnsureLocalGhosts)
      c.defer_local = false;
      c.local_ghosts_stale = true;
    } else if (c.uniform_halo) {
      for (Variable *v : c.vars) {
        const pb2_pack_geom g = md->Geometry(*v);
        PB2_CHECK(pb2_halo_copy_uniform(&g, v->data(), c.halo_nbr.get<int32_t>(), st));
      }
      c.local_ghosts_stale = false;
    } else if (c.sparse) {
      // (Cache() above rebuilt the tables if ReceiveBoundBufs allocated anything; the flags
      // are indexed by channel, which allocation does not change)
      PB2_CHECK(pb2_copy_select(c.copy_local, c.sparse_flags.get<int32_t>(), st));
    } else {
      PB2_CHECK(pb2_copy(c.copy_local, nullptr, st));
    }
    c.elements_local = c.plan->local_elements;
    for (int p = 0; p < pm->DefaultNumPartitions(); ++p) {
      MeshData<Real> *smd =
          p == md->partition_id() ? md.get() : pm->mesh_data.GetOrAdd(md->label(), p).get();
      c.consumed_generation[p] = smd->bvars().send_generation;
    }
    if (pm->multilevel) { // :338-346
      PB2_CHECK(pb2_restrict(c.restrict_set[0], st));
      PB2_CHECK(pb2_restrict_te(c.te_restrict_set[0], st));
    }
  }
  if (DoesNonlocal(bt) && c.plan->recv_elements > 0) {
    if (c.push_mode) {
      // the peers stored this halo into our receive slab (or, direct variant, into our ghost
      // cells) themselves: wait for their arrival flags and unpack, on the communication stream
      // if the consumer defers (it waits for `unpacked` when it reaches the blocks with remote
      // faces), else on the compute stream
      const bool defer = c.defer_remote;
      c.defer_remote = false;
      pb2_stream_t ws = defer ? pm->comm_stream : st;
      PB2_CHECK(pb2_peer_wait(c.push_flags.get<int32_t>(), c.push_peer_ids.get<int32_t>(),
                              c.push_npeers, pm->nranks, c.push_seq, ws));
      if (!c.push_direct) PB2_CHECK(pb2_unpack(c.unpack, c.recv_slab.get<Real>(), nullptr, ws));
      PB2_CHECK(pb2_event_record(c.unpacked, ws));
      if (defer) c.remote_pending = true;
    } else if (c.defer_remote && !pm->multilevel && !c.sparse) {
      // unpack on the communication stream, in order behind the exchange that fills the slab;
      // the compute stream keeps going and waits for `unpacked` when it needs these ghosts
      c.defer_remote = false;
      pb2_stream_t cs = pm->comm_stream;
      PB2_CHECK(pb2_unpack(c.unpack, c.recv_slab.get<Real>(), nullptr, cs));
      PB2_CHECK(pb2_event_record(c.unpacked, cs));
      c.remote_pending = true;
    } else {
      PB2_CHECK(pb2_stream_wait_event(st, c.received));
      PB2_CHECK(pb2_unpack(c.unpack, c.recv_slab.get<Real>(),
                           c.sparse ? c.recv_flags.get<int32_t>() : nullptr, st));
      PB2_CHECK(pb2_event_record(c.unpacked, st));
    }
    c.unpacked_valid = true;
    c.elements_nonlocal = c.plan->recv_elements;
    if (pm->multilevel) {
      PB2_CHECK(pb2_restrict(c.restrict_set[1], st));
      PB2_CHECK(pb2_restrict_te(c.te_restrict_set[1], st));
    }
  }
  return TaskStatus::complete;
}

template <BoundaryType bt>
TaskStatus ProlongateBounds(std::shared_ptr<MeshData<Real>> &md) {
  BvarsCache &c = Cache(md);
  Mesh *pm = md->GetMeshPointer();
  if (!pm->multilevel) return TaskStatus::complete;
  for (int cls = 0; cls < 2; ++cls) {
    if (cls == 0 && !DoesLocal(bt)) continue;
    if (cls == 1 && !DoesNonlocal(bt)) continue;
    for (int o = 0; o < 3; ++o) PB2_CHECK(pb2_prolongate(c.prolongate[cls][o], o, md->stream()));
    for (int o = 0; o < 3; ++o)
      PB2_CHECK(pb2_prolongate_te(c.te_prolongate[cls][o], o, md->stream()));
  }
  // face / edge / node fields: the fine elements inside coarse edges, faces and cells come
  // after ALL elements that fine and coarse grids share (boundary_communication.cpp:384-389) —
  // an internal element of a local region may average shared elements that a nonlocal region
  // prolongates
  for (int cls = 0; cls < 2; ++cls) {
    if (cls == 0 && !DoesLocal(bt)) continue;
    if (cls == 1 && !DoesNonlocal(bt)) continue;
    PB2_CHECK(pb2_prolongate_internal(c.te_internal[cls], md->stream()));
    PB2_CHECK(pb2_prolongate_toth_roe(c.te_toth_roe[cls], md->stream()));
  }
  return TaskStatus::complete;
}

#define PB2_INSTANTIATE(bt)                                                               \
  template TaskStatus StartReceiveBoundBufs<bt>(std::shared_ptr<MeshData<Real>> &);       \
  template TaskStatus SendBoundBufs<bt>(std::shared_ptr<MeshData<Real>> &);               \
  template TaskStatus ReceiveBoundBufs<bt>(std::shared_ptr<MeshData<Real>> &);            \
  template TaskStatus SetBounds<bt>(std::shared_ptr<MeshData<Real>> &);                   \
  template TaskStatus ProlongateBounds<bt>(std::shared_ptr<MeshData<Real>> &);
PB2_INSTANTIATE(BoundaryType::any)
PB2_INSTANTIATE(BoundaryType::local)
PB2_INSTANTIATE(BoundaryType::nonlocal)
#undef PB2_INSTANTIATE

// ---------------------------------------------------------------------------------------
// flux correction (flxcor_send / flxcor_recv of the reference, boundary_communication.cpp:
// 454-461; region selection loop_utils.hpp:145-160)
// ---------------------------------------------------------------------------------------
namespace {

int FaceDir(const int off[3]) { // CellCentOffsets::IsFace
  const int nz = (off[0] != 0) + (off[1] != 0) + (off[2] != 0);
  if (nz != 1) return -1;
  return off[0] != 0 ? 0 : (off[1] != 0 ? 1 : 2);
}

struct FlxChannel {
  int seg;
  int sender_gid, receiver_gid, var, offset_index;
  pb2_flxcor_region send{}; // sender side (restrict into slab)
  pb2_bnd_region recv{};    // receiver side (unpack from slab)
  int64_t n = 0;
};

void RebuildFluxCorrection(MeshData<Real> *md, BvarsCache &c) {
  Mesh *pm = md->GetMeshPointer();
  const int V = pm->virtual_ranks > 1 ? pm->virtual_ranks : 1;
  const int npeers = V > 1 ? V * V : pm->nranks;
  std::vector<Variable *> vars;
  for (Variable *v : md->GetVariablesByFlag({Metadata::WithFluxes}))
    if (v->topological_type() == TopologicalType::Cell) vars.push_back(v);
  std::vector<pb2_flxcor_region> local;
  std::vector<FlxChannel> send, recv;
  c.flxcor_local_elements = 0;
  for (auto &pmb : md->GetBlockList()) {
    const int my_vr = pm->VirtualRankOf(pmb->gid);
    for (auto &nb : pmb->neighbors) {
      const int dir = FaceDir(nb.offsets);
      if (dir < 0 || dir >= pm->ndim) continue;
      const int nb_vr = nb.rank == pm->my_rank ? pm->VirtualRankOf(nb.gid) : 0;
      const bool is_local = nb.rank == pm->my_rank && nb_vr == my_vr;
      const IndexBox box = CalcIndicesFlux(nb, pmb.get());
      for (size_t iv = 0; iv < vars.size(); ++iv) {
        Variable &v = *vars[iv];
        if (nb.loc.level == pmb->loc.level + 1) {
          // this block is the coarser RECEIVER of nb's restricted fluxes
          if (is_local) {
            const MeshBlock *sb = pm->block_list[nb.lid].get();
            const NeighborBlock *q = MatchingNeighbor(sb, pmb->gid, nb.offsets);
            PARTHENON_REQUIRE(q != nullptr, "no matching flux-correction sender");
            const IndexBox sbox = CalcIndicesFlux(*q, sb);
            Variable &sv = ContainerOf(md, sb)->Get(v.label());
            pb2_flxcor_region r{};
            r.fine = sv.flux(dir + 1) + sb->pack_index * sv.block_stride;
            r.coarse = v.flux(dir + 1) + pmb->pack_index * v.block_stride;
            r.dir = dir;
            r.ndim = pm->ndim;
            for (int d = 0; d < 3; ++d) {
              PARTHENON_REQUIRE(sbox.n(d) == box.n(d), "flux-correction extents differ");
              // coarse index -> fine index of the first child face (pr_ops.hpp:129-131)
              const int cis = sb->c_cellbounds.Bounds(d, IndexDomain::interior).s;
              const int fis = sb->cellbounds.Bounds(d, IndexDomain::interior).s;
              r.fs[d] = sb->block_size.symmetry_[d] ? fis : (sbox.s[d] - cis) * 2 + fis;
              r.ds[d] = box.s[d];
              r.n[d] = box.n(d);
            }
            r.ncomp = v.NumComponents();
            r.fine_stride_j = sv.ni;
            r.fine_stride_k = sv.ni * sv.nj;
            r.fine_stride_c = static_cast<int32_t>(sv.comp_stride);
            r.coarse_stride_j = v.ni;
            r.coarse_stride_k = v.ni * v.nj;
            r.coarse_stride_c = static_cast<int32_t>(v.comp_stride);
            // sparse fields: an unallocated sender sends a null message, which leaves the
            // receiver's flux alone (no default fill for flxcor_recv); an unallocated receiver
            // sets nothing
            r.status = sv.IsAllocated(sb->pack_index) && v.IsAllocated(pmb->pack_index)
                           ? PB2_REGION_ALLOCATED
                           : 0u;
            const auto dx = sb->coords.Dx();
            r.area = dir == 0 ? dx[1] * dx[2] : (dir == 1 ? dx[0] * dx[2] : dx[0] * dx[1]);
            c.flxcor_local_elements += static_cast<int64_t>(r.ncomp) * r.n[0] * r.n[1] * r.n[2];
            local.push_back(r);
          } else {
            FlxChannel ch;
            ch.seg = V > 1 ? nb_vr * V + my_vr : nb.rank;
            ch.sender_gid = nb.gid;
            ch.receiver_gid = pmb->gid;
            ch.var = static_cast<int>(iv);
            ch.offset_index = OffsetIndexOf(-nb.offsets[0], -nb.offsets[1], -nb.offsets[2]);
            pb2_bnd_region &r = ch.recv;
            r.var = v.flux(dir + 1) + pmb->pack_index * v.block_stride;
            r.stride_j = v.ni;
            r.stride_k = v.ni * v.nj;
            r.stride_c = static_cast<int32_t>(v.comp_stride);
            for (int d = 0; d < 3; ++d) {
              r.s[d] = box.s[d];
              r.n[d] = box.n(d);
            }
            r.ncomp = v.NumComponents();
            r.flag_slot = -1;
            r.status = PB2_REGION_ALLOCATED | PB2_REGION_BUF_ALLOCATED;
            ch.n = box.size() * r.ncomp;
            recv.push_back(ch);
          }
        } else if (nb.loc.level == pmb->loc.level - 1 && !is_local) {
          // this block is the finer SENDER towards another device
          FlxChannel ch;
          ch.seg = V > 1 ? my_vr * V + nb_vr : nb.rank;
          ch.sender_gid = pmb->gid;
          ch.receiver_gid = nb.gid;
          ch.var = static_cast<int>(iv);
          ch.offset_index = nb.OffsetIndex();
          pb2_flxcor_region &r = ch.send;
          r.fine = v.flux(dir + 1) + pmb->pack_index * v.block_stride;
          r.coarse = nullptr;
          r.dir = dir;
          r.ndim = pm->ndim;
          for (int d = 0; d < 3; ++d) {
            const int cis = pmb->c_cellbounds.Bounds(d, IndexDomain::interior).s;
            const int fis = pmb->cellbounds.Bounds(d, IndexDomain::interior).s;
            r.fs[d] = pmb->block_size.symmetry_[d] ? fis : (box.s[d] - cis) * 2 + fis;
            r.ds[d] = 0;
            r.n[d] = box.n(d);
          }
          r.ncomp = v.NumComponents();
          r.fine_stride_j = v.ni;
          r.fine_stride_k = v.ni * v.nj;
          r.fine_stride_c = static_cast<int32_t>(v.comp_stride);
          r.status = PB2_REGION_ALLOCATED;
          const auto dx = pmb->coords.Dx();
          r.area = dir == 0 ? dx[1] * dx[2] : (dir == 1 ? dx[0] * dx[2] : dx[0] * dx[1]);
          ch.n = box.size() * r.ncomp;
          send.push_back(ch);
        }
      }
    }
  }
  // fluxes of face fields: edge-centred flux fields (Metadata::Flux | Edge), one plan per field
  // (the slab offsets depend on its components per element)
  const std::vector<Variable *> fvars = md->GetVariablesByFlag({Metadata::Flux});
  std::vector<EdgeFluxPlan> fplans;
  if (pm->multilevel)
    for (Variable *fv : fvars) {
      PARTHENON_REQUIRE(fv->topological_type() == TopologicalType::Edge,
                        "flux fields other than the edge-centred flux of a face field are not built");
      fplans.push_back(BuildEdgeFluxPlan(pm, md->GetBlockList(), fv->TensorComponents()));
    }
  // a peer segment of the slabs: [face fluxes of cell-centred fields | edge fluxes of face field
  // 0 | of face field 1 ...]; both sides order each part by the same key => identical offsets,
  // no handshake
  std::vector<int64_t> te_send_size(npeers, 0), te_recv_size(npeers, 0);
  for (const EdgeFluxPlan &fp : fplans)
    for (int p = 0; p < npeers; ++p) {
      te_send_size[p] += fp.send_off[p + 1] - fp.send_off[p];
      te_recv_size[p] += fp.recv_off[p + 1] - fp.recv_off[p];
    }
  std::vector<int64_t> cell_send_size, cell_recv_size;
  auto layout = [&](std::vector<FlxChannel> &chs, std::vector<int64_t> &seg_off, int64_t &total,
                    bool is_send, const std::vector<int64_t> &extra,
                    std::vector<int64_t> &seg_size) {
    std::stable_sort(chs.begin(), chs.end(), [](const FlxChannel &a, const FlxChannel &b) {
      return std::make_tuple(a.seg, a.sender_gid, a.receiver_gid, a.var, a.offset_index) <
             std::make_tuple(b.seg, b.sender_gid, b.receiver_gid, b.var, b.offset_index);
    });
    seg_size.assign(npeers, 0);
    for (auto &ch : chs) {
      (is_send ? ch.send.buf_off : ch.recv.buf_off) = seg_size[ch.seg];
      seg_size[ch.seg] += ch.n + (ch.n & 1);
    }
    seg_off.assign(npeers + 1, 0);
    for (int p = 0; p < npeers; ++p) seg_off[p + 1] = seg_off[p] + seg_size[p] + extra[p];
    for (auto &ch : chs) (is_send ? ch.send.buf_off : ch.recv.buf_off) += seg_off[ch.seg];
    total = seg_off[npeers];
  };
  layout(send, c.flxcor_send_off, c.flxcor_send_elements, true, te_send_size, cell_send_size);
  layout(recv, c.flxcor_recv_off, c.flxcor_recv_elements, false, te_recv_size, cell_recv_size);
  PARTHENON_REQUIRE(c.flxcor_send_elements + c.flxcor_recv_elements == 0 ||
                        pm->DefaultNumPartitions() == 1,
                    "inter-device flux correction needs one MeshData per rank");
  std::vector<pb2_flxcor_region> packs;
  std::vector<pb2_bnd_region> unpacks;
  for (auto &ch : send) packs.push_back(ch.send);
  for (auto &ch : recv) unpacks.push_back(ch.recv);
  PB2_CHECK(pb2_flxcor_table_create(&c.flxcor_local, local.data(), static_cast<int64_t>(local.size())));
  PB2_CHECK(pb2_flxcor_table_create(&c.flxcor_pack, packs.data(), static_cast<int64_t>(packs.size())));
  PB2_CHECK(pb2_bnd_table_create(&c.flxcor_unpack, unpacks.data(), static_cast<int64_t>(unpacks.size())));
  if (c.flxcor_send_elements > 0)
    c.flxcor_send_slab.Allocate(sizeof(Real) * static_cast<size_t>(c.flxcor_send_elements), md->stream());
  if (c.flxcor_recv_elements > 0)
    c.flxcor_recv_slab.Allocate(sizeof(Real) * static_cast<size_t>(c.flxcor_recv_elements), md->stream());

  // ---- edge-centred fluxes of face fields: regions of the plans built above ----
  std::vector<pb2_prores_region> te_restricts, te_send_restricts;
  std::vector<pb2_copy_region> te_copies[2];
  std::vector<pb2_bnd_region> te_packs, te_unpacks[2];
  c.teflx_elements = 0;
  std::array<bool, 27> all_true;
  all_true.fill(true);
  std::vector<int64_t> te_send_base(npeers, 0), te_recv_base(npeers, 0); // of the current field
  for (size_t ifv = 0; ifv < fplans.size(); ++ifv) {
    Variable *fv = fvars[ifv];
    const EdgeFluxPlan &fplan = fplans[ifv];
    const TE edge_els[3] = {TE::E1, TE::E2, TE::E3};
    const int nc = fv->TensorComponents();
    for (const EdgeFluxRestrict &rr : fplan.restricts) {
      // the fine sender may sit in another partition of this device; all partitions share the
      // stream, so restricting its coarse buffer from here is ordered before the copies below
      const MeshBlock *sb = pm->block_list[pm->GetLid(rr.gid)].get();
      AddTeRegions(te_restricts, ContainerOf(md, sb)->Get(fv->label()), sb, rr.el,
                   edge_els[rr.el], nullptr, rr.box, all_true, pm->ndim);
    }
    for (const EdgeFluxRestrict &rr : fplan.send_restricts) {
      const MeshBlock *sb = pm->block_list[pm->GetLid(rr.gid)].get();
      AddTeRegions(te_send_restricts, *fv, sb, rr.el, edge_els[rr.el], nullptr, rr.box, all_true,
                   pm->ndim);
    }
    // pieces that cross devices: packed from the sender's coarse buffer, unpacked into the
    // receiver's flux array — block-edge messages ([0]) before face messages ([1]), like the
    // same-device copies
    for (const EdgeFluxPiece &pc : fplan.send) {
      const MeshBlock *sb = pm->block_list[pm->GetLid(pc.sender_gid)].get();
      pb2_bnd_region r{};
      r.var = fv->coarse() + sb->pack_index * fv->cblock_stride +
              static_cast<int64_t>(pc.el) * nc * fv->ccomp_stride;
      r.stride_j = fv->cni;
      r.stride_k = fv->cni * fv->cnj;
      r.stride_c = static_cast<int32_t>(fv->ccomp_stride);
      r.buf_off = c.flxcor_send_off[pc.seg] + cell_send_size[pc.seg] + te_send_base[pc.seg] +
                  pc.slab_off;
      for (int d = 0; d < 3; ++d) {
        r.s[d] = pc.send_box.s[d];
        r.n[d] = pc.send_box.n(d);
      }
      r.ncomp = nc;
      r.flag_slot = -1;
      r.status = PB2_REGION_ALLOCATED;
      te_packs.push_back(r);
    }
    for (const EdgeFluxPiece &pc : fplan.recv) {
      const MeshBlock *rb = pm->block_list[pm->GetLid(pc.receiver_gid)].get();
      pb2_bnd_region r{};
      r.var = fv->data() + rb->pack_index * fv->block_stride +
              static_cast<int64_t>(pc.el) * nc * fv->comp_stride;
      r.stride_j = fv->ni;
      r.stride_k = fv->ni * fv->nj;
      r.stride_c = static_cast<int32_t>(fv->comp_stride);
      r.buf_off = c.flxcor_recv_off[pc.seg] + cell_recv_size[pc.seg] + te_recv_base[pc.seg] +
                  pc.slab_off;
      for (int d = 0; d < 3; ++d) {
        r.s[d] = pc.recv_box.s[d];
        r.n[d] = pc.recv_box.n(d);
      }
      r.ncomp = nc;
      r.flag_slot = -1;
      r.status = PB2_REGION_ALLOCATED | PB2_REGION_BUF_ALLOCATED;
      c.teflx_elements += static_cast<int64_t>(nc) * r.n[0] * r.n[1] * r.n[2];
      te_unpacks[pc.pass].push_back(r);
    }
    for (int p = 0; p < npeers; ++p) {
      te_send_base[p] += fplan.send_off[p + 1] - fplan.send_off[p];
      te_recv_base[p] += fplan.recv_off[p + 1] - fplan.recv_off[p];
    }
    for (const EdgeFluxPiece &pc : fplan.pieces) {
      const MeshBlock *sb = pm->block_list[pm->GetLid(pc.sender_gid)].get();
      const MeshBlock *rb = pm->block_list[pm->GetLid(pc.receiver_gid)].get();
      Variable &sv = ContainerOf(md, sb)->Get(fv->label());
      pb2_copy_region r{};
      r.src = sv.coarse() + sb->pack_index * sv.cblock_stride +
              static_cast<int64_t>(pc.el) * nc * sv.ccomp_stride;
      r.dst = fv->data() + rb->pack_index * fv->block_stride +
              static_cast<int64_t>(pc.el) * nc * fv->comp_stride;
      for (int d = 0; d < 3; ++d) {
        r.ss[d] = pc.send_box.s[d];
        r.ds[d] = pc.recv_box.s[d];
        r.n[d] = pc.recv_box.n(d);
      }
      r.ncomp = nc;
      r.src_stride_j = sv.cni;
      r.src_stride_k = sv.cni * sv.cnj;
      r.src_stride_c = static_cast<int32_t>(sv.ccomp_stride);
      r.dst_stride_j = fv->ni;
      r.dst_stride_k = fv->ni * fv->nj;
      r.dst_stride_c = static_cast<int32_t>(fv->comp_stride);
      r.flag_slot = -1;
      r.status = PB2_REGION_ALLOCATED;
      c.teflx_elements += static_cast<int64_t>(nc) * r.n[0] * r.n[1] * r.n[2];
      te_copies[pc.pass].push_back(r);
    }
  }
  PB2_CHECK(pb2_prores_table_create(&c.teflx_restrict_send, te_send_restricts.data(),
                                    static_cast<int64_t>(te_send_restricts.size())));
  PB2_CHECK(pb2_bnd_table_create(&c.teflx_pack, te_packs.data(),
                                 static_cast<int64_t>(te_packs.size())));
  for (int pass = 0; pass < 2; ++pass)
    PB2_CHECK(pb2_bnd_table_create(&c.teflx_unpack[pass], te_unpacks[pass].data(),
                                   static_cast<int64_t>(te_unpacks[pass].size())));
  PB2_CHECK(pb2_prores_table_create(&c.teflx_restrict, te_restricts.data(),
                                    static_cast<int64_t>(te_restricts.size())));
  for (int pass = 0; pass < 2; ++pass)
    PB2_CHECK(pb2_copy_table_create(&c.teflx_copy[pass], te_copies[pass].data(),
                                    static_cast<int64_t>(te_copies[pass].size())));
  if (!c.flxcor_packed) {
    PB2_CHECK(pb2_event_create(&c.flxcor_packed));
    PB2_CHECK(pb2_event_create(&c.flxcor_received));
  }
  c.flxcor_built = true;
}

BvarsCache &FlxCache(MeshData<Real> *md) {
  BvarsCache &c = GetBvarsCache(md);
  if (!c.flxcor_built) RebuildFluxCorrection(md, c);
  return c;
}

// SendBoundBufs<flxcor_send>: restrict the fine face fluxes that face another device into the
// peer slabs and ship them
void SendFluxCorrections(MeshData<Real> *md) {
  Mesh *pm = md->GetMeshPointer();
  if (!pm->multilevel) return;
  BvarsCache &c = FlxCache(md);
  if (c.flxcor_send_elements + c.flxcor_recv_elements == 0) return;
  pb2_stream_t st = md->stream(), cs = pm->comm_stream;
  PB2_CHECK(pb2_flux_correct(c.flxcor_pack, c.flxcor_send_slab.get<Real>(), st));
  // edge-centred fluxes of face fields: restrict into the coarse buffers, pack what this rank's
  // blocks own into the same slab
  PB2_CHECK(pb2_restrict_te(c.teflx_restrict_send, st));
  PB2_CHECK(pb2_pack(c.teflx_pack, c.flxcor_send_slab.get<Real>(), nullptr, st));
  PB2_CHECK(pb2_event_record(c.flxcor_packed, st));
  PB2_CHECK(pb2_stream_wait_event(cs, c.flxcor_packed));
  if (pm->nranks > 1) {
    PARTHENON_REQUIRE(pm->comm != nullptr, "multi-rank mesh without a communicator");
    PB2_CHECK(pb2_comm_exchange(pm->comm, c.flxcor_send_slab.get<Real>(), c.flxcor_send_off.data(),
                                c.flxcor_recv_slab.get<Real>(), c.flxcor_recv_off.data(), cs));
  } else {
    PB2_CHECK(pb2_memcpy_d2d(c.flxcor_recv_slab.get(), c.flxcor_send_slab.get(),
                             sizeof(Real) * static_cast<size_t>(c.flxcor_send_elements), cs));
  }
  PB2_CHECK(pb2_event_record(c.flxcor_received, cs));
  c.flxcor_in_flight = true;
}

// SetBounds<flxcor_recv>: the coarser block's face flux is overwritten by the restricted one
void SetFluxCorrectionsImpl(MeshData<Real> *md) {
  Mesh *pm = md->GetMeshPointer();
  if (!pm->multilevel) return;
  BvarsCache &c = FlxCache(md);
  pb2_stream_t st = md->stream();
  PB2_CHECK(pb2_flux_correct(c.flxcor_local, nullptr, st));
  // edge-centred fluxes of face fields: restrict on the fine blocks, then deliver — across block
  // edges first, across faces second (BvarsCache::teflx_copy)
  // — whether a message came from this device (copy) or another one (unpack)
  PB2_CHECK(pb2_restrict_te(c.teflx_restrict, st));
  const bool remote = c.flxcor_in_flight;
  if (remote) PB2_CHECK(pb2_stream_wait_event(st, c.flxcor_received));
  const Real *slab = c.flxcor_recv_slab.get<Real>();
  for (int pass = 0; pass < 2; ++pass) {
    PB2_CHECK(pb2_copy(c.teflx_copy[pass], nullptr, st));
    if (remote) PB2_CHECK(pb2_unpack(c.teflx_unpack[pass], slab, nullptr, st));
  }
  if (remote) {
    PB2_CHECK(pb2_unpack(c.flxcor_unpack, slab, nullptr, st));
    c.flxcor_in_flight = false;
  }
}

} // namespace

TaskStatus StartReceiveFluxCorrections(std::shared_ptr<MeshData<Real>> &) {
  return TaskStatus::complete; // receives are posted with the sends inside one NCCL group
}
TaskStatus LoadAndSendFluxCorrections(std::shared_ptr<MeshData<Real>> &md) {
  SendFluxCorrections(md.get());
  return TaskStatus::complete;
}
TaskStatus ReceiveFluxCorrections(std::shared_ptr<MeshData<Real>> &) {
  return TaskStatus::complete; // completion is a stream-side event wait in SetFluxCorrections
}
TaskStatus SetFluxCorrections(std::shared_ptr<MeshData<Real>> &md) {
  SetFluxCorrectionsImpl(md.get());
  return TaskStatus::complete;
}
void FluxCorrection(MeshData<Real> *md) {
  SendFluxCorrections(md);
  SetFluxCorrectionsImpl(md);
}

TaskStatus ApplyBoundaryConditions(std::shared_ptr<MeshBlockData<Real>> &) {
  // the batch form below does every block of the MeshData in one launch per direction;
  // drivers written against this framework call it instead of looping over blocks
  return TaskStatus::complete;
}
TaskStatus ApplyBoundaryConditionsOnCoarseOrFineMD(std::shared_ptr<MeshData<Real>> &md,
                                                   bool coarse) {
  BvarsCache &c = Cache(md);
  Mesh *pm = md->GetMeshPointer();
  // faces in BoundaryFace order; inner and outer slabs of one direction are disjoint, so the
  // stock conditions of a direction are one launch (boundary_conditions.cpp:47-55); user
  // conditions of that direction follow before the next direction reads their ghosts
  if (c.has_bcs) EnsureRemoteGhosts(md.get()); // edges next to a remote face copy unpacked cells
  for (int d = 0; d < pm->ndim; ++d) {
    if (c.has_bcs) PB2_CHECK(pb2_apply_bcs(c.bc[coarse ? 1 : 0][d], md->stream()));
    for (int f = 2 * d; f < 2 * d + 2; ++f)
      if (pm->user_bcs[f]) pm->user_bcs[f](md, coarse);
  }
  return TaskStatus::complete;
}
TaskStatus ApplyBoundaryConditionsMD(std::shared_ptr<MeshData<Real>> &md) {
  return ApplyBoundaryConditionsOnCoarseOrFineMD(md, false);
}

void EnsureRemoteGhosts(MeshData<Real> *md) {
  BvarsCache &c = md->bvars();
  if (!c.remote_pending) return;
  PB2_CHECK(pb2_stream_wait_event(md->stream(), c.unpacked));
  c.remote_pending = false;
}

void EnsureLocalGhosts(MeshData<Real> *md) {
  EnsureRemoteGhosts(md);
  BvarsCache &c = md->bvars();
  if (!c.local_ghosts_stale) return;
  PARTHENON_REQUIRE(c.uniform_halo, "stale ghosts on a container without the uniform ghost fill");
  for (Variable *v : c.vars) {
    const pb2_pack_geom g = md->Geometry(*v);
    PB2_CHECK(pb2_halo_copy_uniform(&g, v->data(), c.halo_nbr.get<int32_t>(), md->stream()));
  }
  c.local_ghosts_stale = false;
  // physical boundaries copy from cells the exchange fills (edges and corners next to a
  // neighbour): same order as after every exchange
  Mesh *pm = md->GetMeshPointer();
  auto &sp = pm->mesh_data.GetOrAdd(md->label(), md->partition_id());
  ApplyBoundaryConditionsOnCoarseOrFineMD(sp, false);
}

void EnsureLocalGhosts(Mesh *pm) {
  for (auto &kv : pm->mesh_data.All()) EnsureLocalGhosts(kv.second.get());
}

TaskID AddBoundaryExchangeTasks(TaskID dependency, TaskList &tl,
                                std::shared_ptr<MeshData<Real>> &md, bool multilevel) {
  const auto any = BoundaryType::any;
  auto send = tl.AddTask(dependency, SendBoundBufs<any>, md);
  auto recv = tl.AddTask(dependency | send, ReceiveBoundBufs<any>, md);
  auto set = tl.AddTask(recv, SetBounds<any>, md);
  auto pro = set;
  if (multilevel) {
    auto cbound = tl.AddTask(set, ApplyBoundaryConditionsOnCoarseOrFineMD, md, true);
    pro = tl.AddTask(cbound, ProlongateBounds<any>, md);
  }
  return tl.AddTask(pro, ApplyBoundaryConditionsOnCoarseOrFineMD, md, false);
}

void CommunicateBoundaries(std::shared_ptr<MeshData<Real>> &md, bool prolongate) {
  // all partitions of the container send first, then receive/set (mesh.cpp:640-706)
  Mesh *pm = md->GetMeshPointer();
  std::vector<std::shared_ptr<MeshData<Real>>> parts;
  for (int p = 0; p < pm->DefaultNumPartitions(); ++p)
    parts.push_back(pm->mesh_data.GetOrAdd(md->label(), p));
  for (auto &m : parts) SendBoundBufs<BoundaryType::any>(m);
  for (auto &m : parts) {
    PARTHENON_REQUIRE(ReceiveBoundBufs<BoundaryType::any>(m) == TaskStatus::complete,
                      "local boundary buffers were not published");
    SetBounds<BoundaryType::any>(m);
  }
  // mesh.cpp:698-706: coarse BCs + prolongation, then the fine BCs
  for (auto &m : parts) {
    if (prolongate && pm->multilevel) {
      ApplyBoundaryConditionsOnCoarseOrFineMD(m, true);
      ProlongateBounds<BoundaryType::any>(m);
    }
    ApplyBoundaryConditionsOnCoarseOrFineMD(m, false);
  }
}

} // namespace parthenon
