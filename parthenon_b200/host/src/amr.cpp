// amr.cpp — adaptive mesh refinement on one device: tagging bookkeeping, tree update and the
// data movement of a remesh.
//
// Reference: MeshRefinement::SetRefinement (mesh/mesh_refinement.cpp:81-118),
// Mesh::UpdateMeshBlockTree (mesh/mesh-amr_loadbalance.cpp:496-625) with Tree::Refine /
// Tree::Derefine (mesh/forest/tree.cpp:93-143, 229-275), Mesh::RedistributeAndRefineMeshBlocks
// (:663-1010) and Refinement::Tag (amr_criteria/refinement_package.cpp:150-172).
//
// B200 shape of the data movement: the state lives in one slab per field, so a remesh builds a
// new slab and fills it with at most four launches per field —
//   kept blocks       one copy launch (whole blocks, ghosts included)
//   merged blocks     pb2_restrict of the children's interiors into their coarse buffers, then
//                     the same copy launch moves them into the parent's quadrants / octants
//   split blocks      the copy launch fills the children's coarse buffers (entire extents) from
//                     the parent, pb2_prolongate covers interior + ghosts
// followed by one full ghost exchange on the new mesh.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <unordered_set>

#include "pb2/bvals.hpp"
#include "pb2/parthenon.hpp"

namespace parthenon {

void Mesh::SetRefinement(int lid, AmrTag flag) {
  MeshBlock *pmb = block_list[lid].get();
  const int aret = std::max(-1, static_cast<int>(flag));
  if (aret == 0) pmb->refine_flag = 0;
  if (aret >= 0) pmb->deref_count = 0;
  if (aret > 0) {
    pmb->refine_flag = pmb->loc.level == max_level ? 0 : 1;
  } else if (aret < 0) {
    if (pmb->loc.level == root_level) {
      pmb->refine_flag = 0;
      pmb->deref_count = 0;
    } else {
      pmb->deref_count++;
      int ec = 0;
      for (const auto &nb : pmb->neighbors)
        if (nb.loc.level > pmb->loc.level) ec++;
      if (ec > 0)
        pmb->refine_flag = 0;
      else
        pmb->refine_flag = pmb->deref_count >= derefine_count ? -1 : 0;
    }
  }
}

namespace {
using LeafSet = std::unordered_set<LogicalLocation, LogicalLocationHash>;

LogicalLocation Daughter(const LogicalLocation &p, int q, int ndim) {
  LogicalLocation d;
  d.level = p.level + 1;
  for (int k = 0; k < 3; ++k) d.lx[k] = k < ndim ? (p.lx[k] << 1) + ((q >> k) & 1) : 0;
  return d;
}
} // namespace

bool Mesh::UpdateMeshBlockTree(std::vector<LogicalLocation> &new_leaves, int &nnew, int &ndel) {
  const int nleaf = 1 << ndim;
  std::vector<LogicalLocation> lref, lderef, clderef;
  // the refinement flags of every block of the mesh, in gid order (MPI_Allgatherv of the
  // flagged locations in the reference, :508-565)
  std::vector<Real> flags(nbtotal, 0.0);
  for (auto &pmb : block_list) flags[pmb->gid] = pmb->refine_flag;
  AllReduceSum(flags);
  int tnderef = 0;
  for (int g = 0; g < nbtotal; ++g) tnderef += flags[g] == -1.0;
  for (int g = 0; g < nbtotal; ++g) {
    if (flags[g] == 1.0) lref.push_back(loclist[g]);
    if (flags[g] == -1.0 && tnderef >= nleaf) lderef.push_back(loclist[g]);
  }
  if (lref.empty() && tnderef < nleaf) return false; // nothing to do (:525-527)

  // the list of newly derefined blocks: all siblings flagged, consecutive in gid order (:569-596)
  if (tnderef >= nleaf) {
    const int lk = ndim > 2, lj = ndim > 1;
    for (int n = 0; n < tnderef; ++n) {
      if ((lderef[n].lx[0] & 1) || (lderef[n].lx[1] & 1) || (lderef[n].lx[2] & 1)) continue;
      int r = n, rr = 0;
      for (int64_t k = 0; k <= lk; ++k)
        for (int64_t j = 0; j <= lj; ++j)
          for (int64_t i = 0; i <= 1; ++i) {
            if (r < tnderef) {
              if (lderef[n].lx[0] + i == lderef[r].lx[0] && lderef[n].lx[1] + j == lderef[r].lx[1] &&
                  lderef[n].lx[2] + k == lderef[r].lx[2] && lderef[n].level == lderef[r].level)
                rr++;
              r++;
            }
          }
      if (rr == nleaf) {
        LogicalLocation p = lderef[n].GetParent();
        for (int d = ndim; d < 3; ++d) p.lx[d] = 0;
        clderef.push_back(p);
      }
    }
  }
  // :597-602 sorts by level, finest first — all entries but the last one
  if (clderef.size() > 1)
    std::stable_sort(clderef.begin(), clderef.end() - 1,
                     [](const LogicalLocation &a, const LogicalLocation &b) { return a.level > b.level; });

  LeafSet leaves(loclist.begin(), loclist.end());
  // internal nodes of the tree (Tree::internal_nodes): kept in step with the leaf set
  LeafSet internal;
  for (const auto &kv : internal_) internal.insert(kv.first);
  // Tree::Refine with proper nesting (tree.cpp:93-143)
  std::function<int(const LogicalLocation &)> refine = [&](const LogicalLocation &ref) -> int {
    if (!leaves.count(ref)) return 0;
    leaves.erase(ref);
    internal.insert(ref);
    for (int q = 0; q < nleaf; ++q) leaves.insert(Daughter(ref, q, ndim));
    int nadded = nleaf - 1;
    if (ref.level <= root_level) return nadded; // no leaves above the root grid
    LogicalLocation par = ref.GetParent();
    for (int d = ndim; d < 3; ++d) par.lx[d] = 0;
    const int ox[3] = {static_cast<int>(ref.lx[0] - (par.lx[0] << 1)),
                       static_cast<int>(ref.lx[1] - (par.lx[1] << 1)),
                       static_cast<int>(ref.lx[2] - (par.lx[2] << 1))};
    for (int k = 0; k < (ndim > 2 ? 2 : 1); ++k)
      for (int j = 0; j < (ndim > 1 ? 2 : 1); ++j)
        for (int i = 0; i < 2; ++i) {
          LogicalLocation neigh = par, w;
          neigh.lx[0] += i + ox[0] - 1;
          neigh.lx[1] += j + ox[1] - (ndim > 1);
          neigh.lx[2] += k + ox[2] - (ndim > 2);
          if (!WrapLocation(neigh, w)) continue;
          nadded += refine(w);
        }
    return nadded;
  };
  // Tree::Derefine (tree.cpp:229-275)
  auto derefine = [&](const LogicalLocation &ref) -> int {
    for (int q = 0; q < nleaf; ++q) {
      const LogicalLocation d = Daughter(ref, q, ndim);
      if (!leaves.count(d)) return 0;
      for (int k = (ndim > 2 ? -1 : 0); k <= (ndim > 2 ? 1 : 0); ++k)
        for (int j = (ndim > 1 ? -1 : 0); j <= (ndim > 1 ? 1 : 0); ++j)
          for (int i = -1; i <= 1; ++i) {
            LogicalLocation neigh = d, w;
            neigh.lx[0] += i;
            neigh.lx[1] += j;
            neigh.lx[2] += k;
            if (!WrapLocation(neigh, w)) continue;
            if (internal.count(w)) return 0; // would abut a block two levels finer
          }
    }
    for (int q = 0; q < nleaf; ++q) leaves.erase(Daughter(ref, q, ndim));
    internal.erase(ref);
    leaves.insert(ref);
    return nleaf - 1;
  };
  for (auto &l : lref) nnew += refine(l);
  for (auto &l : clderef) ndel += derefine(l);
  if (nnew == 0 && ndel == 0) return false;
  new_leaves.assign(leaves.begin(), leaves.end());
  return true;
}

void Mesh::RebuildFromLeaves(const std::vector<LogicalLocation> &new_leaves,
                             const BlockList_t &old_blocks) {
  static const bool timing = std::getenv("PB2_TIME_HOST") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  BuildTree(nullptr, new_leaves);
  const auto t1 = std::chrono::steady_clock::now();
  multilevel = true;
  std::vector<double> cost(nbtotal, 1.0);
  AssignBlocks(cost, nranks, ranklist);
  nblist.assign(nranks, 0);
  for (int r : ranklist) nblist[r]++;
  nslist.assign(nranks, 0);
  for (int r = 1; r < nranks; ++r) nslist[r] = nslist[r - 1] + nblist[r - 1];
  BuildBlockList(&old_blocks);
  if (timing) {
    const auto t2 = std::chrono::steady_clock::now();
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    std::fprintf(stderr, "[pb2 rebuild] tree %.2f ms, assign + block list %.2f ms\n", ms(t0, t1), ms(t1, t2));
  }
}

// the host half of a remesh alone (tree update, re-partition, new block list): what
// pb2h_topology_regrid drives in device-free tests
bool Mesh::RegridTopologyOnly() {
  PARTHENON_REQUIRE(nranks == 1, "topology-only regrid is a single-rank test facility");
  int nnew = 0, ndel = 0;
  std::vector<LogicalLocation> new_leaves;
  static const bool timing = std::getenv("PB2_TIME_HOST") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  if (!UpdateMeshBlockTree(new_leaves, nnew, ndel)) return false;
  const auto t1 = std::chrono::steady_clock::now();
  const BlockList_t old_blocks = block_list;
  std::unordered_set<LogicalLocation, LogicalLocationHash> old(loclist.begin(), loclist.end());
  RebuildFromLeaves(new_leaves, old_blocks);
  if (timing) {
    const auto t2 = std::chrono::steady_clock::now();
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    std::fprintf(stderr, "[pb2 regrid] tree update %.2f ms, rebuild from leaves %.2f ms\n", ms(t0, t1), ms(t1, t2));
  }
  for (auto &pmb : block_list) {
    pmb->refine_flag = 0;
    if (!old.count(pmb->loc)) pmb->deref_count = 0;
  }
  return true;
}

void Mesh::RedistributeAndRefineMeshBlocks(const std::vector<LogicalLocation> &new_leaves) {
  PARTHENON_REQUIRE(DefaultNumPartitions() == 1,
                    "remeshing needs one MeshData per rank (parthenon/mesh/pack_size = -1)");
  static const bool timing = std::getenv("PB2_TIME_REMESH") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  const auto t0 = now();
  // only the base container survives a remesh (:667)
  mesh_data.PurgeNonBase();
  std::shared_ptr<MeshData<Real>> old_md = mesh_data.GetOrAdd("base", 0);
  const BlockList_t old_blocks = block_list;
  const std::vector<LogicalLocation> old_loclist = loclist;
  const std::vector<int> old_ranklist = ranklist, old_nslist = nslist;
  std::unordered_map<LogicalLocation, int, LogicalLocationHash> old_gid;
  for (int g = 0; g < static_cast<int>(old_loclist.size()); ++g) old_gid[old_loclist[g]] = g;
  auto old_local = [&](int g) { return g - old_nslist[my_rank]; }; // index into the old slab
  // derefinement counters travel with kept blocks (SendSameToSame packs them into the message,
  // :253-283): every rank learns every old counter
  std::vector<Real> old_count(old_loclist.size(), 0.0);
  for (auto &pmb : old_blocks) old_count[pmb->gid] = pmb->deref_count;
  AllReduceSum(old_count);

  // new tree, rank assignment (AssignBlocks with unit costs), block list, empty base container
  RebuildFromLeaves(new_leaves, old_blocks);
  for (auto &pmb : block_list) {
    auto it = old_gid.find(pmb->loc);
    pmb->refine_flag = 0;
    pmb->deref_count = it != old_gid.end() ? static_cast<int>(old_count[it->second]) : 0;
    if (it == old_gid.end() || old_ranklist[it->second] != my_rank)
      pmb->SetAllowedDt(std::numeric_limits<Real>::max()); // MeshBlock::Make: no time step yet
  }
  const auto t1 = now();
  mesh_data.Clear();
  std::shared_ptr<MeshData<Real>> new_md = mesh_data.GetOrAdd("base", 0);
  const auto t2 = now();

  const int nleaf = 1 << ndim;
  const int ng = Globals::nghost;
  pb2_stream_t st = stream;
  auto daughter_of = [&](const LogicalLocation &p, int q) { return Daughter(p, q, ndim); };
  auto parent_of = [&](const LogicalLocation &l) {
    LogicalLocation par = l.GetParent();
    for (int d = ndim; d < 3; ++d) par.lx[d] = 0;
    return par;
  };

  // face / edge / node fields: the fine elements inside coarse edges, faces and cells of NEW
  // children are prolongated only after an exchange in which older fine blocks own the shared
  // elements (mesh-amr_loadbalance.cpp:958-990); the regions are collected here
  bool has_te = false;
  std::vector<pb2_prores_region> te_internal, te_toth_roe;
  const TE te_containers[8] = {TE::NN, TE::E3, TE::E2, TE::E1, TE::F1, TE::F2, TE::F3, TE::CC};
  auto is_submanifold = [](TE f, TE cnt) { // basic_types.hpp:207-232
    int more = 0;
    for (int d = 0; d < 3; ++d) {
      if (TopologicalOffset(cnt, d) && !TopologicalOffset(f, d)) return false;
      more += TopologicalOffset(f, d) && !TopologicalOffset(cnt, d);
    }
    return more > 0;
  };

  for (auto &nvp : new_md->GetVariableVector()) {
    Variable &nv = *nvp;
    // the fields a remesh carries over: pmb->vars_cc_ (Independent / FillGhost cell-centred)
    // (+ Metadata::ForceRemeshComm, mesh-amr_loadbalance.cpp:680-690)
    if (!(nv.IsSet(Metadata::Independent) || nv.IsSet(Metadata::FillGhost) ||
          nv.IsSet(Metadata::ForceRemeshComm)))
      continue;
    // sparse fields (mesh-amr_loadbalance.cpp:700-1000 over allocated variables only): a kept
    // block keeps its allocation status and deallocation counter, a new child exists where its
    // parent did, a new parent where ANY daughter did (the quadrants of the others stay at the
    // zero a fresh allocation holds)
    const bool sparse = nv.metadata().IsSparse() && sparse_config.enabled;
    PARTHENON_REQUIRE(!sparse || nranks == 1,
                      "adaptive remeshing of sparse fields needs a single device in this build");
    // blocks of face / edge / node fields that change device: the 2-GPU run of round 2 did not
    // reproduce the reference's dumps (profiles/multigpu_check_r02_n2.txt), so refuse rather
    // than return wrong shared elements; one device is bit-exact (tests/test_tecomm_gpu.py)
    PARTHENON_REQUIRE(nranks == 1 || nv.topological_type() == TopologicalType::Cell,
                      "adaptive remeshing of face / edge / node fields needs a single device in "
                      "this build (" + nv.label() + ")");
    Variable &ov = old_md->Get(nv.label());
    const int64_t fsz = nv.block_stride + (nv.block_stride & 1);   // slab entries, 16-B aligned
    const int64_t csz = nv.cblock_stride + (nv.cblock_stride & 1);

    // ---- what moves where.  Every rank walks the NEW block list of the whole mesh in gid order,
    // so senders and receivers lay out the per-peer slabs identically without a handshake.
    struct Piece { // one old block (or its coarse buffer) feeding one new block
      int new_gid, old_gid, q; // q: daughter index for merged blocks, else -1
      int kind;                // 0 kept, 1 split (parent -> child), 2 merged (child -> parent)
    };
    std::vector<Piece> pieces;
    for (int n = 0; n < nbtotal; ++n) {
      const LogicalLocation &nl = loclist[n];
      auto same = old_gid.find(nl);
      if (same != old_gid.end()) {
        pieces.push_back({n, same->second, -1, 0});
        continue;
      }
      auto up = nl.level > 0 ? old_gid.find(parent_of(nl)) : old_gid.end();
      if (up != old_gid.end()) {
        pieces.push_back({n, up->second, -1, 1});
        continue;
      }
      for (int q = 0; q < nleaf; ++q) {
        auto ch = old_gid.find(daughter_of(nl, q));
        PARTHENON_REQUIRE(ch != old_gid.end(), "remesh cannot find the origin of a new block");
        pieces.push_back({n, ch->second, q, 2});
      }
    }
    std::vector<int64_t> send_off(nranks + 1, 0), recv_off(nranks + 1, 0);
    std::vector<int64_t> piece_send(pieces.size(), -1), piece_recv(pieces.size(), -1);
    {
      std::vector<int64_t> ssz(nranks, 0), rsz(nranks, 0);
      for (size_t i = 0; i < pieces.size(); ++i) {
        const int src = old_ranklist[pieces[i].old_gid], dst = ranklist[pieces[i].new_gid];
        if (src == dst) continue;
        const int64_t sz = pieces[i].kind == 2 ? csz : fsz;
        if (src == my_rank) {
          piece_send[i] = ssz[dst];
          ssz[dst] += sz;
        }
        if (dst == my_rank) {
          piece_recv[i] = rsz[src];
          rsz[src] += sz;
        }
      }
      for (int r = 0; r < nranks; ++r) {
        send_off[r + 1] = send_off[r] + ssz[r];
        recv_off[r + 1] = recv_off[r] + rsz[r];
      }
      for (size_t i = 0; i < pieces.size(); ++i) {
        if (piece_send[i] >= 0) piece_send[i] += send_off[ranklist[pieces[i].new_gid]];
        if (piece_recv[i] >= 0) piece_recv[i] += recv_off[old_ranklist[pieces[i].old_gid]];
      }
    }
    DeviceBuffer send_slab, recv_slab;
    if (send_off[nranks] > 0) send_slab.Allocate(sizeof(Real) * send_off[nranks], st);
    if (recv_off[nranks] > 0) recv_slab.Allocate(sizeof(Real) * recv_off[nranks], st);

    std::vector<pb2_copy_region> copies, sends;
    std::vector<pb2_prores_region> restricts, prolongs, te_restricts, te_prolongs;
    const bool te = nv.topological_type() != TopologicalType::Cell;
    has_te = has_te || te;
    const std::vector<TE> els = GetTopologicalElements(nv.topological_type());
    const int tnc = nv.TensorComponents();
    // one region of element number e (storage order) of block `lidx` of variable v, box given
    // in the coarse index space
    auto prores_te = [&](Variable &v, const MeshBlock *pmb, int lidx, int e, TE fel, const TE *cel,
                         const int bs[3], const int bn[3]) {
      pb2_prores_region p{};
      p.fine = v.data() + lidx * v.block_stride + static_cast<int64_t>(e) * tnc * v.comp_stride;
      p.coarse =
          v.coarse() + lidx * v.cblock_stride + static_cast<int64_t>(e) * tnc * v.ccomp_stride;
      const UniformCartesian cc(pmb->coords, 2);
      for (int d = 0; d < 3; ++d) {
        p.s[d] = bs[d];
        p.n[d] = bn[d];
        p.fine_is[d] = pmb->cellbounds.Bounds(d, IndexDomain::interior).s;
        p.coarse_is[d] = pmb->c_cellbounds.Bounds(d, IndexDomain::interior).s;
        p.fine_xmin[d] = pmb->coords.GetXmin()[d];
        p.fine_dx[d] = pmb->coords.Dx()[d];
        p.coarse_xmin[d] = cc.GetXmin()[d];
        p.coarse_dx[d] = cc.Dx()[d];
        p.ftop[d] = TopologicalOffset(fel, d);
        p.ctop[d] = cel ? TopologicalOffset(*cel, d) : 0;
      }
      p.ncomp = tnc;
      p.fine_stride_j = v.ni;
      p.fine_stride_k = v.ni * v.nj;
      p.fine_stride_c = static_cast<int32_t>(v.comp_stride);
      p.coarse_stride_j = v.cni;
      p.coarse_stride_k = v.cni * v.cnj;
      p.coarse_stride_c = static_cast<int32_t>(v.ccomp_stride);
      p.ndim = ndim;
      p.status = PB2_REGION_ALLOCATED;
      return p;
    };
    auto whole_block = [&](pb2_copy_region &r, const Real *src, Real *dst, bool coarse) {
      r.src = src;
      r.dst = dst;
      r.ncomp = nv.NumComponents();
      r.flag_slot = -1;
      r.status = PB2_REGION_ALLOCATED;
      r.n[0] = coarse ? nv.cni : nv.ni;
      r.n[1] = coarse ? nv.cnj : nv.nj;
      r.n[2] = coarse ? nv.cnk : nv.nk;
      r.src_stride_j = r.dst_stride_j = r.n[0];
      r.src_stride_k = r.dst_stride_k = r.n[0] * r.n[1];
      r.src_stride_c = r.dst_stride_c =
          static_cast<int32_t>(coarse ? nv.ccomp_stride : nv.comp_stride);
    };
    for (size_t i = 0; i < pieces.size(); ++i) {
      const Piece &pc = pieces[i];
      const int src_rank = old_ranklist[pc.old_gid], dst_rank = ranklist[pc.new_gid];
      const bool mine_old = src_rank == my_rank, mine_new = dst_rank == my_rank;
      if (!mine_old && !mine_new) continue;
      if (sparse) { // (single device: both sides are ours)
        const int ol = old_local(pc.old_gid);
        if (!ov.IsAllocated(ol)) continue;
        const int nbi = block_list[pc.new_gid - nslist[my_rank]]->pack_index;
        if (!nv.IsAllocated(nbi)) {
          nv.SetAllocated(nbi, true); // the fresh slab is zero-filled
          new_md->alloc_generation++;
        }
        if (pc.kind == 0) nv.dealloc_count(nbi) = ov.dealloc_count(ol);
      }
      // -- sender side: merged children are restricted first (GetInteriorRestrict), then the
      //    data leaves in a slab if the new owner is another device
      if (mine_old && pc.kind == 2 && te) {
        // RestrictAverage per element over its coarse interior (GetInteriorRestrict)
        const int ol = old_local(pc.old_gid);
        const MeshBlock *ob = old_blocks[ol].get();
        for (size_t e = 0; e < els.size(); ++e) {
          int bs[3], bn[3];
          for (int d = 0; d < 3; ++d) {
            const IndexRange cb = ob->c_cellbounds.Bounds(d, IndexDomain::interior, els[e]);
            bs[d] = cb.s;
            bn[d] = cb.e - cb.s + 1;
          }
          te_restricts.push_back(
              prores_te(ov, ob, ol, static_cast<int>(e), els[e], nullptr, bs, bn));
        }
      } else if (mine_old && pc.kind == 2) {
        const int ol = old_local(pc.old_gid);
        const MeshBlock *ob = old_blocks[ol].get();
        pb2_prores_region p{};
        p.fine = ov.data() + ol * ov.block_stride;
        p.coarse = ov.coarse() + ol * ov.cblock_stride;
        for (int d = 0; d < 3; ++d) {
          const IndexRange cb = ob->c_cellbounds.Bounds(d, IndexDomain::interior);
          p.s[d] = cb.s;
          p.n[d] = cb.e - cb.s + 1;
          p.fine_is[d] = ob->cellbounds.Bounds(d, IndexDomain::interior).s;
          p.coarse_is[d] = cb.s;
          p.fine_xmin[d] = ob->coords.GetXmin()[d];
          p.fine_dx[d] = ob->coords.Dx()[d];
        }
        p.ncomp = ov.NumComponents();
        p.fine_stride_j = ov.ni;
        p.fine_stride_k = ov.ni * ov.nj;
        p.fine_stride_c = static_cast<int32_t>(ov.comp_stride);
        p.coarse_stride_j = ov.cni;
        p.coarse_stride_k = ov.cni * ov.cnj;
        p.coarse_stride_c = static_cast<int32_t>(ov.ccomp_stride);
        p.ndim = ndim;
        p.status = PB2_REGION_ALLOCATED;
        restricts.push_back(p);
      }
      if (mine_old && !mine_new) {
        const int ol = old_local(pc.old_gid);
        pb2_copy_region r{};
        if (pc.kind == 2)
          whole_block(r, ov.coarse() + ol * ov.cblock_stride, send_slab.get<Real>() + piece_send[i], true);
        else
          whole_block(r, ov.data() + ol * ov.block_stride, send_slab.get<Real>() + piece_send[i], false);
        sends.push_back(r);
        continue;
      }
      // -- receiver side: the source is the old slab (same device) or the slab that arrived
      MeshBlock *pmb = block_list[pc.new_gid - nslist[my_rank]].get();
      const int nb = pmb->pack_index;
      const Real *src_fine = mine_old ? ov.data() + old_local(pc.old_gid) * ov.block_stride
                                      : recv_slab.get<Real>() + piece_recv[i];
      const Real *src_coarse = mine_old && pc.kind == 2
                                   ? ov.coarse() + old_local(pc.old_gid) * ov.cblock_stride
                                   : src_fine;
      if (pc.kind == 0) { // kept block: every cell, ghosts included
        pb2_copy_region r{};
        whole_block(r, src_fine, nv.data() + nb * nv.block_stride, false);
        copies.push_back(r);
      } else if (pc.kind == 1) {
        // split: coarse buffer (entire extents) <- the parent's fine data, shifted into the
        // quadrant / octant this child covers (TryRecvCoarseToFine :117-137)
        pb2_copy_region r{};
        whole_block(r, src_fine, nv.coarse() + nb * nv.cblock_stride, true);
        for (int d = 0; d < 3; ++d)
          r.ss[d] = (d < ndim && (pmb->loc.lx[d] & 1)) ? base_block_size.nx_[d] / 2 : 0;
        r.src_stride_j = nv.ni;
        r.src_stride_k = nv.ni * nv.nj;
        r.src_stride_c = static_cast<int32_t>(nv.comp_stride);
        copies.push_back(r);
        if (te) {
          // ProlongateShared per element over its coarse interior +- nghost / 2; the internal
          // elements follow after the first exchange (see below)
          for (size_t e = 0; e < els.size(); ++e) {
            auto box_of = [&](TE el, int bs[3], int bn[3]) {
              for (int d = 0; d < 3; ++d) {
                const IndexRange cb = pmb->c_cellbounds.Bounds(d, IndexDomain::interior, el);
                const int g2 = d < ndim ? ng / 2 : 0;
                bs[d] = cb.s - g2;
                bn[d] = cb.e - cb.s + 1 + 2 * g2;
              }
            };
            int bs[3], bn[3];
            box_of(els[e], bs, bn);
            const int ei = static_cast<int>(e);
            te_prolongs.push_back(prores_te(nv, pmb, nb, ei, els[e], nullptr, bs, bn));
            if (nv.metadata().InternalProlongationOp() == 1) {
              const TE ccel = TE::CC;
              box_of(ccel, bs, bn);
              pb2_prores_region q = prores_te(nv, pmb, nb, ei, els[e], &ccel, bs, bn);
              q.fine -= static_cast<int64_t>(ei) * tnc * nv.comp_stride; // points at element F1
              te_toth_roe.push_back(q);
            } else {
              for (const TE &cel : te_containers)
                if (is_submanifold(els[e], cel)) {
                  box_of(cel, bs, bn);
                  te_internal.push_back(prores_te(nv, pmb, nb, ei, els[e], &cel, bs, bn));
                }
            }
          }
          continue;
        }
        // ProlongateShared over GetInteriorProlongate: coarse interior +- nghost / 2
        // (bnd_info.cpp:199-203)
        pb2_prores_region p{};
        p.fine = nv.data() + nb * nv.block_stride;
        p.coarse = nv.coarse() + nb * nv.cblock_stride;
        const UniformCartesian cc(pmb->coords, 2);
        for (int d = 0; d < 3; ++d) {
          const IndexRange cb = pmb->c_cellbounds.Bounds(d, IndexDomain::interior);
          const int g2 = d < ndim ? ng / 2 : 0;
          p.s[d] = cb.s - g2;
          p.n[d] = cb.e - cb.s + 1 + 2 * g2;
          p.fine_is[d] = pmb->cellbounds.Bounds(d, IndexDomain::interior).s;
          p.coarse_is[d] = cb.s;
          p.fine_xmin[d] = pmb->coords.GetXmin()[d];
          p.fine_dx[d] = pmb->coords.Dx()[d];
          p.coarse_xmin[d] = cc.GetXmin()[d];
          p.coarse_dx[d] = cc.Dx()[d];
        }
        p.ncomp = nv.NumComponents();
        p.fine_stride_j = nv.ni;
        p.fine_stride_k = nv.ni * nv.nj;
        p.fine_stride_c = static_cast<int32_t>(nv.comp_stride);
        p.coarse_stride_j = nv.cni;
        p.coarse_stride_k = nv.cni * nv.cnj;
        p.coarse_stride_c = static_cast<int32_t>(nv.ccomp_stride);
        p.ndim = ndim;
        p.status = PB2_REGION_ALLOCATED;
        prolongs.push_back(p);
      } else {
        // merged: the child's restricted interior lands in the parent's quadrant / octant
        // (TryRecvFineToCoarse :225-246)
        const LogicalLocation dloc = daughter_of(pmb->loc, pc.q);
        if (te) {
          // per element; shared elements come from the upper daughter: the lower one leaves out
          // its last entry along every direction the element is displaced in (:216-228)
          for (size_t e = 0; e < els.size(); ++e) {
            pb2_copy_region r{};
            r.src = src_coarse + static_cast<int64_t>(e) * tnc * nv.ccomp_stride;
            r.dst = nv.data() + nb * nv.block_stride + static_cast<int64_t>(e) * tnc * nv.comp_stride;
            r.ncomp = tnc;
            r.flag_slot = -1;
            r.status = PB2_REGION_ALLOCATED;
            for (int d = 0; d < 3; ++d) {
              const IndexRange cb = pmb->c_cellbounds.Bounds(d, IndexDomain::interior, els[e]);
              const int top = d < ndim ? TopologicalOffset(els[e], d) : 0;
              const bool upper = d < ndim && (dloc.lx[d] & 1);
              const int last = cb.e - (d < ndim && !upper ? top : 0);
              r.ss[d] = cb.s;
              r.n[d] = last - cb.s + 1;
              r.ds[d] = cb.s + (upper ? r.n[d] - top : 0);
            }
            r.src_stride_j = nv.cni;
            r.src_stride_k = nv.cni * nv.cnj;
            r.src_stride_c = static_cast<int32_t>(nv.ccomp_stride);
            r.dst_stride_j = nv.ni;
            r.dst_stride_k = nv.ni * nv.nj;
            r.dst_stride_c = static_cast<int32_t>(nv.comp_stride);
            copies.push_back(r);
          }
          continue;
        }
        pb2_copy_region r{};
        whole_block(r, src_coarse, nv.data() + nb * nv.block_stride, true);
        for (int d = 0; d < 3; ++d) {
          const IndexRange cb = pmb->c_cellbounds.Bounds(d, IndexDomain::interior);
          const bool upper = d < ndim && (dloc.lx[d] & 1);
          r.ss[d] = cb.s;
          r.ds[d] = cb.s + (upper ? cb.e - cb.s + 1 : 0);
          r.n[d] = cb.e - cb.s + 1;
        }
        r.dst_stride_j = nv.ni;
        r.dst_stride_k = nv.ni * nv.nj;
        r.dst_stride_c = static_cast<int32_t>(nv.comp_stride);
        copies.push_back(r);
      }
    }
    pb2_bnd_table *t_res = nullptr, *t_send = nullptr, *t_copy = nullptr, *t_pro = nullptr;
    pb2_bnd_table *t_res_te = nullptr, *t_pro_te = nullptr;
    PB2_CHECK(pb2_prores_table_create(&t_res_te, te_restricts.data(),
                                      static_cast<int64_t>(te_restricts.size())));
    PB2_CHECK(pb2_prores_table_create(&t_pro_te, te_prolongs.data(),
                                      static_cast<int64_t>(te_prolongs.size())));
    PB2_CHECK(pb2_prores_table_create(&t_res, restricts.data(), static_cast<int64_t>(restricts.size())));
    PB2_CHECK(pb2_copy_table_create(&t_send, sends.data(), static_cast<int64_t>(sends.size())));
    PB2_CHECK(pb2_copy_table_create(&t_copy, copies.data(), static_cast<int64_t>(copies.size())));
    PB2_CHECK(pb2_prores_table_create(&t_pro, prolongs.data(), static_cast<int64_t>(prolongs.size())));
    PB2_CHECK(pb2_restrict(t_res, st));
    PB2_CHECK(pb2_restrict_te(t_res_te, st));
    PB2_CHECK(pb2_copy(t_send, nullptr, st));
    if (nranks > 1) // blocks that change device: one grouped NCCL send/recv per peer
      PB2_CHECK(pb2_comm_exchange(comm, send_slab.get<Real>(), send_off.data(),
                                  recv_slab.get<Real>(), recv_off.data(), st));
    PB2_CHECK(pb2_copy(t_copy, nullptr, st));
    PB2_CHECK(pb2_prolongate(t_pro, nv.metadata().ProlongationOp(), st));
    PB2_CHECK(pb2_prolongate_te(t_pro_te, nv.metadata().ProlongationOp(), st));
    PB2_CHECK(pb2_stream_sync(st));
    pb2_bnd_table_destroy(t_res_te);
    pb2_bnd_table_destroy(t_pro_te);
    pb2_bnd_table_destroy(t_res);
    pb2_bnd_table_destroy(t_send);
    pb2_bnd_table_destroy(t_copy);
    pb2_bnd_table_destroy(t_pro);
  }
  const auto t3 = now();
  old_md.reset(); // the old slabs go away here
  const auto t4 = now();
  if (has_te) {
    // A newly refined block may have neighbours that were fine already: its prolongated shared
    // elements must give way to theirs.  One exchange with an ownership that ranks new fine
    // blocks below old ones (:958-984), then the internal elements of the new blocks (:986-990)
    for (int n = 0; n < nbtotal; ++n) {
      const LogicalLocation &nl = loclist[n];
      if (old_gid.count(nl) == 0 && nl.level > 0 && old_gid.count(parent_of(nl)) > 0)
        newly_refined_.insert(nl);
    }
    ownership_.clear();
    CommunicateBoundaries(new_md, true);
    pb2_bnd_table *t_int = nullptr, *t_tr = nullptr;
    PB2_CHECK(pb2_prores_table_create(&t_int, te_internal.data(),
                                      static_cast<int64_t>(te_internal.size())));
    PB2_CHECK(pb2_prores_table_create(&t_tr, te_toth_roe.data(),
                                      static_cast<int64_t>(te_toth_roe.size())));
    PB2_CHECK(pb2_prolongate_internal(t_int, st));
    PB2_CHECK(pb2_prolongate_toth_roe(t_tr, st));
    PB2_CHECK(pb2_stream_sync(st));
    pb2_bnd_table_destroy(t_int);
    pb2_bnd_table_destroy(t_tr);
    // the regular ownership again (:992-996): plan and tables of the exchange are rebuilt
    newly_refined_.clear();
    ownership_.clear();
    plan_cache.clear(); // (plans are shared through the mesh: the one just used ranked the new
                        // blocks below the old ones and must not be found again)
    new_md->bvars().Invalidate();
  }
  // PreCommFillDerived; CommunicateBoundaries; FillDerived (:1000-1003)
  Update::PreCommFillDerived(new_md.get());
  CommunicateBoundaries(new_md, true);
  Update::FillDerived(new_md.get());
  PB2_CHECK(pb2_stream_sync(st));
  if (timing)
    std::fprintf(stderr, "remesh: purge+tree+blocks %.2f ms, new container %.2f, move data %.2f, "
                         "free old %.2f, exchange+derived %.2f (nblocks %d)\n",
                 ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4), ms(t4, now()), nbtotal);
}

void Mesh::LoadBalancingAndAdaptiveMeshRefinement(ParameterInput *, ApplicationInput *) {
  modified = false;
  if (!adaptive) return;
  int nnew = 0, ndel = 0;
  std::vector<LogicalLocation> new_leaves;
  if (!UpdateMeshBlockTree(new_leaves, nnew, ndel)) return;
  nbnew += nnew;
  nbdel += ndel;
  RedistributeAndRefineMeshBlocks(new_leaves);
  modified = true;
}

namespace Refinement {
// amr_criteria/refinement_package.cpp:150-172: every package votes, the strongest tag wins
// (CheckAllRefinement :40-76), then MeshRefinement::SetRefinement
TaskStatus Tag(MeshData<Real> *rc) {
  Mesh *pm = rc->GetMeshPointer();
  std::vector<AmrTag> tags(rc->NumBlocks(), AmrTag::derefine);
  bool any = false;
  for (const auto &pkg : pm->packages.AllPackages()) {
    if (pkg.second->CheckRefinementMesh == nullptr) continue;
    std::vector<AmrTag> t(rc->NumBlocks(), AmrTag::derefine);
    pkg.second->CheckRefinementMesh(rc, t);
    for (size_t b = 0; b < t.size(); ++b) tags[b] = any ? std::max(tags[b], t[b]) : t[b];
    any = true;
  }
  // the stock criteria of <parthenon/refinementN> (:72-88): a "refine" at or above the
  // criterion's max level counts as "same"
  if (!pm->amr_criteria.empty()) {
    const int nb = rc->NumBlocks();
    DeviceBuffer dev;
    dev.Allocate(sizeof(Real) * nb, rc->stream());
    std::vector<Real> maxd(nb);
    for (const auto &c : pm->amr_criteria) {
      if (!rc->HasVariable(c.field)) {
        // amr_criteria.cpp:89-91: a block without the field votes "same" (not "derefine")
        for (int b = 0; b < nb; ++b) tags[b] = std::max(tags[b], AmrTag::same);
        continue;
      }
      Variable &v = rc->Get(c.field);
      const pb2_pack_geom g = rc->Geometry(v);
      PB2_CHECK(pb2_block_derivative(&g, v.data(), c.comp, c.order, dev.get<Real>(), rc->stream()));
      PB2_CHECK(pb2_memcpy_d2h(maxd.data(), dev.get(), sizeof(Real) * nb, rc->stream()));
      PB2_CHECK(pb2_stream_sync(rc->stream()));
      for (int b = 0; b < nb; ++b) {
        if (!v.IsAllocated(b)) { // ... and so does a block on which it is not allocated
          tags[b] = std::max(tags[b], AmrTag::same);
          continue;
        }
        AmrTag t = maxd[b] > c.refine_criteria
                       ? AmrTag::refine
                       : (maxd[b] < c.derefine_criteria ? AmrTag::derefine : AmrTag::same);
        if (t == AmrTag::refine && rc->GetBlock(b)->loc.level >= c.max_level) t = AmrTag::same;
        tags[b] = std::max(tags[b], t);
      }
    }
    any = true;
  }
  if (!any) std::fill(tags.begin(), tags.end(), AmrTag::derefine);
  for (int b = 0; b < rc->NumBlocks(); ++b) pm->SetRefinement(rc->GetBlock(b)->lid, tags[b]);
  return TaskStatus::complete;
}
} // namespace Refinement

} // namespace parthenon
