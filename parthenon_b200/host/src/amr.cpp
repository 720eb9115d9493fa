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
  int tnderef = 0;
  for (auto &pmb : block_list) tnderef += pmb->refine_flag == -1;
  for (auto &pmb : block_list) {
    if (pmb->refine_flag == 1) lref.push_back(pmb->loc);
    if (pmb->refine_flag == -1 && tnderef >= nleaf) lderef.push_back(pmb->loc);
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

void Mesh::RedistributeAndRefineMeshBlocks(const std::vector<LogicalLocation> &new_leaves) {
  PARTHENON_REQUIRE(nranks == 1 && DefaultNumPartitions() == 1,
                    "remeshing needs one device and one MeshData per rank (pack_size = -1)");
  // only the base container survives a remesh (:667)
  mesh_data.PurgeNonBase();
  std::shared_ptr<MeshData<Real>> old_md = mesh_data.GetOrAdd("base", 0);
  const BlockList_t old_blocks = block_list;
  std::unordered_map<LogicalLocation, int, LogicalLocationHash> old_index;
  for (auto &pmb : old_blocks) old_index[pmb->loc] = pmb->pack_index;

  // new tree, block list and (empty) base container
  BuildTree(nullptr, new_leaves);
  multilevel = true;
  std::vector<double> cost(nbtotal, 1.0);
  AssignBlocks(cost, nranks, ranklist);
  nblist.assign(nranks, 0);
  for (int r : ranklist) nblist[r]++;
  nslist.assign(nranks, 0);
  for (int r = 1; r < nranks; ++r) nslist[r] = nslist[r - 1] + nblist[r - 1];
  BuildBlockList(&old_blocks);
  for (auto &pmb : block_list)
    if (!old_index.count(pmb->loc)) { // MeshBlock::Make: fresh counters, no time step yet
      pmb->refine_flag = 0;
      pmb->deref_count = 0;
      pmb->SetAllowedDt(std::numeric_limits<Real>::max());
    }
  mesh_data.Clear();
  std::shared_ptr<MeshData<Real>> new_md = mesh_data.GetOrAdd("base", 0);

  const int nleaf = 1 << ndim;
  const int ng = Globals::nghost;
  pb2_stream_t st = stream;
  for (auto &nvp : new_md->GetVariableVector()) {
    Variable &nv = *nvp;
    // the fields a remesh carries over: pmb->vars_cc_ (meshblock.cpp: Independent / FillGhost /
    // ForceRemeshComm cell-centred variables)
    if (!(nv.IsSet(Metadata::Independent) || nv.IsSet(Metadata::FillGhost))) continue;
    PARTHENON_REQUIRE(!nv.metadata().IsSparse(), "sparse fields cannot be remeshed in this build");
    Variable &ov = old_md->Get(nv.label());
    std::vector<pb2_copy_region> copies;
    std::vector<pb2_prores_region> restricts, prolongs;
    auto whole = [&](pb2_copy_region &r) {
      r.ncomp = nv.NumComponents();
      r.flag_slot = -1;
      r.status = PB2_REGION_ALLOCATED;
    };
    for (auto &pmb : block_list) {
      const int nb = pmb->pack_index;
      auto same = old_index.find(pmb->loc);
      if (same != old_index.end()) {
        // kept block: every cell, ghosts included
        pb2_copy_region r{};
        whole(r);
        r.src = ov.data() + same->second * ov.block_stride;
        r.dst = nv.data() + nb * nv.block_stride;
        r.n[0] = nv.ni;
        r.n[1] = nv.nj;
        r.n[2] = nv.nk;
        r.src_stride_j = r.dst_stride_j = nv.ni;
        r.src_stride_k = r.dst_stride_k = nv.ni * nv.nj;
        r.src_stride_c = r.dst_stride_c = static_cast<int32_t>(nv.comp_stride);
        copies.push_back(r);
        continue;
      }
      LogicalLocation par = pmb->loc.GetParent();
      for (int d = ndim; d < 3; ++d) par.lx[d] = 0;
      auto up = pmb->loc.level > 0 ? old_index.find(par) : old_index.end();
      if (up != old_index.end()) {
        // split: coarse buffer (entire extents) <- the parent's fine data, shifted into the
        // quadrant / octant this child covers (TryRecvCoarseToFine :117-137)
        pb2_copy_region r{};
        whole(r);
        r.src = ov.data() + up->second * ov.block_stride;
        r.dst = nv.coarse() + nb * nv.cblock_stride;
        for (int d = 0; d < 3; ++d) {
          const bool upper = d < ndim && (pmb->loc.lx[d] & 1);
          r.ss[d] = upper ? base_block_size.nx_[d] / 2 : 0;
          r.ds[d] = 0;
        }
        r.n[0] = nv.cni;
        r.n[1] = nv.cnj;
        r.n[2] = nv.cnk;
        r.src_stride_j = nv.ni;
        r.src_stride_k = nv.ni * nv.nj;
        r.src_stride_c = static_cast<int32_t>(nv.comp_stride);
        r.dst_stride_j = nv.cni;
        r.dst_stride_k = nv.cni * nv.cnj;
        r.dst_stride_c = static_cast<int32_t>(nv.ccomp_stride);
        copies.push_back(r);
        // ProlongateShared over GetInteriorProlongate: coarse interior +- nghost / 2
        // (bnd_info.cpp:199-203)
        pb2_prores_region p{};
        p.fine = nv.data() + nb * nv.block_stride;
        p.coarse = nv.coarse() + nb * nv.cblock_stride;
        for (int d = 0; d < 3; ++d) {
          const IndexRange cb = pmb->c_cellbounds.Bounds(d, IndexDomain::interior);
          const int g2 = d < ndim ? ng / 2 : 0;
          p.s[d] = cb.s - g2;
          p.n[d] = cb.e - cb.s + 1 + 2 * g2;
          p.fine_is[d] = pmb->cellbounds.Bounds(d, IndexDomain::interior).s;
          p.coarse_is[d] = cb.s;
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
        const UniformCartesian cc(pmb->coords, 2);
        for (int d = 0; d < 3; ++d) {
          p.fine_xmin[d] = pmb->coords.GetXmin()[d];
          p.fine_dx[d] = pmb->coords.Dx()[d];
          p.coarse_xmin[d] = cc.GetXmin()[d];
          p.coarse_dx[d] = cc.Dx()[d];
        }
        prolongs.push_back(p);
        continue;
      }
      // merged: every child restricts its interior into ITS coarse buffer (GetInteriorRestrict),
      // which then lands in the parent's quadrant / octant (TryRecvFineToCoarse :225-246)
      for (int q = 0; q < nleaf; ++q) {
        const LogicalLocation dloc = Daughter(pmb->loc, q, ndim);
        auto ch = old_index.find(dloc);
        PARTHENON_REQUIRE(ch != old_index.end(), "remesh cannot find the origin of a new block");
        const MeshBlock *ob = old_blocks[ch->second].get();
        pb2_prores_region p{};
        p.fine = ov.data() + ch->second * ov.block_stride;
        p.coarse = ov.coarse() + ch->second * ov.cblock_stride;
        pb2_copy_region r{};
        whole(r);
        r.src = p.coarse;
        r.dst = nv.data() + nb * nv.block_stride;
        for (int d = 0; d < 3; ++d) {
          const IndexRange cb = ob->c_cellbounds.Bounds(d, IndexDomain::interior);
          p.s[d] = cb.s;
          p.n[d] = cb.e - cb.s + 1;
          p.fine_is[d] = ob->cellbounds.Bounds(d, IndexDomain::interior).s;
          p.coarse_is[d] = cb.s;
          const bool upper = d < ndim && (dloc.lx[d] & 1);
          r.ss[d] = cb.s;
          r.ds[d] = cb.s + (upper ? cb.e - cb.s + 1 : 0);
          r.n[d] = cb.e - cb.s + 1;
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
        for (int d = 0; d < 3; ++d) {
          p.fine_xmin[d] = ob->coords.GetXmin()[d];
          p.fine_dx[d] = ob->coords.Dx()[d];
        }
        restricts.push_back(p);
        r.src_stride_j = ov.cni;
        r.src_stride_k = ov.cni * ov.cnj;
        r.src_stride_c = static_cast<int32_t>(ov.ccomp_stride);
        r.dst_stride_j = nv.ni;
        r.dst_stride_k = nv.ni * nv.nj;
        r.dst_stride_c = static_cast<int32_t>(nv.comp_stride);
        copies.push_back(r);
      }
    }
    pb2_bnd_table *t_res = nullptr, *t_copy = nullptr, *t_pro = nullptr;
    PB2_CHECK(pb2_prores_table_create(&t_res, restricts.data(), static_cast<int64_t>(restricts.size())));
    PB2_CHECK(pb2_copy_table_create(&t_copy, copies.data(), static_cast<int64_t>(copies.size())));
    PB2_CHECK(pb2_prores_table_create(&t_pro, prolongs.data(), static_cast<int64_t>(prolongs.size())));
    PB2_CHECK(pb2_restrict(t_res, st));
    PB2_CHECK(pb2_copy(t_copy, nullptr, st));
    PB2_CHECK(pb2_prolongate(t_pro, nv.metadata().ProlongationOp(), st));
    PB2_CHECK(pb2_stream_sync(st));
    pb2_bnd_table_destroy(t_res);
    pb2_bnd_table_destroy(t_copy);
    pb2_bnd_table_destroy(t_pro);
  }
  old_md.reset(); // the old slabs go away here
  // PreCommFillDerived; CommunicateBoundaries; FillDerived (:1000-1003)
  Update::PreCommFillDerived(new_md.get());
  CommunicateBoundaries(new_md, true);
  Update::FillDerived(new_md.get());
  PB2_CHECK(pb2_stream_sync(st));
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
  if (!any) std::fill(tags.begin(), tags.end(), AmrTag::same);
  for (int b = 0; b < rc->NumBlocks(); ++b) pm->SetRefinement(rc->GetBlock(b)->lid, tags[b]);
  return TaskStatus::complete;
}
} // namespace Refinement

} // namespace parthenon
