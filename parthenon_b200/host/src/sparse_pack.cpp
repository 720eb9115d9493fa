// sparse_pack.cpp — descriptor resolution and table build of SparsePack (see pb2/sparse_pack.hpp).
#include "pb2/sparse_pack.hpp"

#include <algorithm>

namespace parthenon {

namespace impl {

PackDescriptor::PackDescriptor(const std::vector<const StateDescriptor *> &packages,
                               const std::vector<std::string> &group_names,
                               const Selector &selector, const std::set<PDOpt> &options)
    : nvar_groups(static_cast<int>(group_names.size())), var_group_names(group_names),
      var_groups(group_names.size()), with_fluxes(options.count(PDOpt::WithFluxes) > 0),
      coarse(options.count(PDOpt::Coarse) > 0), flat(options.count(PDOpt::Flatten) > 0) {
  PARTHENON_REQUIRE(!(with_fluxes && coarse),
                    "Probably shouldn't be making a coarse pack with fine fluxes.");
  struct Key {
    std::string base;
    int sparse_id;
    std::string label;
  };
  std::vector<std::vector<Key>> groups(nvar_groups);
  for (const StateDescriptor *psd : packages) {
    for (const FieldEntry &f : psd->AllFields()) {
      for (int i = 0; i < nvar_groups; ++i) {
        if (!selector(i, f)) continue;
        std::string base = f.name;
        if (f.sparse_id >= 0) base = f.name.substr(0, f.name.rfind('_'));
        groups[i].push_back(Key{base, f.sparse_id, f.name});
      }
    }
  }
  for (int i = 0; i < nvar_groups; ++i) {
    std::stable_sort(groups[i].begin(), groups[i].end(), [](const Key &a, const Key &b) {
      if (a.base == b.base) return a.sparse_id < b.sparse_id;
      return a.base < b.base;
    });
    for (const Key &k : groups[i]) {
      var_groups[i].push_back(k.label);
      identifier += k.label + "_";
      ++nvar_tot;
    }
    identifier += "|";
  }
  identifier += std::to_string(with_fluxes) + std::to_string(coarse) + std::to_string(flat);
}

} // namespace impl

SparsePack::Descriptor MakePackDescriptor(const std::vector<const StateDescriptor *> &packages,
                                          const std::vector<std::string> &vars,
                                          const std::vector<bool> &use_regex,
                                          const std::vector<MetadataFlag> &flags,
                                          const std::set<PDOpt> &options) {
  PARTHENON_REQUIRE(vars.size() == use_regex.size(),
                    "Vargroup names and use_regex need to be the same size.");
  auto selector = [&](int vidx, const FieldEntry &f) {
    for (const auto &flag : flags)
      if (!f.m.IsSet(flag)) return false;
    if (use_regex[vidx]) return std::regex_match(f.name, std::regex(vars[vidx]));
    if (vars[vidx] == f.name) return true;
    // a sparse pool is selected by its base name (make_pack_descriptor.hpp:61)
    return f.sparse_id >= 0 && vars[vidx] == f.name.substr(0, f.name.rfind('_'));
  };
  return SparsePack::Descriptor(impl::PackDescriptor(packages, vars, selector, options));
}

SparsePack::Descriptor MakePackDescriptor(MeshData<Real> *md, const std::vector<std::string> &vars,
                                          const std::vector<MetadataFlag> &flags,
                                          const std::set<PDOpt> &options) {
  std::vector<const StateDescriptor *> pk;
  const Packages_t &packages = md->GetMeshPointer()->packages;
  for (const std::string &name : packages.Order()) pk.push_back(packages.Get(name).get());
  return MakePackDescriptor(pk, vars, std::vector<bool>(vars.size(), false), flags, options);
}

namespace {

struct Entry { // one component of the pack on one block
  Variable *v;
  int comp;
};

std::shared_ptr<SparsePackStorage> Build(MeshData<Real> *md, const impl::PackDescriptor &desc,
                                         const std::vector<bool> &include_block,
                                         const std::vector<uint8_t> &alloc_status) {
  auto s = std::make_shared<SparsePackStorage>();
  s->alloc_status = alloc_status;
  s->include_block = include_block;
  const int nvar = desc.nvar_groups;
  // blocks of the pack (sparse_pack_base.cpp ForEachBlock)
  std::vector<int> blocks;
  for (int b = 0; b < md->NumBlocks(); ++b)
    if (include_block.empty() || include_block[b]) blocks.push_back(b);
  const int nb = static_cast<int>(blocks.size());

  // resolve the groups against the container: variables it does not hold are skipped
  std::vector<std::vector<Variable *>> groups(nvar);
  TopologicalType tt = TopologicalType::Cell;
  bool have_tt = false;
  Variable *first = nullptr;
  for (int i = 0; i < nvar; ++i) {
    for (const std::string &label : desc.var_groups[i]) {
      if (!md->HasVariable(label)) continue;
      Variable *v = &md->Get(label);
      groups[i].push_back(v);
      const TopologicalType t = v->topological_type();
      const bool cell_like = t == TopologicalType::Cell;
      if (!have_tt) {
        tt = t;
        have_tt = true;
        first = v;
      } else {
        PARTHENON_REQUIRE((tt == TopologicalType::Cell) == cell_like && v->ni == first->ni &&
                              v->nj == first->nj && v->nk == first->nk,
                          "a pack holds fields of one topological type (their arrays share "
                          "one set of extents): " + v->label());
      }
    }
  }
  const bool multi_el = have_tt && (tt == TopologicalType::Face || tt == TopologicalType::Edge);
  // leading ("type") dimension: the field (3 element arrays for face / edge fields), then fluxes
  const int flx_idx = multi_el ? 3 : 1;
  const int ntypes = flx_idx + (desc.with_fluxes ? 3 : 0);

  // pass 1: sizes and bounds
  std::vector<std::vector<Entry>> entries(nb); // per pack block (non-flat) in pack-index order
  std::vector<Entry> flat_entries;
  std::vector<int> flat_block; // block of every flat entry (for coords)
  s->bounds_h.assign(static_cast<size_t>(2) * nb * (nvar + 1), 0);
  auto bnd = [&](int w, int b, int v) -> int32_t & {
    return s->bounds_h[(static_cast<size_t>(w) * nb + b) * (nvar + 1) + v];
  };
  int idx = 0, max_size = 0, total = 0;
  for (int bi = 0; bi < nb; ++bi) {
    const int b = blocks[bi];
    if (!desc.flat) idx = 0;
    for (int i = 0; i < nvar; ++i) {
      bnd(0, bi, i) = idx;
      for (Variable *v : groups[i]) {
        if (!v->IsAllocated(b)) continue;
        // the tensor components of one element; face / edge fields put their elements in the
        // type dimension
        const int nc = multi_el ? v->TensorComponents() : v->NumComponents();
        for (int c = 0; c < nc; ++c) {
          if (desc.flat) {
            flat_entries.push_back(Entry{v, c});
            flat_block.push_back(b);
          } else {
            entries[bi].push_back(Entry{v, c});
          }
          ++idx;
          ++total;
        }
      }
      bnd(1, bi, i) = idx - 1;
      if (bnd(1, bi, i) < bnd(0, bi, i)) { // nothing allocated that meets the criteria
        bnd(0, bi, i) = -1;
        bnd(1, bi, i) = -2;
      }
    }
    bnd(1, bi, nvar) = idx - 1;
    max_size = std::max(max_size, idx);
  }
  const int pack_blocks = desc.flat ? 1 : nb;
  const int maxvars = std::max(max_size, 1);

  // pass 2: pointer table, labels, coordinates, block properties
  std::vector<Real *> ptr(static_cast<size_t>(ntypes) * pack_blocks * maxvars, nullptr);
  s->labels_h.assign(static_cast<size_t>(pack_blocks) * maxvars, "");
  auto fill = [&](int pb, int n, const Entry &e, int b) {
    Variable &v = *e.v;
    Real *base = desc.coarse ? v.coarse() : v.data();
    const int64_t bs = desc.coarse ? v.cblock_stride : v.block_stride;
    const int64_t cs = desc.coarse ? v.ccomp_stride : v.comp_stride;
    const int nel = multi_el ? v.NumElements() : 1;
    const int ntc = v.TensorComponents();
    for (int el = 0; el < nel; ++el) {
      const int comp = multi_el ? el * ntc + e.comp : e.comp;
      ptr[(static_cast<size_t>(el) * pack_blocks + pb) * maxvars + n] = base + b * bs + comp * cs;
    }
    // (the flux of a face field is the separate edge field "bnd_flux::<name>": pack it by name)
    if (desc.with_fluxes && v.IsSet(Metadata::WithFluxes) &&
        v.topological_type() == TopologicalType::Cell) {
      for (int d = 1; d <= 3; ++d) {
        if (d > md->GetMeshPointer()->ndim) continue;
        ptr[(static_cast<size_t>(flx_idx + d - 1) * pack_blocks + pb) * maxvars + n] =
            v.flux(d) + b * v.block_stride + e.comp * v.comp_stride;
      }
    }
    s->labels_h[static_cast<size_t>(pb) * maxvars + n] = v.label();
  };
  std::vector<pb2_pack_coords> coords(desc.flat ? maxvars : std::max(nb, 1));
  auto block_coords = [&](int b) {
    const MeshBlock *pmb = md->GetBlock(b);
    pb2_pack_coords c{};
    for (int d = 0; d < 3; ++d) {
      c.xmin[d] = pmb->block_size.xmin_[d];
      c.dx[d] = pmb->coords.Dx()[d] * (desc.coarse && !pmb->block_size.symmetry_[d] ? 2 : 1);
    }
    return c;
  };
  if (desc.flat) {
    for (size_t n = 0; n < flat_entries.size(); ++n) {
      fill(0, static_cast<int>(n), flat_entries[n], flat_block[n]);
      coords[n] = block_coords(flat_block[n]);
    }
  } else {
    for (int bi = 0; bi < nb; ++bi) {
      for (size_t n = 0; n < entries[bi].size(); ++n)
        fill(bi, static_cast<int>(n), entries[bi][n], blocks[bi]);
      coords[bi] = block_coords(blocks[bi]);
    }
  }
  s->block_props_h.assign(static_cast<size_t>(std::max(nb, 1)) * 28, 0);
  for (int bi = 0; bi < nb; ++bi) {
    const MeshBlock *pmb = md->GetBlock(blocks[bi]);
    for (int n = 0; n < 27; ++n) s->block_props_h[bi * 28 + n] = pmb->loc.level;
    s->block_props_h[bi * 28 + 27] = pmb->gid;
    for (const NeighborBlock &nbk : pmb->neighbors)
      s->block_props_h[bi * 28 + (nbk.offsets[2] + 1) +
                       3 * ((nbk.offsets[1] + 1) + 3 * (nbk.offsets[0] + 1))] = nbk.loc.level;
  }

  pb2_stream_t st = md->stream();
  auto upload = [&](DeviceBuffer &d, const void *src, size_t bytes) {
    d.Allocate(std::max<size_t>(bytes, 8), st);
    if (bytes) PB2_CHECK(pb2_memcpy_h2d(d.get(), src, bytes, st));
  };
  upload(s->ptr, ptr.data(), ptr.size() * sizeof(Real *));
  upload(s->bounds, s->bounds_h.data(), s->bounds_h.size() * sizeof(int32_t));
  upload(s->coords, coords.data(), coords.size() * sizeof(pb2_pack_coords));
  upload(s->block_props, s->block_props_h.data(), s->block_props_h.size() * sizeof(int32_t));
  PB2_CHECK(pb2_stream_sync(st)); // the host vectors go out of scope

  pb2_sparse_pack &p = s->view;
  p.ptr = s->ptr.get<double *const>();
  p.bounds = s->bounds.get<int32_t>();
  p.coords = s->coords.get<pb2_pack_coords>();
  p.nblocks = pack_blocks;
  p.nblocks_md = nb;
  p.maxvars = maxvars;
  p.nvar = nvar;
  p.size = total;
  p.flat = desc.flat;
  p.with_fluxes = desc.with_fluxes;
  p.coarse = desc.coarse;
  const MeshBlock *pmb0 = md->NumBlocks() ? md->GetBlock(0) : nullptr;
  if (first && pmb0) {
    p.ni = desc.coarse ? first->cni : first->ni;
    p.nj = desc.coarse ? first->cnj : first->nj;
    p.nk = desc.coarse ? first->cnk : first->nk;
    const IndexShape &cb = desc.coarse ? pmb0->c_cellbounds : pmb0->cellbounds;
    p.is = cb.is(IndexDomain::interior);
    p.ie = cb.ie(IndexDomain::interior);
    p.js = cb.js(IndexDomain::interior);
    p.je = cb.je(IndexDomain::interior);
    p.ks = cb.ks(IndexDomain::interior);
    p.ke = cb.ke(IndexDomain::interior);
  }
  return s;
}

} // namespace

SparsePack SparsePack::Descriptor::GetPack(MeshData<Real> *md,
                                           const std::vector<bool> &include_block) const {
  PARTHENON_REQUIRE(include_block.empty() ||
                        static_cast<int>(include_block.size()) == md->NumBlocks(),
                    "Passed wrong size block include list.");
  // allocation status of every selected variable on every block (SparsePackBase::GetAllocStatus)
  std::vector<uint8_t> status;
  for (int b = 0; b < md->NumBlocks(); ++b) {
    if (!include_block.empty() && !include_block[b]) continue;
    for (const auto &group : var_groups)
      for (const std::string &label : group)
        status.push_back(md->HasVariable(label) && md->Get(label).IsAllocated(b) ? 1 : 0);
  }
  auto &cache = md->GetSparsePackCache();
  auto it = cache.find(identifier);
  if (it != cache.end() && it->second->alloc_status == status &&
      it->second->include_block == include_block)
    return SparsePack(it->second);
  auto s = Build(md, *this, include_block, status);
  cache[identifier] = s;
  return SparsePack(s);
}

} // namespace parthenon
