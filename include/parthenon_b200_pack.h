/* parthenon_b200_pack.h — device-side variable packs for user kernels.
 *
 * What a downstream kernel includes to index a pack of fields over every block of a MeshData
 * batch the way Parthenon's SparsePack is indexed (src/interface/sparse_pack.hpp:53-420):
 *
 *     pack(b, n, k, j, i)              component n of block b
 *     pack(b, PackIdx, k, j, i)        by variable (descriptor order) + component offset
 *     pack.flux(b, dir, n, k, j, i)    face flux, dir = 1..3
 *     pack.GetLowerBound(b, var) / GetUpperBound(b, var) / Contains(b, var) / GetSize(b, var)
 *     pack.GetNBlocks(), GetMaxNumberOfVars(), GetSize(), GetCoordinates(b)
 *
 * The reference keeps, per (block, component), a 136-byte view handle that every access
 * dereferences (variable_pack.hpp:259).  Fields here live in one slab per field
 * ([block][component][k][j][i]), so a pack is a table of ONE pointer per (block, pack index) —
 * null where a sparse variable is not allocated on that block — plus the reference's bounds
 * table; the cell index is plain arithmetic on extents that are the same for every entry.
 *
 * The struct is POD and is passed to kernels BY VALUE; the tables it points to are owned by the
 * host-side parthenon::SparsePack (host/pb2/sparse_pack.hpp) and stay valid until the
 * MeshData's allocation status changes (the host object is then rebuilt from its descriptor).
 * Plain C declarations first (the C ABI hands this struct out: pb2h_sim_sparse_pack), the
 * accessors follow for C++ / CUDA.
 */
#ifndef PARTHENON_B200_PACK_H_
#define PARTHENON_B200_PACK_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pb2_pack_coords {
  double xmin[3]; /* lower corner of the block's interior */
  double dx[3];   /* cell widths */
} pb2_pack_coords;

typedef struct pb2_sparse_pack {
  /* [ntypes][nblocks][maxvars] component base pointers (device), type 0 = the field, types
   * 1..3 = its x1 / x2 / x3 face fluxes (only with PDOpt::WithFluxes); NULL = not allocated */
  double *const *ptr;
  /* [2][nblocks][nvar + 1] inclusive component ranges (device): bounds[0] = lower, [1] = upper
   * per (block, variable); entry nvar = the whole block.  An absent variable has lower = -1,
   * upper = -2 as in the reference (sparse_pack_base.cpp:222-228). */
  const int32_t *bounds;
  const pb2_pack_coords *coords; /* [nblocks] (device) */
  int32_t nblocks;   /* 1 for a flattened pack */
  int32_t nblocks_md; /* blocks of the MeshData (== nblocks unless flattened) */
  int32_t maxvars;   /* extent of the pack-index dimension */
  int32_t nvar;      /* variables in the descriptor */
  int32_t size;      /* total components in the pack */
  int32_t flat, with_fluxes, coarse;
  int32_t ni, nj, nk;             /* array extents of one component */
  int32_t is, ie, js, je, ks, ke; /* interior index bounds (inclusive) */
} pb2_sparse_pack;

#ifdef __cplusplus
} /* extern "C" */

#if defined(__CUDACC__)
#define PB2_PACK_HD __host__ __device__ __forceinline__
#else
#define PB2_PACK_HD inline
#endif

namespace pb2 {

/* sparse_pack_base.hpp: index of a variable in the descriptor + a component offset */
struct PackIdx {
  int var, off;
  PB2_PACK_HD explicit PackIdx(int v, int o = 0) : var(v), off(o) {}
  PB2_PACK_HD PackIdx operator+(int o) const { return PackIdx(var, off + o); }
  PB2_PACK_HD int VariableIdx() const { return var; }
  PB2_PACK_HD int Offset() const { return off; }
};

struct SparsePackView : pb2_sparse_pack {
  SparsePackView() = default;
  PB2_PACK_HD explicit SparsePackView(const pb2_sparse_pack &p) : pb2_sparse_pack(p) {}

  PB2_PACK_HD int GetNBlocks() const { return nblocks; }
  PB2_PACK_HD int GetMaxNumberOfVars() const { return maxvars; }
  PB2_PACK_HD int GetSize() const { return size; }
  PB2_PACK_HD const pb2_pack_coords &GetCoordinates(int b = 0) const { return coords[b]; }

  /* bounds of block b (sparse_pack.hpp:131-147); for a flattened pack b counts the MeshData's
   * blocks inside the single unified index space */
  PB2_PACK_HD int bound(int which, int b, int v) const {
    return bounds[((int64_t)which * nblocks_md + b) * (nvar + 1) + v];
  }
  PB2_PACK_HD int GetLowerBound(int b) const {
    return (flat && b > 0) ? bound(1, b - 1, nvar) + 1 : 0;
  }
  PB2_PACK_HD int GetUpperBound(int b) const { return bound(1, b, nvar); }
  PB2_PACK_HD int GetLowerBound(int b, PackIdx v) const { return bound(0, b, v.var); }
  PB2_PACK_HD int GetUpperBound(int b, PackIdx v) const { return bound(1, b, v.var); }
  PB2_PACK_HD int GetSize(int b, PackIdx v) const {
    return GetUpperBound(b, v) - GetLowerBound(b, v) + 1;
  }
  PB2_PACK_HD bool Contains(int b) const { return GetUpperBound(b) >= 0; }
  PB2_PACK_HD bool Contains(int b, PackIdx v) const { return GetUpperBound(b, v) >= 0; }

  /* component base pointer (NULL if the variable is not allocated on that block) */
  PB2_PACK_HD double *Component(int type, int b, int n) const {
    return ptr[((int64_t)type * nblocks + b) * maxvars + n];
  }
  PB2_PACK_HD bool IsAllocated(int b, int n) const { return Component(0, b, n) != nullptr; }
  PB2_PACK_HD int64_t Cell(int k, int j, int i) const {
    return ((int64_t)k * nj + j) * ni + i;
  }

  PB2_PACK_HD double &operator()(int b, int n, int k, int j, int i) const {
    return Component(0, b, n)[Cell(k, j, i)];
  }
  PB2_PACK_HD double &operator()(int b, PackIdx v, int k, int j, int i) const {
    return Component(0, b, bound(0, b, v.var) + v.off)[Cell(k, j, i)];
  }
  /* flattened packs: one unified outer index (sparse_pack.hpp:309-313) */
  PB2_PACK_HD double &operator()(int n, int k, int j, int i) const {
    return Component(0, 0, n)[Cell(k, j, i)];
  }
  PB2_PACK_HD double &flux(int b, int dir, int n, int k, int j, int i) const {
    return Component(dir, b, n)[Cell(k, j, i)];
  }
  PB2_PACK_HD double &flux(int b, int dir, PackIdx v, int k, int j, int i) const {
    return Component(dir, b, bound(0, b, v.var) + v.off)[Cell(k, j, i)];
  }
  PB2_PACK_HD double &flux(int dir, int n, int k, int j, int i) const {
    return Component(dir, 0, n)[Cell(k, j, i)];
  }
};

} /* namespace pb2 */
#endif /* __cplusplus */
#endif /* PARTHENON_B200_PACK_H_ */
