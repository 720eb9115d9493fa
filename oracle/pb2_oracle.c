/* pb2_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See pb2_oracle.h.
 *
 * Plain-C restatement of the reference's ghost-zone hot path: mesh topology for a
 * single-tree forest (forests of differently oriented trees: topology in oracle/forest.py,
 * handed over through orc_mesh_create_custom), boundary index boxes, pack/unpack — through the
 * neighbour tree's LogicalCoordinateTransformation where there is one —,
 * restriction/prolongation at
 * fine-coarse boundaries, flux correction (face fluxes of cell-centred fields and the
 * edge-centred fluxes of face fields), physical boundaries, the benchmarks/burgers RK2
 * cycle, example/advection and example/sparse_advection (uniform, statically and adaptively
 * refined meshes), and the exchange / remesh of face, edge and node fields with block
 * ownership.  Every section is pinned against dumps of the reference (tests/golden/).
 * Must be compiled with -ffp-contract=off (the reference CPU build emits no FMAs).
 */
#include "pb2_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXNB 64

typedef struct {
  int level;
  long lx[3];
} Loc;

typedef struct {
  int gid;
  Loc loc;        /* wrapped location (as stored in the tree) */
  Loc origin_loc; /* location in the frame of the owning block (may lie outside) */
  int off[3];
  /* forests of differently oriented trees: NeighborBlock::lcoord_trans
   * (logical_coordinate_transformation.hpp:36-112); tr_on == 0: identity */
  int tr_on, tr_dir[3], tr_flip[3];
} Neighbor;

typedef struct {
  Loc loc;
  int nnb;
  Neighbor nb[MAXNB];
  double xmin[3], xmax[3]; /* interior bounds */
  double dx[3];            /* UniformCartesian dx_ */
  double cxmin[3];         /* UniformCartesian xmin_ (includes ghost offset) */
  double ccxmin[3], cdx[3]; /* coarse coords (uniform_cartesian.hpp:41-55) */
  int bc[6]; /* custom (forest) meshes: MeshBlock::boundary_flag per face, 0 = none */
} Block;

struct OrcMesh {
  int ndim, nx[3], ng, nrb[3], root_level, nblocks, multilevel;
  int periodic[3];
  int bc[6]; /* per mesh face (inner_x1, outer_x1, ...): 0 periodic, 1 outflow, 2 reflect */
  double xmin[3], xmax[3];
  int custom;          /* blocks, neighbours and block boundary flags were supplied (forest) */
  int prolong_constant; /* cell-centred prolongation is ProlongatePiecewiseConstant */
  int is[3], ie[3], n[3];    /* fine interior bounds and full extents (i,j,k order) */
  int cis[3], cie[3], cn[3]; /* coarse */
  Block *blocks;
};

/* ------------------------------------------------------------------------------------ */
/* topology */

/* src/mesh/forest/logical_location.cpp:61-74 */
static double index_to_symmetrized_coordinate(long index, int bloc, long nrange) {
  long noffset = index - nrange / 2;
  long noffset_ceil = index - (nrange + 1) / 2;
  return (double)(noffset + noffset_ceil + (long)bloc) / (2.0 * (double)nrange);
}
/* src/defs.hpp:98-101 */
static double logical_to_actual(double u, double xmin, double xmax) {
  return 0.5 * (xmin + xmax) + (u * xmax - u * xmin);
}

/* z-order key; src/utils/morton_number.hpp:43 (x in the lowest interleaved bit) */
static uint64_t morton_key(const Loc *l, int maxlevel) {
  uint64_t key = 0;
  int sh = maxlevel - l->level;
  for (int bit = 0; bit < maxlevel; ++bit)
    for (int d = 0; d < 3; ++d) {
      uint64_t c = ((uint64_t)l->lx[d]) << sh;
      key |= ((c >> bit) & 1ull) << (3 * bit + d);
    }
  return key;
}

typedef struct {
  uint64_t key;
  Loc loc;
} SortEnt;
static int cmp_ent(const void *a, const void *b) {
  const SortEnt *x = (const SortEnt *)a, *y = (const SortEnt *)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return x->loc.level - y->loc.level;
}

static int find_leaf(const OrcMesh *m, const Loc *l) {
  for (int b = 0; b < m->nblocks; ++b) {
    const Loc *q = &m->blocks[b].loc;
    if (q->level == l->level && q->lx[0] == l->lx[0] && q->lx[1] == l->lx[1] &&
        q->lx[2] == l->lx[2])
      return b;
  }
  return -1;
}
/* is `l` an internal node, i.e. does some leaf lie strictly below it */
static int is_internal(const OrcMesh *m, const Loc *l) {
  for (int b = 0; b < m->nblocks; ++b) {
    const Loc *q = &m->blocks[b].loc;
    if (q->level <= l->level) continue;
    int sh = q->level - l->level;
    if ((q->lx[0] >> sh) == l->lx[0] && (q->lx[1] >> sh) == l->lx[1] &&
        (q->lx[2] >> sh) == l->lx[2])
      return 1;
  }
  return 0;
}

static long nblocks_at(const OrcMesh *m, int level, int d) {
  /* number of blocks spanning the tree in direction d at `level`; the single tree
   * covers nrb root blocks at root_level */
  if (d >= m->ndim) return 1;
  return ((long)m->nrb[d]) << (level - m->root_level);
}

/* wrap an out-of-tree same-level location through the periodic boundary; returns 0 if
 * it falls outside a non-periodic boundary */
static int wrap_loc(const OrcMesh *m, const Loc *in, Loc *out) {
  *out = *in;
  for (int d = 0; d < 3; ++d) {
    long n = nblocks_at(m, in->level, d);
    if (in->lx[d] < 0 || in->lx[d] >= n) {
      if (d >= m->ndim || !m->periodic[d]) return 0;
      out->lx[d] = ((in->lx[d] % n) + n) % n;
    }
  }
  return 1;
}

/* src/mesh/forest/logical_location.cpp:110-129 (on unwrapped coordinates) */
static int is_neighbor(const Loc *a, const Loc *in) {
  int max_level = a->level > in->level ? a->level : in->level;
  long bs_in = 1L << (max_level - in->level), bs_this = 1L << (max_level - a->level);
  for (int d = 0; d < 3; ++d) {
    long low = a->lx[d] * bs_this - 1, hi = low + bs_this + 1;
    long low_in = in->lx[d] * bs_in, hi_in = low_in + bs_in - 1;
    if (hi < low_in || low > hi_in) return 0;
  }
  return 1;
}

static long floor_shift(long v, int sh) { return v >> sh; /* arithmetic shift */ }

/* src/mesh/forest/logical_location.cpp:98-108 */
static void same_level_offsets(const Loc *me, const Loc *nb, int off[3]) {
  int sn = nb->level - me->level > 0 ? nb->level - me->level : 0;
  int sm = me->level - nb->level > 0 ? me->level - nb->level : 0;
  for (int d = 0; d < 3; ++d)
    off[d] = (int)(floor_shift(nb->lx[d], sn) - floor_shift(me->lx[d], sm));
}

static void add_neighbor(const OrcMesh *m, Block *blk, int gid, const Loc *wrapped,
                         const Loc *origin) {
  if (blk->nnb >= MAXNB) {
    fprintf(stderr, "oracle: too many neighbors\n");
    abort();
  }
  Neighbor *nb = &blk->nb[blk->nnb++];
  nb->gid = gid;
  nb->loc = *wrapped;
  nb->origin_loc = *origin;
  same_level_offsets(&blk->loc, origin, nb->off); /* mesh-gmg.cpp:64 */
  (void)m;
}

/* src/mesh/forest/tree.cpp:139-226 (leaf grid) */
static void find_neighbors(OrcMesh *m, int b) {
  Block *blk = &m->blocks[b];
  const Loc *loc = &blk->loc;
  blk->nnb = 0;
  /* Indexer3D: first index (ox1) slowest?  tree.cpp:141-148 iterates offsets(o) over a
   * 3D indexer built as ({ox1 range},{ox2 range},{ox3 range}) whose LAST entry varies
   * fastest (indexer.hpp:117-144). */
  int lo[3], hi[3];
  for (int d = 0; d < 3; ++d) {
    lo[d] = d < m->ndim ? -1 : 0;
    hi[d] = d < m->ndim ? 1 : 0;
  }
  for (int o1 = lo[0]; o1 <= hi[0]; ++o1)
    for (int o2 = lo[1]; o2 <= hi[1]; ++o2)
      for (int o3 = lo[2]; o3 <= hi[2]; ++o3) {
        if (o1 == 0 && o2 == 0 && o3 == 0) continue;
        Loc neigh = *loc, w;
        neigh.lx[0] += o1;
        neigh.lx[1] += o2;
        neigh.lx[2] += o3;
        if (!wrap_loc(m, &neigh, &w)) continue;
        int g = find_leaf(m, &w);
        if (g >= 0) {
          add_neighbor(m, blk, g, &w, &neigh);
        } else if (is_internal(m, &w)) {
          /* daughters of the (unwrapped) neighbor that touch this block */
          int nd[3] = {m->ndim > 0 ? 2 : 1, m->ndim > 1 ? 2 : 1, m->ndim > 2 ? 2 : 1};
          /* GetDaughters order: logical_location.cpp (ox3 outer, ox2, ox1 inner) */
          for (int d3 = 0; d3 < nd[2]; ++d3)
            for (int d2 = 0; d2 < nd[1]; ++d2)
              for (int d1 = 0; d1 < nd[0]; ++d1) {
                Loc dn, dw;
                dn.level = neigh.level + 1;
                dn.lx[0] = (neigh.lx[0] << 1) + d1;
                dn.lx[1] = m->ndim > 1 ? (neigh.lx[1] << 1) + d2 : neigh.lx[1];
                dn.lx[2] = m->ndim > 2 ? (neigh.lx[2] << 1) + d3 : neigh.lx[2];
                if (m->ndim < 2) dn.lx[1] = 0;
                if (m->ndim < 3) dn.lx[2] = 0;
                if (!is_neighbor(loc, &dn)) continue;
                wrap_loc(m, &dn, &dw);
                int gd = find_leaf(m, &dw);
                if (gd < 0) {
                  fprintf(stderr, "oracle: mesh violates 2:1 nesting\n");
                  abort();
                }
                add_neighbor(m, blk, gd, &dw, &dn);
              }
        } else {
          /* coarser neighbor: parent of neigh */
          Loc par = neigh, pw;
          par.level = neigh.level - 1;
          for (int d = 0; d < 3; ++d) par.lx[d] = floor_shift(neigh.lx[d], 1);
          if (m->ndim < 2) par.lx[1] = 0;
          if (m->ndim < 3) par.lx[2] = 0;
          if (!wrap_loc(m, &par, &pw)) continue;
          int gp = find_leaf(m, &pw);
          if (gp < 0) continue;
          int so[3];
          same_level_offsets(loc, &par, so);
          if (so[0] == o1 && so[1] == o2 && so[2] == o3) add_neighbor(m, blk, gp, &pw, &par);
        }
      }
}

/* mesh boundary flags used by the NEXT orc_mesh_create* call (test convenience) */
static int g_next_bc[6] = {0, 0, 0, 0, 0, 0};
void orc_mesh_set_next_bcs(const int bc[6]) {
  for (int f = 0; f < 6; ++f) g_next_bc[f] = bc ? bc[f] : 0;
}

OrcMesh *orc_mesh_create(int ndim, const int nx[3], int ng, const int nrb[3],
                         const double xmin[3], const double xmax[3], int nleaf,
                         const int *leaves) {
  OrcMesh *m = (OrcMesh *)calloc(1, sizeof(OrcMesh));
  m->ndim = ndim;
  m->ng = ng;
  int maxrb = 1;
  for (int d = 0; d < 3; ++d) {
    m->nx[d] = d < ndim ? nx[d] : 1;
    m->nrb[d] = d < ndim ? nrb[d] : 1;
    m->xmin[d] = xmin[d];
    m->xmax[d] = xmax[d];
    m->bc[2 * d] = g_next_bc[2 * d];
    m->bc[2 * d + 1] = g_next_bc[2 * d + 1];
    m->periodic[d] = m->bc[2 * d] == 0;
    if (m->nrb[d] > maxrb) maxrb = m->nrb[d];
  }
  /* single tree: all non-symmetry directions must hold the same power-of-two number of
   * root blocks (forest.cpp:73-105: ntree = nblock / max common power-of-2 divisor) */
  int rl = 0;
  while ((1 << rl) < maxrb) ++rl;
  for (int d = 0; d < ndim; ++d)
    if (m->nrb[d] != (1 << rl)) {
      fprintf(stderr, "oracle: only single-tree (2^n cubed root grid) meshes supported\n");
      free(m);
      return NULL;
    }
  m->root_level = rl;
  m->nblocks = nleaf;
  m->blocks = (Block *)calloc((size_t)nleaf, sizeof(Block));
  SortEnt *ent = (SortEnt *)malloc(sizeof(SortEnt) * (size_t)nleaf);
  int maxlevel = 0;
  for (int i = 0; i < nleaf; ++i)
    if (leaves[4 * i] > maxlevel) maxlevel = leaves[4 * i];
  for (int i = 0; i < nleaf; ++i) {
    ent[i].loc.level = leaves[4 * i];
    for (int d = 0; d < 3; ++d) ent[i].loc.lx[d] = leaves[4 * i + 1 + d];
    ent[i].key = morton_key(&ent[i].loc, maxlevel);
    if (ent[i].loc.level != rl) m->multilevel = 1;
  }
  qsort(ent, (size_t)nleaf, sizeof(SortEnt), cmp_ent);
  for (int d = 0; d < 3; ++d) {
    int sym = d >= ndim;
    m->is[d] = sym ? 0 : ng;
    m->ie[d] = sym ? 0 : ng + m->nx[d] - 1;
    m->n[d] = sym ? 1 : m->nx[d] + 2 * ng;
    int cnx = sym ? 0 : (m->nx[d] / 2 > 1 ? m->nx[d] / 2 : 1); /* meshblock.cpp:204-216 */
    m->cis[d] = sym ? 0 : ng;
    m->cie[d] = sym ? 0 : ng + cnx - 1;
    m->cn[d] = sym ? 1 : cnx + 2 * ng;
  }
  for (int b = 0; b < nleaf; ++b) {
    Block *blk = &m->blocks[b];
    blk->loc = ent[b].loc;
    for (int d = 0; d < 3; ++d) {
      if (d < ndim) {
        /* tree.cpp:297-312 + logical_location.cpp:76-79 */
        long nb_tot = 1L << (blk->loc.level > 0 ? blk->loc.level : 0);
        /* the tree domain equals the mesh domain (single tree) */
        double ul = index_to_symmetrized_coordinate(blk->loc.lx[d], 0, nb_tot);
        double ur = index_to_symmetrized_coordinate(blk->loc.lx[d], 2, nb_tot);
        blk->xmin[d] = logical_to_actual(ul, xmin[d], xmax[d]);
        blk->xmax[d] = logical_to_actual(ur, xmin[d], xmax[d]);
      } else {
        blk->xmin[d] = xmin[d];
        blk->xmax[d] = xmax[d];
      }
      /* uniform_cartesian.hpp:30-40 */
      blk->dx[d] = (blk->xmax[d] - blk->xmin[d]) / m->nx[d];
      int istart = d < ndim ? ng : 0;
      blk->cxmin[d] = blk->xmin[d] - istart * blk->dx[d];
      /* uniform_cartesian.hpp:41-55, coarsen = 2 */
      blk->ccxmin[d] = blk->cxmin[d] + istart * blk->dx[d] * (1 - 2);
      blk->cdx[d] = blk->dx[d] * ((d == 0 || istart > 0) ? 2 : 1);
    }
  }
  free(ent);
  for (int b = 0; b < nleaf; ++b) find_neighbors(m, b);
  return m;
}

OrcMesh *orc_mesh_create_uniform(int ndim, const int nx[3], int ng, const int nrb[3],
                                 const double xmin[3], const double xmax[3]) {
  int n[3] = {nrb[0], ndim > 1 ? nrb[1] : 1, ndim > 2 ? nrb[2] : 1};
  int nleaf = n[0] * n[1] * n[2];
  int rl = 0;
  while ((1 << rl) < n[0]) ++rl;
  int *leaves = (int *)malloc(sizeof(int) * 4 * (size_t)nleaf);
  int c = 0;
  for (int k = 0; k < n[2]; ++k)
    for (int j = 0; j < n[1]; ++j)
      for (int i = 0; i < n[0]; ++i) {
        leaves[4 * c] = rl;
        leaves[4 * c + 1] = i;
        leaves[4 * c + 2] = j;
        leaves[4 * c + 3] = k;
        ++c;
      }
  OrcMesh *m = orc_mesh_create(ndim, nx, ng, nrb, xmin, xmax, nleaf, leaves);
  free(leaves);
  return m;
}

/* A mesh given block by block (test infrastructure for forests of differently oriented
 * trees, mesh/forest/forest.cpp:204-297): the caller — oracle/forest.py, which restates the
 * forest topology — supplies every block's location, boundary flags and neighbour list with
 * each neighbour's LogicalCoordinateTransformation; everything downstream (index boxes, pack,
 * unpack, restriction, prolongation, physical boundaries) is the code pinned on single trees.
 * blocks: nblocks x 4 ints (level, lx1, lx2, lx3) in gid order; block_bc: nblocks x 6;
 * nnb: neighbours per block; nbs: per neighbour 18 ints
 *   gid, level, lx[3] (as stored), origin level, origin lx[3], tr_on, tr_dir[3], tr_flip[3],
 * and two spares, in the order the reference's FindNeighbors lists them. */
OrcMesh *orc_mesh_create_custom(int ndim, const int nx[3], int ng, int nblocks, const int *blocks,
                                const int *block_bc, const int *nnb, const int *nbs,
                                int multilevel, int prolong_constant) {
  OrcMesh *m = (OrcMesh *)calloc(1, sizeof(OrcMesh));
  m->ndim = ndim;
  m->ng = ng;
  m->custom = 1;
  m->multilevel = multilevel;
  m->prolong_constant = prolong_constant;
  m->nblocks = nblocks;
  for (int d = 0; d < 3; ++d) {
    m->nx[d] = d < ndim ? nx[d] : 1;
    m->nrb[d] = 1;
    m->xmin[d] = 0.0;
    m->xmax[d] = 1.0;
    int sym = d >= ndim;
    m->is[d] = sym ? 0 : ng;
    m->ie[d] = sym ? 0 : ng + m->nx[d] - 1;
    m->n[d] = sym ? 1 : m->nx[d] + 2 * ng;
    int cnx = sym ? 0 : (m->nx[d] / 2 > 1 ? m->nx[d] / 2 : 1);
    m->cis[d] = sym ? 0 : ng;
    m->cie[d] = sym ? 0 : ng + cnx - 1;
    m->cn[d] = sym ? 1 : cnx + 2 * ng;
  }
  m->blocks = (Block *)calloc((size_t)nblocks, sizeof(Block));
  const int *q = nbs;
  for (int b = 0; b < nblocks; ++b) {
    Block *blk = &m->blocks[b];
    blk->loc.level = blocks[4 * b];
    for (int d = 0; d < 3; ++d) {
      blk->loc.lx[d] = blocks[4 * b + 1 + d];
      blk->xmin[d] = 0.0; /* unit blocks: coordinates play no role in what this mesh is for */
      blk->xmax[d] = 1.0;
      blk->dx[d] = 1.0 / m->nx[d];
      int istart = d < ndim ? ng : 0;
      blk->cxmin[d] = -istart * blk->dx[d];
      blk->ccxmin[d] = blk->cxmin[d] + istart * blk->dx[d] * (1 - 2);
      blk->cdx[d] = blk->dx[d] * ((d == 0 || istart > 0) ? 2 : 1);
    }
    for (int f = 0; f < 6; ++f) blk->bc[f] = block_bc[6 * b + f];
    if (nnb[b] > MAXNB) {
      fprintf(stderr, "oracle: too many neighbors\n");
      abort();
    }
    blk->nnb = nnb[b];
    for (int n = 0; n < nnb[b]; ++n, q += 18) {
      Neighbor *nb = &blk->nb[n];
      nb->gid = q[0];
      nb->loc.level = q[1];
      nb->origin_loc.level = q[5];
      for (int d = 0; d < 3; ++d) {
        nb->loc.lx[d] = q[2 + d];
        nb->origin_loc.lx[d] = q[6 + d];
        nb->tr_dir[d] = q[10 + d];
        nb->tr_flip[d] = q[13 + d];
      }
      nb->tr_on = q[9];
      same_level_offsets(&blk->loc, &nb->origin_loc, nb->off); /* mesh-gmg.cpp:64 */
    }
  }
  return m;
}

void orc_mesh_destroy(OrcMesh *m) {
  if (!m) return;
  free(m->blocks);
  free(m);
}
int orc_mesh_nblocks(const OrcMesh *m) { return m->nblocks; }
int orc_mesh_multilevel(const OrcMesh *m) { return m->multilevel; }
void orc_mesh_dims(const OrcMesh *m, int dims[3], int cdims[3]) {
  for (int d = 0; d < 3; ++d) {
    dims[d] = m->n[2 - d];
    cdims[d] = m->cn[2 - d];
  }
}
void orc_mesh_block_loc(const OrcMesh *m, int b, int loc[4]) {
  loc[0] = m->blocks[b].loc.level;
  for (int d = 0; d < 3; ++d) loc[1 + d] = (int)m->blocks[b].loc.lx[d];
}
void orc_mesh_block_bounds(const OrcMesh *m, int b, double xmin[3], double xmax[3]) {
  for (int d = 0; d < 3; ++d) {
    xmin[d] = m->blocks[b].xmin[d];
    xmax[d] = m->blocks[b].xmax[d];
  }
}
int orc_mesh_num_neighbors(const OrcMesh *m, int b) { return m->blocks[b].nnb; }
void orc_mesh_neighbor(const OrcMesh *m, int b, int n, int out[5]) {
  const Neighbor *nb = &m->blocks[b].nb[n];
  out[0] = nb->gid;
  out[1] = nb->loc.level;
  out[2] = nb->off[0];
  out[3] = nb->off[1];
  out[4] = nb->off[2];
}

/* ------------------------------------------------------------------------------------ */
/* index boxes: src/bvals/comms/bnd_info.cpp:105-252, cell centred (top_offset = 0),
 * non-flux, identity logical coordinate transform, full ownership. */
enum { IR_SEND = 0, IR_RECV = 1 };

void orc_calc_indices(const OrcMesh *m, int b, int n, int ir_type, int prores, int s[3],
                      int e[3]) {
  const Block *blk = &m->blocks[b];
  const Neighbor *nb = &blk->nb[n];
  const Loc *loc = &blk->loc;
  const int ng = m->ng;
  int use_coarse = prores || nb->loc.level < loc->level; /* :124-125 */
  int bs[3], be[3], nbs[3], nbe[3];
  int coarse_fac = nb->loc.level > loc->level ? 2 : 1; /* :130 */
  for (int d = 0; d < 3; ++d) {
    bs[d] = use_coarse ? m->cis[d] : m->is[d];
    be[d] = use_coarse ? m->cie[d] : m->ie[d];
    /* neighbor_shape = IndexShape(nb.block_size.nx / coarse_fac, nghost)  :131-135 */
    int sym = d >= m->ndim;
    int nnx = m->nx[d] / coarse_fac;
    nbs[d] = sym ? 0 : ng;
    nbe[d] = sym ? 0 : ng + nnx - 1;
    if (sym) nbe[d] = 0;
  }
  int interior_offset = ir_type == IR_SEND ? ng : 0; /* :157-160 */
  int exterior_offset = ir_type == IR_RECV ? ng : 0;
  if (prores) exterior_offset /= 2; /* :161-166 */
  for (int d = 0; d < 3; ++d) {
    int not_sym = d < m->ndim;
    if (nb->off[d] == 0) {
      s[d] = bs[d];
      e[d] = be[d];
      if (loc->level < nb->origin_loc.level && not_sym) { /* :173-192 */
        int extra = (be[d] - bs[d] + 1) - (nbe[d] - nbs[d] + 1);
        s[d] += (nb->origin_loc.lx[d] % 2 + 2) % 2 == 1 ? extra - interior_offset : 0;
        e[d] -= (nb->origin_loc.lx[d] % 2 + 2) % 2 == 0 ? extra - interior_offset : 0;
      }
      if (loc->level > nb->origin_loc.level && not_sym) { /* :193-204 */
        s[d] -= loc->lx[d] % 2 == 1 ? exterior_offset : 0;
        e[d] += loc->lx[d] % 2 == 0 ? exterior_offset : 0;
      }
    } else if (nb->off[d] > 0) { /* :211-214 */
      s[d] = be[d] + (-interior_offset + 1);
      e[d] = be[d] + exterior_offset;
    } else { /* :215-218 */
      s[d] = bs[d] - exterior_offset;
      e[d] = bs[d] + (interior_offset - 1);
    }
  }
}

/* ------------------------------------------------------------------------------------ */
/* exchange */

static inline size_t fidx(const OrcMesh *m, int ncomp, int b, int c, int k, int j, int i) {
  return ((((size_t)b * ncomp + c) * m->n[2] + k) * m->n[1] + j) * m->n[0] + i;
}
static inline size_t cidx(const OrcMesh *m, int ncomp, int b, int c, int k, int j, int i) {
  return ((((size_t)b * ncomp + c) * m->cn[2] + k) * m->cn[1] + j) * m->cn[0] + i;
}

int64_t orc_count_regions(const OrcMesh *m) {
  int64_t n = 0;
  for (int b = 0; b < m->nblocks; ++b) n += m->blocks[b].nnb;
  return n;
}

/* RestrictAverage::Do  src/prolong_restrict/pr_ops.hpp:105-165 (cell centred, DIM =
 * m->ndim).  Volumes are the (constant) fine cell volume of the block. */
static void restrict_cell(const OrcMesh *m, const double *U, double *Uc, int ncomp, int b,
                          int c, int ck, int cj, int ci) {
  const Block *blk = &m->blocks[b];
  const int DIM = m->ndim;
  const int i = (ci - m->cis[0]) * 2 + m->is[0];
  const int j = DIM > 1 ? (cj - m->cis[1]) * 2 + m->is[1] : m->is[1];
  const int k = DIM > 2 ? (ck - m->cis[2]) * 2 + m->is[2] : m->is[2];
  double vol[2][2][2], terms[2][2][2];
  memset(vol, 0, sizeof(vol));
  memset(terms, 0, sizeof(terms));
  const double cellvol = blk->dx[0] * blk->dx[1] * blk->dx[2]; /* uniform_cartesian.hpp:39 */
  for (int ok = 0; ok < 1 + (DIM > 2); ++ok)
    for (int oj = 0; oj < 1 + (DIM > 1); ++oj)
      for (int oi = 0; oi < 2; ++oi) {
        vol[ok][oj][oi] = cellvol;
        terms[ok][oj][oi] = vol[ok][oj][oi] * U[fidx(m, ncomp, b, c, k + ok, j + oj, i + oi)];
      }
  const double tvol = ((vol[0][0][0] + vol[0][1][0]) + (vol[0][0][1] + vol[0][1][1])) +
                      ((vol[1][0][0] + vol[1][1][0]) + (vol[1][0][1] + vol[1][1][1]));
  Uc[cidx(m, ncomp, b, c, ck, cj, ci)] =
      (((terms[0][0][0] + terms[0][1][0]) + (terms[0][0][1] + terms[0][1][1])) +
       ((terms[1][0][0] + terms[1][1][0]) + (terms[1][0][1] + terms[1][1][1]))) /
      tvol;
}

static void restrict_region(const OrcMesh *m, const double *U, double *Uc, int ncomp, int b,
                            const int s[3], const int e[3]) {
  for (int c = 0; c < ncomp; ++c)
    for (int k = s[2]; k <= e[2]; ++k)
      for (int j = s[1]; j <= e[1]; ++j)
        for (int i = s[0]; i <= e[0]; ++i) restrict_cell(m, U, Uc, ncomp, b, c, k, j, i);
}

/* ProResInfo::GetSend (bnd_info.cpp:387-403) + refinement::Restrict called from
 * SendBoundBufs (boundary_communication.cpp:82-87) */
void orc_restrict_send(const OrcMesh *m, const double *U, double *Uc, int ncomp) {
#pragma omp parallel for schedule(dynamic)
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    for (int n = 0; n < blk->nnb; ++n) {
      if (blk->nb[n].origin_loc.level < blk->loc.level) {
        int s[3], e[3];
        orc_calc_indices(m, b, n, IR_SEND, 1, s, e);
        restrict_region(m, U, Uc, ncomp, b, s, e);
      }
    }
  }
}

/* pack kernel: boundary_communication.cpp:95-140; region source selection
 * bnd_info.cpp:285-289 (coarse buffer if the neighbor is coarser) */
int64_t orc_pack(const OrcMesh *m, const double *U, const double *Uc, int ncomp,
                 double *buf, int64_t *buf_off) {
  int64_t off = 0, r = 0;
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    for (int n = 0; n < blk->nnb; ++n) {
      int s[3], e[3];
      orc_calc_indices(m, b, n, IR_SEND, 0, s, e);
      buf_off[r++] = off;
      off += (int64_t)ncomp * (e[2] - s[2] + 1) * (e[1] - s[1] + 1) * (e[0] - s[0] + 1);
    }
  }
  buf_off[r] = off;
  if (!buf) return off;
  int64_t *first = (int64_t *)malloc(sizeof(int64_t) * (size_t)(m->nblocks + 1));
  first[0] = 0;
  for (int b = 0; b < m->nblocks; ++b) first[b + 1] = first[b] + m->blocks[b].nnb;
#pragma omp parallel for schedule(dynamic)
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    for (int n = 0; n < blk->nnb; ++n) {
      int s[3], e[3];
      orc_calc_indices(m, b, n, IR_SEND, 0, s, e);
      int coarse = blk->nb[n].origin_loc.level < blk->loc.level;
      double *p = buf + buf_off[first[b] + n];
      for (int c = 0; c < ncomp; ++c)
        for (int k = s[2]; k <= e[2]; ++k)
          for (int j = s[1]; j <= e[1]; ++j)
            for (int i = s[0]; i <= e[0]; ++i)
              *p++ = coarse ? Uc[cidx(m, ncomp, b, c, k, j, i)]
                            : U[fidx(m, ncomp, b, c, k, j, i)];
    }
  }
  free(first);
  return off;
}

/* unpack kernel: boundary_communication.cpp:273-334; the sending buffer is the one keyed
 * (sender gid, receiver gid, reverse offset index): bvals_utils.hpp:43-67 */
void orc_unpack(const OrcMesh *m, double *U, double *Uc, int ncomp, const double *buf,
                const int64_t *buf_off) {
  int64_t *first = (int64_t *)malloc(sizeof(int64_t) * (size_t)(m->nblocks + 1));
  first[0] = 0;
  for (int b = 0; b < m->nblocks; ++b) first[b + 1] = first[b] + m->blocks[b].nnb;
#pragma omp parallel for schedule(dynamic)
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    for (int n = 0; n < blk->nnb; ++n) {
      const Neighbor *nb = &blk->nb[n];
      /* find the sender's matching region: the receive key carries the offsets transformed
       * into the sender's frame, reversed (ReceiveKey, bvals_utils.hpp:57-67;
       * LogicalCoordinateTransformation::Transform(CellCentOffsets),
       * logical_coordinate_transformation.cpp:92-99) */
      int toff[3] = {nb->off[0], nb->off[1], nb->off[2]};
      if (nb->tr_on)
        for (int d = 0; d < 3; ++d)
          toff[nb->tr_dir[d]] = nb->tr_flip[d] ? -nb->off[d] : nb->off[d];
      const Block *sb = &m->blocks[nb->gid];
      int sn = -1;
      for (int q = 0; q < sb->nnb; ++q)
        if (sb->nb[q].gid == b && sb->nb[q].off[0] == -toff[0] &&
            sb->nb[q].off[1] == -toff[1] && sb->nb[q].off[2] == -toff[2]) {
          sn = q;
          break;
        }
      if (sn < 0) {
        fprintf(stderr, "oracle: no matching send region for block %d nb %d\n", b, n);
        abort();
      }
      int s[3], e[3];
      orc_calc_indices(m, b, n, IR_RECV, 0, s, e);
      int64_t r = first[nb->gid] + sn;
      int64_t cnt = (int64_t)ncomp * (e[2] - s[2] + 1) * (e[1] - s[1] + 1) * (e[0] - s[0] + 1);
      if (cnt != buf_off[r + 1] - buf_off[r]) {
        fprintf(stderr, "oracle: send/recv size mismatch block %d nb %d: %ld vs %ld\n", b, n,
                (long)cnt, (long)(buf_off[r + 1] - buf_off[r]));
        abort();
      }
      int coarse = nb->origin_loc.level < blk->loc.level;
      const double *p = buf + buf_off[r];
      if (nb->tr_on) {
        /* bnd_info.cpp:216-228: the box goes to the sender's logical coordinates (Transform of
         * both corners with ncell = var.GetDim(1), re-sorted), SetBounds walks it in buffer
         * order and writes every element at InverseTransform (boundary_communication.cpp:
         * 282-308; cell-centred fields: no sign factor) */
        const int ncell = coarse ? m->cn[0] : m->n[0];
        int ts[3], te[3];
        for (int d = 0; d < 3; ++d) {
          const int o = nb->tr_dir[d];
          ts[o] = nb->tr_flip[d] ? ncell - 1 - s[d] : s[d];
          te[o] = nb->tr_flip[d] ? ncell - 1 - e[d] : e[d];
        }
        for (int d = 0; d < 3; ++d)
          if (ts[d] > te[d]) {
            const int t = ts[d];
            ts[d] = te[d];
            te[d] = t;
          }
        for (int c = 0; c < ncomp; ++c)
          for (int k = ts[2]; k <= te[2]; ++k)
            for (int j = ts[1]; j <= te[1]; ++j)
              for (int i = ts[0]; i <= te[0]; ++i) {
                const int in[3] = {i, j, k};
                int out[3];
                for (int d = 0; d < 3; ++d)
                  out[d] = nb->tr_flip[d] ? ncell - 1 - in[nb->tr_dir[d]] : in[nb->tr_dir[d]];
                if (coarse)
                  Uc[cidx(m, ncomp, b, c, out[2], out[1], out[0])] = *p++;
                else
                  U[fidx(m, ncomp, b, c, out[2], out[1], out[0])] = *p++;
              }
        continue;
      }
      for (int c = 0; c < ncomp; ++c)
        for (int k = s[2]; k <= e[2]; ++k)
          for (int j = s[1]; j <= e[1]; ++j)
            for (int i = s[0]; i <= e[0]; ++i) {
              if (coarse)
                Uc[cidx(m, ncomp, b, c, k, j, i)] = *p++;
              else
                U[fidx(m, ncomp, b, c, k, j, i)] = *p++;
            }
    }
  }
  free(first);
}

/* ProResInfo::GetSet (bnd_info.cpp:405-448) restriction part + SetBounds :338-346 */
void orc_restrict_set(const OrcMesh *m, const double *U, double *Uc, int ncomp) {
  if (!m->multilevel) return;
#pragma omp parallel for schedule(dynamic)
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    int restricted = 0;
    if (blk->loc.level > 0)
      for (int n = 0; n < blk->nnb; ++n)
        restricted = restricted || (blk->nb[n].origin_loc.level == blk->loc.level - 1);
    if (!restricted) continue;
    for (int n = 0; n < blk->nnb; ++n) {
      if (blk->nb[n].origin_loc.level < blk->loc.level) continue;
      int s[3], e[3];
      orc_calc_indices(m, b, n, IR_RECV, 1, s, e);
      restrict_region(m, U, Uc, ncomp, b, s, e);
    }
  }
}

static inline double sgn(double x) { return (x > 0) - (x < 0); } /* utils/utils.hpp SIGN */

/* util::GradMinMod pr_ops.hpp:95-101 */
static inline double grad_minmod(double fc, double fm, double fp, double dxm, double dxp) {
  double gxm = (fc - fm) / dxm;
  double gxp = (fp - fc) / dxp;
  return 0.5 * (sgn(gxm) + sgn(gxp)) * fmin(fabs(gxm), fabs(gxp));
}

/* ProlongateSharedGeneral<true,false>::Do pr_ops.hpp:167-280 (MinMod, cell centred) with
 * util::GetGridSpacings pr_ops.hpp:76-93 */
static void prolongate_cell(const OrcMesh *m, double *U, const double *Uc, int ncomp, int b,
                            int c, int k, int j, int i) {
  const Block *blk = &m->blocks[b];
  const int DIM = m->ndim;
  const int fi = (i - m->cis[0]) * 2 + m->is[0];
  const int fj = DIM > 1 ? (j - m->cis[1]) * 2 + m->is[1] : m->is[1];
  const int fk = DIM > 2 ? (k - m->cis[2]) * 2 + m->is[2] : m->is[2];
  const double fc = Uc[cidx(m, ncomp, b, c, k, j, i)];
  double g[3] = {0, 0, 0}, dxfm[3] = {0, 0, 0}, dxfp[3] = {0, 0, 0};
  const int cc[3] = {i, j, k}, ff[3] = {fi, fj, fk};
  for (int d = 0; d < DIM; ++d) {
    const double xm = blk->ccxmin[d] + ((cc[d] - 1) + 0.5) * blk->cdx[d];
    const double xc = blk->ccxmin[d] + (cc[d] + 0.5) * blk->cdx[d];
    const double xp = blk->ccxmin[d] + ((cc[d] + 1) + 0.5) * blk->cdx[d];
    const double dxm = xc - xm, dxp = xp - xc;
    const double fxm = blk->cxmin[d] + (ff[d] + 0.5) * blk->dx[d];
    const double fxp = blk->cxmin[d] + ((ff[d] + 1) + 0.5) * blk->dx[d];
    dxfm[d] = xc - fxm;
    dxfp[d] = fxp - xc;
    int o[3] = {0, 0, 0};
    o[d] = 1;
    const double fm = Uc[cidx(m, ncomp, b, c, k - o[2], j - o[1], i - o[0])];
    const double fp = Uc[cidx(m, ncomp, b, c, k + o[2], j + o[1], i + o[0])];
    g[d] = grad_minmod(fc, fm, fp, dxm, dxp);
    /* ProlongatePiecewiseConstant = ProlongateSharedGeneral<false, true>: zero slopes
     * (pr_ops.hpp:206-262) */
    if (m->prolong_constant) g[d] = 0.0;
  }
  const double gx1m = g[0], gx1p = g[0], gx2m = g[1], gx2p = g[1], gx3m = g[2], gx3p = g[2];
  const double dx1fm = dxfm[0], dx1fp = dxfp[0], dx2fm = dxfm[1], dx2fp = dxfp[1],
               dx3fm = dxfm[2], dx3fp = dxfp[2];
  U[fidx(m, ncomp, b, c, fk, fj, fi)] = fc - (gx1m * dx1fm + gx2m * dx2fm + gx3m * dx3fm);
  U[fidx(m, ncomp, b, c, fk, fj, fi + 1)] =
      fc + (gx1p * dx1fp - gx2m * dx2fm - gx3m * dx3fm);
  if (DIM > 1) {
    U[fidx(m, ncomp, b, c, fk, fj + 1, fi)] =
        fc - (gx1m * dx1fm - gx2p * dx2fp + gx3m * dx3fm);
    U[fidx(m, ncomp, b, c, fk, fj + 1, fi + 1)] =
        fc + (gx1p * dx1fp + gx2p * dx2fp - gx3m * dx3fm);
  }
  if (DIM > 2) {
    U[fidx(m, ncomp, b, c, fk + 1, fj, fi)] =
        fc - (gx1m * dx1fm + gx2m * dx2fm - gx3p * dx3fp);
    U[fidx(m, ncomp, b, c, fk + 1, fj, fi + 1)] =
        fc + (gx1p * dx1fp - gx2m * dx2fm + gx3p * dx3fp);
    U[fidx(m, ncomp, b, c, fk + 1, fj + 1, fi)] =
        fc - (gx1m * dx1fm - gx2p * dx2fp - gx3p * dx3fp);
    U[fidx(m, ncomp, b, c, fk + 1, fj + 1, fi + 1)] =
        fc + (gx1p * dx1fp + gx2p * dx2fp + gx3p * dx3fp);
  }
}

/* ProlongateBounds boundary_communication.cpp:361-393 over GetSet regions whose
 * neighbor is coarser (bnd_info.cpp:421-446) */
void orc_prolongate(const OrcMesh *m, double *U, const double *Uc, int ncomp) {
  if (!m->multilevel) return;
#pragma omp parallel for schedule(dynamic)
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    for (int n = 0; n < blk->nnb; ++n) {
      if (!(blk->nb[n].origin_loc.level < blk->loc.level)) continue;
      int s[3], e[3];
      orc_calc_indices(m, b, n, IR_RECV, 1, s, e);
      for (int c = 0; c < ncomp; ++c)
        for (int k = s[2]; k <= e[2]; ++k)
          for (int j = s[1]; j <= e[1]; ++j)
            for (int i = s[0]; i <= e[0]; ++i) prolongate_cell(m, U, Uc, ncomp, b, c, k, j, i);
    }
  }
}

/* ------------------------------------------------------------------------------------ */
/* flux correction at fine-coarse faces: SendBoundBufs<flxcor_send> / SetBounds<flxcor_recv>
 * (boundary_communication.cpp:454-461) over the face flux of a cell-centred field.  Flux
 * arrays have the cell array's extents; entry (k,j,i) of direction d is the LOWER d-face of
 * cell (k,j,i), so face indices run is..ie+1 (IndexShape::GetBounds*(interior, F_d),
 * mesh/domain.hpp:162-251). */

static int face_dir(const int off[3]) { /* CellCentOffsets::IsFace: exactly one non-zero */
  int nz = (off[0] != 0) + (off[1] != 0) + (off[2] != 0);
  if (nz != 1) return -1;
  return off[0] != 0 ? 0 : (off[1] != 0 ? 1 : 2);
}

/* CalcIndices (bnd_info.cpp:105-252) for a Metadata::Flux face field, element F_dir with
 * dir the direction of the neighbour offset (GetFluxCorrectionElements :71-83).
 * ir_type IR_SEND: the finer block, in its COARSE index space (nb coarser => c_cellbounds,
 * :124-125); IR_RECV: the coarser block, in its fine index space. */
static void calc_indices_flux(const OrcMesh *m, int b, int n, int ir_type, int s[3],
                              int e[3]) {
  const Block *blk = &m->blocks[b];
  const Neighbor *nb = &blk->nb[n];
  const Loc *loc = &blk->loc;
  const int dir = face_dir(nb->off);
  const int use_coarse = nb->loc.level < loc->level;
  const int coarse_fac = nb->loc.level > loc->level ? 2 : 1;
  for (int d = 0; d < 3; ++d) {
    const int not_sym = d < m->ndim;
    const int top = (d == dir && not_sym) ? 1 : 0; /* TopologicalOffset of F_dir */
    const int bs = use_coarse ? m->cis[d] : m->is[d];
    const int be = (use_coarse ? m->cie[d] : m->ie[d]) + top;
    const int nbn = not_sym ? m->nx[d] / coarse_fac + top : 1;
    if (nb->off[d] == 0) {
      s[d] = bs;
      e[d] = be;
      if (loc->level < nb->origin_loc.level && not_sym) { /* :173-192, interior_offset = 0
          for the receiving side; the sending side never has a finer neighbour here */
        const int extra = (be - bs + 1) - nbn;
        s[d] += (nb->origin_loc.lx[d] % 2 + 2) % 2 == 1 ? extra : 0;
        e[d] -= (nb->origin_loc.lx[d] % 2 + 2) % 2 == 0 ? extra : 0;
      }
      /* :193-204 only moves the box by exterior_offset, which is 0 for a sender */
    } else if (nb->off[d] > 0) { /* :207-208, flux: shared element only */
      s[d] = be;
      e[d] = be;
    } else { /* :210-211 */
      s[d] = bs;
      e[d] = bs;
    }
  }
  (void)ir_type;
}

void orc_calc_indices_flux(const OrcMesh *m, int b, int n, int s[3], int e[3]) {
  const Block *blk = &m->blocks[b];
  calc_indices_flux(m, b, n, blk->nb[n].loc.level < blk->loc.level ? IR_SEND : IR_RECV, s, e);
}

/* RestrictAverage::Do pr_ops.hpp:105-165 with el = F_dir: average over the fine faces that
 * tile one coarse face, weights coords.Volume<F_dir> = face area
 * (uniform_cartesian.hpp:36-38, 248-258); fixed summation tree :155-162 */
static double restrict_face(const OrcMesh *m, const double *F, int ncomp, int b, int c,
                            int dir, int ck, int cj, int ci) {
  const Block *blk = &m->blocks[b];
  const int DIM = m->ndim;
  const int inc[3] = {DIM > 0 && dir != 0, DIM > 1 && dir != 1, DIM > 2 && dir != 2};
  const int i = (ci - m->cis[0]) * 2 + m->is[0];
  const int j = DIM > 1 ? (cj - m->cis[1]) * 2 + m->is[1] : m->is[1];
  const int k = DIM > 2 ? (ck - m->cis[2]) * 2 + m->is[2] : m->is[2];
  const double area[3] = {blk->dx[1] * blk->dx[2], blk->dx[0] * blk->dx[2],
                          blk->dx[0] * blk->dx[1]};
  double vol[2][2][2], terms[2][2][2];
  memset(vol, 0, sizeof(vol));
  memset(terms, 0, sizeof(terms));
  for (int ok = 0; ok < 1 + inc[2]; ++ok)
    for (int oj = 0; oj < 1 + inc[1]; ++oj)
      for (int oi = 0; oi < 1 + inc[0]; ++oi) {
        vol[ok][oj][oi] = area[dir];
        terms[ok][oj][oi] = vol[ok][oj][oi] * F[fidx(m, ncomp, b, c, k + ok, j + oj, i + oi)];
      }
  const double tvol = ((vol[0][0][0] + vol[0][1][0]) + (vol[0][0][1] + vol[0][1][1])) +
                      ((vol[1][0][0] + vol[1][1][0]) + (vol[1][0][1] + vol[1][1][1]));
  return (((terms[0][0][0] + terms[0][1][0]) + (terms[0][0][1] + terms[0][1][1])) +
          ((terms[1][0][0] + terms[1][1][0]) + (terms[1][0][1] + terms[1][1][1]))) /
         tvol;
}

/* Every coarse block overwrites the part of a face flux shared with a finer neighbour by
 * the area-weighted average of that neighbour's fine fluxes.  Region selection:
 * ForEachBoundary<flxcor_recv> (loop_utils.hpp:145-160): face neighbours one level finer. */
/* alloc (optional, stride astride, entry alloc[b * astride]): sparse fields — an unallocated
 * sender sends a null message, which leaves the receiver's flux alone (no default fill for
 * flxcor_recv, boundary_communication.cpp:311); an unallocated receiver sets nothing */
static int64_t flux_correct_masked(const OrcMesh *m, double *const F[3], int ncomp,
                                   const unsigned char *alloc, int astride) {
  int64_t moved = 0;
  if (!m->multilevel) return 0;
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    if (alloc != NULL && !alloc[b * astride]) continue;
    for (int n = 0; n < blk->nnb; ++n) {
      const Neighbor *nb = &blk->nb[n];
      const int dir = face_dir(nb->off);
      if (dir < 0 || dir >= m->ndim) continue;
      if (nb->loc.level != blk->loc.level + 1) continue;
      if (alloc != NULL && !alloc[nb->gid * astride]) continue;
      /* the sender's matching region */
      const Block *sb = &m->blocks[nb->gid];
      int sn = -1;
      for (int q = 0; q < sb->nnb; ++q)
        if (sb->nb[q].gid == b && sb->nb[q].off[0] == -nb->off[0] &&
            sb->nb[q].off[1] == -nb->off[1] && sb->nb[q].off[2] == -nb->off[2]) {
          sn = q;
          break;
        }
      if (sn < 0) {
        fprintf(stderr, "oracle: no matching flux-correction sender for block %d nb %d\n", b, n);
        abort();
      }
      int ss[3], se[3], rs[3], re[3];
      calc_indices_flux(m, nb->gid, sn, IR_SEND, ss, se);
      calc_indices_flux(m, b, n, IR_RECV, rs, re);
      for (int d = 0; d < 3; ++d)
        if (se[d] - ss[d] != re[d] - rs[d]) {
          fprintf(stderr, "oracle: flux-correction extent mismatch block %d nb %d dir %d\n", b, n, d);
          abort();
        }
      for (int c = 0; c < ncomp; ++c)
        for (int k = 0; k <= re[2] - rs[2]; ++k)
          for (int j = 0; j <= re[1] - rs[1]; ++j)
            for (int i = 0; i <= re[0] - rs[0]; ++i) {
              F[dir][fidx(m, ncomp, b, c, rs[2] + k, rs[1] + j, rs[0] + i)] =
                  restrict_face(m, F[dir], ncomp, nb->gid, c, dir, ss[2] + k, ss[1] + j, ss[0] + i);
              ++moved;
            }
    }
  }
  return moved;
}
int64_t orc_flux_correct(const OrcMesh *m, double *const F[3], int ncomp) {
  return flux_correct_masked(m, F, ncomp, NULL, 0);
}

void orc_apply_bcs_coarse(const OrcMesh *m, double *Uc, int ncomp);

int64_t orc_exchange(const OrcMesh *m, double *U, double *Uc, int ncomp, int prolongate) {
  int64_t nreg = orc_count_regions(m);
  int64_t *off = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nreg + 1));
  int64_t total = orc_pack(m, U, Uc, ncomp, NULL, off);
  double *buf = (double *)malloc(sizeof(double) * (size_t)(total > 0 ? total : 1));
  if (m->multilevel) orc_restrict_send(m, U, Uc, ncomp);
  orc_pack(m, U, Uc, ncomp, buf, off);
  orc_unpack(m, U, Uc, ncomp, buf, off);
  orc_restrict_set(m, U, Uc, ncomp);
  if (prolongate && m->multilevel && Uc) {
    orc_apply_bcs_coarse(m, Uc, ncomp); /* ApplyBoundaryConditionsOnCoarseOrFineMD(md, true) */
    orc_prolongate(m, U, Uc, ncomp);
  }
  free(buf);
  free(off);
  return total;
}

/* ------------------------------------------------------------------------------------ */
/* ghost exchange of NON-CELL-CENTRED fields (face / edge / node) between blocks of the SAME
 * level: element-aware index boxes (bnd_info.cpp:105-252 with TopologicalOffset{I,J,K},
 * basic_types.hpp:195-203, and the element bounds of mesh/domain.hpp:183-251), one buffer
 * section per topological element (boundary_communication.cpp:108-137), and the ownership
 * mask of the unpack (:296-300; mesh/forest/block_ownership.cpp:42-140; utils/indexer.hpp:
 * 163-175): an element shared by several blocks takes the value of the block that owns it —
 * highest level first, then highest (tree, Morton) number, i.e. highest gid on one level.
 *
 * Array layout: [block][element][comp][k][j][i] with every non-symmetry extent one longer than
 * the cell-centred one (metadata.cpp:383-387). */
enum { ORC_TE_CELL = 0, ORC_TE_FACE = 1, ORC_TE_EDGE = 2, ORC_TE_NODE = 3 };

int orc_te_num_elements(int kind) { return kind == ORC_TE_FACE || kind == ORC_TE_EDGE ? 3 : 1; }
/* TopologicalOffsetI/J/K of element `el` of a field of `kind` */
static void te_top_offset(int kind, int el, int top[3]) {
  for (int d = 0; d < 3; ++d) {
    if (kind == ORC_TE_CELL)
      top[d] = 0;
    else if (kind == ORC_TE_FACE)
      top[d] = d == el; /* F1: I, F2: J, F3: K */
    else if (kind == ORC_TE_EDGE)
      top[d] = d != el; /* E1: J and K, E2: I and K, E3: I and J */
    else
      top[d] = 1;
  }
}
void orc_te_extents(const OrcMesh *m, int kind, int pn[3]) {
  for (int d = 0; d < 3; ++d) pn[d] = m->n[d] + (kind != ORC_TE_CELL && m->n[d] > 1 ? 1 : 0);
}

/* LogicalLocation::IsNeighborOfTE (mesh/forest/logical_location.cpp:131-158) */
static int is_neighbor_of_te(const Loc *self, const Loc *in, const int te[3]) {
  const int maxl = in->level > self->level ? in->level : self->level;
  const long bs_in = 1L << (maxl - in->level), bs_this = 1L << (maxl - self->level);
  for (int d = 0; d < 3; ++d) {
    long low = self->lx[d] * bs_this, hi = low + bs_this - 1;
    if (te[d] == -1) {
      low -= 1;
      hi = low + 1;
    } else if (te[d] == 1) {
      hi += 1;
      low = hi - 1;
    }
    const long low_in = in->lx[d] * bs_in, hi_in = low_in + bs_in - 1;
    if (hi < low_in || low > hi_in) return 0;
  }
  return 1;
}

/* DetermineOwnership (block_ownership.cpp:42-83) of leaf block g, no newly refined blocks:
 * owns[(o1+1) + 3 (o2+1) + 9 (o3+1)].  On one tree (level, Morton number) orders leaves like
 * (level, gid). */
/* blocks created by refinement in the remesh in progress (NULL otherwise): they rank below
 * older blocks of their level, above their parents' level (block_ownership.cpp:48-54) */
static const int *g_newly_refined = NULL;
static int ownership_level(const Loc *l, int gid) {
  return g_newly_refined != NULL && g_newly_refined[gid] ? 2 * l->level - 1 : 2 * l->level;
}
static void determine_ownership(const OrcMesh *m, int g, int owns[27]) {
  const Block *blk = &m->blocks[g];
  for (int o1 = -1; o1 <= 1; ++o1)
    for (int o2 = -1; o2 <= 1; ++o2)
      for (int o3 = -1; o3 <= 1; ++o3) {
        int own = 1;
        const int te[3] = {o1, o2, o3};
        for (int n = 0; n < blk->nnb && own; ++n) {
          const Neighbor *nb = &blk->nb[n];
          const int la = ownership_level(&blk->loc, g), lb = ownership_level(&nb->loc, nb->gid);
          const int less = la != lb ? la < lb : g < nb->gid;
          if (less && is_neighbor_of_te(&blk->loc, &nb->origin_loc, te)) own = 0;
        }
        owns[(o1 + 1) + 3 * (o2 + 1) + 9 * (o3 + 1)] = own;
      }
}

/* GetIndexRangeMaskFromOwnership (block_ownership.cpp:85-140) */
#define OWN(a, i, j, k) (a)[((i) + 1) + 3 * ((j) + 1) + 9 * ((k) + 1)]
static void index_range_mask(const int top[3], const int sender[27], const int sox[3],
                             int mask[27]) {
  for (int i = -1; i <= 1; ++i)
    for (int j = -1; j <= 1; ++j)
      for (int k = -1; k <= 1; ++k)
        OWN(mask, i, j, k) = OWN(sender, top[0] ? i : 0, top[1] ? j : 0, top[2] ? k : 0);
  if (sox[0] != 0)
    for (int j = -1; j <= 1; ++j)
      for (int k = -1; k <= 1; ++k) OWN(mask, -sox[0], j, k) = OWN(mask, 0, j, k);
  if (sox[1] != 0)
    for (int i = -1; i <= 1; ++i)
      for (int k = -1; k <= 1; ++k) OWN(mask, i, -sox[1], k) = OWN(mask, i, 0, k);
  if (sox[2] != 0)
    for (int i = -1; i <= 1; ++i)
      for (int j = -1; j <= 1; ++j) OWN(mask, i, j, -sox[2]) = OWN(mask, i, j, 0);
}

/* CalcIndices for element (kind, el), same-level neighbour, non-flux (bnd_info.cpp:205-218) */
void orc_calc_indices_te(const OrcMesh *m, int b, int n, int kind, int el, int ir_type,
                         int s[3], int e[3]) {
  const Neighbor *nb = &m->blocks[b].nb[n];
  int top[3];
  te_top_offset(kind, el, top);
  const int interior_offset = ir_type == IR_SEND ? m->ng : 0;
  const int exterior_offset = ir_type == IR_RECV ? m->ng : 0;
  for (int d = 0; d < 3; ++d) {
    /* IndexShape::GetBounds(interior, el): domain.hpp:183-251 */
    const int bs = m->is[d], be = m->n[d] == 1 ? 0 : m->ie[d] + top[d];
    if (nb->off[d] == 0) {
      s[d] = bs;
      e[d] = be;
    } else if (nb->off[d] > 0) {
      s[d] = be + (-interior_offset + 1 - top[d]);
      e[d] = be + exterior_offset;
    } else {
      s[d] = bs - exterior_offset;
      e[d] = bs + (interior_offset - 1 + top[d]);
    }
  }
}

/* the ownership mask the receiving region (b, n) applies to element (kind, el):
 * bnd_info.cpp:232-248 */
void orc_te_recv_mask(const OrcMesh *m, int b, int n, int kind, int el, int mask[27]) {
  const Neighbor *nb = &m->blocks[b].nb[n];
  int top[3], sender[27];
  te_top_offset(kind, el, top);
  determine_ownership(m, nb->gid, sender);
  const int sox[3] = {-nb->off[0], -nb->off[1], -nb->off[2]};
  index_range_mask(top, sender, sox, mask);
}

/* ---- multilevel: the general CalcIndices for one element (bnd_info.cpp:105-252), the
 * element forms of RestrictAverage (pr_ops.hpp:105-165), ProlongateSharedMinMod (:167-280)
 * and ProlongateInternalAverage (:291-382), applied in the order of SendBoundBufs (restrict
 * send regions, pack), SetBounds (masked unpack, restrict set regions) and ProlongateBounds
 * (shared, then internal; boundary_communication.cpp:361-393) ---- */
typedef struct {
  int s[3], e[3];
  int mask[27];
} TeBox;
static inline int te_active(const TeBox *bx, int k, int j, int i) {
  const int ii = (i == bx->e[0]) - (i == bx->s[0]), jj = (j == bx->e[1]) - (j == bx->s[1]),
            kk = (k == bx->e[2]) - (k == bx->s[2]);
  return OWN(bx->mask, ii, jj, kk);
}

/* el is a TopologicalElement in its own right here (kind, el): the box of a cell / face /
 * edge / node element whatever the field holds (ProResInfo carries all ten, bnd_info.cpp:
 * 440-446) */
/* flux = true in CalcIndices (bnd_info.cpp:207-218): fluxes are only communicated on the
 * elements two blocks share — along a direction the neighbour is offset in, the box is the one
 * plane of the element's interior range that lies on the boundary */
static int g_te_flux_boxes = 0;
static void calc_indices_te_general(const OrcMesh *m, int b, int n, int kind, int el,
                                    int ir_type, int prores, TeBox *out) {
  const Block *blk = &m->blocks[b];
  const Neighbor *nb = &blk->nb[n];
  const Loc *loc = &blk->loc;
  const int ng = m->ng;
  int top[3];
  te_top_offset(kind, el, top);
  const int use_coarse = prores || nb->loc.level < loc->level;
  const int coarse_fac = nb->loc.level > loc->level ? 2 : 1;
  const int interior_offset = ir_type == IR_SEND ? ng : 0;
  int exterior_offset = ir_type == IR_RECV ? ng : 0;
  if (prores) exterior_offset /= 2;
  for (int d = 0; d < 3; ++d) {
    const int sym = d >= m->ndim;
    const int bs = use_coarse ? m->cis[d] : m->is[d];
    const int be = sym ? 0 : (use_coarse ? m->cie[d] : m->ie[d]) + top[d];
    const int nbs = sym ? 0 : ng, nbe = sym ? 0 : ng + m->nx[d] / coarse_fac - 1 + top[d];
    int *s = &out->s[d], *e = &out->e[d];
    if (nb->off[d] == 0) {
      *s = bs;
      *e = be;
      if (loc->level < nb->origin_loc.level && !sym) {
        const int extra = (be - bs + 1) - (nbe - nbs + 1);
        const int odd = (int)((nb->origin_loc.lx[d] % 2 + 2) % 2);
        *s += odd == 1 ? extra - interior_offset : 0;
        *e -= odd == 0 ? extra - interior_offset : 0;
      }
      if (loc->level > nb->origin_loc.level && !sym) {
        *s -= loc->lx[d] % 2 == 1 ? exterior_offset : 0;
        *e += loc->lx[d] % 2 == 0 ? exterior_offset : 0;
      }
    } else if (nb->off[d] > 0) {
      *s = be + (-interior_offset + 1 - top[d]);
      *e = be + exterior_offset;
      if (g_te_flux_boxes) *s = *e = be;
    } else {
      *s = bs - exterior_offset;
      *e = bs + (interior_offset - 1 + top[d]);
      if (g_te_flux_boxes) *s = *e = bs;
    }
  }
  for (int q = 0; q < 27; ++q) out->mask[q] = 1;
  if (ir_type == IR_RECV) { /* :232-248 */
    int sox[3] = {-nb->off[0], -nb->off[1], -nb->off[2]}, sender[27];
    if (nb->origin_loc.level < loc->level)
      for (int d = 0; d < 3; ++d)
        if (sox[d] == 0) sox[d] = loc->lx[d] % 2 == 1 ? 1 : -1;
    determine_ownership(m, nb->gid, sender);
    index_range_mask(top, sender, sox, out->mask);
  }
}

typedef struct {
  const OrcMesh *m;
  double *U, *Uc;
  int ncomp, nel, pn[3], cpn[3];
  size_t blk_sz, cblk_sz;
} TeField;
static inline double *te_f(const TeField *f, int b, int el, int c, int k, int j, int i) {
  return f->U + (size_t)b * f->blk_sz +
         ((((size_t)el * f->ncomp + c) * f->pn[2] + k) * f->pn[1] + j) * f->pn[0] + i;
}
static inline double *te_c(const TeField *f, int b, int el, int c, int k, int j, int i) {
  return f->Uc + (size_t)b * f->cblk_sz +
         ((((size_t)el * f->ncomp + c) * f->cpn[2] + k) * f->cpn[1] + j) * f->cpn[0] + i;
}

/* RestrictAverage::Do for element (kind, el): averages over the directions the element is
 * centred in (INCLUDE_Xd), weights = Volume<el> of uniform_cartesian.hpp:247-270 */
static void te_restrict(const TeField *f, int kind, int b, int el, int c, int ck, int cj,
                        int ci) {
  const OrcMesh *m = f->m;
  const Block *blk = &m->blocks[b];
  const int DIM = m->ndim;
  int top[3];
  te_top_offset(kind, el, top);
  const int inc[3] = {DIM > 0 && !top[0], DIM > 1 && !top[1], DIM > 2 && !top[2]};
  const int i = (ci - m->cis[0]) * 2 + m->is[0];
  const int j = DIM > 1 ? (cj - m->cis[1]) * 2 + m->is[1] : m->is[1];
  const int k = DIM > 2 ? (ck - m->cis[2]) * 2 + m->is[2] : m->is[2];
  const double *dx = blk->dx;
  double v;
  if (kind == ORC_TE_CELL)
    v = dx[0] * dx[1] * dx[2];
  else if (kind == ORC_TE_FACE)
    v = el == 0 ? dx[1] * dx[2] : (el == 1 ? dx[0] * dx[2] : dx[0] * dx[1]);
  else if (kind == ORC_TE_EDGE)
    v = dx[el];
  else
    v = 1.0;
  double vol[2][2][2], terms[2][2][2];
  memset(vol, 0, sizeof(vol));
  memset(terms, 0, sizeof(terms));
  for (int ok = 0; ok < 1 + inc[2]; ++ok)
    for (int oj = 0; oj < 1 + inc[1]; ++oj)
      for (int oi = 0; oi < 1 + inc[0]; ++oi) {
        vol[ok][oj][oi] = v;
        terms[ok][oj][oi] = vol[ok][oj][oi] * *te_f(f, b, el, c, k + ok, j + oj, i + oi);
      }
  const double tvol = ((vol[0][0][0] + vol[0][1][0]) + (vol[0][0][1] + vol[0][1][1])) +
                      ((vol[1][0][0] + vol[1][1][0]) + (vol[1][0][1] + vol[1][1][1]));
  *te_c(f, b, el, c, ck, cj, ci) =
      (((terms[0][0][0] + terms[0][1][0]) + (terms[0][0][1] + terms[0][1][1])) +
       ((terms[1][0][0] + terms[1][1][0]) + (terms[1][0][1] + terms[1][1][1]))) /
      tvol;
}

/* ProlongateSharedGeneral<true,false>::Do for element (kind, el): the fine elements that
 * coincide with coarse element (k, j, i); slopes only in the directions the element is
 * centred in (where GetGridSpacings uses cell-centre positions) */
static void te_prolongate_shared(const TeField *f, int kind, int b, int el, int c, int k, int j,
                                 int i, int op) {
  const OrcMesh *m = f->m;
  const Block *blk = &m->blocks[b];
  const int DIM = m->ndim;
  int top[3];
  te_top_offset(kind, el, top);
  const int inc[3] = {DIM > 0 && !top[0], DIM > 1 && !top[1], DIM > 2 && !top[2]};
  const int fi = (i - m->cis[0]) * 2 + m->is[0];
  const int fj = DIM > 1 ? (j - m->cis[1]) * 2 + m->is[1] : m->is[1];
  const int fk = DIM > 2 ? (k - m->cis[2]) * 2 + m->is[2] : m->is[2];
  const double fc = *te_c(f, b, el, c, k, j, i);
  double g[3] = {0, 0, 0}, gm[3] = {0, 0, 0}, gp[3] = {0, 0, 0}, dxfm[3] = {0, 0, 0},
         dxfp[3] = {0, 0, 0};
  const int cc[3] = {i, j, k}, ff[3] = {fi, fj, fk};
  for (int d = 0; d < 3; ++d) {
    if (!inc[d]) continue;
    const double xm = blk->ccxmin[d] + ((cc[d] - 1) + 0.5) * blk->cdx[d];
    const double xc = blk->ccxmin[d] + (cc[d] + 0.5) * blk->cdx[d];
    const double xp = blk->ccxmin[d] + ((cc[d] + 1) + 0.5) * blk->cdx[d];
    const double dxm = xc - xm, dxp = xp - xc;
    const double fxm = blk->cxmin[d] + (ff[d] + 0.5) * blk->dx[d];
    const double fxp = blk->cxmin[d] + ((ff[d] + 1) + 0.5) * blk->dx[d];
    dxfm[d] = xc - fxm;
    dxfp[d] = fxp - xc;
    int o[3] = {0, 0, 0};
    o[d] = 1;
    const double fm = *te_c(f, b, el, c, k - o[2], j - o[1], i - o[0]);
    const double fp = *te_c(f, b, el, c, k + o[2], j + o[1], i + o[0]);
    g[d] = grad_minmod(fc, fm, fp, dxm, dxp);
    /* ProlongateSharedGeneral<use_minmod_slope, piecewise_constant> pr_ops.hpp:206-262:
     * 0 MinMod (both one-sided slopes replaced by the limited one), 1 Linear (the one-sided
     * slopes themselves), 2 PiecewiseConstant (zero slopes) */
    gm[d] = op == 0 ? g[d] : (op == 1 ? (fc - fm) / dxm : 0.0);
    gp[d] = op == 0 ? g[d] : (op == 1 ? (fp - fc) / dxp : 0.0);
  }
  const double gx1m = gm[0], gx1p = gp[0], gx2m = gm[1], gx2p = gp[1], gx3m = gm[2],
               gx3p = gp[2];
  const double dx1fm = dxfm[0], dx1fp = dxfp[0], dx2fm = dxfm[1], dx2fp = dxfp[1],
               dx3fm = dxfm[2], dx3fp = dxfp[2];
  *te_f(f, b, el, c, fk, fj, fi) = fc - (gx1m * dx1fm + gx2m * dx2fm + gx3m * dx3fm);
  if (inc[0])
    *te_f(f, b, el, c, fk, fj, fi + 1) = fc + (gx1p * dx1fp - gx2m * dx2fm - gx3m * dx3fm);
  if (inc[1])
    *te_f(f, b, el, c, fk, fj + 1, fi) = fc - (gx1m * dx1fm - gx2p * dx2fp + gx3m * dx3fm);
  if (inc[1] && inc[0])
    *te_f(f, b, el, c, fk, fj + 1, fi + 1) = fc + (gx1p * dx1fp + gx2p * dx2fp - gx3m * dx3fm);
  if (inc[2])
    *te_f(f, b, el, c, fk + 1, fj, fi) = fc - (gx1m * dx1fm + gx2m * dx2fm - gx3p * dx3fp);
  if (inc[2] && inc[0])
    *te_f(f, b, el, c, fk + 1, fj, fi + 1) = fc + (gx1p * dx1fp - gx2m * dx2fm + gx3p * dx3fp);
  if (inc[2] && inc[1])
    *te_f(f, b, el, c, fk + 1, fj + 1, fi) = fc - (gx1m * dx1fm - gx2p * dx2fp - gx3p * dx3fp);
  if (inc[2] && inc[1] && inc[0])
    *te_f(f, b, el, c, fk + 1, fj + 1, fi + 1) =
        fc + (gx1p * dx1fp + gx2p * dx2fp + gx3p * dx3fp);
}

/* IsSubmanifold(containee, container), basic_types.hpp:207-232, on (kind, el) pairs: an
 * element is a boundary of another iff it is displaced in every direction the container is
 * and in at least one more */
static int te_is_submanifold(const int ftop[3], const int ctop[3]) {
  int more = 0;
  for (int d = 0; d < 3; ++d) {
    if (ctop[d] && !ftop[d]) return 0;
    more += ftop[d] && !ctop[d];
  }
  return more > 0;
}

/* ProlongateInternalAverage::Do<DIM, fel, cel>: fine elements of the field (ftop) strictly
 * inside coarse element cel (ctop) at coarse position (k, j, i) = the average of the fine
 * shared elements around them */
static void te_prolongate_internal(const TeField *f, const int ftop[3], const int ctop[3], int b,
                                   int el, int c, int k, int j, int i) {
  const OrcMesh *m = f->m;
  const int DIM = m->ndim;
  const int fi = (i - m->cis[0]) * 2 + m->is[0];
  const int fj = DIM > 1 ? (j - m->cis[1]) * 2 + m->is[1] : m->is[1];
  const int fk = DIM > 2 ? (k - m->cis[2]) * 2 + m->is[2] : m->is[2];
  int center[3], stencil[3];
  for (int d = 0; d < 3; ++d) {
    center[d] = d < DIM && !ftop[d];
    stencil[d] = d < DIM && !center[d] && !ctop[d];
  }
  const double w = 1.0 / ((1.0 + stencil[2]) * (1.0 + stencil[1]) * (1.0 + stencil[0]));
  for (int ok = 0; ok < 1 + center[2]; ++ok)
    for (int oj = 0; oj < 1 + center[1]; ++oj)
      for (int oi = 0; oi < 1 + center[0]; ++oi) {
        const int tk = fk + ok + stencil[2], tj = fj + oj + stencil[1],
                  ti = fi + oi + stencil[0];
        double v = 0.0;
        for (int stk = -stencil[2]; stk <= stencil[2]; stk += 2)
          for (int stj = -stencil[1]; stj <= stencil[1]; stj += 2)
            for (int sti = -stencil[0]; sti <= stencil[0]; sti += 2)
              v += w * *te_f(f, b, el, c, tk + stk, tj + stj, ti + sti);
        *te_f(f, b, el, c, tk, tj, ti) = v;
      }
}

/* ProlongateInternalTothAndRoe::Do<DIM, fel, CC> pr_ops.hpp:384-470: the fine faces of
 * element el (0..2 = F1..F3) inside coarse cell (k, j, i) from the fine faces on the cell's
 * surface so that the fine divergence equals the coarse one (Toth & Roe 2002); written for the
 * x-component, the others by cyclic permutation */
static void te_toth_roe(const TeField *f, int b, int el, int c, int k, int j, int i) {
  const OrcMesh *m = f->m;
  const Block *blk = &m->blocks[b];
  const int DIM = m->ndim;
  const int fi = (i - m->cis[0]) * 2 + m->is[0];
  const int fj = DIM > 1 ? (j - m->cis[1]) * 2 + m->is[1] : m->is[1];
  const int fk = DIM > 2 ? (k - m->cis[2]) * 2 + m->is[2] : m->is[2];
  const int g3 = DIM > 2, g2 = DIM > 1;
#define FP(eidx, ok, oj, oi)                                                                  \
  (el == 0   ? te_f(f, b, (el + (eidx)) % 3, c, fk + (ok)*g3, fj + (oj)*g2, fi + (oi))         \
   : el == 1 ? te_f(f, b, (el + (eidx)) % 3, c, fk + (oj)*g3, fj + (oi)*g2, fi + (ok))         \
             : te_f(f, b, (el + (eidx)) % 3, c, fk + (oi)*g3, fj + (ok)*g2, fi + (oj)))
#define SG(o) ((o) == 0 ? -1.0 : 1.0)
  double Uxx = 0.0, Vxyz = 0.0, Wxyz = 0.0;
  for (int v = 0; v <= 1; ++v)
    for (int u = 0; u <= 2; u += 2)
      for (int t = 0; t <= 1; ++t) {
        const double fine2 = *FP(1, v, u, t);
        const double fine3 = *FP(2, u, v, t);
        Uxx += SG(t) * SG(u) * (fine2 + fine3);
        Vxyz += SG(t) * SG(u) * SG(v) * fine2;
        Wxyz += SG(t) * SG(u) * SG(v) * fine3;
      }
  Uxx *= 0.125;
  const double d1 = blk->cdx[el], d2 = blk->cdx[(el + 1) % 3], d3 = blk->cdx[(el + 2) % 3];
  const double dx2 = d1 * d1, dy2 = d2 * d2, dz2 = d3 * d3; /* std::pow(x, 2) */
  Vxyz *= 0.125 * dz2 / (dx2 + dz2);
  Wxyz *= 0.125 * dy2 / (dx2 + dy2);
  for (int ok = 0; ok <= 1; ++ok)
    for (int oj = 0; oj <= 1; ++oj)
      *FP(0, ok, oj, 1) =
          0.5 * (*FP(0, ok, oj, 0) + *FP(0, ok, oj, 2)) + Uxx + SG(ok) * Vxyz + SG(oj) * Wxyz;
#undef FP
#undef SG
}

/* GenericBC<DIR, SIDE, Outflow | Reflect> for the elements of a face / edge / node field
 * (boundary_conditions_generic.hpp:174-268; no Metadata::Vector components, so reflections
 * keep the sign): for element el the reference index along the boundary direction is the
 * first / last INTERIOR entry of the element (the boundary face itself for an element
 * displaced in that direction) and the ghost slab covers the element's whole index range in
 * the other directions (mesh/domain.hpp:183-251) */
static void te_apply_bcs(const TeField *f, int kind, int coarse) {
  const OrcMesh *m = f->m;
  const int *is = coarse ? m->cis : m->is, *ie = coarse ? m->cie : m->ie,
            *nn = coarse ? m->cn : m->n;
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    for (int face = 0; face < 6; ++face) {
      const int d = face / 2, inner = (face % 2) == 0;
      if (d >= m->ndim || m->bc[face] == 0) continue;
      const long nb_d = nblocks_at(m, blk->loc.level, d);
      if (inner ? blk->loc.lx[d] != 0 : blk->loc.lx[d] != nb_d - 1) continue;
      for (int el = 0; el < f->nel; ++el) {
        int top[3], lo[3], hi[3];
        te_top_offset(kind, el, top);
        for (int q = 0; q < 3; ++q) {
          lo[q] = 0;
          hi[q] = nn[q] == 1 ? 0 : nn[q] - 1 + top[q];
        }
        const int ref = inner ? is[d] : ie[d] + top[d];
        const int offset = 2 * ref + (inner ? -1 : 1);
        if (inner)
          hi[d] = is[d] - 1;
        else
          lo[d] = ie[d] + 1 + top[d];
        for (int c = 0; c < f->ncomp; ++c)
          for (int k = lo[2]; k <= hi[2]; ++k)
            for (int j = lo[1]; j <= hi[1]; ++j)
              for (int i = lo[0]; i <= hi[0]; ++i) {
                int sidx[3] = {i, j, k};
                sidx[d] = m->bc[face] == 2 ? offset - sidx[d] : ref;
                if (coarse)
                  *te_c(f, b, el, c, k, j, i) = *te_c(f, b, el, c, sidx[2], sidx[1], sidx[0]);
                else
                  *te_f(f, b, el, c, k, j, i) = *te_f(f, b, el, c, sidx[2], sidx[1], sidx[0]);
              }
      }
    }
  }
}

/* kinds and in-kind element numbers of the ten TopologicalElements in the order the internal
 * prolongation visits the containers: NN, E3, E2, E1, F1, F2, F3, CC (pr_loops.hpp:82-108) */
static const int kCelKind[8] = {ORC_TE_NODE, ORC_TE_EDGE, ORC_TE_EDGE, ORC_TE_EDGE,
                                ORC_TE_FACE, ORC_TE_FACE, ORC_TE_FACE, ORC_TE_CELL};
static const int kCelEl[8] = {0, 2, 1, 0, 0, 1, 2, 0};

static int64_t exchange_te_impl(const OrcMesh *m, double *U, double *Uc, int ncomp, int kind,
                                int toth_roe, int shared_op);
int64_t orc_exchange_te_ml(const OrcMesh *m, double *U, double *Uc, int ncomp, int kind) {
  return exchange_te_impl(m, U, Uc, ncomp, kind, 0, 0);
}
/* shared_op: 0 ProlongateSharedMinMod, 1 ProlongateSharedLinear, 2 ProlongatePiecewiseConstant */
int64_t orc_exchange_te_ml_op(const OrcMesh *m, double *U, double *Uc, int ncomp, int kind,
                              int shared_op) {
  return exchange_te_impl(m, U, Uc, ncomp, kind, 0, shared_op);
}
/* the general index box and receive mask of region (b, n), element (kind, el), for tests */
void orc_calc_indices_te_general(const OrcMesh *m, int b, int n, int kind, int el, int ir_type,
                                 int prores, int s[3], int e[3], int mask[27]) {
  TeBox bx;
  calc_indices_te_general(m, b, n, kind, el, ir_type, prores, &bx);
  for (int d = 0; d < 3; ++d) {
    s[d] = bx.s[d];
    e[d] = bx.e[d];
  }
  for (int q = 0; q < 27; ++q) mask[q] = bx.mask[q];
}
/* a face field that registered ProlongateInternalTothAndRoe as its internal prolongation */
int64_t orc_exchange_te_ml_toth_roe(const OrcMesh *m, double *U, double *Uc, int ncomp) {
  return exchange_te_impl(m, U, Uc, ncomp, ORC_TE_FACE, 1, 0);
}
static int64_t exchange_te_impl(const OrcMesh *m, double *U, double *Uc, int ncomp, int kind,
                                int toth_roe, int shared_op) {
  TeField F;
  F.m = m;
  F.U = U;
  F.Uc = Uc;
  F.ncomp = ncomp;
  F.nel = orc_te_num_elements(kind);
  orc_te_extents(m, kind, F.pn);
  for (int d = 0; d < 3; ++d) F.cpn[d] = m->cn[d] + (kind != ORC_TE_CELL && m->cn[d] > 1 ? 1 : 0);
  F.blk_sz = (size_t)F.nel * ncomp * F.pn[2] * F.pn[1] * F.pn[0];
  F.cblk_sz = (size_t)F.nel * ncomp * F.cpn[2] * F.cpn[1] * F.cpn[0];
  const int nel = F.nel;
  if (m->multilevel && !Uc) {
    fprintf(stderr, "oracle: multilevel exchange needs coarse buffers\n");
    abort();
  }
  int64_t nreg = orc_count_regions(m);
  int64_t *off = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nreg + 1));
  int64_t *first = (int64_t *)malloc(sizeof(int64_t) * (size_t)(m->nblocks + 1));
  first[0] = 0;
  for (int b = 0; b < m->nblocks; ++b) first[b + 1] = first[b] + m->blocks[b].nnb;
  TeBox bx;
  /* SendBoundBufs: restriction of the send regions that face a coarser block (:82-87) */
  if (m->multilevel)
    for (int b = 0; b < m->nblocks; ++b) {
      const Block *blk = &m->blocks[b];
      for (int n = 0; n < blk->nnb; ++n) {
        if (!(blk->nb[n].origin_loc.level < blk->loc.level)) continue;
        for (int el = 0; el < nel; ++el) {
          calc_indices_te_general(m, b, n, kind, el, IR_SEND, 1, &bx);
          for (int c = 0; c < ncomp; ++c)
            for (int k = bx.s[2]; k <= bx.e[2]; ++k)
              for (int j = bx.s[1]; j <= bx.e[1]; ++j)
                for (int i = bx.s[0]; i <= bx.e[0]; ++i) te_restrict(&F, kind, b, el, c, k, j, i);
        }
      }
    }
  /* pack */
  int64_t total = 0, r = 0;
  for (int b = 0; b < m->nblocks; ++b)
    for (int n = 0; n < m->blocks[b].nnb; ++n) {
      off[r++] = total;
      for (int el = 0; el < nel; ++el) {
        calc_indices_te_general(m, b, n, kind, el, IR_SEND, 0, &bx);
        total += (int64_t)ncomp * (bx.e[2] - bx.s[2] + 1) * (bx.e[1] - bx.s[1] + 1) *
                 (bx.e[0] - bx.s[0] + 1);
      }
    }
  off[r] = total;
  double *buf = (double *)malloc(sizeof(double) * (size_t)(total > 0 ? total : 1));
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    for (int n = 0; n < blk->nnb; ++n) {
      double *p = buf + off[first[b] + n];
      const int coarse = blk->nb[n].origin_loc.level < blk->loc.level;
      for (int el = 0; el < nel; ++el) {
        calc_indices_te_general(m, b, n, kind, el, IR_SEND, 0, &bx);
        for (int c = 0; c < ncomp; ++c)
          for (int k = bx.s[2]; k <= bx.e[2]; ++k)
            for (int j = bx.s[1]; j <= bx.e[1]; ++j)
              for (int i = bx.s[0]; i <= bx.e[0]; ++i)
                *p++ = coarse ? *te_c(&F, b, el, c, k, j, i) : *te_f(&F, b, el, c, k, j, i);
      }
    }
  }
  /* SetBounds: unpack under the ownership mask (:296-300) */
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    for (int n = 0; n < blk->nnb; ++n) {
      const Neighbor *nb = &blk->nb[n];
      const Block *sb = &m->blocks[nb->gid];
      int sn = -1;
      for (int q = 0; q < sb->nnb; ++q)
        if (sb->nb[q].gid == b && sb->nb[q].off[0] == -nb->off[0] &&
            sb->nb[q].off[1] == -nb->off[1] && sb->nb[q].off[2] == -nb->off[2]) {
          sn = q;
          break;
        }
      if (sn < 0) abort();
      const double *p = buf + off[first[nb->gid] + sn];
      const int coarse = nb->origin_loc.level < blk->loc.level;
      for (int el = 0; el < nel; ++el) {
        TeBox sx;
        calc_indices_te_general(m, b, n, kind, el, IR_RECV, 0, &bx);
        calc_indices_te_general(m, nb->gid, sn, kind, el, IR_SEND, 0, &sx);
        for (int d = 0; d < 3; ++d)
          if (bx.e[d] - bx.s[d] != sx.e[d] - sx.s[d]) {
            fprintf(stderr, "oracle: send/recv box mismatch (block %d nb %d el %d dir %d)\n", b,
                    n, el, d);
            abort();
          }
        for (int c = 0; c < ncomp; ++c)
          for (int k = bx.s[2]; k <= bx.e[2]; ++k)
            for (int j = bx.s[1]; j <= bx.e[1]; ++j)
              for (int i = bx.s[0]; i <= bx.e[0]; ++i, ++p)
                if (te_active(&bx, k, j, i)) {
                  if (coarse)
                    *te_c(&F, b, el, c, k, j, i) = *p;
                  else
                    *te_f(&F, b, el, c, k, j, i) = *p;
                }
      }
    }
  }
  if (m->multilevel) {
    /* restriction of the received regions of blocks that have a coarser neighbour
     * (ProResInfo::GetSet, bnd_info.cpp:405-431; SetBounds :338-346) */
    for (int b = 0; b < m->nblocks; ++b) {
      const Block *blk = &m->blocks[b];
      int restricted = 0;
      if (blk->loc.level > 0)
        for (int n = 0; n < blk->nnb; ++n)
          restricted = restricted || (blk->nb[n].origin_loc.level == blk->loc.level - 1);
      if (!restricted) continue;
      for (int n = 0; n < blk->nnb; ++n) {
        if (blk->nb[n].origin_loc.level < blk->loc.level) continue;
        for (int el = 0; el < nel; ++el) {
          calc_indices_te_general(m, b, n, kind, el, IR_RECV, 1, &bx);
          for (int c = 0; c < ncomp; ++c)
            for (int k = bx.s[2]; k <= bx.e[2]; ++k)
              for (int j = bx.s[1]; j <= bx.e[1]; ++j)
                for (int i = bx.s[0]; i <= bx.e[0]; ++i)
                  if (te_active(&bx, k, j, i)) te_restrict(&F, kind, b, el, c, k, j, i);
        }
      }
    }
    /* physical boundaries of the coarse buffers (boundary_communication.cpp:445-449) */
    te_apply_bcs(&F, kind, 1);
    /* ProlongateBounds: shared elements first ... */
    for (int b = 0; b < m->nblocks; ++b) {
      const Block *blk = &m->blocks[b];
      for (int n = 0; n < blk->nnb; ++n) {
        if (!(blk->nb[n].origin_loc.level < blk->loc.level)) continue;
        for (int el = 0; el < nel; ++el) {
          calc_indices_te_general(m, b, n, kind, el, IR_RECV, 1, &bx);
          for (int c = 0; c < ncomp; ++c)
            for (int k = bx.s[2]; k <= bx.e[2]; ++k)
              for (int j = bx.s[1]; j <= bx.e[1]; ++j)
                for (int i = bx.s[0]; i <= bx.e[0]; ++i)
                  if (te_active(&bx, k, j, i))
                    te_prolongate_shared(&F, kind, b, el, c, k, j, i, shared_op);
        }
      }
    }
    /* ... then the fine elements inside coarse edges, faces and cells */
    for (int b = 0; b < m->nblocks; ++b) {
      const Block *blk = &m->blocks[b];
      for (int n = 0; n < blk->nnb; ++n) {
        if (!(blk->nb[n].origin_loc.level < blk->loc.level)) continue;
        for (int el = 0; el < nel; ++el) {
          int ftop[3];
          te_top_offset(kind, el, ftop);
          for (int q = 0; q < 8; ++q) {
            int ctop[3];
            te_top_offset(kCelKind[q], kCelEl[q], ctop);
            /* OperationRequired: Toth & Roe fills coarse CELLS only (pr_ops.hpp:390-393) */
            if (toth_roe ? kCelKind[q] != ORC_TE_CELL : !te_is_submanifold(ftop, ctop)) continue;
            calc_indices_te_general(m, b, n, kCelKind[q], kCelEl[q], IR_RECV, 1, &bx);
            for (int c = 0; c < ncomp; ++c)
              for (int k = bx.s[2]; k <= bx.e[2]; ++k)
                for (int j = bx.s[1]; j <= bx.e[1]; ++j)
                  for (int i = bx.s[0]; i <= bx.e[0]; ++i)
                    if (te_active(&bx, k, j, i)) {
                      if (toth_roe)
                        te_toth_roe(&F, b, el, c, k, j, i);
                      else
                        te_prolongate_internal(&F, ftop, ctop, b, el, c, k, j, i);
                    }
          }
        }
      }
    }
  }
  te_apply_bcs(&F, kind, 0); /* and of the fine arrays, after the prolongation */
  free(buf);
  free(off);
  free(first);
  return total;
}
#undef OWN

/* ------------------------------------------------------------------------------------ */
/* Flux correction of a FACE field: its flux is an EDGE field (Metadata::Flux | Edge,
 * state_descriptor.cpp:313-318).  A fine block restricts (RestrictAverage per element) the edge
 * fluxes it shares with a neighbour ONE level coarser — across a face the two edge elements
 * tangent to it, across a block edge the one element along it (GetFluxCorrectionElements,
 * bnd_info.cpp:71-103; neighbour filter of ForEachBoundary<flxcor_*>, loop_utils.hpp:134-158) —
 * into its coarse buffer (ProResInfo::GetSend :387-403), sends them from there, and the coarse
 * block writes those the sending block owns into its flux array (SetBounds<flxcor_recv> with
 * the mask of block_ownership.cpp:85-140).  F: [block][3 elements][ncomp][k][j][i] with the
 * extents of orc_te_extents(kind = edge); Fc: the coarse buffers. */
static int te_flux_elements(const int off[3], int els[2]) {
  const int nz = (off[0] != 0) + (off[1] != 0) + (off[2] != 0);
  if (nz == 1) { /* face: E2,E3 | E3,E1 | E1,E2 */
    if (off[0]) { els[0] = 1; els[1] = 2; }
    if (off[1]) { els[0] = 2; els[1] = 0; }
    if (off[2]) { els[0] = 0; els[1] = 1; }
    return 2;
  }
  if (nz == 2) { /* edge: the element along it */
    els[0] = off[0] == 0 ? 0 : (off[1] == 0 ? 1 : 2);
    return 1;
  }
  return 0;
}

/* Where the reference is not deterministic: a coarse edge entry on the rim of a face can be
 * delivered twice, by the fine block across the face and by a fine block across the block edge
 * (2-D: corner) next to it.  The mask consults ownership only at the two ends of a box range; a
 * range of ONE entry — every direction a flux box is offset in — counts as interior
 * (indexer.hpp:171-173), so both messages are active, they carry the two fine blocks' own
 * values, and SetBounds unpacks the messages in a shuffled order on parallel teams
 * (bvals_utils.hpp:108-112).  Here the messages across faces are unpacked after those across
 * block edges, so the face neighbour's value stays; wcount (optional, one int per entry of F)
 * counts the deliveries, which lets a test leave doubly delivered entries out of a comparison
 * with a reference dump. */
int64_t orc_flux_correct_edge(const OrcMesh *m, double *F_, double *Fc, int ncomp, int *wcount) {
  const int kind = ORC_TE_EDGE;
  TeField F;
  F.m = m;
  F.U = F_;
  F.Uc = Fc;
  F.ncomp = ncomp;
  F.nel = 3;
  orc_te_extents(m, kind, F.pn);
  for (int d = 0; d < 3; ++d) F.cpn[d] = m->cn[d] + (m->cn[d] > 1 ? 1 : 0);
  F.blk_sz = (size_t)F.nel * ncomp * F.pn[2] * F.pn[1] * F.pn[0];
  F.cblk_sz = (size_t)F.nel * ncomp * F.cpn[2] * F.cpn[1] * F.cpn[0];
  if (!m->multilevel) return 0;
  g_te_flux_boxes = 1;
  TeBox bx;
  int64_t *first = (int64_t *)malloc(sizeof(int64_t) * (size_t)(m->nblocks + 1));
  first[0] = 0;
  for (int b = 0; b < m->nblocks; ++b) first[b + 1] = first[b] + m->blocks[b].nnb;
  int64_t *off = (int64_t *)calloc((size_t)first[m->nblocks] + 1, sizeof(int64_t));
  /* send side: restrict, then size and fill the buffers */
  int64_t total = 0;
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    for (int n = 0; n < blk->nnb; ++n) {
      off[first[b] + n] = total;
      const Neighbor *nb = &blk->nb[n];
      int els[2];
      const int ne = nb->loc.level == blk->loc.level - 1 ? te_flux_elements(nb->off, els) : 0;
      for (int q = 0; q < ne; ++q) {
        calc_indices_te_general(m, b, n, kind, els[q], IR_SEND, 1, &bx);
        for (int c = 0; c < ncomp; ++c)
          for (int k = bx.s[2]; k <= bx.e[2]; ++k)
            for (int j = bx.s[1]; j <= bx.e[1]; ++j)
              for (int i = bx.s[0]; i <= bx.e[0]; ++i) te_restrict(&F, kind, b, els[q], c, k, j, i);
        calc_indices_te_general(m, b, n, kind, els[q], IR_SEND, 0, &bx);
        total += (int64_t)ncomp * (bx.e[2] - bx.s[2] + 1) * (bx.e[1] - bx.s[1] + 1) *
                 (bx.e[0] - bx.s[0] + 1);
      }
    }
  }
  double *buf = (double *)malloc(sizeof(double) * (size_t)(total > 0 ? total : 1));
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    for (int n = 0; n < blk->nnb; ++n) {
      const Neighbor *nb = &blk->nb[n];
      int els[2];
      const int ne = nb->loc.level == blk->loc.level - 1 ? te_flux_elements(nb->off, els) : 0;
      double *p = buf + off[first[b] + n];
      for (int q = 0; q < ne; ++q) {
        calc_indices_te_general(m, b, n, kind, els[q], IR_SEND, 0, &bx);
        for (int c = 0; c < ncomp; ++c)
          for (int k = bx.s[2]; k <= bx.e[2]; ++k)
            for (int j = bx.s[1]; j <= bx.e[1]; ++j)
              for (int i = bx.s[0]; i <= bx.e[0]; ++i) *p++ = *te_c(&F, b, els[q], c, k, j, i);
      }
    }
  }
  /* receive side: the coarser block, under the ownership mask of the sender; messages across
   * block edges first, then those across faces */
  for (int pass = 0; pass < 2; ++pass)
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    for (int n = 0; n < blk->nnb; ++n) {
      const Neighbor *nb = &blk->nb[n];
      int els[2];
      const int ne = nb->loc.level == blk->loc.level + 1 ? te_flux_elements(nb->off, els) : 0;
      if (ne == 0 || ne != pass + 1) continue; /* one element: block edge; two: face */
      const Block *sb = &m->blocks[nb->gid];
      int sn = -1;
      for (int q = 0; q < sb->nnb; ++q)
        if (sb->nb[q].gid == b && sb->nb[q].off[0] == -nb->off[0] &&
            sb->nb[q].off[1] == -nb->off[1] && sb->nb[q].off[2] == -nb->off[2]) {
          sn = q;
          break;
        }
      if (sn < 0) {
        fprintf(stderr, "oracle: no matching flux-correction sender (block %d nb %d)\n", b, n);
        abort();
      }
      const double *p = buf + off[first[nb->gid] + sn];
      for (int q = 0; q < ne; ++q) {
        TeBox sx;
        calc_indices_te_general(m, b, n, kind, els[q], IR_RECV, 0, &bx);
        calc_indices_te_general(m, nb->gid, sn, kind, els[q], IR_SEND, 0, &sx);
        for (int d = 0; d < 3; ++d)
          if (bx.e[d] - bx.s[d] != sx.e[d] - sx.s[d]) {
            fprintf(stderr, "oracle: flux-correction box mismatch (block %d nb %d el %d dir %d)\n",
                    b, n, els[q], d);
            abort();
          }
        for (int c = 0; c < ncomp; ++c)
          for (int k = bx.s[2]; k <= bx.e[2]; ++k)
            for (int j = bx.s[1]; j <= bx.e[1]; ++j)
              for (int i = bx.s[0]; i <= bx.e[0]; ++i, ++p)
                if (te_active(&bx, k, j, i)) {
                  double *dst = te_f(&F, b, els[q], c, k, j, i);
                  *dst = *p;
                  if (wcount) wcount[dst - F.U]++;
                }
      }
    }
  }
  g_te_flux_boxes = 0;
  free(buf);
  free(off);
  free(first);
  return total;
}

int64_t orc_exchange_te(const OrcMesh *m, double *U, int ncomp, int kind) {
  if (m->multilevel) {
    fprintf(stderr, "oracle: orc_exchange_te is the uniform-mesh form; use orc_exchange_te_ml\n");
    abort();
  }
  return orc_exchange_te_ml(m, U, NULL, ncomp, kind);
}

/* ------------------------------------------------------------------------------------ */
/* physical boundary conditions: ApplyBoundaryConditionsOnCoarseOrFine
 * (bvals/boundary_conditions.cpp:36-58) with the generic outflow / reflect functions
 * (boundary_conditions_generic.hpp:174-268) on the fine arrays of cell-centred fields
 * without a Metadata::Vector component (no sign flip).  Faces in BoundaryFace order; the
 * ghost slab of a face spans the ENTIRE extents of the other directions
 * (mesh/domain.hpp:183-251), so edges and corners outside the mesh come out right when the
 * faces are applied in order. */
static void apply_bcs_generic(const OrcMesh *m, double *A, int ncomp, int coarse) {
  const int *is = coarse ? m->cis : m->is, *ie = coarse ? m->cie : m->ie,
            *nn = coarse ? m->cn : m->n;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    for (int face = 0; face < 6; ++face) {
      const int d = face / 2, inner = (face % 2) == 0;
      int flag = m->bc[face];
      if (m->custom) {
        flag = blk->bc[face]; /* the block's own flags (Tree::GetBlockBCs, tree.cpp:320-332) */
        if (d >= m->ndim || flag == 0) continue;
      } else {
        if (d >= m->ndim || m->bc[face] == 0) continue;
        /* MeshBlock::boundary_flag: the mesh flag where the block touches the mesh boundary */
        const long nb_d = nblocks_at(m, blk->loc.level, d);
        if (inner ? blk->loc.lx[d] != 0 : blk->loc.lx[d] != nb_d - 1) continue;
      }
      const int ref = inner ? is[d] : ie[d];
      const int offset = 2 * ref + (inner ? -1 : 1);
      int lo[3] = {0, 0, 0}, hi[3] = {nn[0] - 1, nn[1] - 1, nn[2] - 1};
      if (inner)
        hi[d] = is[d] - 1;
      else
        lo[d] = ie[d] + 1;
      for (int c = 0; c < ncomp; ++c)
        for (int k = lo[2]; k <= hi[2]; ++k)
          for (int j = lo[1]; j <= hi[1]; ++j)
            for (int i = lo[0]; i <= hi[0]; ++i) {
              int s[3] = {i, j, k};
              s[d] = flag == 2 ? offset - s[d] : ref;
              if (coarse)
                A[cidx(m, ncomp, b, c, k, j, i)] = A[cidx(m, ncomp, b, c, s[2], s[1], s[0])];
              else
                A[fidx(m, ncomp, b, c, k, j, i)] = A[fidx(m, ncomp, b, c, s[2], s[1], s[0])];
            }
    }
  }
}
void orc_apply_bcs(const OrcMesh *m, double *U, int ncomp) { apply_bcs_generic(m, U, ncomp, 0); }
/* the same on the coarse buffers (coarse = true, c_cellbounds): done between SetBounds and
 * ProlongateBounds on multilevel meshes (boundary_communication.cpp:445-449, mesh.cpp:701-703) */
void orc_apply_bcs_coarse(const OrcMesh *m, double *Uc, int ncomp) {
  apply_bcs_generic(m, Uc, ncomp, 1);
}

/* ------------------------------------------------------------------------------------ */
/* burgers: benchmarks/burgers/recon.hpp, burgers_package.{hpp,cpp} */

#define ORC_EPS (10.0 * DBL_EPSILON) /* src/utils/robust.hpp:39-42 */

/* recon.hpp:27-32 */
static inline double mc(double dm, double dp) {
  const double dc = (dm * dp > 0.0) * 0.5 * (dm + dp);
  return copysign(fmin(fabs(dc), 2.0 * fmin(fabs(dm), fabs(dp))), dc);
}

/* recon.hpp:34-40 */
void orc_linear(double qm, double q0, double qp, double *ql, double *qr) {
  double dq = qp - q0;
  dq = 0.5 * mc(q0 - qm, dq);
  *ql = q0 + dq;
  *qr = q0 - dq;
}

/* recon.hpp:42-99 */
void orc_weno5z(double q0, double q1, double q2, double q3, double q4, double *pql,
                double *pqr) {
  static const double w5alpha[3][3] = {{1.0 / 3.0, -7.0 / 6.0, 11.0 / 6.0},
                                       {-1.0 / 6.0, 5.0 / 6.0, 1.0 / 3.0},
                                       {1.0 / 3.0, 5.0 / 6.0, -1.0 / 6.0}};
  static const double w5gamma[3] = {0.1, 0.6, 0.3};
  const double eps = ORC_EPS;
  const double thirteen_thirds = 13.0 / 3.0;
  double ql, qr;

  double a = q0 - 2 * q1 + q2;
  double b = q0 - 4.0 * q1 + 3.0 * q2;
  double beta0 = thirteen_thirds * a * a + b * b + eps;
  a = q1 - 2.0 * q2 + q3;
  b = q3 - q1;
  double beta1 = thirteen_thirds * a * a + b * b + eps;
  a = q2 - 2.0 * q3 + q4;
  b = q4 - 4.0 * q3 + 3.0 * q2;
  double beta2 = thirteen_thirds * a * a + b * b + eps;
  const double tau5 = fabs(beta2 - beta0);

  beta0 = (beta0 + tau5) / beta0;
  beta1 = (beta1 + tau5) / beta1;
  beta2 = (beta2 + tau5) / beta2;

  double w0 = w5gamma[0] * beta0 + eps;
  double w1 = w5gamma[1] * beta1 + eps;
  double w2 = w5gamma[2] * beta2 + eps;
  double wsum = 1.0 / (w0 + w1 + w2);
  ql = w0 * (w5alpha[0][0] * q0 + w5alpha[0][1] * q1 + w5alpha[0][2] * q2);
  ql += w1 * (w5alpha[1][0] * q1 + w5alpha[1][1] * q2 + w5alpha[1][2] * q3);
  ql += w2 * (w5alpha[2][0] * q2 + w5alpha[2][1] * q3 + w5alpha[2][2] * q4);
  ql *= wsum;
  const double alpha_l =
      3.0 * wsum * w0 * w1 * w2 /
          (w5gamma[2] * w0 * w1 + w5gamma[1] * w0 * w2 + w5gamma[0] * w1 * w2) +
      eps;

  w0 = w5gamma[0] * beta2 + eps;
  w1 = w5gamma[1] * beta1 + eps;
  w2 = w5gamma[2] * beta0 + eps;
  wsum = 1.0 / (w0 + w1 + w2);
  qr = w0 * (w5alpha[0][0] * q4 + w5alpha[0][1] * q3 + w5alpha[0][2] * q2);
  qr += w1 * (w5alpha[1][0] * q3 + w5alpha[1][1] * q2 + w5alpha[1][2] * q1);
  qr += w2 * (w5alpha[2][0] * q2 + w5alpha[2][1] * q1 + w5alpha[2][2] * q0);
  qr *= wsum;
  const double alpha_r =
      3.0 * wsum * w0 * w1 * w2 /
          (w5gamma[2] * w0 * w1 + w5gamma[1] * w0 * w2 + w5gamma[0] * w1 * w2) +
      eps;

  double dq = q3 - q2;
  dq = 0.5 * mc(q2 - q1, dq);

  const double alpha_lin = 2.0 * alpha_l * alpha_r / (alpha_l + alpha_r);
  ql = alpha_lin * ql + (1.0 - alpha_lin) * (q2 + dq);
  qr = alpha_lin * qr + (1.0 - alpha_lin) * (q2 - dq);
  *pql = ql;
  *pqr = qr;
}

/* burgers_package.hpp:31-43 */
void orc_lr_to_flux(double uxl, double uxr, double uyl, double uyr, double uzl,
                    double uzr, double upl, double upr, double *psl, double *psr,
                    double *fux, double *fuy, double *fuz) {
  const double sl = fmin(fmin(upl, upr), 0.0);
  const double sr = fmax(fmax(upl, upr), 0.0);
  const double islsr = 1.0 / (sr - sl + (sl * sr == 0.0));
  *fux = 0.5 * (sr * uxl * upl - sl * uxr * upr + sl * sr * (uxr - uxl)) * islsr;
  *fuy = 0.5 * (sr * uyl * upl - sl * uyr * upr + sl * sr * (uyr - uyl)) * islsr;
  *fuz = 0.5 * (sr * uzl * upl - sl * uzr * upr + sl * sr * (uzr - uzl)) * islsr;
  *psl = sl;
  *psr = sr;
}

struct OrcBurgers {
  const OrcMesh *m;
  int ncomp, recon;
  double cfl;
  size_t nfield; /* nblocks*ncomp*ncell */
  double *U;     /* base */
  double *U1;    /* stage "1" */
  double *dUdt;
  double *rec[6]; /* Ulx Urx Uly Ury Ulz Urz */
  double *flux[3];
  double *derived;
  double *Uc; /* coarse-buffer scratch of the container being exchanged (multilevel) */
  double dt, time, allowed_dt;
  int ncycle;
};

/* uniform_cartesian.hpp:94-97 */
static inline double xc(const Block *blk, int d, int idx) {
  return blk->cxmin[d] + (idx + 0.5) * blk->dx[d];
}

/* benchmarks/burgers/parthenon_app_inputs.cpp:35-78 (kx_fact etc. are unused there) */
void orc_burgers_ic(const OrcMesh *m, double *U, int ncomp) {
#pragma omp parallel for schedule(static)
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    for (int k = m->is[2]; k <= m->ie[2]; ++k)
      for (int j = m->is[1]; j <= m->ie[1]; ++j)
        for (int i = m->is[0]; i <= m->ie[0]; ++i) {
          const double x = xc(blk, 0, i), y = xc(blk, 1, j), z = xc(blk, 2, k);
          const double qx = (tanh(-20.0 * x) * cos(M_PI * x) + 1.0) * exp(-30.0 * y * y) *
                            exp(-30.0 * z * z);
          const double qy =
              (sin(M_PI * y) + 0.2) * exp(-30.0 * x * x) * exp(-30.0 * z * z);
          const double qz = (tanh(-20. * z) * cos(M_PI * z) + 0.5) * exp(-30.0 * x * x) *
                            exp(-30.0 * y * y);
          U[fidx(m, ncomp, b, 0, k, j, i)] = qx;
          U[fidx(m, ncomp, b, 1, k, j, i)] = qy;
          U[fidx(m, ncomp, b, 2, k, j, i)] = qz;
          for (int n = 3; n < ncomp; ++n) {
            double q = 1;
            if (fabs(x) < 0.025 && fabs(y) < 0.15 && fabs(z) < 0.025) q += 10.0;
            U[fidx(m, ncomp, b, n, k, j, i)] = q;
          }
        }
  }
}

OrcBurgers *orc_burgers_create(const OrcMesh *m, int num_scalars, int recon, double cfl) {
  OrcBurgers *s = (OrcBurgers *)calloc(1, sizeof(OrcBurgers));
  s->m = m;
  s->ncomp = 3 + num_scalars;
  s->recon = recon;
  s->cfl = cfl;
  size_t ncell = (size_t)m->n[0] * m->n[1] * m->n[2];
  s->nfield = (size_t)m->nblocks * s->ncomp * ncell;
  s->U = (double *)calloc(s->nfield, sizeof(double));
  s->U1 = (double *)calloc(s->nfield, sizeof(double));
  s->dUdt = (double *)calloc(s->nfield, sizeof(double));
  for (int r = 0; r < 6; ++r) s->rec[r] = (double *)calloc(s->nfield, sizeof(double));
  for (int d = 0; d < 3; ++d) s->flux[d] = (double *)calloc(s->nfield, sizeof(double));
  s->derived = (double *)calloc((size_t)m->nblocks * ncell, sizeof(double));
  if (m->multilevel)
    s->Uc = (double *)calloc((size_t)m->nblocks * s->ncomp * m->cn[0] * m->cn[1] * m->cn[2],
                             sizeof(double));
  s->dt = DBL_MAX;
  return s;
}
void orc_burgers_destroy(OrcBurgers *s) {
  if (!s) return;
  free(s->U);
  free(s->U1);
  free(s->dUdt);
  for (int r = 0; r < 6; ++r) free(s->rec[r]);
  for (int d = 0; d < 3; ++d) free(s->flux[d]);
  free(s->derived);
  free(s->Uc);
  free(s);
}
double *orc_burgers_U(OrcBurgers *s) { return s->U; }
double *orc_burgers_derived(OrcBurgers *s) { return s->derived; }
double *orc_burgers_flux(OrcBurgers *s, int dir) { return s->flux[dir]; }
double orc_burgers_dt(const OrcBurgers *s) { return s->dt; }
double orc_burgers_time(const OrcBurgers *s) { return s->time; }
int orc_burgers_cycle(const OrcBurgers *s) { return s->ncycle; }

/* CalculateFluxes burgers_package.cpp:202-404 (3-D, ndim-aware like the reference) */
void orc_burgers_calculate_fluxes(OrcBurgers *st, const double *U) {
  const OrcMesh *m = st->m;
  const int nc = st->ncomp, ndim = m->ndim;
  const int is = m->is[0], ie = m->ie[0], js = m->is[1], je = m->ie[1], ks = m->is[2],
            ke = m->ie[2];
  const int dk = ndim > 2 ? 1 : 0, dj = ndim > 1 ? 1 : 0;
  const size_t sj = (size_t)m->n[0], sk = (size_t)m->n[0] * m->n[1];
  double *Ulx = st->rec[0], *Urx = st->rec[1], *Uly = st->rec[2], *Ury = st->rec[3],
         *Ulz = st->rec[4], *Urz = st->rec[5];
  /* reconstruction :236-303 */
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < m->nblocks; ++b)
    for (int k = ks - dk; k <= ke + dk; ++k)
      for (int j = js - dj; j <= je + dj; ++j) {
        const int xrec = (k >= ks && k <= ke) && (j >= js && j <= je);
        const int yrec = (k >= ks && k <= ke) && (ndim > 1);
        const int zrec = (j >= js && j <= je) && (ndim > 2);
        for (int n = 0; n < nc; ++n) {
          const size_t row = fidx(m, nc, b, n, k, j, 0);
          const double *pq = U + row;
          if (xrec)
            for (int i = is - 1; i <= ie + 1; ++i) {
              if (st->recon == 0)
                orc_weno5z(pq[i - 2], pq[i - 1], pq[i], pq[i + 1], pq[i + 2],
                           &Ulx[row + i + 1], &Urx[row + i]);
              else
                orc_linear(pq[i - 1], pq[i], pq[i + 1], &Ulx[row + i + 1], &Urx[row + i]);
            }
          if (yrec)
            for (int i = is; i <= ie; ++i) {
              if (st->recon == 0)
                orc_weno5z(pq[i - 2 * sj], pq[i - sj], pq[i], pq[i + sj], pq[i + 2 * sj],
                           &Uly[row + sj + i], &Ury[row + i]);
              else
                orc_linear(pq[i - sj], pq[i], pq[i + sj], &Uly[row + sj + i], &Ury[row + i]);
            }
          if (zrec)
            for (int i = is; i <= ie; ++i) {
              if (st->recon == 0)
                orc_weno5z(pq[i - 2 * sk], pq[i - sk], pq[i], pq[i + sk], pq[i + 2 * sk],
                           &Ulz[row + sk + i], &Urz[row + i]);
              else
                orc_linear(pq[i - sk], pq[i], pq[i + sk], &Ulz[row + sk + i], &Urz[row + i]);
            }
        }
      }
  /* Riemann :307-401 */
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < m->nblocks; ++b)
    for (int k = ks; k <= ke + dk; ++k)
      for (int j = js; j <= je + dj; ++j) {
        const int xflux = (k <= ke && j <= je);
        const int yflux = (ndim > 1 && k <= ke);
        const int zflux = (ndim > 2 && j <= je);
        const size_t r0 = fidx(m, nc, b, 0, k, j, 0);
        const size_t cs = (size_t)m->n[0] * m->n[1] * m->n[2]; /* component stride */
        for (int dir = 0; dir < 3; ++dir) {
          if (dir == 0 && !xflux) continue;
          if (dir == 1 && !yflux) continue;
          if (dir == 2 && !zflux) continue;
          const double *L = st->rec[2 * dir], *R = st->rec[2 * dir + 1];
          double *F = st->flux[dir];
          const int iend = dir == 0 ? ie + 1 : ie;
          for (int i = is; i <= iend; ++i) {
            const size_t p = r0 + i;
            double sl, sr;
            orc_lr_to_flux(L[p], R[p], L[p + cs], R[p + cs], L[p + 2 * cs], R[p + 2 * cs],
                           L[p + dir * cs], R[p + dir * cs], &sl, &sr, &F[p], &F[p + cs],
                           &F[p + 2 * cs]);
            const double upl = L[p + dir * cs], upr = R[p + dir * cs];
            for (int n = 3; n < nc; ++n) {
              const double ql = L[p + n * cs], qr = R[p + n * cs];
              F[p + n * cs] = (sr * upl * ql - sl * upr * qr + sl * sr * (qr - ql)) /
                              (sr - sl + (sl * sr == 0.0));
            }
          }
        }
      }
}

/* CalculateDerived burgers_package.cpp:143-167 : pack {derived, U} => d = U component 3 */
static void calculate_derived(OrcBurgers *st, const double *U) {
  const OrcMesh *m = st->m;
  const int nc = st->ncomp;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < m->nblocks; ++b)
    for (int k = m->is[2]; k <= m->ie[2]; ++k)
      for (int j = m->is[1]; j <= m->ie[1]; ++j)
        for (int i = m->is[0]; i <= m->ie[0]; ++i) {
          const double u1 = U[fidx(m, nc, b, 0, k, j, i)];
          const double u2 = U[fidx(m, nc, b, 1, k, j, i)];
          const double u3 = U[fidx(m, nc, b, 2, k, j, i)];
          const double d = U[fidx(m, nc, b, 3, k, j, i)];
          st->derived[fidx(m, 1, b, 0, k, j, i)] = 0.5 * d * (u1 * u1 + u2 * u2 + u3 * u3);
        }
}

/* EstimateTimestepMesh burgers_package.cpp:170-200 */
static double estimate_timestep(const OrcBurgers *st, const double *U) {
  const OrcMesh *m = st->m;
  const int nc = st->ncomp, ndim = m->ndim;
  double min_dt = DBL_MAX;
#pragma omp parallel for collapse(2) schedule(static) reduction(min : min_dt)
  for (int b = 0; b < m->nblocks; ++b)
    for (int k = m->is[2]; k <= m->ie[2]; ++k)
      for (int j = m->is[1]; j <= m->ie[1]; ++j) {
        const Block *blk = &m->blocks[b];
        for (int i = m->is[0]; i <= m->ie[0]; ++i) {
          const double v =
              1.0 / ((fabs(U[fidx(m, nc, b, 0, k, j, i)])) / blk->dx[0] +
                     (ndim > 1) * (fabs(U[fidx(m, nc, b, 1, k, j, i)])) / blk->dx[1] +
                     (ndim > 2) * (fabs(U[fidx(m, nc, b, 2, k, j, i)])) / blk->dx[2]);
          min_dt = fmin(min_dt, v);
        }
      }
  return st->cfl * min_dt;
}

/* FluxDivergence update.cpp:63-86 with FluxDivHelper update.hpp:43-58 */
static void flux_divergence_generic(const OrcMesh *m, int nc, double *const flux[3],
                                    double *dUdt) {
  const int ndim = m->ndim;
  const size_t sj = (size_t)m->n[0], sk = (size_t)m->n[0] * m->n[1];
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < m->nblocks; ++b)
    for (int l = 0; l < nc; ++l) {
      const Block *blk = &m->blocks[b];
      const double a1 = blk->dx[1] * blk->dx[2], a2 = blk->dx[0] * blk->dx[2],
                   a3 = blk->dx[0] * blk->dx[1];
      const double vol = blk->dx[0] * blk->dx[1] * blk->dx[2];
      for (int k = m->is[2]; k <= m->ie[2]; ++k)
        for (int j = m->is[1]; j <= m->ie[1]; ++j)
          for (int i = m->is[0]; i <= m->ie[0]; ++i) {
            const size_t p = fidx(m, nc, b, l, k, j, i);
            double du = (a1 * flux[0][p + 1] - a1 * flux[0][p]);
            if (ndim >= 2) du += (a2 * flux[1][p + sj] - a2 * flux[1][p]);
            if (ndim == 3) du += (a3 * flux[2][p + sk] - a3 * flux[2][p]);
            dUdt[p] = -du / vol;
          }
    }
}
static void flux_divergence(OrcBurgers *st) {
  flux_divergence_generic(st->m, st->ncomp, st->flux, st->dUdt);
}

/* WeightedSumData update.hpp:71-91: z = w1*x + w2*y over the ENTIRE extents */
static void weighted_sum(size_t n, const double *x, const double *y, double w1, double w2,
                         double *z) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) z[i] = w1 * x[i] + w2 * y[i];
}

/* one stage of BurgersDriver::MakeTaskCollection burgers_driver.cpp:53-148 with rk2
 * coefficients (low_storage_integrator.cpp:67-86) and stage names
 * (staged_integrator.cpp:23-30): stage 1: mc0 = base, mc1 = "1"; stage 2: mc0 = "1",
 * mc1 = base. */
void orc_burgers_stage(OrcBurgers *st, int stage) {
  const double beta = stage == 1 ? 1.0 : 0.5;
  double *mc0 = stage == 1 ? st->U : st->U1;
  double *mc1 = stage == 1 ? st->U1 : st->U;
  double *base = st->U;
  orc_burgers_calculate_fluxes(st, mc0);
  /* LoadAndSendFluxCorrections / SetFluxCorrections burgers_driver.cpp:94-96 */
  if (st->m->multilevel) orc_flux_correct(st->m, st->flux, st->ncomp);
  flux_divergence(st);
  weighted_sum(st->nfield, mc0, base, beta, 1.0 - beta, mc0);        /* AverageIndependentData */
  weighted_sum(st->nfield, mc0, st->dUdt, 1.0, beta * st->dt, mc1); /* UpdateIndependentData */
  /* Send/Receive/SetBounds<local|nonlocal>  burgers_driver.cpp:106-119.  The reference's
   * stage list has NO ProlongateBounds task: on a multilevel mesh the coarse buffers are
   * filled and restricted, but fine ghosts facing a coarser block keep the values they had
   * (prolongated once by Mesh::Initialize, then carried through the full-extent
   * WeightedSumData passes).  Reproduced as is. */
  orc_exchange(st->m, mc1, st->Uc, st->ncomp, 0);
  orc_apply_bcs(st->m, mc1, st->ncomp); /* ApplyBoundaryConditions burgers_driver.cpp:137 */
  calculate_derived(st, mc1);
  if (stage == 2) st->allowed_dt = estimate_timestep(st, mc1);
}

/* EvolutionDriver::SetGlobalTimeStep driver.cpp:210-270 with default limits */
static void set_global_timestep(OrcBurgers *st) {
  if (st->dt < 0.1 * DBL_MAX) st->dt *= 2.0;
  st->dt = fmin(st->dt, st->allowed_dt);
  st->allowed_dt = DBL_MAX;
}

void orc_burgers_init(OrcBurgers *st) {
  orc_burgers_ic(st->m, st->U, st->ncomp);
  /* Mesh::Initialize -> CommunicateBoundaries (mesh.cpp:640-706) does prolongate */
  orc_exchange(st->m, st->U, st->Uc, st->ncomp, st->m->multilevel);
  orc_apply_bcs(st->m, st->U, st->ncomp); /* mesh.cpp:705 */
  calculate_derived(st, st->U);
  st->allowed_dt = estimate_timestep(st, st->U); /* InitializeBlockTimeSteps driver.cpp:194 */
  st->dt = DBL_MAX;
  set_global_timestep(st);
  st->time = 0;
  st->ncycle = 0;
}

void orc_burgers_step(OrcBurgers *st) {
  orc_burgers_stage(st, 1);
  orc_burgers_stage(st, 2);
  st->ncycle++;
  st->time += st->dt;
  set_global_timestep(st);
}

/* MassHistory burgers_package.cpp:406-439, octant order :94-107 */
void orc_burgers_history(const OrcBurgers *st, double out[8]) {
  const OrcMesh *m = st->m;
  const int nc = st->ncomp;
  double mesh_vol = 1;
  double mid[3];
  for (int d = 0; d < 3; ++d) {
    mesh_vol *= (m->xmax[d] - m->xmin[d]);
    mid[d] = 0.5 * (m->xmin[d] + m->xmax[d]);
  }
  int oct = 0;
  for (int s1 = 0; s1 < 2; ++s1)
    for (int s2 = 0; s2 < 2; ++s2)
      for (int s3 = 0; s3 < 2; ++s3) {
        const double lo[3] = {s1 ? mid[0] : m->xmin[0], s2 ? mid[1] : m->xmin[1],
                              s3 ? mid[2] : m->xmin[2]};
        const double hi[3] = {s1 ? m->xmax[0] : mid[0], s2 ? m->xmax[1] : mid[1],
                              s3 ? m->xmax[2] : mid[2]};
        double result = 0.0;
#pragma omp parallel for collapse(2) schedule(static) reduction(+ : result)
        for (int b = 0; b < m->nblocks; ++b)
          for (int v = 0; v < nc; ++v) {
            const Block *blk = &m->blocks[b];
            const double vol = blk->dx[0] * blk->dx[1] * blk->dx[2];
            const double weight = vol / (mesh_vol + 1e-20);
            for (int k = m->is[2]; k <= m->ie[2]; ++k)
              for (int j = m->is[1]; j <= m->ie[1]; ++j)
                for (int i = m->is[0]; i <= m->ie[0]; ++i) {
                  const double x1 = xc(blk, 0, i), x2 = xc(blk, 1, j), x3 = xc(blk, 2, k);
                  const double mask = (lo[0] <= x1) && (x1 <= hi[0]) && (lo[1] <= x2) &&
                                      (x2 <= hi[1]) && (lo[2] <= x3) && (x3 <= hi[2]);
                  const double q = st->U[fidx(m, nc, b, v, k, j, i)];
                  result += mask * q * q * weight;
                }
          }
        out[oct++] = result;
      }
}

/* ------------------------------------------------------------------------------------ */
/* example/advection (constant velocity, v_const = true): advection_package.cpp,
 * advection_driver.cpp, parthenon_app_inputs.cpp.  One field "advected" of vec_size
 * components; fill_derived = false. */
struct OrcAdvection {
  const OrcMesh *m;
  int ncomp, profile;
  double amp, v[3], cfl;
  size_t nfield;
  double *U, *U1, *dUdt, *flux[3], *Uc;
  double dt, time, allowed_dt;
  int ncycle;
};

OrcAdvection *orc_advection_create(const OrcMesh *m, int vec_size, int profile, double amp,
                                   const double v[3], double cfl) {
  OrcAdvection *s = (OrcAdvection *)calloc(1, sizeof(OrcAdvection));
  s->m = m;
  s->ncomp = vec_size;
  s->profile = profile;
  s->amp = amp;
  for (int d = 0; d < 3; ++d) s->v[d] = v[d];
  s->cfl = cfl;
  size_t ncell = (size_t)m->n[0] * m->n[1] * m->n[2];
  s->nfield = (size_t)m->nblocks * s->ncomp * ncell;
  s->U = (double *)calloc(s->nfield, sizeof(double));
  s->U1 = (double *)calloc(s->nfield, sizeof(double));
  s->dUdt = (double *)calloc(s->nfield, sizeof(double));
  for (int d = 0; d < 3; ++d) s->flux[d] = (double *)calloc(s->nfield, sizeof(double));
  s->Uc = (double *)calloc((size_t)m->nblocks * s->ncomp * m->cn[0] * m->cn[1] * m->cn[2],
                           sizeof(double));
  s->dt = DBL_MAX;
  return s;
}
void orc_advection_destroy(OrcAdvection *s) {
  if (!s) return;
  free(s->U);
  free(s->U1);
  free(s->dUdt);
  for (int d = 0; d < 3; ++d) free(s->flux[d]);
  free(s->Uc);
  free(s);
}
double *orc_advection_U(OrcAdvection *s) { return s->U; }
double *orc_advection_flux(OrcAdvection *s, int dir) { return s->flux[dir]; }
double orc_advection_dt(const OrcAdvection *s) { return s->dt; }
double orc_advection_time(const OrcAdvection *s) { return s->time; }

/* ProblemGenerator parthenon_app_inputs.cpp:40-100: profiles smooth_gaussian (1) and
 * hard_sphere (2), interior cells */
static void advection_ic(OrcAdvection *st) {
  const OrcMesh *m = st->m;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    for (int n = 0; n < st->ncomp; ++n)
      for (int k = m->is[2]; k <= m->ie[2]; ++k)
        for (int j = m->is[1]; j <= m->ie[1]; ++j)
          for (int i = m->is[0]; i <= m->ie[0]; ++i) {
            const double x = xc(blk, 0, i), y = xc(blk, 1, j), z = xc(blk, 2, k);
            const double rsq = x * x + y * y + z * z;
            double q;
            if (st->profile == 1)
              q = 1. + st->amp * exp(-100.0 * rsq);
            else
              q = (rsq < 0.15 * 0.15 ? 1.0 : 0.0);
            st->U[fidx(m, st->ncomp, b, n, k, j, i)] = q;
          }
  }
}

/* CalculateFluxes advection_package.cpp:540-646 with DonorCellX1/2/3
 * (reconstruct/dc_inline.hpp:31-71): upwind cell value times the constant velocity */
void orc_advection_calculate_fluxes(OrcAdvection *st, const double *U) {
  const OrcMesh *m = st->m;
  const int nc = st->ncomp, ndim = m->ndim;
  const size_t str[3] = {1, (size_t)m->n[0], (size_t)m->n[0] * m->n[1]};
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < m->nblocks; ++b)
    for (int n = 0; n < nc; ++n)
      for (int dir = 0; dir < ndim; ++dir) {
        const int ke = m->ie[2] + (dir == 2), je = m->ie[1] + (dir == 1),
                  ie = m->ie[0] + (dir == 0);
        const double vel = st->v[dir];
        for (int k = m->is[2]; k <= ke; ++k)
          for (int j = m->is[1]; j <= je; ++j)
            for (int i = m->is[0]; i <= ie; ++i) {
              const size_t p = fidx(m, nc, b, n, k, j, i);
              st->flux[dir][p] = vel > 0.0 ? U[p - str[dir]] * vel : U[p] * vel;
            }
      }
}

/* EstimateTimestepBlock advection_package.cpp:505-536, min over blocks
 * (Update::EstimateTimestep update.hpp:268-278, driver.cpp:210-270) */
static double advection_estimate_timestep(const OrcAdvection *st) {
  const OrcMesh *m = st->m;
  double min_dt = DBL_MAX;
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    double bdt = DBL_MAX;
    for (int d = 0; d < 3; ++d)
      if (st->v[d] != 0.0) bdt = fmin(bdt, blk->dx[d] / fabs(st->v[d]));
    min_dt = fmin(min_dt, st->cfl * bdt);
  }
  return min_dt;
}

/* one stage of AdvectionDriver::MakeTaskCollection advection_driver.cpp:56-163 */
void orc_advection_stage(OrcAdvection *st, int stage) {
  const double beta = stage == 1 ? 1.0 : 0.5;
  double *mc0 = stage == 1 ? st->U : st->U1;
  double *mc1 = stage == 1 ? st->U1 : st->U;
  orc_advection_calculate_fluxes(st, mc0);
  if (st->m->multilevel) orc_flux_correct(st->m, st->flux, st->ncomp); /* :123 */
  flux_divergence_generic(st->m, st->ncomp, st->flux, st->dUdt);
  weighted_sum(st->nfield, mc0, st->U, beta, 1.0 - beta, mc0);
  weighted_sum(st->nfield, mc0, st->dUdt, 1.0, beta * st->dt, mc1);
  /* AddBoundaryExchangeTasks(update, tl, mc1, multilevel) :136: prolongates */
  orc_exchange(st->m, mc1, st->Uc, st->ncomp, st->m->multilevel);
  orc_apply_bcs(st->m, mc1, st->ncomp); /* fbound of AddBoundaryExchangeTasks, :449 */
  if (stage == 2) st->allowed_dt = advection_estimate_timestep(st);
}

void orc_advection_init(OrcAdvection *st) {
  advection_ic(st);
  orc_exchange(st->m, st->U, st->Uc, st->ncomp, st->m->multilevel);
  orc_apply_bcs(st->m, st->U, st->ncomp); /* mesh.cpp:705 */
  st->allowed_dt = advection_estimate_timestep(st);
  st->dt = DBL_MAX;
  if (st->dt < 0.1 * DBL_MAX) st->dt *= 2.0;
  st->dt = fmin(st->dt, st->allowed_dt);
  st->allowed_dt = DBL_MAX;
  st->time = 0;
  st->ncycle = 0;
}

void orc_advection_step(OrcAdvection *st) {
  orc_advection_stage(st, 1);
  orc_advection_stage(st, 2);
  st->ncycle++;
  st->time += st->dt;
  if (st->dt < 0.1 * DBL_MAX) st->dt *= 2.0; /* SetGlobalTimeStep driver.cpp:210-270 */
  st->dt = fmin(st->dt, st->allowed_dt);
  st->allowed_dt = DBL_MAX;
}

/* ------------------------------------------------------------------------------------ */
/* adaptive mesh refinement for example/advection (refinement = adaptive):
 *   tagging      Refinement::Tag -> advection_package::CheckRefinement
 *                (advection_package.cpp:239-273) -> MeshRefinement::SetRefinement
 *                (mesh/mesh_refinement.cpp:81-118)
 *   tree update  Mesh::UpdateMeshBlockTree (mesh-amr_loadbalance.cpp:496-625) with
 *                Tree::Refine / Tree::Derefine (mesh/forest/tree.cpp:93-143, 229-275:
 *                proper nesting enforced)
 *   data         Mesh::RedistributeAndRefineMeshBlocks (:663-1010): kept blocks keep their
 *                arrays, derefined blocks are restricted (GetInteriorRestrict) and copied into
 *                the new parent (TryRecvFineToCoarse :196-251), refined blocks get the
 *                parent's data in their coarse buffer (TryRecvCoarseToFine :87-149) and are
 *                prolongated over interior + ghosts (GetInteriorProlongate), then
 *                CommunicateBoundaries
 *   driver       EvolutionDriver::Execute (driver.cpp:99-150), Mesh::Initialize's refinement
 *                loop (mesh.cpp:745-860)
 * Single process (no load balancing across ranks). */
typedef struct {
  Loc *leaf;
  int n, cap;
} LeafSet;

static int loc_eq(const Loc *a, const Loc *b) {
  return a->level == b->level && a->lx[0] == b->lx[0] && a->lx[1] == b->lx[1] &&
         a->lx[2] == b->lx[2];
}
static int ls_find(const LeafSet *t, const Loc *l) {
  for (int i = 0; i < t->n; ++i)
    if (loc_eq(&t->leaf[i], l)) return i;
  return -1;
}
static int ls_internal(const LeafSet *t, const Loc *l) { /* some leaf lies strictly below l */
  for (int i = 0; i < t->n; ++i) {
    const Loc *q = &t->leaf[i];
    if (q->level <= l->level) continue;
    const int sh = q->level - l->level;
    if ((q->lx[0] >> sh) == l->lx[0] && (q->lx[1] >> sh) == l->lx[1] &&
        (q->lx[2] >> sh) == l->lx[2])
      return 1;
  }
  return 0;
}
static void ls_add(LeafSet *t, const Loc *l) {
  if (t->n == t->cap) {
    t->cap = t->cap ? 2 * t->cap : 64;
    t->leaf = (Loc *)realloc(t->leaf, sizeof(Loc) * (size_t)t->cap);
  }
  t->leaf[t->n++] = *l;
}
static void ls_remove(LeafSet *t, int i) { t->leaf[i] = t->leaf[--t->n]; }
static int ndaughters(const OrcMesh *m) { return 1 << m->ndim; }
static Loc daughter(const OrcMesh *m, const Loc *p, int q) { /* ox1 fastest */
  Loc d;
  d.level = p->level + 1;
  for (int k = 0; k < 3; ++k) d.lx[k] = k < m->ndim ? (p->lx[k] << 1) + ((q >> k) & 1) : 0;
  return d;
}
static Loc parent_of(const OrcMesh *m, const Loc *l) {
  Loc p;
  p.level = l->level - 1;
  for (int k = 0; k < 3; ++k) p.lx[k] = k < m->ndim ? (l->lx[k] >> 1) : 0;
  return p;
}

/* Tree::Refine tree.cpp:93-143 */
static int tree_refine(const OrcMesh *m, LeafSet *t, const Loc *ref) {
  const int i0 = ls_find(t, ref);
  if (i0 < 0) return 0; /* can't refine a block that doesn't exist */
  ls_remove(t, i0);
  const int nd = ndaughters(m);
  for (int q = 0; q < nd; ++q) {
    Loc d = daughter(m, ref, q);
    ls_add(t, &d);
  }
  int nadded = nd - 1;
  if (ref->level <= m->root_level) return nadded; /* no leaves above the root grid */
  /* proper nesting: the same-level neighbours of the PARENT on the side of this block */
  const Loc par = parent_of(m, ref);
  const int ox[3] = {(int)(ref->lx[0] - (par.lx[0] << 1)), (int)(ref->lx[1] - (par.lx[1] << 1)),
                     (int)(ref->lx[2] - (par.lx[2] << 1))};
  for (int k = 0; k < (m->ndim > 2 ? 2 : 1); ++k)
    for (int j = 0; j < (m->ndim > 1 ? 2 : 1); ++j)
      for (int i = 0; i < 2; ++i) {
        Loc neigh = par, w;
        neigh.lx[0] += i + ox[0] - 1;
        neigh.lx[1] += j + ox[1] - (m->ndim > 1);
        neigh.lx[2] += k + ox[2] - (m->ndim > 2);
        if (!wrap_loc(m, &neigh, &w)) continue; /* no tree on that side */
        nadded += tree_refine(m, t, &w);
      }
  return nadded;
}

/* Tree::Derefine tree.cpp:229-275 */
static int tree_derefine(const OrcMesh *m, LeafSet *t, const Loc *ref) {
  const int nd = ndaughters(m);
  for (int q = 0; q < nd; ++q) {
    const Loc d = daughter(m, ref, q);
    if (ls_find(t, &d) < 0) return 0;
    for (int k = (m->ndim > 2 ? -1 : 0); k <= (m->ndim > 2 ? 1 : 0); ++k)
      for (int j = (m->ndim > 1 ? -1 : 0); j <= (m->ndim > 1 ? 1 : 0); ++j)
        for (int i = -1; i <= 1; ++i) {
          Loc neigh = d, w;
          neigh.lx[0] += i;
          neigh.lx[1] += j;
          neigh.lx[2] += k;
          if (!wrap_loc(m, &neigh, &w)) continue;
          if (ls_internal(t, &w)) return 0; /* would abut a block two levels finer */
        }
  }
  for (int q = 0; q < nd; ++q) {
    const Loc d = daughter(m, ref, q);
    ls_remove(t, ls_find(t, &d));
  }
  ls_add(t, ref);
  return nd - 1;
}

struct OrcAmr {
  /* mesh definition */
  int ndim, nx[3], ng, nrb[3], bc[6];
  double xmin[3], xmax[3];
  int max_level, deref_threshold;
  double refine_tol, derefine_tol;
  OrcMesh *m;
  OrcAdvection *adv;  /* state on the current mesh */
  int *deref_count;   /* per block: MeshRefinement::deref_count_ */
  int *refine_flag;   /* per block: MeshRefinement::refine_flag_ */
  int *last_tag;      /* per block: the raw AmrTag of the last tagging pass (for tests) */
  int profile;
  double amp, v[3], cfl;
  int ncomp;
  /* app = 1: benchmarks/burgers with the generic <parthenon/refinementN> criterion
   * derivative_order_1 (amr_criteria.cpp:92-101) on one component of U */
  int app;
  OrcBurgers *bur;
  int num_scalars, recon;
  int crit_comp, crit_max_level; /* vector_i; max_level + root_level (parthenon_manager.cpp:216-220) */
  /* app = 2: the non-cell-centred field application of tests/golden/refgen/teamr_dump_main.cpp:
   * a face (2 components), an edge and a node field that never evolve; the mesh follows a
   * geometric criterion that moves with te_cycle */
  double *te_U[3], *te_Uc[3];
  int te_cycle;
  int te_toth_roe; /* the face field registered ProlongateInternalTothAndRoe */
  /* app = 3: example/sparse_advection with refinement = adaptive */
  struct OrcSparse *sp;
  double sp_speed, sp_alloc_thr, sp_dealloc_thr;
  int sp_dealloc_count;
};
static const int kTeNcomp[3] = {2, 1, 1};
static void te_alloc(const OrcMesh *m, int f, double **U, double **Uc);
static void te_ic(struct OrcAmr *a);
static void te_remesh(struct OrcAmr *a, OrcMesh *nm);
static void sparse_remesh(struct OrcAmr *a, OrcMesh *nm);
static void sparse_minmax(const struct OrcSparse *st, int b, double *mn, double *mx);
static void amr_tag(struct OrcAmr *a, const double *U);
static int amr_remesh(struct OrcAmr *a);

static OrcMesh *amr_make_mesh(const struct OrcAmr *a, const LeafSet *t) {
  int *lv = (int *)malloc(sizeof(int) * 4 * (size_t)t->n);
  for (int i = 0; i < t->n; ++i) {
    lv[4 * i] = t->leaf[i].level;
    for (int d = 0; d < 3; ++d) lv[4 * i + 1 + d] = (int)t->leaf[i].lx[d];
  }
  orc_mesh_set_next_bcs(a->bc);
  OrcMesh *m = orc_mesh_create(a->ndim, a->nx, a->ng, a->nrb, a->xmin, a->xmax, t->n, lv);
  orc_mesh_set_next_bcs(NULL);
  free(lv);
  if (m) m->multilevel = 1; /* refinement != none (mesh.cpp:118-140) */
  return m;
}

struct OrcAmr *orc_amr_create(int ndim, const int nx[3], int ng, const int nrb[3],
                              const double xmin[3], const double xmax[3], int numlevel,
                              int derefine_count, double refine_tol, double derefine_tol,
                              int vec_size, int profile, double amp, const double v[3],
                              double cfl) {
  struct OrcAmr *a = (struct OrcAmr *)calloc(1, sizeof(struct OrcAmr));
  a->ndim = ndim;
  a->ng = ng;
  for (int d = 0; d < 3; ++d) {
    a->nx[d] = nx[d];
    a->nrb[d] = nrb[d];
    a->xmin[d] = xmin[d];
    a->xmax[d] = xmax[d];
    a->v[d] = v[d];
  }
  a->deref_threshold = derefine_count;
  a->refine_tol = refine_tol;
  a->derefine_tol = derefine_tol;
  a->profile = profile;
  a->amp = amp;
  a->cfl = cfl;
  a->ncomp = vec_size;
  /* root grid */
  LeafSet t = {0, 0, 0};
  int rl = 0;
  while ((1 << rl) < nrb[0]) ++rl;
  for (int k = 0; k < (ndim > 2 ? nrb[2] : 1); ++k)
    for (int j = 0; j < (ndim > 1 ? nrb[1] : 1); ++j)
      for (int i = 0; i < nrb[0]; ++i) {
        Loc l = {rl, {i, j, k}};
        ls_add(&t, &l);
      }
  a->max_level = numlevel + rl - 1; /* mesh.cpp:125 */
  a->m = amr_make_mesh(a, &t);
  free(t.leaf);
  a->adv = orc_advection_create(a->m, vec_size, profile, amp, v, cfl);
  a->deref_count = (int *)calloc((size_t)a->m->nblocks, sizeof(int));
  a->refine_flag = (int *)calloc((size_t)a->m->nblocks, sizeof(int));
  a->last_tag = (int *)calloc((size_t)a->m->nblocks, sizeof(int));
  return a;
}
/* the non-cell-centred field application (app = 2) */
struct OrcAmr *orc_amr_create_te(int ndim, const int nx[3], int ng, const int nrb[3],
                                 const double xmin[3], const double xmax[3], int numlevel,
                                 int derefine_count) {
  const double v0[3] = {0, 0, 0};
  struct OrcAmr *a = orc_amr_create(ndim, nx, ng, nrb, xmin, xmax, numlevel, derefine_count, 0.0,
                                    0.0, 1, 2, 0.0, v0, 0.0);
  orc_advection_destroy(a->adv);
  a->adv = NULL;
  a->app = 2;
  for (int f = 0; f < 3; ++f) te_alloc(a->m, f, &a->te_U[f], &a->te_Uc[f]);
  return a;
}
/* Mesh::Initialize: problem generator, exchange, tag, remesh until the mesh stops changing */
static void amr_init_te(struct OrcAmr *a) {
  int done;
  a->te_cycle = 0;
  do {
    te_ic(a);
    for (int f = 0; f < 3; ++f)
      exchange_te_impl(a->m, a->te_U[f], a->te_Uc[f], kTeNcomp[f], f + 1,
                       f == 0 && a->te_toth_roe, 0);
    amr_tag(a, NULL);
    done = !amr_remesh(a);
  } while (!done);
}
/* one "cycle" of the fixture generator: tag with the criterion of `cycle`, remesh */
int orc_amr_te_cycle(struct OrcAmr *a, int cycle) {
  a->te_cycle = cycle;
  amr_tag(a, NULL);
  return amr_remesh(a);
}
const double *orc_amr_te_field(const struct OrcAmr *a, int f) { return a->te_U[f]; }
void orc_amr_te_set_toth_roe(struct OrcAmr *a, int on) { a->te_toth_roe = on; }

/* benchmarks/burgers, refinement = adaptive: criterion = derivative_order_1 on U(vector_i) */
struct OrcAmr *orc_amr_create_burgers(int ndim, const int nx[3], int ng, const int nrb[3],
                                      const double xmin[3], const double xmax[3], int numlevel,
                                      int derefine_count, double refine_tol, double derefine_tol,
                                      int vector_i, int num_scalars, int recon, double cfl) {
  const double v0[3] = {0, 0, 0};
  struct OrcAmr *a = orc_amr_create(ndim, nx, ng, nrb, xmin, xmax, numlevel, derefine_count,
                                    refine_tol, derefine_tol, 3 + num_scalars, 2, 0.0, v0, cfl);
  orc_advection_destroy(a->adv);
  a->adv = NULL;
  a->app = 1;
  a->num_scalars = num_scalars;
  a->recon = recon;
  a->crit_comp = vector_i;
  a->crit_max_level = numlevel + a->m->root_level;
  a->bur = orc_burgers_create(a->m, num_scalars, recon, cfl);
  if (!a->bur->Uc)
    a->bur->Uc = (double *)calloc((size_t)a->m->nblocks * a->ncomp * a->m->cn[0] * a->m->cn[1] * a->m->cn[2],
                                  sizeof(double));
  return a;
}

void orc_amr_destroy(struct OrcAmr *a) {
  if (!a) return;
  orc_sparse_destroy(a->sp);
  for (int f = 0; f < 3; ++f) {
    free(a->te_U[f]);
    free(a->te_Uc[f]);
  }
  orc_burgers_destroy(a->bur);
  orc_advection_destroy(a->adv);
  orc_mesh_destroy(a->m);
  free(a->deref_count);
  free(a->refine_flag);
  free(a->last_tag);
  free(a);
}
const OrcMesh *orc_amr_mesh(const struct OrcAmr *a) { return a->m; }
void orc_amr_deref_counts(const struct OrcAmr *a, int *out) {
  for (int b = 0; b < a->m->nblocks; ++b) out[b] = a->deref_count[b];
}
void orc_amr_tags(const struct OrcAmr *a, int *out) {
  for (int b = 0; b < a->m->nblocks; ++b) out[b] = a->last_tag[b];
}
double *orc_amr_U(struct OrcAmr *a) { return a->app ? a->bur->U : a->adv->U; }
double orc_amr_dt(const struct OrcAmr *a) { return a->app ? a->bur->dt : a->adv->dt; }
double orc_amr_time(const struct OrcAmr *a) { return a->app ? a->bur->time : a->adv->time; }

/* CheckRefinement + SetRefinement for every block, on container U */
static void amr_tag(struct OrcAmr *a, const double *U) {
  const OrcMesh *m = a->m;
  const size_t per_block = (size_t)a->ncomp * m->n[0] * m->n[1] * m->n[2];
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    double mn = DBL_MAX, mx = -DBL_MAX; /* Kokkos::MinMax identity */
    const double *p = U + (size_t)b * per_block;
    for (size_t q = 0; U != NULL && q < per_block; ++q) {
      mn = p[q] < mn ? p[q] : mn;
      mx = p[q] > mx ? p[q] : mx;
    }
    if (a->app == 3) sparse_minmax(a->sp, b, &mn, &mx); /* sparse_advection_package.cpp:110-143 */
    int aret = 0; /* AmrTag: derefine -1, same 0, refine 1 */
    if (a->app == 2) {
      /* TagByPosition of the fixture generator */
      const double xc = 0.5 * (blk->xmin[0] + blk->xmax[0]), yc = 0.5 * (blk->xmin[1] + blk->xmax[1]);
      const double zc = m->ndim > 2 ? 0.5 * (blk->xmin[2] + blk->xmax[2]) : 0.0;
      const double px = -0.25 + 0.125 * a->te_cycle, py = -0.125 + 0.0625 * a->te_cycle,
                   pz = m->ndim > 2 ? 0.125 : 0.0;
      const double r2 = (xc - px) * (xc - px) + (yc - py) * (yc - py) + (zc - pz) * (zc - pz);
      aret = r2 < 0.2 * 0.2 ? 1 : -1;
    } else if (a->app == 1) {
      /* Refinement::FirstDerivative refinement_package.cpp:92-122 over the interior of one
       * component; CheckAllRefinement :55-90: a "refine" at or above the criterion's max level
       * becomes "same" (:78-81); packages without a CheckRefinementBlock vote "derefine" */
      const size_t str[3] = {1, (size_t)m->n[0], (size_t)m->n[0] * m->n[1]};
      double maxd = 0.0;
      for (int k = m->is[2]; k <= m->ie[2]; ++k)
        for (int j = m->is[1]; j <= m->ie[1]; ++j)
          for (int i = m->is[0]; i <= m->ie[0]; ++i) {
            const double *q = U + fidx(m, a->ncomp, b, a->crit_comp, k, j, i);
            const double scale = fabs(q[0]);
            for (int d = 0; d < m->ndim; ++d) {
              const double dd = 0.5 * fabs(q[str[d]] - q[-(long)str[d]]) / (scale + 1.0e-20);
              maxd = dd > maxd ? dd : maxd;
            }
          }
      aret = maxd > a->refine_tol ? 1 : (maxd < a->derefine_tol ? -1 : 0);
      if (aret == 1 && blk->loc.level >= a->crit_max_level) aret = 0;
    } else if (mx > a->refine_tol && mn < a->derefine_tol)
      aret = 1;
    else if (mx < a->derefine_tol)
      aret = -1;
    int *flag = &a->refine_flag[b], *cnt = &a->deref_count[b];
    a->last_tag[b] = aret;
    if (aret == 0) *flag = 0;
    if (aret >= 0) *cnt = 0;
    if (aret > 0) {
      *flag = blk->loc.level == a->max_level ? 0 : 1;
    } else if (aret < 0) {
      if (blk->loc.level == m->root_level) {
        *flag = 0;
        *cnt = 0;
      } else {
        (*cnt)++;
        int ec = 0;
        for (int n = 0; n < blk->nnb; ++n)
          if (blk->nb[n].loc.level > blk->loc.level) ec++;
        if (ec > 0)
          *flag = 0;
        else
          *flag = *cnt >= a->deref_threshold ? -1 : 0;
      }
    }
  }
}


/* ---- remesh of face / edge / node fields (mesh-amr_loadbalance.cpp:663-1010) ---- */
static void te_field_init(TeField *F, const OrcMesh *m, double *U, double *Uc, int ncomp, int kind) {
  F->m = m;
  F->U = U;
  F->Uc = Uc;
  F->ncomp = ncomp;
  F->nel = orc_te_num_elements(kind);
  orc_te_extents(m, kind, F->pn);
  for (int d = 0; d < 3; ++d) F->cpn[d] = m->cn[d] + (kind != ORC_TE_CELL && m->cn[d] > 1 ? 1 : 0);
  F->blk_sz = (size_t)F->nel * ncomp * F->pn[2] * F->pn[1] * F->pn[0];
  F->cblk_sz = (size_t)F->nel * ncomp * F->cpn[2] * F->cpn[1] * F->cpn[0];
}
static void te_alloc(const OrcMesh *m, int f, double **U, double **Uc) {
  TeField F;
  te_field_init(&F, m, NULL, NULL, kTeNcomp[f], f + 1);
  *U = (double *)calloc((size_t)m->nblocks * F.blk_sz, sizeof(double));
  *Uc = (double *)calloc((size_t)m->nblocks * F.cblk_sz, sizeof(double));
}
/* the fixture generator's problem generator: a code in every entry of every array */
static void te_ic(struct OrcAmr *a) {
  for (int f = 0; f < 3; ++f) {
    TeField F;
    te_field_init(&F, a->m, a->te_U[f], a->te_Uc[f], kTeNcomp[f], f + 1);
    const size_t per = (size_t)F.pn[2] * F.pn[1] * F.pn[0];
    for (int b = 0; b < a->m->nblocks; ++b)
      for (int el = 0; el < F.nel; ++el)
        for (int c = 0; c < F.ncomp; ++c)
          for (size_t q = 0; q < per; ++q)
            F.U[(size_t)b * F.blk_sz + ((size_t)el * F.ncomp + c) * per + q] =
                (b + 1) * 1.0e6 + el * 1.0e5 + c * 5.0e4 + (double)q;
  }
}

static void te_remesh(struct OrcAmr *a, OrcMesh *nm) {
  const OrcMesh *m = a->m;
  const int nleaf = ndaughters(m);
  int *ncount = (int *)calloc((size_t)nm->nblocks, sizeof(int));
  int *nflag = (int *)calloc((size_t)nm->nblocks, sizeof(int));
  int *newly = (int *)calloc((size_t)nm->nblocks, sizeof(int));
  double *nU[3], *nUc[3];
  for (int f = 0; f < 3; ++f) {
    const int kind = f + 1;
    te_alloc(nm, f, &nU[f], &nUc[f]);
    TeField O, N;
    te_field_init(&O, m, a->te_U[f], a->te_Uc[f], kTeNcomp[f], kind);
    te_field_init(&N, nm, nU[f], nUc[f], kTeNcomp[f], kind);
    const int nc = O.ncomp;
    for (int nb = 0; nb < nm->nblocks; ++nb) {
      const Loc *nl = &nm->blocks[nb].loc;
      int ob = find_leaf(m, nl);
      if (ob >= 0) { /* kept: the block object with all its data and counters */
        memcpy(N.U + (size_t)nb * N.blk_sz, O.U + (size_t)ob * O.blk_sz, O.blk_sz * sizeof(double));
        memcpy(N.Uc + (size_t)nb * N.cblk_sz, O.Uc + (size_t)ob * O.cblk_sz,
               O.cblk_sz * sizeof(double));
        ncount[nb] = a->deref_count[ob];
        nflag[nb] = a->refine_flag[ob];
        continue;
      }
      if (nl->level > 0) {
        const Loc par = parent_of(m, nl);
        ob = find_leaf(m, &par);
        if (ob >= 0) { /* refined */
          newly[nb] = 1;
          for (int el = 0; el < O.nel; ++el) {
            int top[3], sh[3], hi[3];
            te_top_offset(kind, el, top);
            for (int d = 0; d < 3; ++d) {
              /* TryRecvCoarseToFine :118-141 */
              const int nint = m->n[d] == 1 ? 1 : m->ie[d] - m->is[d] + 1 + top[d];
              sh[d] = (d < m->ndim && (nl->lx[d] & 1L)) ? nint / 2 : 0;
              hi[d] = m->cn[d] == 1 ? 0 : m->cn[d] - 1 + top[d];
            }
            for (int c = 0; c < nc; ++c)
              for (int k = 0; k <= hi[2]; ++k)
                for (int j = 0; j <= hi[1]; ++j)
                  for (int i = 0; i <= hi[0]; ++i)
                    *te_c(&N, nb, el, c, k, j, i) =
                        *te_f(&O, ob, el, c, k + sh[2], j + sh[1], i + sh[0]);
          }
          continue;
        }
      }
      /* derefined: the daughters were leaves */
      for (int q = 0; q < nleaf; ++q) {
        const Loc dl = daughter(m, nl, q);
        ob = find_leaf(m, &dl);
        if (ob < 0) {
          fprintf(stderr, "oracle: remesh cannot find the origin of a new block\n");
          abort();
        }
        for (int el = 0; el < O.nel; ++el) {
          int top[3], s[3], e[3], sh[3];
          te_top_offset(kind, el, top);
          /* Restrict over GetInteriorRestrict: the coarse interior of the element */
          for (int d = 0; d < 3; ++d) {
            s[d] = m->cis[d];
            e[d] = m->cn[d] == 1 ? 0 : m->cie[d] + top[d];
          }
          for (int c = 0; c < nc; ++c)
            for (int k = s[2]; k <= e[2]; ++k)
              for (int j = s[1]; j <= e[1]; ++j)
                for (int i = s[0]; i <= e[0]; ++i) te_restrict(&O, kind, ob, el, c, k, j, i);
          /* TryRecvFineToCoarse :210-232: shared elements come from the upper daughter */
          for (int d = 0; d < 3; ++d) {
            const int ox = (int)(dl.lx[d] & 1L);
            if (d < m->ndim && ox == 0) e[d] -= top[d];
            sh[d] = (ox == 0 || d >= m->ndim) ? 0 : (e[d] - s[d] + 1 - top[d]);
          }
          for (int c = 0; c < nc; ++c)
            for (int k = s[2]; k <= e[2]; ++k)
              for (int j = s[1]; j <= e[1]; ++j)
                for (int i = s[0]; i <= e[0]; ++i)
                  *te_f(&N, nb, el, c, k + sh[2], j + sh[1], i + sh[0]) =
                      *te_c(&O, ob, el, c, k, j, i);
        }
      }
    }
    /* ProlongateShared over GetInteriorProlongate of the new fine blocks (:925-937):
     * CalcIndices(InteriorRecv, prores) = the coarse interior of the element +- ng / 2 */
    for (int nb = 0; nb < nm->nblocks; ++nb) {
      if (!newly[nb]) continue;
      for (int el = 0; el < N.nel; ++el) {
        int top[3], s[3], e[3];
        te_top_offset(kind, el, top);
        for (int d = 0; d < 3; ++d) {
          const int g2 = d < nm->ndim ? nm->ng / 2 : 0;
          s[d] = nm->cis[d] - g2;
          e[d] = (nm->cn[d] == 1 ? 0 : nm->cie[d] + top[d]) + g2;
        }
        for (int c = 0; c < nc; ++c)
          for (int k = s[2]; k <= e[2]; ++k)
            for (int j = s[1]; j <= e[1]; ++j)
              for (int i = s[0]; i <= e[0]; ++i) te_prolongate_shared(&N, kind, nb, el, c, k, j, i, 0);
      }
    }
  }
  /* :958-990: shared elements of new fine blocks from neighbours that were fine already
   * (ownership favours old blocks), then the internal elements of the new fine blocks */
  g_newly_refined = newly;
  for (int f = 0; f < 3; ++f)
    exchange_te_impl(nm, nU[f], nUc[f], kTeNcomp[f], f + 1, f == 0 && a->te_toth_roe, 0);
  g_newly_refined = NULL;
  for (int f = 0; f < 3; ++f) {
    const int kind = f + 1;
    TeField N;
    te_field_init(&N, nm, nU[f], nUc[f], kTeNcomp[f], kind);
    for (int nb = 0; nb < nm->nblocks; ++nb) {
      if (!newly[nb]) continue;
      for (int el = 0; el < N.nel; ++el) {
        int ftop[3];
        te_top_offset(kind, el, ftop);
        for (int q = 0; q < 8; ++q) {
          int ctop[3], s[3], e[3];
          te_top_offset(kCelKind[q], kCelEl[q], ctop);
          const int tr = f == 0 && a->te_toth_roe;
          if (tr ? kCelKind[q] != ORC_TE_CELL : !te_is_submanifold(ftop, ctop)) continue;
          for (int d = 0; d < 3; ++d) {
            const int g2 = d < nm->ndim ? nm->ng / 2 : 0;
            s[d] = nm->cis[d] - g2;
            e[d] = (nm->cn[d] == 1 ? 0 : nm->cie[d] + ctop[d]) + g2;
          }
          for (int c = 0; c < N.ncomp; ++c)
            for (int k = s[2]; k <= e[2]; ++k)
              for (int j = s[1]; j <= e[1]; ++j)
                for (int i = s[0]; i <= e[0]; ++i) {
                  if (tr)
                    te_toth_roe(&N, nb, el, c, k, j, i);
                  else
                    te_prolongate_internal(&N, ftop, ctop, nb, el, c, k, j, i);
                }
        }
      }
    }
  }
  /* :992-1003: the regular ownership again, then the exchange of everything */
  for (int f = 0; f < 3; ++f)
    exchange_te_impl(nm, nU[f], nUc[f], kTeNcomp[f], f + 1, f == 0 && a->te_toth_roe, 0);
  for (int f = 0; f < 3; ++f) {
    free(a->te_U[f]);
    free(a->te_Uc[f]);
    a->te_U[f] = nU[f];
    a->te_Uc[f] = nUc[f];
  }
  orc_mesh_destroy(a->m);
  free(a->deref_count);
  free(a->refine_flag);
  free(a->last_tag);
  free(newly);
  a->m = nm;
  a->deref_count = ncount;
  a->refine_flag = nflag;
  a->last_tag = (int *)calloc((size_t)nm->nblocks, sizeof(int));
}

static int cmp_level_desc(const void *x, const void *y) {
  return ((const Loc *)y)->level - ((const Loc *)x)->level;
}

/* LoadBalancingAndAdaptiveMeshRefinement: returns 1 if the mesh changed */
static int amr_remesh(struct OrcAmr *a) {
  const OrcMesh *m = a->m;
  const int nleaf = ndaughters(m);
  int tnref = 0, tnderef = 0;
  for (int b = 0; b < m->nblocks; ++b) {
    tnref += a->refine_flag[b] == 1;
    tnderef += a->refine_flag[b] == -1;
  }
  if (tnref == 0 && tnderef < nleaf) return 0;
  Loc *lref = (Loc *)malloc(sizeof(Loc) * (size_t)(tnref + 1));
  Loc *lderef = (Loc *)malloc(sizeof(Loc) * (size_t)(tnderef + 1));
  Loc *clderef = (Loc *)malloc(sizeof(Loc) * (size_t)(tnderef / nleaf + 1));
  int ir = 0, id = 0;
  for (int b = 0; b < m->nblocks; ++b) {
    if (a->refine_flag[b] == 1) lref[ir++] = m->blocks[b].loc;
    if (a->refine_flag[b] == -1 && tnderef >= nleaf) lderef[id++] = m->blocks[b].loc;
  }
  int ctnd = 0;
  if (tnderef >= nleaf) {
    const int lk = m->ndim > 2, lj = m->ndim > 1;
    for (int n = 0; n < tnderef; ++n) {
      if ((lderef[n].lx[0] & 1L) || (lderef[n].lx[1] & 1L) || (lderef[n].lx[2] & 1L)) continue;
      int r = n, rr = 0;
      for (long k = 0; k <= lk; ++k)
        for (long j = 0; j <= lj; ++j)
          for (long i = 0; i <= 1; ++i) {
            if (r < tnderef) {
              if (lderef[n].lx[0] + i == lderef[r].lx[0] && lderef[n].lx[1] + j == lderef[r].lx[1] &&
                  lderef[n].lx[2] + k == lderef[r].lx[2] && lderef[n].level == lderef[r].level)
                rr++;
              r++;
            }
          }
      if (rr == nleaf) clderef[ctnd++] = parent_of(m, &lderef[n]);
    }
  }
  /* :597-602 sorts [first, last) with last = &clderef[ctnd - 1]: all but the final entry */
  if (ctnd > 1) qsort(clderef, (size_t)(ctnd - 1), sizeof(Loc), cmp_level_desc);

  LeafSet t = {0, 0, 0};
  for (int b = 0; b < m->nblocks; ++b) ls_add(&t, &m->blocks[b].loc);
  int nnew = 0, ndel = 0;
  for (int n = 0; n < tnref; ++n) nnew += tree_refine(m, &t, &lref[n]);
  for (int n = 0; n < ctnd; ++n) ndel += tree_derefine(m, &t, &clderef[n]);
  free(lref);
  free(lderef);
  free(clderef);
  if (nnew == 0 && ndel == 0) {
    free(t.leaf);
    return 0;
  }

  /* ---- RedistributeAndRefineMeshBlocks ---- */
  OrcMesh *nm = amr_make_mesh(a, &t);
  free(t.leaf);
  if (a->app == 2) {
    te_remesh(a, nm);
    return 1;
  }
  if (a->app == 3) {
    sparse_remesh(a, nm);
    return 1;
  }
  /* the application state on the new mesh; only the base container is carried over (:667) */
  typedef struct {
    double *U, *Uc;
  } BaseFields;
  BaseFields oa_, na_, *oa = &oa_, *na = &na_;
  OrcAdvection *nadv = NULL;
  OrcBurgers *nbur = NULL;
  if (a->app == 1) {
    nbur = orc_burgers_create(nm, a->num_scalars, a->recon, a->cfl);
    if (!nbur->Uc)
      nbur->Uc = (double *)calloc((size_t)nm->nblocks * a->ncomp * nm->cn[0] * nm->cn[1] * nm->cn[2],
                                  sizeof(double));
    nbur->dt = a->bur->dt;
    nbur->time = a->bur->time;
    nbur->ncycle = a->bur->ncycle;
    nbur->allowed_dt = a->bur->allowed_dt;
    oa_.U = a->bur->U;
    oa_.Uc = a->bur->Uc;
    na_.U = nbur->U;
    na_.Uc = nbur->Uc;
  } else {
    nadv = orc_advection_create(nm, a->ncomp, a->profile, a->amp, a->v, a->cfl);
    nadv->dt = a->adv->dt;
    nadv->time = a->adv->time;
    nadv->ncycle = a->adv->ncycle;
    nadv->allowed_dt = a->adv->allowed_dt;
    oa_.U = a->adv->U;
    oa_.Uc = a->adv->Uc;
    na_.U = nadv->U;
    na_.Uc = nadv->Uc;
  }
  int *ncount = (int *)calloc((size_t)nm->nblocks, sizeof(int));
  int *nflag = (int *)calloc((size_t)nm->nblocks, sizeof(int));
  const int nc = a->ncomp;
  const size_t per_block = (size_t)nc * m->n[0] * m->n[1] * m->n[2];
  const size_t cper_block = (size_t)nc * m->cn[0] * m->cn[1] * m->cn[2];
  double *ctmp = (double *)calloc((size_t)m->nblocks * cper_block, sizeof(double));
  for (int nb = 0; nb < nm->nblocks; ++nb) {
    const Loc *nl = &nm->blocks[nb].loc;
    /* same location in the old mesh: the block object is kept (all data, counters) */
    int ob = find_leaf(m, nl);
    if (ob >= 0) {
      memcpy(na->U + (size_t)nb * per_block, oa->U + (size_t)ob * per_block,
             per_block * sizeof(double));
      ncount[nb] = a->deref_count[ob];
      nflag[nb] = a->refine_flag[ob];
      continue;
    }
    /* refined: the old parent was a leaf */
    if (nl->level > 0) {
      const Loc par = parent_of(m, nl);
      ob = find_leaf(m, &par);
      if (ob >= 0) {
        /* TryRecvCoarseToFine :87-149: coarse buffer (entire extents) <- parent fine data */
        int sh[3];
        for (int d = 0; d < 3; ++d)
          sh[d] = (d < m->ndim && (nl->lx[d] & 1L)) ? (m->ie[d] - m->is[d] + 1) / 2 : 0;
        for (int c = 0; c < nc; ++c)
          for (int k = 0; k < m->cn[2]; ++k)
            for (int j = 0; j < m->cn[1]; ++j)
              for (int i = 0; i < m->cn[0]; ++i)
                na->Uc[cidx(nm, nc, nb, c, k, j, i)] =
                    oa->U[fidx(m, nc, ob, c, k + sh[2], j + sh[1], i + sh[0])];
        /* ProlongateShared over GetInteriorProlongate: coarse interior +- ng/2
         * (bnd_info.cpp:199-203) */
        int s[3], e[3];
        for (int d = 0; d < 3; ++d) {
          const int g2 = d < m->ndim ? m->ng / 2 : 0;
          s[d] = nm->cis[d] - g2;
          e[d] = nm->cie[d] + g2;
        }
        for (int c = 0; c < nc; ++c)
          for (int k = s[2]; k <= e[2]; ++k)
            for (int j = s[1]; j <= e[1]; ++j)
              for (int i = s[0]; i <= e[0]; ++i) prolongate_cell(nm, na->U, na->Uc, nc, nb, c, k, j, i);
        continue;
      }
    }
    /* derefined: the daughters were leaves */
    for (int q = 0; q < nleaf; ++q) {
      const Loc d = daughter(m, nl, q);
      ob = find_leaf(m, &d);
      if (ob < 0) {
        fprintf(stderr, "oracle: remesh cannot find the origin of a new block\n");
        abort();
      }
      /* Restrict over GetInteriorRestrict (coarse interior) into the child's coarse buffer */
      int s[3] = {m->cis[0], m->cis[1], m->cis[2]}, e[3] = {m->cie[0], m->cie[1], m->cie[2]};
      restrict_region(m, oa->U, ctmp, nc, ob, s, e);
      /* TryRecvFineToCoarse :196-251: parent fine sub-box <- child's coarse interior */
      int sh[3];
      for (int dd = 0; dd < 3; ++dd)
        sh[dd] = (dd < m->ndim && (d.lx[dd] & 1L)) ? (m->cie[dd] - m->cis[dd] + 1) : 0;
      for (int c = 0; c < nc; ++c)
        for (int k = m->cis[2]; k <= m->cie[2]; ++k)
          for (int j = m->cis[1]; j <= m->cie[1]; ++j)
            for (int i = m->cis[0]; i <= m->cie[0]; ++i)
              na->U[fidx(nm, nc, nb, c, k + sh[2], j + sh[1], i + sh[0])] =
                  ctmp[cidx(m, nc, ob, c, k, j, i)];
    }
  }
  free(ctmp);
  if (a->app == 1) {
    orc_burgers_destroy(a->bur);
    a->bur = nbur;
  } else {
    orc_advection_destroy(a->adv);
    a->adv = nadv;
  }
  orc_mesh_destroy(a->m);
  free(a->deref_count);
  free(a->refine_flag);
  free(a->last_tag);
  a->m = nm;
  a->deref_count = ncount;
  a->refine_flag = nflag;
  a->last_tag = (int *)calloc((size_t)nm->nblocks, sizeof(int));
  /* PreCommFillDerived; CommunicateBoundaries; FillDerived  (:1000-1003) */
  orc_exchange(nm, na->U, na->Uc, nc, 1);
  orc_apply_bcs(nm, na->U, nc);
  if (a->app == 1) calculate_derived(a->bur, a->bur->U);
  return 1;
}

/* Mesh::Initialize mesh.cpp:745-860 + EvolutionDriver::InitializeBlockTimeStepsAndBoundaries */
static void amr_init_burgers(struct OrcAmr *a) {
  int done;
  do {
    OrcBurgers *st = a->bur;
    orc_burgers_ic(a->m, st->U, st->ncomp);
    orc_exchange(a->m, st->U, st->Uc, st->ncomp, 1);
    orc_apply_bcs(a->m, st->U, st->ncomp);
    calculate_derived(st, st->U);
    amr_tag(a, st->U);
    done = !amr_remesh(a);
  } while (!done);
  OrcBurgers *st = a->bur;
  st->allowed_dt = estimate_timestep(st, st->U);
  st->dt = DBL_MAX;
  set_global_timestep(st);
  st->time = 0;
  st->ncycle = 0;
}

void orc_amr_init(struct OrcAmr *a) {
  if (a->app == 1) {
    amr_init_burgers(a);
    return;
  }
  if (a->app == 2) {
    amr_init_te(a);
    return;
  }
  int done;
  do {
    OrcAdvection *st = a->adv;
    advection_ic(st);
    orc_exchange(a->m, st->U, st->Uc, st->ncomp, 1);
    orc_apply_bcs(a->m, st->U, st->ncomp);
    amr_tag(a, st->U);
    done = !amr_remesh(a);
  } while (!done);
  OrcAdvection *st = a->adv;
  st->allowed_dt = advection_estimate_timestep(st);
  st->dt = fmin(DBL_MAX, st->allowed_dt);
  st->allowed_dt = DBL_MAX;
  st->time = 0;
  st->ncycle = 0;
}

/* one pass of the main loop, driver.cpp:99-150, in two halves so that a caller can look at
 * the state where the reference's PostStepUserWorkInLoop hook sees it (after Step, before
 * the mesh is adapted):
 *   orc_amr_step    Step (both stages, tagging on the last stage's container,
 *                   advection_driver.cpp:158), ncycle / time advance
 *   orc_amr_regrid  LoadBalancingAndAdaptiveMeshRefinement, InitializeBlockTimeSteps if the
 *                   mesh changed, SetGlobalTimeStep; returns 1 if the mesh changed */
void orc_amr_step(struct OrcAmr *a) {
  if (a->app == 1) {
    OrcBurgers *bs = a->bur;
    orc_burgers_stage(bs, 1);
    orc_burgers_stage(bs, 2);
    amr_tag(a, bs->U); /* Refinement::Tag after ApplyBoundaryConditions, burgers_driver.cpp:137-144 */
    bs->ncycle++;
    bs->time += bs->dt;
    return;
  }
  OrcAdvection *st = a->adv;
  orc_advection_stage(st, 1);
  orc_advection_stage(st, 2);
  amr_tag(a, st->U);
  st->ncycle++;
  st->time += st->dt;
}
int orc_amr_regrid(struct OrcAmr *a) {
  const int changed = amr_remesh(a);
  if (a->app == 1) {
    OrcBurgers *bs = a->bur;
    if (changed) bs->allowed_dt = estimate_timestep(bs, bs->U); /* InitializeBlockTimeSteps */
    set_global_timestep(bs);
    return changed;
  }
  OrcAdvection *st = a->adv;
  if (changed) st->allowed_dt = advection_estimate_timestep(st);
  if (st->dt < 0.1 * DBL_MAX) st->dt *= 2.0;
  st->dt = fmin(st->dt, st->allowed_dt);
  st->allowed_dt = DBL_MAX;
  return changed;
}

/* ------------------------------------------------------------------------------------ */
/* example/sparse_advection on a uniform mesh (2-D in the reference):
 * NF sparse fields "sparse_f" with one component each; a field exists on a block only
 * where it is ALLOCATED.  Sparse semantics of the exchange:
 *   send    boundary_communication.cpp:95-159  flag = allocated && any |x| >= alloc threshold
 *           in the send box; flag false => SendNull
 *   receive :201-235  a non-null message for an unallocated field allocates it on that block
 *           in every stage (MeshBlock::AllocateSparse meshblock.cpp:277-325; arrays start
 *           zeroed, variable.cpp:112-127)
 *   set     :273-334  allocated receiver: data if the message was non-null, else the sparse
 *           default value (0)
 *   Update::SparseDealloc update.cpp:143-217 after the last stage. */
#define ORC_NF 4
struct OrcSparse {
  const OrcMesh *m;
  double cfl, alloc_thr, dealloc_thr, init_size;
  int dealloc_count;
  double vx[ORC_NF], vy[ORC_NF], vz[ORC_NF], x0[ORC_NF], y0[ORC_NF];
  size_t ncell, nfield; /* per field: nblocks * ncell */
  double *U[ORC_NF], *U1[ORC_NF], *dUdt[ORC_NF], *flux[ORC_NF][3];
  double *Uc[ORC_NF];   /* coarse buffers (multilevel meshes) */
  unsigned char *alloc; /* [nblocks][NF] */
  int *counter;         /* [nblocks][NF] dealloc_count of the control variable */
  double dt, time, allowed_dt;
  int ncycle;
};

OrcSparse *orc_sparse_create(const OrcMesh *m, double speed, double cfl, double alloc_thr,
                             double dealloc_thr, int dealloc_count) {
  OrcSparse *s = (OrcSparse *)calloc(1, sizeof(OrcSparse));
  s->m = m;
  s->cfl = cfl;
  s->alloc_thr = alloc_thr;
  s->dealloc_thr = dealloc_thr;
  s->dealloc_count = dealloc_count;
  s->init_size = 0.1;
  /* sparse_advection_package.cpp:60-70 */
  const double pos = 0.8, sp = speed / sqrt(2.0);
  const double x0[4] = {pos, -pos, -pos, pos}, y0[4] = {pos, pos, -pos, -pos};
  const double vx[4] = {-sp, sp, sp, -sp}, vy[4] = {-sp, -sp, sp, sp};
  s->ncell = (size_t)m->n[0] * m->n[1] * m->n[2];
  s->nfield = (size_t)m->nblocks * s->ncell;
  for (int f = 0; f < ORC_NF; ++f) {
    s->x0[f] = x0[f];
    s->y0[f] = y0[f];
    s->vx[f] = vx[f];
    s->vy[f] = vy[f];
    s->vz[f] = 0.0; /* sparse_advection_package.cpp:70 */
    s->U[f] = (double *)calloc(s->nfield, sizeof(double));
    s->U1[f] = (double *)calloc(s->nfield, sizeof(double));
    s->dUdt[f] = (double *)calloc(s->nfield, sizeof(double));
    for (int d = 0; d < 3; ++d) s->flux[f][d] = (double *)calloc(s->nfield, sizeof(double));
    s->Uc[f] = m->multilevel ? (double *)calloc((size_t)m->nblocks * m->cn[0] * m->cn[1] * m->cn[2],
                                                sizeof(double))
                             : NULL;
  }
  s->alloc = (unsigned char *)calloc((size_t)m->nblocks * ORC_NF, 1);
  s->counter = (int *)calloc((size_t)m->nblocks * ORC_NF, sizeof(int));
  s->dt = DBL_MAX;
  return s;
}
void orc_sparse_destroy(OrcSparse *s) {
  if (!s) return;
  for (int f = 0; f < ORC_NF; ++f) {
    free(s->U[f]);
    free(s->U1[f]);
    free(s->dUdt[f]);
    free(s->Uc[f]);
    for (int d = 0; d < 3; ++d) free(s->flux[f][d]);
  }
  free(s->alloc);
  free(s->counter);
  free(s);
}
double *orc_sparse_U(OrcSparse *s, int f) { return s->U[f]; }
const unsigned char *orc_sparse_alloc(const OrcSparse *s) { return s->alloc; }
double orc_sparse_dt(const OrcSparse *s) { return s->dt; }
double orc_sparse_time(const OrcSparse *s) { return s->time; }

/* MeshBlock::AllocateSparse: the field appears, zero-filled, in every stage */
static void sparse_allocate(OrcSparse *st, int b, int f) {
  st->alloc[b * ORC_NF + f] = 1;
  const size_t o = (size_t)b * st->ncell, n = st->ncell * sizeof(double);
  memset(st->U[f] + o, 0, n);
  memset(st->U1[f] + o, 0, n);
  memset(st->dUdt[f] + o, 0, n);
  for (int d = 0; d < 3; ++d) memset(st->flux[f][d] + o, 0, n);
  if (st->Uc[f]) {
    const size_t cn = (size_t)st->m->cn[0] * st->m->cn[1] * st->m->cn[2];
    memset(st->Uc[f] + (size_t)b * cn, 0, cn * sizeof(double));
  }
}

/* the same exchange on a MULTILEVEL mesh: the allocation-aware forms of restriction (send and
 * set), pack / unpack through the coarse buffers, and prolongation (ProResInfo::allocated,
 * pr_loops.hpp:43-46 DoRefinementOp) */
static void sparse_exchange_ml(OrcSparse *st, int f, double *U) {
  const OrcMesh *m = st->m;
  double *Uc = st->Uc[f];
  const unsigned char *al = st->alloc;
#define ALLOC(b) (al[(b)*ORC_NF + f])
  int64_t nreg = orc_count_regions(m);
  int64_t *off = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nreg + 1));
  int64_t total = orc_pack(m, U, Uc, 1, NULL, off);
  double *buf = (double *)malloc(sizeof(double) * (size_t)(total > 0 ? total : 1));
  unsigned char *flag = (unsigned char *)calloc((size_t)nreg, 1);
  int64_t *first = (int64_t *)malloc(sizeof(int64_t) * (size_t)(m->nblocks + 1));
  first[0] = 0;
  for (int b = 0; b < m->nblocks; ++b) first[b + 1] = first[b] + m->blocks[b].nnb;
  /* restriction of the send regions that face a coarser block */
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    if (!ALLOC(b)) continue;
    for (int n = 0; n < blk->nnb; ++n)
      if (blk->nb[n].origin_loc.level < blk->loc.level) {
        int s[3], e[3];
        orc_calc_indices(m, b, n, IR_SEND, 1, s, e);
        restrict_region(m, U, Uc, 1, b, s, e);
      }
  }
  orc_pack(m, U, Uc, 1, buf, off);
  for (int b = 0; b < m->nblocks; ++b)
    for (int n = 0; n < m->blocks[b].nnb; ++n) {
      const int64_t r = first[b] + n;
      if (!ALLOC(b)) continue;
      for (int64_t q = off[r]; q < off[r + 1]; ++q)
        if (fabs(buf[q]) >= st->alloc_thr) {
          flag[r] = 1;
          break;
        }
    }
  /* receive: allocate on the first non-null message; set */
  int *sender_region = (int *)malloc(sizeof(int) * (size_t)nreg);
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    for (int n = 0; n < blk->nnb; ++n) {
      const Neighbor *nb = &blk->nb[n];
      const Block *sb = &m->blocks[nb->gid];
      int sn = -1;
      for (int q = 0; q < sb->nnb; ++q)
        if (sb->nb[q].gid == b && sb->nb[q].off[0] == -nb->off[0] &&
            sb->nb[q].off[1] == -nb->off[1] && sb->nb[q].off[2] == -nb->off[2]) {
          sn = q;
          break;
        }
      if (sn < 0) abort();
      sender_region[first[b] + n] = (int)(first[nb->gid] + sn);
      if (flag[first[nb->gid] + sn] && !ALLOC(b)) sparse_allocate(st, b, f);
    }
  }
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    if (!ALLOC(b)) continue;
    for (int n = 0; n < blk->nnb; ++n) {
      const Neighbor *nb = &blk->nb[n];
      int s[3], e[3];
      orc_calc_indices(m, b, n, IR_RECV, 0, s, e);
      const int64_t r = sender_region[first[b] + n];
      const int coarse = nb->origin_loc.level < blk->loc.level;
      const double *p = buf + off[r];
      for (int k = s[2]; k <= e[2]; ++k)
        for (int j = s[1]; j <= e[1]; ++j)
          for (int i = s[0]; i <= e[0]; ++i) {
            const double val = flag[r] ? *p : 0.0; /* sparse_default_val */
            ++p;
            if (coarse)
              Uc[cidx(m, 1, b, 0, k, j, i)] = val;
            else
              U[fidx(m, 1, b, 0, k, j, i)] = val;
          }
    }
  }
  /* restriction of the received regions of blocks with a coarser neighbour */
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    if (!ALLOC(b)) continue;
    int restricted = 0;
    if (blk->loc.level > 0)
      for (int n = 0; n < blk->nnb; ++n)
        restricted = restricted || (blk->nb[n].origin_loc.level == blk->loc.level - 1);
    if (!restricted) continue;
    for (int n = 0; n < blk->nnb; ++n) {
      if (blk->nb[n].origin_loc.level < blk->loc.level) continue;
      int s[3], e[3];
      orc_calc_indices(m, b, n, IR_RECV, 1, s, e);
      restrict_region(m, U, Uc, 1, b, s, e);
    }
  }
  /* prolongation into the ghosts that face a coarser block */
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    if (!ALLOC(b)) continue;
    for (int n = 0; n < blk->nnb; ++n) {
      if (!(blk->nb[n].origin_loc.level < blk->loc.level)) continue;
      int s[3], e[3];
      orc_calc_indices(m, b, n, IR_RECV, 1, s, e);
      for (int k = s[2]; k <= e[2]; ++k)
        for (int j = s[1]; j <= e[1]; ++j)
          for (int i = s[0]; i <= e[0]; ++i) prolongate_cell(m, U, Uc, 1, b, 0, k, j, i);
    }
  }
#undef ALLOC
  free(sender_region);
  free(first);
  free(flag);
  free(buf);
  free(off);
}

/* the exchange of ONE sparse field of one container (see the section comment) */
static void sparse_exchange(OrcSparse *st, int f, double *U) {
  const OrcMesh *m = st->m;
  if (m->multilevel) {
    sparse_exchange_ml(st, f, U);
    return;
  }
  int64_t nreg = orc_count_regions(m);
  int64_t *off = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nreg + 1));
  int64_t total = orc_pack(m, U, NULL, 1, NULL, off);
  double *buf = (double *)malloc(sizeof(double) * (size_t)(total > 0 ? total : 1));
  unsigned char *flag = (unsigned char *)calloc((size_t)nreg, 1);
  int64_t *first = (int64_t *)malloc(sizeof(int64_t) * (size_t)(m->nblocks + 1));
  first[0] = 0;
  for (int b = 0; b < m->nblocks; ++b) first[b + 1] = first[b] + m->blocks[b].nnb;
  orc_pack(m, U, NULL, 1, buf, off);
  for (int b = 0; b < m->nblocks; ++b)
    for (int n = 0; n < m->blocks[b].nnb; ++n) {
      const int64_t r = first[b] + n;
      if (!st->alloc[b * ORC_NF + f]) continue;
      for (int64_t q = off[r]; q < off[r + 1]; ++q)
        if (fabs(buf[q]) >= st->alloc_thr) {
          flag[r] = 1;
          break;
        }
    }
  /* receive: allocate on the first non-null message */
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    for (int n = 0; n < blk->nnb; ++n) {
      const Neighbor *nb = &blk->nb[n];
      const Block *sb = &m->blocks[nb->gid];
      for (int q = 0; q < sb->nnb; ++q)
        if (sb->nb[q].gid == b && sb->nb[q].off[0] == -nb->off[0] &&
            sb->nb[q].off[1] == -nb->off[1] && sb->nb[q].off[2] == -nb->off[2]) {
          if (flag[first[nb->gid] + q] && !st->alloc[b * ORC_NF + f]) sparse_allocate(st, b, f);
          break;
        }
    }
  }
  /* set */
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    if (!st->alloc[b * ORC_NF + f]) continue;
    for (int n = 0; n < blk->nnb; ++n) {
      const Neighbor *nb = &blk->nb[n];
      const Block *sb = &m->blocks[nb->gid];
      int sn = -1;
      for (int q = 0; q < sb->nnb; ++q)
        if (sb->nb[q].gid == b && sb->nb[q].off[0] == -nb->off[0] &&
            sb->nb[q].off[1] == -nb->off[1] && sb->nb[q].off[2] == -nb->off[2]) {
          sn = q;
          break;
        }
      int s[3], e[3];
      orc_calc_indices(m, b, n, IR_RECV, 0, s, e);
      const int64_t r = first[nb->gid] + sn;
      const double *p = buf + off[r];
      for (int k = s[2]; k <= e[2]; ++k)
        for (int j = s[1]; j <= e[1]; ++j)
          for (int i = s[0]; i <= e[0]; ++i) {
            const double val = *p++;
            U[fidx(m, 1, b, 0, k, j, i)] = flag[r] ? val : 0.0; /* sparse_default_val */
          }
    }
  }
  free(first);
  free(flag);
  free(buf);
  free(off);
}

/* ProblemGenerator example/sparse_advection/parthenon_app_inputs.cpp:42-105 */
static void sparse_ic(OrcSparse *st) {
  const OrcMesh *m = st->m;
  const double size = st->init_size * st->init_size;
  for (int f = 0; f < ORC_NF; ++f)
    for (int b = 0; b < m->nblocks; ++b) {
      const Block *blk = &m->blocks[b];
      int any = 0;
      for (int k = m->is[2]; k <= m->ie[2] && !any; ++k)
        for (int j = m->is[1]; j <= m->ie[1] && !any; ++j)
          for (int i = m->is[0]; i <= m->ie[0] && !any; ++i) {
            const double x = xc(blk, 0, i) - st->x0[f], y = xc(blk, 1, j) - st->y0[f],
                         z = xc(blk, 2, k);
            if (x * x + y * y + z * z < size) any = 1;
          }
      if (!any) continue;
      if (!st->alloc[b * ORC_NF + f]) sparse_allocate(st, b, f);
      for (int k = m->is[2]; k <= m->ie[2]; ++k)
        for (int j = m->is[1]; j <= m->ie[1]; ++j)
          for (int i = m->is[0]; i <= m->ie[0]; ++i) {
            const double x = xc(blk, 0, i) - st->x0[f], y = xc(blk, 1, j) - st->y0[f],
                         z = xc(blk, 2, k);
            st->U[f][fidx(m, 1, b, 0, k, j, i)] = (x * x + y * y + z * z < size ? 1.0 : 0.0);
          }
    }
}

/* EstimateTimestepBlock sparse_advection_package.cpp:136-168: all fields vote, allocated
 * or not */
static double sparse_estimate_timestep(const OrcSparse *st) {
  const OrcMesh *m = st->m;
  double min_dt = DBL_MAX;
  for (int b = 0; b < m->nblocks; ++b) {
    const Block *blk = &m->blocks[b];
    double bdt = DBL_MAX;
    for (int f = 0; f < ORC_NF; ++f) {
      if (st->vx[f] != 0.0) bdt = fmin(bdt, blk->dx[0] / fabs(st->vx[f]));
      if (st->vy[f] != 0.0) bdt = fmin(bdt, blk->dx[1] / fabs(st->vy[f]));
      if (st->vz[f] != 0.0) bdt = fmin(bdt, blk->dx[2] / fabs(st->vz[f]));
    }
    min_dt = fmin(min_dt, st->cfl * bdt);
  }
  return min_dt;
}

/* Update::SparseDealloc update.cpp:143-217 on container `U` (the stage's mc1) */
static void sparse_dealloc(OrcSparse *st, double *const U[ORC_NF]) {
  const OrcMesh *m = st->m;
  for (int b = 0; b < m->nblocks; ++b)
    for (int f = 0; f < ORC_NF; ++f) {
      if (!st->alloc[b * ORC_NF + f]) continue;
      int all_zero = 1;
      const double *p = U[f] + (size_t)b * st->ncell;
      for (size_t q = 0; q < st->ncell; ++q)
        if (fabs(p[q]) > st->dealloc_thr) {
          all_zero = 0;
          break;
        }
      int *counter = &st->counter[b * ORC_NF + f];
      if (all_zero)
        (*counter)++;
      else
        *counter = 0;
      if (*counter > st->dealloc_count) {
        *counter = 0;
        st->alloc[b * ORC_NF + f] = 0;
      }
    }
}

/* one stage of SparseAdvectionDriver::MakeTaskCollection sparse_advection_driver.cpp:56-149 */
static void sparse_stage(OrcSparse *st, int stage) {
  const OrcMesh *m = st->m;
  const double beta = stage == 1 ? 1.0 : 0.5;
  const size_t sj = (size_t)m->n[0], sk = (size_t)m->n[0] * m->n[1];
  const int dim3 = m->ndim > 2; /* the reference stops at 2-D (:256-257); the x3 terms below
                                   are the same donor-cell / divergence formulas one
                                   direction further (no reference parity in 3-D) */
  for (int f = 0; f < ORC_NF; ++f) {
    double *mc0 = stage == 1 ? st->U[f] : st->U1[f];
    double *mc1 = stage == 1 ? st->U1[f] : st->U[f];
    double *base = st->U[f];
    const double v[3] = {st->vx[f], st->vy[f], st->vz[f]};
    for (int b = 0; b < m->nblocks; ++b) {
      if (!st->alloc[b * ORC_NF + f]) continue; /* IsAllocated guards everywhere */
      /* CalculateFluxes sparse_advection_package.cpp:173-258 (donor cell) */
      for (int k = m->is[2]; k <= m->ie[2] + dim3; ++k)
        for (int j = m->is[1]; j <= m->ie[1] + 1; ++j)
          for (int i = m->is[0]; i <= m->ie[0] + 1; ++i) {
            const size_t p = fidx(m, 1, b, 0, k, j, i);
            if (j <= m->ie[1] && k <= m->ie[2])
              st->flux[f][0][p] = (v[0] > 0.0 ? mc0[p - 1] : mc0[p]) * v[0];
            if (i <= m->ie[0] && k <= m->ie[2])
              st->flux[f][1][p] = (v[1] > 0.0 ? mc0[p - sj] : mc0[p]) * v[1];
            if (dim3 && i <= m->ie[0] && j <= m->ie[1])
              st->flux[f][2][p] = (v[2] > 0.0 ? mc0[p - sk] : mc0[p]) * v[2];
          }
    }
    /* AddFluxCorrectionTasks sparse_advection_driver.cpp:99-104 (multilevel meshes) */
    if (m->multilevel) flux_correct_masked(m, st->flux[f], 1, st->alloc + f, ORC_NF);
    for (int b = 0; b < m->nblocks; ++b) {
      if (!st->alloc[b * ORC_NF + f]) continue;
      const Block *blk = &m->blocks[b];
      /* FluxDivergence update.cpp:63-86 */
      const double a1 = blk->dx[1] * blk->dx[2], a2 = blk->dx[0] * blk->dx[2],
                   a3 = blk->dx[0] * blk->dx[1];
      const double vol = blk->dx[0] * blk->dx[1] * blk->dx[2];
      for (int k = m->is[2]; k <= m->ie[2]; ++k)
        for (int j = m->is[1]; j <= m->ie[1]; ++j)
          for (int i = m->is[0]; i <= m->ie[0]; ++i) {
            const size_t p = fidx(m, 1, b, 0, k, j, i);
            double du = (a1 * st->flux[f][0][p + 1] - a1 * st->flux[f][0][p]);
            du += (a2 * st->flux[f][1][p + sj] - a2 * st->flux[f][1][p]);
            if (dim3) du += (a3 * st->flux[f][2][p + sk] - a3 * st->flux[f][2][p]);
            st->dUdt[f][p] = -du / vol;
          }
      /* Average / UpdateIndependentData update.hpp:71-91 over the entire extents */
      const size_t o = (size_t)b * st->ncell;
      for (size_t q = o; q < o + st->ncell; ++q) mc0[q] = beta * mc0[q] + (1.0 - beta) * base[q];
      for (size_t q = o; q < o + st->ncell; ++q)
        mc1[q] = 1.0 * mc0[q] + (beta * st->dt) * st->dUdt[f][q];
    }
  }
  for (int f = 0; f < ORC_NF; ++f) sparse_exchange(st, f, stage == 1 ? st->U1[f] : st->U[f]);
  if (stage == 2) {
    sparse_dealloc(st, st->U);
    st->allowed_dt = sparse_estimate_timestep(st);
  }
}

void orc_sparse_init(OrcSparse *st) {
  sparse_ic(st);
  for (int f = 0; f < ORC_NF; ++f) sparse_exchange(st, f, st->U[f]);
  st->allowed_dt = sparse_estimate_timestep(st);
  st->dt = fmin(DBL_MAX, st->allowed_dt);
  st->allowed_dt = DBL_MAX;
  st->time = 0;
  st->ncycle = 0;
}

void orc_sparse_step(OrcSparse *st) {
  sparse_stage(st, 1);
  sparse_stage(st, 2);
  st->ncycle++;
  st->time += st->dt;
  if (st->dt < 0.1 * DBL_MAX) st->dt *= 2.0;
  st->dt = fmin(st->dt, st->allowed_dt);
  st->allowed_dt = DBL_MAX;
}


/* ---- example/sparse_advection with refinement = adaptive ---- */
static void sparse_minmax(const struct OrcSparse *st, int b, double *mn, double *mx) {
  *mn = DBL_MAX;
  *mx = -DBL_MAX;
  for (int f = 0; f < ORC_NF; ++f) {
    if (!st->alloc[b * ORC_NF + f]) continue;
    const double *p = st->U[f] + (size_t)b * st->ncell;
    for (size_t q = 0; q < st->ncell; ++q) {
      *mn = p[q] < *mn ? p[q] : *mn;
      *mx = p[q] > *mx ? p[q] : *mx;
    }
  }
}

/* RedistributeAndRefineMeshBlocks for the sparse fields: a new child exists where its parent
 * did (TryRecvCoarseToFine :100-103, :137-141), a new parent where any daughter did
 * (TryRecvFineToCoarse :193-195); counters of Update::SparseDealloc live on the block object */
static void sparse_remesh(struct OrcAmr *a, OrcMesh *nm) {
  const OrcMesh *m = a->m;
  OrcSparse *os = a->sp;
  OrcSparse *ns = orc_sparse_create(nm, a->sp_speed, os->cfl, a->sp_alloc_thr, a->sp_dealloc_thr,
                                    a->sp_dealloc_count);
  ns->dt = os->dt;
  ns->time = os->time;
  ns->ncycle = os->ncycle;
  ns->allowed_dt = os->allowed_dt;
  const int nleaf = ndaughters(m);
  int *ncount = (int *)calloc((size_t)nm->nblocks, sizeof(int));
  int *nflag = (int *)calloc((size_t)nm->nblocks, sizeof(int));
  const size_t cper = (size_t)m->cn[0] * m->cn[1] * m->cn[2];
  double *ctmp = (double *)calloc((size_t)m->nblocks * cper, sizeof(double));
  for (int nb = 0; nb < nm->nblocks; ++nb) {
    const Loc *nl = &nm->blocks[nb].loc;
    int ob = find_leaf(m, nl);
    if (ob >= 0) {
      ncount[nb] = a->deref_count[ob];
      nflag[nb] = a->refine_flag[ob];
    }
    for (int f = 0; f < ORC_NF; ++f) {
      double *nU = ns->U[f], *nUc = ns->Uc[f];
      const double *oU = os->U[f];
      if (ob >= 0) { /* kept */
        if (!os->alloc[ob * ORC_NF + f]) continue;
        ns->alloc[nb * ORC_NF + f] = 1;
        ns->counter[nb * ORC_NF + f] = os->counter[ob * ORC_NF + f];
        memcpy(nU + (size_t)nb * ns->ncell, oU + (size_t)ob * os->ncell, os->ncell * sizeof(double));
        continue;
      }
      int pb = -1;
      if (nl->level > 0) {
        const Loc par = parent_of(m, nl);
        pb = find_leaf(m, &par);
      }
      if (pb >= 0) { /* refined */
        if (!os->alloc[pb * ORC_NF + f]) continue;
        sparse_allocate(ns, nb, f);
        int sh[3];
        for (int d = 0; d < 3; ++d)
          sh[d] = (d < m->ndim && (nl->lx[d] & 1L)) ? (m->ie[d] - m->is[d] + 1) / 2 : 0;
        for (int k = 0; k < m->cn[2]; ++k)
          for (int j = 0; j < m->cn[1]; ++j)
            for (int i = 0; i < m->cn[0]; ++i)
              nUc[cidx(nm, 1, nb, 0, k, j, i)] =
                  oU[fidx(m, 1, pb, 0, k + sh[2], j + sh[1], i + sh[0])];
        int s[3], e[3];
        for (int d = 0; d < 3; ++d) {
          const int g2 = d < m->ndim ? m->ng / 2 : 0;
          s[d] = nm->cis[d] - g2;
          e[d] = nm->cie[d] + g2;
        }
        for (int k = s[2]; k <= e[2]; ++k)
          for (int j = s[1]; j <= e[1]; ++j)
            for (int i = s[0]; i <= e[0]; ++i) prolongate_cell(nm, nU, nUc, 1, nb, 0, k, j, i);
        continue;
      }
      /* derefined */
      for (int q = 0; q < nleaf; ++q) {
        const Loc dl = daughter(m, nl, q);
        const int db = find_leaf(m, &dl);
        if (db < 0) {
          fprintf(stderr, "oracle: remesh cannot find the origin of a new block\n");
          abort();
        }
        if (!os->alloc[db * ORC_NF + f]) continue;
        if (!ns->alloc[nb * ORC_NF + f]) sparse_allocate(ns, nb, f);
        int s[3] = {m->cis[0], m->cis[1], m->cis[2]}, e[3] = {m->cie[0], m->cie[1], m->cie[2]};
        restrict_region(m, oU, ctmp, 1, db, s, e);
        int sh[3];
        for (int dd = 0; dd < 3; ++dd)
          sh[dd] = (dd < m->ndim && (dl.lx[dd] & 1L)) ? (m->cie[dd] - m->cis[dd] + 1) : 0;
        for (int k = m->cis[2]; k <= m->cie[2]; ++k)
          for (int j = m->cis[1]; j <= m->cie[1]; ++j)
            for (int i = m->cis[0]; i <= m->cie[0]; ++i)
              nU[fidx(nm, 1, nb, 0, k + sh[2], j + sh[1], i + sh[0])] =
                  ctmp[cidx(m, 1, db, 0, k, j, i)];
      }
    }
  }
  free(ctmp);
  orc_sparse_destroy(os);
  orc_mesh_destroy(a->m);
  free(a->deref_count);
  free(a->refine_flag);
  free(a->last_tag);
  a->m = nm;
  a->sp = ns;
  a->deref_count = ncount;
  a->refine_flag = nflag;
  a->last_tag = (int *)calloc((size_t)nm->nblocks, sizeof(int));
  /* CommunicateBoundaries on the new mesh (:1000-1003) */
  for (int f = 0; f < ORC_NF; ++f) sparse_exchange(ns, f, ns->U[f]);
}

struct OrcAmr *orc_amr_create_sparse(int ndim, const int nx[3], int ng, const int nrb[3],
                                     const double xmin[3], const double xmax[3], int numlevel,
                                     int derefine_count, double refine_tol, double derefine_tol,
                                     double speed, double cfl, double alloc_thr,
                                     double dealloc_thr, int dealloc_count) {
  const double v0[3] = {0, 0, 0};
  struct OrcAmr *a = orc_amr_create(ndim, nx, ng, nrb, xmin, xmax, numlevel, derefine_count,
                                    refine_tol, derefine_tol, 1, 2, 0.0, v0, cfl);
  orc_advection_destroy(a->adv);
  a->adv = NULL;
  a->app = 3;
  a->sp_speed = speed;
  a->sp_alloc_thr = alloc_thr;
  a->sp_dealloc_thr = dealloc_thr;
  a->sp_dealloc_count = dealloc_count;
  a->sp = orc_sparse_create(a->m, speed, cfl, alloc_thr, dealloc_thr, dealloc_count);
  return a;
}
/* Mesh::Initialize: problem generator, exchange, tag, remesh until the mesh stops changing */
void orc_amr_sparse_init(struct OrcAmr *a) {
  int done;
  do {
    sparse_ic(a->sp);
    for (int f = 0; f < ORC_NF; ++f) sparse_exchange(a->sp, f, a->sp->U[f]);
    amr_tag(a, NULL);
    done = !amr_remesh(a);
  } while (!done);
  OrcSparse *st = a->sp;
  st->allowed_dt = sparse_estimate_timestep(st);
  st->dt = fmin(DBL_MAX, st->allowed_dt);
  st->allowed_dt = DBL_MAX;
  st->time = 0;
  st->ncycle = 0;
}
/* Step + tagging, then LoadBalancingAndAdaptiveMeshRefinement + SetGlobalTimeStep */
void orc_amr_sparse_step(struct OrcAmr *a) {
  OrcSparse *st = a->sp;
  sparse_stage(st, 1);
  sparse_stage(st, 2);
  amr_tag(a, NULL);
  st->ncycle++;
  st->time += st->dt;
}
int orc_amr_sparse_regrid(struct OrcAmr *a) {
  const int changed = amr_remesh(a);
  OrcSparse *st = a->sp;
  if (changed) st->allowed_dt = sparse_estimate_timestep(st);
  if (st->dt < 0.1 * DBL_MAX) st->dt *= 2.0;
  st->dt = fmin(st->dt, st->allowed_dt);
  st->allowed_dt = DBL_MAX;
  return changed;
}
struct OrcSparse *orc_amr_sparse_state(struct OrcAmr *a) { return a->sp; }

void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int orc_get_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------------------ */
/* SetBounds through a LogicalCoordinateTransformation (trees of a forest that meet with
 * different orientations): boundary_communication.cpp:282-308 with
 * LogicalCoordinateTransformation::InverseTransform(std::array<int,3>)
 * (mesh/forest/logical_coordinate_transformation.hpp:90-99):
 *   for every box cell (i, j, k) in buffer order (i fastest, then j, k, component slowest)
 *     out[dir] = dir_flip[dir] ? ncell - 1 - in[|dir_connection[dir]|] : in[...]
 *     var(c, out[2], out[1], out[0]) = fac * buf[m]
 * var: one block's array [ncomp][nk][nj][ni] given by its strides; s / n: box start / extent.
 * Stand-alone form of what orc_unpack does for a transformed neighbour (that one is pinned
 * against the reference's forest dumps, tests/golden/forest_*.npz); this entry serves the
 * kernel-level tests of pb2_unpack over all 48 transformations and is checked by its properties
 * (identity = plain unpack, a flip applied twice, an axis permutation and its inverse). */
void orc_unpack_box_transformed(double *var, int64_t stride_j, int64_t stride_k, int64_t stride_c,
                                const int s[3], const int n[3], int ncomp, const double *buf,
                                const int dir_connection[3], const int dir_flip[3], int ncell,
                                double fac) {
  int64_t m = 0;
  for (int c = 0; c < ncomp; ++c)
    for (int k = 0; k < n[2]; ++k)
      for (int j = 0; j < n[1]; ++j)
        for (int i = 0; i < n[0]; ++i, ++m) {
          const int in[3] = {s[0] + i, s[1] + j, s[2] + k};
          int out[3];
          for (int dir = 0; dir < 3; ++dir) {
            const int indir = abs(dir_connection[dir]);
            out[dir] = dir_flip[dir] ? ncell - 1 - in[indir] : in[indir];
          }
          var[(int64_t)c * stride_c + (int64_t)out[2] * stride_k + (int64_t)out[1] * stride_j +
              out[0]] = fac * buf[m];
        }
}
