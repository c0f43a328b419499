// Tile plan of the fused Laplace-Beltrami assembly (gg_lb_tile.cuh): the numeric phase of
// assemble_matrix! / assemble_matrix_and_vector (src/ODEs/ODEs.jl:36-39; form test/Laplacian/LaplaceBeltrami.jl:47) without
// the element-matrix round trip through HBM.
//
// One CTA owns a TILE of up to 128 cells (a TX x TY block of one panel, found from the cells' chart intervals, so it works for
// any cell order and for caller-supplied meshes).  CSR rows whose cells all lie in one tile are COMPLETE: their values are
// accumulated in shared memory straight from the registers that hold the element matrices and written once (k_lb_tile).
// The accumulation is conflict-free without atomics: the cells of a tile are 4-coloured by the parity of their chart
// intervals (cells of one colour share no DOF), warp w holds the cells of colour w, and the four warps add their
// contributions one after the other -- every slot is summed in colour order, deterministically.  The other rows -- the DOFs
// on tile edges and corners, about a tenth of the values -- are grouped by the set of tiles around them: every tile stores
// the element matrices of its border cells to a (sparse) global buffer and a second, small kernel (k_lb_tile_groups) sums
// these rows from it, in ascending cell order, while the data is still in L2.
//
// What the kernels need to know about a row -- which cells around it feed which position -- is a ROW TYPE: the list of its
// columns in CSR order, each with up to four (cell slot, packed element entry) sources.  On a structured panel a handful of
// row types covers every row (vertex, the two edge orientations, cell interior, and their variants along panel borders), so
// the table stays in L1; the group kernel works from them.  A tile PATTERN holds, for every (local cell, local DOF pair), the
// shared-memory slot its element-matrix entry is added to; tiles with the same local connectivity share one pattern (the
// interior tiles of a panel all do), so the patterns stay in L1 / L2.  Nothing in the types or patterns depends on how the DOFs
// are numbered beyond the order of the columns inside a row; the numbering enters through the per-tile row bases only.
#pragma once
#include <memory>
#include <stdint.h>
#include "gg_common.cuh"

namespace gg {

constexpr int TILE_CELLS = 128;     // cells (threads) per tile
constexpr int TILE_ROWCAP = 640;    // rows of a tile the kernel stages in shared memory (a 16 x 8 Q2 tile has 465 complete rows)
constexpr int TILE_MAXLEN = 32;     // longest CSR row a row type can describe (Q2: 25)
constexpr int TILE_MAXCLS = 8;      // row-length classes per pattern

// one row type: len columns; column j sums the sources src[j][0..3] = (cell slot << 6) | packed plane, 0xFF = none, in that
// order (ascending cell); vsrc = element-vector sources (cell slot << 4) | local dof
struct RowType {
  uint8_t src[TILE_MAXLEN][4];
  uint8_t vsrc[4];
  int32_t len;
};
// one row shared by several tiles: first CSR slot, DOF id, row type, and for the type's four cell slots the column of the cell in
// the border buffers (tile * TILE_CELLS + local cell); the rows are sorted by length, a class is a run of equal length
struct __align__(8) GroupRow { int32_t base, dof; int32_t type; int32_t col[4]; int32_t pad; };
// rows of a pattern are sorted by (row length, DOF id); a class is a run of equal length
struct RowClass { int32_t len, first, count, slot0; };   // slot0: first shared-memory slot of the class (tile patterns)
struct TilePat { int32_t row0, nrows, ncls, nslots; RowClass cls[TILE_MAXCLS]; };
struct GroupClasses { int32_t ncls, pad; RowClass cls[TILE_MAXCLS]; int64_t ent0[TILE_MAXCLS + 1]; };   // ent0: first entry (thread) of a class

// Local geometry of Q_p on a quadrilateral (lexicographic DOF I = i1 + (p+1) i2): which element entries can have more than one
// contributing cell.  (I, J) can only be shared if both DOFs lie on one edge of the cell; only a corner's diagonal entry
// (I == J, a vertex) can have more than two cells.
__host__ __device__ constexpr bool tile_on_edge(int p, int I, int e) {   // e: 0 left, 1 right, 2 bottom, 3 top
  return e == 0 ? I % (p + 1) == 0 : e == 1 ? I % (p + 1) == p : e == 2 ? I / (p + 1) == 0 : I / (p + 1) == p;
}
__host__ __device__ constexpr bool tile_shared_pair(int p, int I, int J) {
  return (tile_on_edge(p, I, 0) && tile_on_edge(p, J, 0)) || (tile_on_edge(p, I, 1) && tile_on_edge(p, J, 1)) ||
         (tile_on_edge(p, I, 2) && tile_on_edge(p, J, 2)) || (tile_on_edge(p, I, 3) && tile_on_edge(p, J, 3));
}
__host__ __device__ constexpr bool tile_shared_dof(int p, int I) { return tile_shared_pair(p, I, I); }
__host__ __device__ constexpr bool tile_corner_dof(int p, int I) {
  return (I % (p + 1) == 0 || I % (p + 1) == p) && (I / (p + 1) == 0 || I / (p + 1) == p);
}

// device view handed to the kernels
struct TileDev {
  int32_t n_super, pad0;
  const int32_t *inst_ptr, *inst_tile;     // supertile -> tiles that share its geometry (1 each without geometry classes)
  const int32_t* tile_cells;               // [n_tiles][TILE_CELLS] cell ids, ascending, -1 padded
  const int32_t *tile_pat, *tile_eoff, *tile_voff;
  const int32_t *tile_row0, *trow_base, *trow_dof;   // per tile: first CSR slot and DOF id of its complete rows (numbering-dependent)
  const TilePat* pats;
  // [n_pat][PW][TILE_CELLS], PW = (M*M+1)/2: shared-memory slots of the element entries (I, J) (lexicographic local DOFs, row-major)
  // of a local cell, two 16-bit slots per word; [n_pat][VW][TILE_CELLS], VW = (M+1)/2: rows of its local DOFs.  Slots / rows past
  // the pattern's own counts are dummies (rows that are not complete in this tile, Dirichlet DOFs)
  const uint32_t *pair_idx, *vec_idx;
  // per (tile, local cell), in tile order: chart-interval ids and extents of the cell, and its DOF ids [tile][I][local cell] in
  // lexicographic local order -- what a thread reads first, with one level of indirection less than through the cell id
  const int32_t *tc_ix, *tc_iy, *tc_dofs;
  const double *tc_hx, *tc_hy;
  const uint8_t* bidx;                     // [n_pat][TILE_CELLS] != 0xFF: the local cell is a border cell (its element matrix goes to ebuf_b)
  const RowType* types;
  const GroupRow* grows;                   // rows on tile edges / corners (k_lb_tile_groups)
  int64_t n_grows;
  double *ebuf_b, *vbuf_b;                 // border element matrices [tile * 128 + local cell][plane] / vectors [...][dof] (only border cells are touched)
};

struct TilePlan {
  bool classes = false;
  int M = 0, NP = 0;
  int64_t n_tiles = 0, n_super = 0, n_pat = 0, n_grows = 0, n_types = 0;
  int max_rows = 0, max_group_rows = 0, max_slots = 0;
  DevBuf<int32_t> inst_ptr, inst_tile, tile_cells, tile_pat, tile_eoff, tile_voff;
  DevBuf<int32_t> tile_row0, trow_base, trow_dof;
  DevBuf<TilePat> pats;
  GroupClasses gcls;
  DevBuf<uint32_t> pair_idx, vec_idx;
  DevBuf<int32_t> tc_ix, tc_iy, tc_dofs;
  DevBuf<double> tc_hx, tc_hy;
  DevBuf<GroupRow> grows;
  DevBuf<RowType> types;
  DevBuf<uint8_t> bidx;
  DevBuf<double> ebuf_b, vbuf_b;
  // statistics of the build (for DESIGN.md / bench)
  int64_t complete_slots = 0, group_slots = 0, border_cells = 0, pattern_bytes = 0;
  TileDev dev() const;
};

// nullptr when the matrix / mesh does not fit the tile scheme (the caller falls back to the two-kernel path)
std::shared_ptr<TilePlan> build_tile_plan(gg_csr* A, bool classes);

}  // namespace gg
