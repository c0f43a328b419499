// Tile plan of the fused Laplace-Beltrami assembly (gg_lb_tile.cuh): the numeric phase of
// assemble_matrix! / assemble_matrix_and_vector (src/ODEs/ODEs.jl:36-39; form test/Laplacian/LaplaceBeltrami.jl:47) without
// the element-matrix round trip through HBM.
//
// One CTA owns a TILE of up to 128 cells (a TX x TY block of one panel, found from the cells' chart intervals, so it works for
// any cell order and for caller-supplied meshes).  CSR rows whose cells all lie in one tile are COMPLETE: the CTA sums them
// from the element matrices it holds in shared memory and writes their values once.  The other rows are grouped by the set
// of tiles around them (two tiles across a tile edge, up to four around a tile corner): every tile stores the element matrices
// of its border cells to a small global buffer and bumps the counters of its groups; the CTA that brings a counter to its
// final value sums the group's rows (reading the border data, which is still in L2).  Every slot is summed in ascending cell
// order whoever does it: deterministic, atomic-free values, bit for bit those of the two-kernel path (k_gather_chunks).
//
// What a CTA needs to know about its rows (which local cell / entry feeds which position of which row) is the same for all
// tiles with the same local connectivity, so these PATTERNS are stored once per distinct tile (interior tiles of a panel
// share one) and stay L2-resident; nothing in them depends on how the DOFs are numbered, only the row pointers do.
#pragma once
#include <memory>
#include <stdint.h>
#include "gg_common.cuh"

namespace gg {

constexpr int TILE_CELLS = 128;     // cells (threads) per tile
constexpr int TILE_ROWCAP = 1024;   // rows per pattern the kernel can stage (a 16 x 8 Q2 tile has 465 complete rows)
constexpr uint16_t TILE_NONE = 0xFFFFu;

struct TilePat { int32_t row0, nrows, ent0, nent, nborder, pad; };
struct GroupPat { int32_t row0, nrows, ent0, nent; };
// a group of rows shared by up to four tiles: pattern, arrivals needed, first row in grow_base / grow_dof, and per tile slot the
// offsets of that tile's border buffers and its border-cell count
struct GroupInst { int32_t pat, need, row0, pad; int32_t eoff[4], voff[4], nb[4]; };

// device view handed to the kernel
struct TileDev {
  int32_t n_super;
  const int32_t *inst_ptr, *inst_tile;     // supertile -> tiles that share its geometry (1 each without geometry classes)
  const int32_t* tile_cells;               // [n_tiles][TILE_CELLS] cell ids, ascending, -1 padded
  const int32_t *tile_pat, *tile_eoff, *tile_voff;
  const TilePat* pats;
  const int32_t *tile_row0, *trow_base, *trow_dof;   // per tile: first CSR slot and DOF id of its complete rows (numbering-dependent)
  const ushort4* rowvsrc;                  // complete row k of a pattern: element-vector sources (local dof << 7) | local cell, TILE_NONE padded
  const uint32_t* emeta;                   // entry: (row index in the pattern << 8) | position in the row
  const ushort4* esrc;                     // its sources: plane * TILE_CELLS + local cell, ascending cell, TILE_NONE padded
  const uint8_t* bidx;                     // [n_pat][TILE_CELLS] border index of a local cell (0xFF: interior cell)
  const int32_t *tg_ptr, *tg_idx;          // tile -> groups it contributes to
  const GroupInst* groups;
  int32_t* gcount;                         // arrivals per group (zero between launches)
  const GroupPat* gpats;
  const int32_t *grow_base, *grow_dof;     // per group: first CSR slot and DOF id of its rows
  const ushort4* g_rowvsrc;                // (tile slot << 11) | (border index << 4) | local dof
  const uint32_t* g_emeta;
  const ushort4* g_esrc;                   // (tile slot << 13) | (border index << 6) | plane
  double *ebuf_b, *vbuf_b;                 // border element matrices [tile][plane][border cell] / vectors [tile][dof][border cell]
};

struct TilePlan {
  bool classes = false;
  int M = 0, NP = 0;
  int64_t n_tiles = 0, n_super = 0, n_pat = 0, n_groups = 0, n_gpat = 0;
  int max_rows = 0;
  DevBuf<int32_t> inst_ptr, inst_tile, tile_cells, tile_pat, tile_eoff, tile_voff, tg_ptr, tg_idx, gcount;
  DevBuf<int32_t> tile_row0, trow_base, trow_dof, grow_base, grow_dof;
  DevBuf<TilePat> pats;
  DevBuf<GroupPat> gpats;
  DevBuf<GroupInst> groups;
  DevBuf<ushort4> rowvsrc, esrc, g_rowvsrc, g_esrc;
  DevBuf<uint32_t> emeta, g_emeta;
  DevBuf<uint8_t> bidx;
  DevBuf<double> ebuf_b, vbuf_b;
  // statistics of the build (for DESIGN.md / bench)
  int64_t complete_slots = 0, group_slots = 0, border_cells = 0, pattern_bytes = 0;
  TileDev dev() const;
};

// nullptr when the matrix / mesh does not fit the tile scheme (the caller falls back to the two-kernel path)
std::shared_ptr<TilePlan> build_tile_plan(gg_csr* A, bool classes);

}  // namespace gg
