// Statically condensed element-by-element PCG (gg_ebe.cu): internal interface used by the RK steppers.
#pragma once
#include "gg_common.cuh"

namespace gg {

struct EbeSolver {
  gg_ctx* ctx = nullptr;
  gg_space* sp = nullptr;
  int MB = 0, MI = 0;        // boundary (shared) / interior (cell-private) local DOFs
  int64_t nb = 0, nc = 0;    // boundary free DOFs (numbered first), cells
  DevBuf<double> S, T, W;    // per cell (or geometry class), entry-major: packed S_c, T_c = M_ii^-1 M_if, packed M_ii^-1
  // geometry classes (constant operators on meshes whose cells repeat chart boxes, e.g. the 6 congruent panels of the cubed
  // sphere): S, T, W are stored once per class for the UNSIGNED reference basis; cls[c] = class of cell c, ns = classes
  DevBuf<int32_t> cls;
  DevBuf<int32_t> ccell;     // [ns][6] cells of every class, -1 padded (tiled condensation passes of the solve)
  size_t tile_smem = 0;      // dynamic shared memory of the tiled passes (0: cell loops)
  DevBuf<int32_t> dofT;      // [MB][nc] boundary DOF ids, (dof << 1) | (sign < 0), -1: none (EbeArgs::dofT)
  int64_t ns = 0;            // stride of S, T, W (= nc when not compressed)
  bool compressed = false;
  DevBuf<int32_t> bstart, rowoff, blk_off;
  DevBuf<uint8_t> bsize, blk_size;
  DevBuf<double> binv;
  int64_t n_blocks = 0;
  int bc_n = 0, bc_start[3] = {0, 0, 0}, bc_size[3] = {1, 1, 1}, bc_off[3] = {0, 0, 0}, bc_b0[3] = {0, 0, 0}, cblk = 0;  // runs of equally sized blocks (0: use the arrays)
  DevBuf<int2> pair;         // [nb] first two element sources of every boundary row, -1 padded (EbeArgs::pair)
  int tail = 0;              // some row has more than two sources: the kernel also walks vptr / vsrc beyond the pair
  DevBuf<double> r, r2, z, p, Ap, ev, ev2, partials;
  DevBuf<double> stats;      // [0] iterations of the last solve, [1] its relative residual, [2..5] phase timers (ns) + iterations
  double rtol = 1e-12;
  int max_iter = 1000;
  int grid = 0;
};

bool ebe_supported(const gg_space* sp);
EbeSolver* ebe_create(gg_space* sp, double rtol, int max_iter);
void ebe_set_matrix(EbeSolver* s, const double* ebuf_full);
void ebe_set_scalar_mass(EbeSolver* s, int degree, const double* coef);
// warm: use the content of x as the initial guess (convergence is still measured against the right-hand side norm)
// be != nullptr: the right-hand side in element form (entry-major element vectors [mloc][nc], what gather_vector would sum into
// b) -- the set-up passes of the solve gather it themselves, b is not read
void ebe_solve_async(EbeSolver* s, const double* b, double scale_b, double* x, bool warm = false, const double* be = nullptr);
void ebe_destroy(EbeSolver* s);

}  // namespace gg
