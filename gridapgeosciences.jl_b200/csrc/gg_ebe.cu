// Statically condensed, element-by-element, persistent single-launch PCG for the mass solves of the RK stages.
//
// Stands in for `LUSolver()` / `CGSolver(JacobiLinearSolver(); rtol=1e-12)` at the
// `Algebra.solve!(x, ::LinearSolver, op, cache)` seam (src/ODEs/StageOperators.jl:55-88, src/ODEs/DAE.jl:241-274) on the
// symmetric positive definite mass matrices of the transient drivers (RT_k mass: test/Transient/TransientWaveEquation.jl:63;
// potential-vorticity mass int q p w sqrtg: test/Transient/TransientShallowWater.jl:122).
//
// B200 design.  A CG iteration on the assembled RT_2 mass matrix of a 256x256-per-panel mesh streams 3.4 GB of CSR
// (0.5 ms of HBM time) plus 0.5 GB of block-Jacobi inverses; re-integrating the form every iteration instead costs
// 0.37 ms of FP64 pipe (measured, gg_mf.cu).  Both lose to linear algebra done once per matrix:
//   * the interior DOFs of a cell (12 of the 24 RT_2 moments, 4 of the 16 Q_3 nodes) are private to the cell, so they
//     are eliminated exactly, cell by cell:  S_c = M_ff - M_fi M_ii^-1 M_if.  CG then runs on the facet / vertex / edge
//     DOFs only (3x fewer unknowns for RT_2) and needs 19 instead of 34 iterations for rtol 1e-12;
//   * the condensed operator is never assembled: every iteration applies the packed symmetric S_c (78 doubles per cell for
//     RT_2 / Q_3, stored entry-major so that consecutive threads read consecutive addresses) and sums the element vectors
//     per DOF in fixed cell order (atomic-free, deterministic);
//   * one cooperative launch per solve, 3 grid barriers per iteration (the direction update p = z + beta p is folded into
//     the operator phase), dot products through per-CTA partials that every CTA re-sums in the same order;
//   * block-Jacobi preconditioner over the DOFs of one facet / edge, built from the condensed element matrices.
// The eliminated DOFs are recovered cell by cell after the iteration: x_i = M_ii^-1 b_i - (M_ii^-1 M_if) x_f.
#include <cooperative_groups.h>
#include <cstdlib>
#include <unordered_map>
#include "gg_common.cuh"
#include "gg_ebe.cuh"
#include "gg_refel.cuh"
#include "gg_dist.cuh"
#include "gg_dist_host.cuh"

namespace cg = cooperative_groups;

namespace gg {

static inline unsigned nblk(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

__host__ __device__ constexpr int ebe_sym(int MB, int i, int j) {  // packed upper triangle, i <= j
  return i * MB - (i * (i - 1)) / 2 + (j - i);
}

// ---- condensation: one thread per cell, element matrix entry-major full layout ebuf[(i*m+j)*nc + c] ------------------
template <int MB, int MI>
__global__ void __launch_bounds__(64)
k_ebe_condense(int64_t nc, const double* __restrict__ ebuf, double* __restrict__ S, double* __restrict__ T,
               double* __restrict__ W) {
  constexpr int M = MB + MI;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  auto E = [&](int i, int j) { return ebuf[(int64_t)(i * M + j) * nc + c]; };
  if (MI == 0) {
#pragma unroll
    for (int i = 0; i < MB; ++i)
#pragma unroll
      for (int j = i; j < MB; ++j) S[(int64_t)ebe_sym(MB, i, j) * nc + c] = E(i, j);
    return;
  }
  constexpr int MIa = MI > 0 ? MI : 1;
  // W = M_ii^-1 (Gauss-Jordan; SPD block)
  double a[MIa][MIa], w[MIa][MIa];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < MI; ++j) { a[i][j] = E(MB + i, MB + j); w[i][j] = i == j ? 1.0 : 0.0; }
#pragma unroll
  for (int k = 0; k < MI; ++k) {
    const double d = 1.0 / a[k][k];
#pragma unroll
    for (int j = 0; j < MI; ++j) { a[k][j] *= d; w[k][j] *= d; }
#pragma unroll
    for (int i = 0; i < MI; ++i) {
      if (i == k) continue;
      const double f = a[i][k];
#pragma unroll
      for (int j = 0; j < MI; ++j) { a[i][j] = fma(-f, a[k][j], a[i][j]); w[i][j] = fma(-f, w[k][j], w[i][j]); }
    }
  }
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = i; j < MI; ++j) W[(int64_t)ebe_sym(MI, i, j) * nc + c] = 0.5 * (w[i][j] + w[j][i]);
  // T = W M_if, column by column; S = M_ff - M_fi T (upper triangle)
  for (int j = 0; j < MB; ++j) {
    double v[MIa], t[MIa];
#pragma unroll
    for (int s = 0; s < MI; ++s) v[s] = E(MB + s, j);
#pragma unroll
    for (int r = 0; r < MI; ++r) {
      double acc = 0.0;
#pragma unroll
      for (int s = 0; s < MI; ++s) acc = fma(w[r][s], v[s], acc);
      t[r] = acc;
      T[(int64_t)(r * MB + j) * nc + c] = acc;
    }
    for (int i = 0; i <= j; ++i) {
      double acc = E(i, j);
#pragma unroll
      for (int r = 0; r < MI; ++r) acc = fma(-E(i, MB + r), t[r], acc);
      S[(int64_t)ebe_sym(MB, i, j) * nc + c] = acc;
    }
  }
}

// ---- fused element matrix + condensation for the scalar mass with a coefficient ----------------------------------------
// M_IJ = int coef phi_I phi_J sqrtg on Q_P (the state-dependent block of jac_y, test/Transient/TransientShallowWater.jl:122,
// re-evaluated before every diagnostic solve).  One thread per cell, sum-factorised: with I = (i1,i2), J = (j1,j2) the entry
// only depends on the unordered pairs {i1,j1}, {i2,j2}:
//   B[q1][s2] = sum_q2 c(q1,q2) NN[s2][q2],   A[s1][s2] = sum_q1 NN[s1][q1] B[q1][s2],   NN[s][q] = N_i(x_q) N_j(x_q)
// (NS^2 = 100 accumulators for Q_3 instead of 136 packed entries x 49 points), then the cell-private interior nodes are
// eliminated in registers and only the condensed S_c, T_c and M_ii^-1 are written (entry-major, coalesced).
__constant__ double c_ebe_nn[MAX_NQ * 10];  // [q][s], s = packed (i <= j) pair of 1-D basis functions
__constant__ double c_ebe_w[MAX_NQ];

__host__ __device__ constexpr int q_loc2lex(int P, int a) {
  // Gridap local node order of Q_P on the square -> lexicographic ix + (P+1) iy (same as quad_local_to_lex)
  const int nb = P + 1, ne = P - 1;
  if (a == 0) return 0;
  if (a == 1) return P;
  if (a == 2) return nb * P;
  if (a == 3) return P + nb * P;
  int e = a - 4;
  if (e < ne) return e + 1;
  e -= ne;
  if (e < ne) return e + 1 + nb * P;
  e -= ne;
  if (e < ne) return nb * (e + 1);
  e -= ne;
  if (e < ne) return P + nb * (e + 1);
  e -= ne;
  return (e % ne + 1) + nb * (e / ne + 1);
}

template <int P>
__global__ void __launch_bounds__(64)
k_ebe_scalar_mass(int64_t nc, int nq, CellGeo g, const double* __restrict__ coef, double* __restrict__ S,
                  double* __restrict__ T, double* __restrict__ W) {
  constexpr int NB = P + 1, NS = NB * (NB + 1) / 2, MB = 4 * P, MI = (P - 1) * (P - 1);
  constexpr int MIa = MI > 0 ? MI : 1;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const double HX = g.hx[c], HY = g.hy[c];
  const double* __restrict__ ta = g.tx + (int64_t)g.ix[c] * nq;
  const double* __restrict__ tb = g.ty + (int64_t)g.iy[c] * nq;
  const double* __restrict__ cf = coef + c * (int64_t)(nq * nq);
  const double scale = g.radius * g.radius * HX * HY;
  double acc[NS * NS];
#pragma unroll
  for (int k = 0; k < NS * NS; ++k) acc[k] = 0.0;
#pragma unroll 1
  for (int q1 = 0; q1 < nq; ++q1) {
    const double t = ta[q1], sa = fma(t, t, 1.0), w1 = c_ebe_w[q1] * scale * sa;
    double B[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) B[k] = 0.0;
#pragma unroll 1
    for (int q2 = 0; q2 < nq; ++q2) {
      const double b = tb[q2], sb = fma(b, b, 1.0);
      double ri;
      {
        const double x = sa + b * b;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(ri) : "d"(x));
        const double e = fma(x, -(ri * ri), 1.0);
        ri = fma(fma(e, 0.375, 0.5), ri * e, ri);
      }
      const double cq = w1 * c_ebe_w[q2] * sb * ri * ri * ri * cf[q1 + nq * q2];
#pragma unroll
      for (int k = 0; k < NS; ++k) B[k] = fma(cq, c_ebe_nn[q2 * NS + k], B[k]);
    }
#pragma unroll
    for (int s1 = 0; s1 < NS; ++s1) {
      const double n1 = c_ebe_nn[q1 * NS + s1];
#pragma unroll
      for (int s2 = 0; s2 < NS; ++s2) acc[s1 * NS + s2] = fma(n1, B[s2], acc[s1 * NS + s2]);
    }
  }
  // element matrix entry of local DOFs (I, J)
  auto E = [&](int I, int J) -> double {
    const int li = q_loc2lex(P, I), lj = q_loc2lex(P, J);
    const int i1 = li % NB, i2 = li / NB, j1 = lj % NB, j2 = lj / NB;
    const int a1 = i1 < j1 ? i1 : j1, b1 = i1 < j1 ? j1 : i1, a2 = i2 < j2 ? i2 : j2, b2 = i2 < j2 ? j2 : i2;
    return acc[ebe_sym(NB, a1, b1) * NS + ebe_sym(NB, a2, b2)];
  };
  if (MI == 0) {
#pragma unroll
    for (int i = 0; i < MB; ++i)
#pragma unroll
      for (int j = i; j < MB; ++j) S[(int64_t)ebe_sym(MB, i, j) * nc + c] = E(i, j);
    return;
  }
  double a[MIa][MIa], w[MIa][MIa];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < MI; ++j) { a[i][j] = E(MB + i, MB + j); w[i][j] = i == j ? 1.0 : 0.0; }
#pragma unroll
  for (int k = 0; k < MI; ++k) {
    const double d = 1.0 / a[k][k];
#pragma unroll
    for (int j = 0; j < MI; ++j) { a[k][j] *= d; w[k][j] *= d; }
#pragma unroll
    for (int i = 0; i < MI; ++i) {
      if (i == k) continue;
      const double f = a[i][k];
#pragma unroll
      for (int j = 0; j < MI; ++j) { a[i][j] = fma(-f, a[k][j], a[i][j]); w[i][j] = fma(-f, w[k][j], w[i][j]); }
    }
  }
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = i; j < MI; ++j) W[(int64_t)ebe_sym(MI, i, j) * nc + c] = 0.5 * (w[i][j] + w[j][i]);
#pragma unroll
  for (int j = 0; j < MB; ++j) {
    double tcol[MIa];
#pragma unroll
    for (int r = 0; r < MI; ++r) {
      double v = 0.0;
#pragma unroll
      for (int q = 0; q < MI; ++q) v = fma(w[r][q], E(MB + q, j), v);
      tcol[r] = v;
      T[(int64_t)(r * MB + j) * nc + c] = v;
    }
#pragma unroll
    for (int i = 0; i <= j; ++i) {
      double v = E(i, j);
#pragma unroll
      for (int r = 0; r < MI; ++r) v = fma(-E(i, MB + r), tcol[r], v);
      S[(int64_t)ebe_sym(MB, i, j) * nc + c] = v;
    }
  }
}

// ---- block-Jacobi of the (never assembled) condensed matrix -----------------------------------------------------------
// one thread per boundary DOF row: bmat[rowoff[d] + b] = sum over the cells around d of S_c[local(d)][local(bstart + b)]
__global__ void k_ebe_block_rows(int64_t nb, int64_t nc, int mloc, int MB, const int32_t* __restrict__ vptr,
                                 const int32_t* __restrict__ vsrc, const int32_t* __restrict__ cell_dofs,
                                 const int32_t* __restrict__ bstart, const uint8_t* __restrict__ bsize,
                                 const int32_t* __restrict__ rowoff, const double* __restrict__ S, double* __restrict__ bmat) {
  const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= nb) return;
  const int s = bstart[d], m = bsize[d];
  double* out = bmat + rowoff[d];
  for (int b = 0; b < m; ++b) out[b] = 0.0;
  for (int32_t k = vptr[d]; k < vptr[d + 1]; ++k) {
    const int32_t src = vsrc[k];
    const int ia = (int)(src / nc);
    const int64_t c = src - (int64_t)ia * nc;
    for (int b = 0; b < m; ++b) {
      int ib = -1;
      for (int l = 0; l < MB; ++l)
        if (cell_dofs[c * mloc + l] == s + b) { ib = l; break; }
      if (ib < 0) continue;
      const int lo = ia < ib ? ia : ib, hi = ia < ib ? ib : ia;
      out[b] += S[(int64_t)(lo * MB - (lo * (lo - 1)) / 2 + (hi - lo)) * nc + c];
    }
  }
}

constexpr int EBE_MAX_BLOCK = 8;
__global__ void k_ebe_block_inverse(int64_t n_blocks, const int32_t* __restrict__ blk_off, const uint8_t* __restrict__ blk_size,
                                    double* __restrict__ bmat) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  const int m = blk_size[b];
  double* A = bmat + blk_off[b];
  double a[EBE_MAX_BLOCK][EBE_MAX_BLOCK], v[EBE_MAX_BLOCK][EBE_MAX_BLOCK];
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) { a[i][j] = A[i * m + j]; v[i][j] = i == j ? 1.0 : 0.0; }
  for (int i = 0; i < m; ++i)
    if (a[i][i] == 0.0) a[i][i] = 1.0;  // DOF without cells (sub-mesh)
  for (int k = 0; k < m; ++k) {
    const double d = 1.0 / a[k][k];
    for (int j = 0; j < m; ++j) { a[k][j] *= d; v[k][j] *= d; }
    for (int i = 0; i < m; ++i) {
      if (i == k) continue;
      const double f = a[i][k];
      for (int j = 0; j < m; ++j) { a[i][j] = fma(-f, a[k][j], a[i][j]); v[i][j] = fma(-f, v[k][j], v[i][j]); }
    }
  }
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) A[i * m + j] = v[i][j];
}

// transposed, sign-tagged boundary DOF table of the solver kernel (EbeArgs::dofT)
__global__ void k_ebe_dof_table(int64_t nc, int mloc, int MB, const int32_t* __restrict__ cell_dofs, const int8_t* __restrict__ sg,
                                int32_t* __restrict__ dofT) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * MB) return;
  const int64_t c = t % nc;
  const int j = (int)(t / nc);
  const int32_t d = cell_dofs[c * mloc + j];
  dofT[t] = d < 0 ? -1 : ((d << 1) | ((sg && sg[c * mloc + j] < 0) ? 1 : 0));
}

// first two element sources of every boundary row (EbeArgs::pair)
__global__ void k_ebe_source_pairs(int64_t n, const int32_t* __restrict__ vptr, const int32_t* __restrict__ vsrc, int2* __restrict__ pair) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t lo = vptr[i], hi = vptr[i + 1];
  pair[i] = make_int2(lo < hi ? vsrc[lo] : -1, lo + 1 < hi ? vsrc[lo + 1] : -1);
}

// ---- the solve -------------------------------------------------------------------------------------------------------
struct EbeArgs {
  int64_t nb, nc;
  int mloc;
  const int32_t* cell_dofs;
  // boundary DOF ids of the cells, transposed [MB][nc] so that a warp's 32 cells load 128 consecutive bytes per local DOF (the
  // cell-major table costs 24 L1 wavefronts per load, and phase A is L1-wavefront bound); entry = (dof << 1) | (sign < 0), -1: none
  const int32_t* dofT;
  const double *S, *T, *W;
  int64_t ns;             // stride of S, T, W: cells, or geometry classes when cls != nullptr
  const int32_t* cls;     // cell -> geometry class (nullptr: one operator per cell, DOF signs folded in)
  // class -> its (at most EBE_TILE_REP) cells, -1 padded; when set, the condensation passes of the solve walk the classes in tiles
  // of EBE_TILE_CLS: a CTA stages T (and W) of a tile in shared memory once and serves every congruent cell from there
  const int32_t* ccell;
  const int8_t* sg;       // [nc][mloc] DOF signs applied around the class operators (nullptr: none)
  double* ev;   // [MB][nc]
  double* ev2;  // [MB][nc] second element-vector buffer (S x of the warm start, next to the condensed right-hand side)
  const int32_t *vptr, *vsrc;
  // first two sources of every boundary row, -1 padded (facet DOFs never have more): one 8-byte load instead of the dependent
  // vptr -> vsrc chain in the latency-bound gather phases; the row pointers are only read when some row has a tail (tail != 0)
  const int2* pair;
  int tail;
  const int32_t *bstart, *rowoff;
  const uint8_t* bsize;
  // block classes of the builtin numberings (runs of equally sized blocks: RT_k one class, CG_p vertices + edges): block start,
  // size and offset into binv by arithmetic instead of 9 bytes of index arrays per row; n == 0: read the arrays
  struct { int n; int start[3], size[3], off[3], b0[3]; } bc;   // b0: index of the first block of the run
  int64_t n_blocks;
  int bc_rows;            // the row loops also use the arithmetic lookup (measured: 1 us slower than the arrays; GG_EBE_BLOCK_CLASSES=1)
  int row_setup;          // the two row-wise set-up passes instead of the merged block-wise one (chosen by mesh size)
  int cblk;               // phase C walks the preconditioner blocks (one thread per block) instead of the rows
  const double* binv;
  const double* b;
  // right-hand side in element form instead (b == nullptr): the entry-major element vectors [mloc][nc] that gather_vector would
  // sum into b.  Boundary rows sum their sources (the same `pair` / vsrc indices as ev), interior DOFs have one source each.
  const double* be;
  double scale_b;
  double *x, *ra, *rb, *z, *p, *Ap, *partials;
  double rtol;
  int max_iter;
  int warm;  // start from the content of x instead of 0
  double* stats;
  DistDev dist;  // world == 1: single GPU
};

constexpr int EBE_TILE_CLS = 32, EBE_TILE_REP = 6;
extern __shared__ double ebe_tile[];   // [entries][EBE_TILE_CLS]

// stage `n_ent` entry planes of a class operator (src[ent * ns + k]) for the classes [k0, k0 + EBE_TILE_CLS) at dst
__device__ __forceinline__ void ebe_stage_tile(double* dst, const double* __restrict__ src, int n_ent, int64_t ns, int64_t k0) {
  for (int e = threadIdx.x; e < n_ent * EBE_TILE_CLS; e += blockDim.x) {
    const int ent = e / EBE_TILE_CLS, kk = e - ent * EBE_TILE_CLS;
    dst[e] = k0 + kk < ns ? src[(int64_t)ent * ns + k0 + kk] : 0.0;
  }
}

__device__ __forceinline__ unsigned long long ebe_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ double ebe_block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  }
  return v;  // valid in thread 0
}

__device__ __forceinline__ double ebe_grid_total(const double* partials, int nblocks, double* sh) {
  double v = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) v += partials[i];
  v = ebe_block_sum(v, sh);
  __shared__ double bc;
  if (threadIdx.x == 0) bc = v;
  __syncthreads();
  return bc;
}

// sum of the element contributions of DOF row i (at most 2 sources on the fast path, more in the tail loop)
__device__ __forceinline__ double ebe_gather_row(const EbeArgs& a, int64_t i, const double* __restrict__ ev) {
  const int32_t lo = a.vptr[i], hi = a.vptr[i + 1];
  double acc = lo < hi ? ev[a.vsrc[lo]] : 0.0;
  for (int32_t k = lo + 1; k < hi; ++k) acc += ev[a.vsrc[k]];
  return acc;
}

__device__ __forceinline__ void ebe_block_of(const EbeArgs& a, int64_t i, int& sb, int& mb, int& ro) {
  if (a.bc.n == 0 || !a.bc_rows) { sb = a.bstart[i]; mb = a.bsize[i]; ro = a.rowoff[i]; return; }
  int k = 0;
  if (a.bc.n > 1 && i >= a.bc.start[1]) k = 1;
  if (a.bc.n > 2 && i >= a.bc.start[2]) k = 2;
  const int m = a.bc.size[k], d = (int)i - a.bc.start[k];
  const int g = m == 3 ? d / 3 : (m == 2 ? d >> 1 : (m == 1 ? d : d / m));
  sb = a.bc.start[k] + g * m; mb = m; ro = a.bc.off[k] + (g * m + (d - g * m)) * m;
}

// first two sources of row i (the common case) and the range of the remaining ones
__device__ __forceinline__ void ebe_row_sources(const EbeArgs& a, int64_t i, bool ok, int32_t& lo, int32_t& hi, int32_t& s0, int32_t& s1) {
  const int2 v = ok ? __ldg(a.pair + i) : make_int2(-1, -1);
  s0 = v.x; s1 = v.y;
  lo = hi = 0;
  if (a.tail && ok) { lo = a.vptr[i]; hi = a.vptr[i + 1]; }
}

// S_c p_c for one cell, p gathered through `load`
template <int MB, bool CLS, typename Load>
__device__ __forceinline__ void ebe_apply_cell(const EbeArgs& a, int64_t c, double* __restrict__ out, Load load) {
  const int64_t nc = a.nc, ns = CLS ? a.ns : a.nc;
  const int64_t k = CLS ? a.cls[c] : c;
  double pe[MB], acc[MB];
  unsigned neg = 0;  // bit j: local DOF j carries a -1 sign around the class operator
#pragma unroll
  for (int j = 0; j < MB; ++j) {
    const int32_t e = __ldg(a.dofT + (int64_t)j * nc + c);
    pe[j] = e >= 0 ? load(e >> 1) : 0.0;
    if (CLS && e >= 0 && (e & 1)) { pe[j] = -pe[j]; neg |= 1u << j; }
    acc[j] = 0.0;
  }
#pragma unroll
  for (int i = 0; i < MB; ++i)
#pragma unroll
    for (int j = i; j < MB; ++j) {
      const double sv = a.S[(int64_t)ebe_sym(MB, i, j) * ns + k];
      acc[i] = fma(sv, pe[j], acc[i]);
      if (j != i) acc[j] = fma(sv, pe[i], acc[j]);
    }
#pragma unroll
  for (int j = 0; j < MB; ++j) out[(int64_t)j * nc + c] = (neg >> j) & 1u ? -acc[j] : acc[j];
}

// Partitioned runs (a.dist.world > 1): rows of ghost DOFs are neither updated nor counted in the dot products; the owner of a
// row stores its new z straight into the neighbours' exchange slots the moment it is computed (NVLink stores from inside
// phase C), the dot products are summed over the ranks in rank order (gg_dist.cuh), and the ghost values are copied from
// the slot into z after the round -- 2 rounds per iteration, no host, no NCCL.  (z itself stays in ordinary device memory:
// writing the whole vector into the IPC-exported window made phase C 2x slower on B200.)
#ifndef GG_EBE_MIN_BLOCKS
#define GG_EBE_MIN_BLOCKS 2
#endif
template <int MB, int MI, bool CLS>
__global__ void __launch_bounds__(256, GG_EBE_MIN_BLOCKS)
k_pcg_ebe(const EbeArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sh[32];
  constexpr int MIa = MI > 0 ? MI : 1;
  constexpr int UNR = 4;
  constexpr int UNRH = 2;
  constexpr int MBU = 3;   // preconditioner block rows loaded without a loop (facet blocks of RT_0..2, vertex / edge blocks of CG_1..3)
#ifndef GG_EBE_UNRB
#define GG_EBE_UNRB 4
#endif
  constexpr int UNRB = GG_EBE_UNRB;   // rows in flight per thread in the gather phase (latency-bound at 16 warps per SM)
  const int nbk = gridDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  const int64_t n = a.nb, nc = a.nc;
  const bool multi = a.dist.world > 1 || a.dist.force_multi;
  const uint8_t* __restrict__ owned = multi ? a.dist.owned : nullptr;
  unsigned long long epoch = multi ? a.dist.win[a.dist.rank]->epoch : 0ull;
  double* P0 = a.partials;
  double* P1 = a.partials + nbk;
  double* P2 = a.partials + 2 * nbk;
  double* r_old = a.ra;
  double* r_new = a.rb;

  const unsigned long long tp0 = ebe_now();
  // ---- one pass over the cells: condensed right-hand side  ev = T_c^T (s b_i)  and, for a warm start, ev2 = S_c x_c ----
  // condensed right-hand side of cell c, T_c read through ldT(entry)
  auto condensed_rhs = [&](int64_t c, auto ldT) {
    double bi[MIa];
    const int8_t* __restrict__ sg = CLS ? a.sg + c * a.mloc : nullptr;
#pragma unroll
    for (int t = 0; t < MI; ++t)
      bi[t] = a.scale_b * (a.be ? a.be[(int64_t)(MB + t) * nc + c] : a.b[n + c * MI + t]) * (sg ? (double)sg[MB + t] : 1.0);
#pragma unroll
    for (int j = 0; j < MB; ++j) {
      double acc = 0.0;
#pragma unroll
      for (int t = 0; t < MI; ++t) acc = fma(ldT(t * MB + j), bi[t], acc);
      a.ev[(int64_t)j * nc + c] = sg ? acc * (double)sg[j] : acc;
    }
  };
  const bool tiles = CLS && MI > 0 && a.ccell != nullptr;
  if (MI > 0 || a.warm) {
    if (tiles) {
      const int64_t ntiles = (a.ns + EBE_TILE_CLS - 1) / EBE_TILE_CLS;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t k0 = tile * EBE_TILE_CLS;
        __syncthreads();
        ebe_stage_tile(ebe_tile, a.T, MI * MB, a.ns, k0);
        __syncthreads();
        const int kk = threadIdx.x % EBE_TILE_CLS;
        for (int r = threadIdx.x / EBE_TILE_CLS; r < EBE_TILE_REP; r += blockDim.x / EBE_TILE_CLS) {
          const int64_t c = k0 + kk < a.ns ? a.ccell[(k0 + kk) * EBE_TILE_REP + r] : -1;
          if (c < 0) continue;
          condensed_rhs(c, [&](int ent) { return ebe_tile[ent * EBE_TILE_CLS + kk]; });
          if (a.warm) ebe_apply_cell<MB, CLS>(a, c, a.ev2, [&](int32_t d) { return a.x[d]; });
        }
      }
    } else {
      for (int64_t c = tid; c < nc; c += nth) {
        if (MI > 0) {
          const int64_t k = CLS ? a.cls[c] : c, ns = CLS ? a.ns : nc;
          condensed_rhs(c, [&](int ent) { return a.T[(int64_t)ent * ns + k]; });
        }
        // initial guess = what the output vector holds (the solution of the previous stage / step)
        if (a.warm) ebe_apply_cell<MB, CLS>(a, c, a.ev2, [&](int32_t d) { return a.x[d]; });
      }
    }
    grid.sync();
  }
  const unsigned long long tp1 = ebe_now();
  double bb = 0.0, rz = 0.0, rr = 0.0;
  int it = 0;
  unsigned long long tp2 = 0;
  if (a.cblk && !a.row_setup) {
    // ---- block-wise set-up: bt = s b_f - gather(ev), |bt|^2, r = bt - gather(ev2) AND z = Binv r, r.z, r.r in one pass: the rows of a
    //      preconditioner block belong to one thread, so no grid barrier is needed between the residual and the preconditioner ----
    double lbb = 0.0, lrz = 0.0, lrr0 = 0.0;
    for (int64_t b = tid; b < a.n_blocks; b += nth) {
      int k = 0;
      if (a.bc.n > 1 && b >= a.bc.b0[1]) k = 1;
      if (a.bc.n > 2 && b >= a.bc.b0[2]) k = 2;
      const int m = a.bc.size[k], g = (int)b - a.bc.b0[k];
      const int64_t r0 = a.bc.start[k] + (int64_t)g * m;
      const double* __restrict__ B = a.binv + a.bc.off[k] + (int64_t)g * m * m;
      int2 pr[3];
      double bt[3], rw[3], Bm[9], r3[3];
      uint8_t ob[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const bool on = j < m;
        pr[j] = on ? __ldg(a.pair + r0 + j) : make_int2(-1, -1);
        bt[j] = (on && !a.be) ? a.scale_b * a.b[r0 + j] : 0.0;
        ob[j] = on ? (owned ? owned[r0 + j] : (uint8_t)1) : (uint8_t)0;
        rw[j] = 0.0;
      }
      if (a.be) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          double v = 0.0;
          if (pr[j].x >= 0) v = a.be[pr[j].x];
          if (pr[j].y >= 0) v += a.be[pr[j].y];
          bt[j] = a.scale_b * v;
        }
      }
#pragma unroll
      for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int l = 0; l < 3; ++l) Bm[j * 3 + l] = (j < m && l < m) ? B[j * m + l] : 0.0;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (MI > 0) {
          if (pr[j].x >= 0) bt[j] -= a.ev[pr[j].x];
          if (pr[j].y >= 0) bt[j] -= a.ev[pr[j].y];
        }
        if (a.warm) {
          if (pr[j].x >= 0) rw[j] = a.ev2[pr[j].x];
          if (pr[j].y >= 0) rw[j] += a.ev2[pr[j].y];
        }
      }
      if (a.tail) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (j >= m) continue;
          for (int32_t q = a.vptr[r0 + j] + 2; q < a.vptr[r0 + j + 1]; ++q) {
            if (a.be) bt[j] = fma(a.scale_b, a.be[a.vsrc[q]], bt[j]);
            if (MI > 0) bt[j] -= a.ev[a.vsrc[q]];
            if (a.warm) rw[j] += a.ev2[a.vsrc[q]];
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        r3[j] = bt[j] - rw[j];
        if (j >= m) continue;
        r_old[r0 + j] = r3[j];
        a.p[r0 + j] = 0.0;
        if (!a.warm) a.x[r0 + j] = 0.0;
        if (ob[j]) lbb = fma(bt[j], bt[j], lbb);
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (!ob[j]) continue;
        double zj = 0.0;
#pragma unroll
        for (int l = 0; l < 3; ++l)
          if (l < m) zj = fma(Bm[j * 3 + l], r3[l], zj);
        a.z[r0 + j] = zj;
        lrz = fma(r3[j], zj, lrz);
        lrr0 = fma(r3[j], r3[j], lrr0);
      }
    }
    lbb = ebe_block_sum(lbb, sh);
    lrz = ebe_block_sum(lrz, sh);
    lrr0 = ebe_block_sum(lrr0, sh);
    if (threadIdx.x == 0) { P0[blockIdx.x] = lrz; P1[blockIdx.x] = lbb; P2[blockIdx.x] = lrr0; }
    grid.sync();
    if (multi) { dist_push(a.dist, epoch, a.z, tid, nth, n); grid.sync(); }
    rz = ebe_grid_total(P0, nbk, sh);
    bb = ebe_grid_total(P1, nbk, sh);
    rr = ebe_grid_total(P2, nbk, sh);
    if (multi) {
      double v[3] = {rz, rr, bb};
      dist_allreduce<3>(grid, a.dist, epoch, v);
      rz = v[0]; rr = v[1]; bb = v[2];
      dist_unpack(a.dist, epoch, a.z, tid, nth, n);
    }
    tp2 = ebe_now();
  } else {
    // ---- one pass over the rows: bt = s b_f - gather(ev), |bt|^2, r = bt - gather(ev2)  (4 rows in flight per thread) ----
    double lbb = 0.0;
    for (int64_t i0 = tid; i0 < n; i0 += UNR * nth) {
      int32_t lo[UNR], hi[UNR], s0[UNR], s1[UNR];
      double bt[UNR], rw[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int64_t i = i0 + u * nth;
        const bool ok = i < n;
        ebe_row_sources(a, i, ok, lo[u], hi[u], s0[u], s1[u]);
        bt[u] = (ok && !a.be) ? a.scale_b * a.b[i] : 0.0;
        rw[u] = 0.0;
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (a.be) {
          double v = 0.0;
          if (s0[u] >= 0) v = a.be[s0[u]];
          if (s1[u] >= 0) v += a.be[s1[u]];
          bt[u] = a.scale_b * v;
        }
        if (MI > 0) {
          if (s0[u] >= 0) bt[u] -= a.ev[s0[u]];
          if (s1[u] >= 0) bt[u] -= a.ev[s1[u]];
        }
        if (a.warm) {
          if (s0[u] >= 0) rw[u] = a.ev2[s0[u]];
          if (s1[u] >= 0) rw[u] += a.ev2[s1[u]];
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int64_t i = i0 + u * nth;
        if (i >= n) continue;
        for (int32_t k = lo[u] + 2; k < hi[u]; ++k) {
          if (a.be) bt[u] = fma(a.scale_b, a.be[a.vsrc[k]], bt[u]);
          if (MI > 0) bt[u] -= a.ev[a.vsrc[k]];
          if (a.warm) rw[u] += a.ev2[a.vsrc[k]];
        }
        r_old[i] = bt[u] - rw[u];
        if (!owned || owned[i]) lbb = fma(bt[u], bt[u], lbb);
        if (!a.warm) a.x[i] = 0.0;
      }
    }
    lbb = ebe_block_sum(lbb, sh);
    if (threadIdx.x == 0) P1[blockIdx.x] = lbb;
    grid.sync();
    bb = ebe_grid_total(P1, nbk, sh);
    if (multi) {
      double v[1] = {bb};
      dist_allreduce<1>(grid, a.dist, epoch, v);
      bb = v[0];
    }
    tp2 = ebe_now();
    double lrz = 0.0, lrr0 = 0.0;
    for (int64_t i0 = tid; i0 < n; i0 += UNR * nth) {
      int sb[UNR], mb[UNR], ro[UNR];
      double zi[UNR], ri[UNR];
      bool own[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int64_t i = i0 + u * nth;
        const bool ok = i < n;
        own[u] = ok && (!owned || owned[i]);
        sb[u] = 0; mb[u] = 0; ro[u] = 0;
        if (ok) ebe_block_of(a, i, sb[u], mb[u], ro[u]);
        if (!own[u]) mb[u] = 0;
        ri[u] = ok ? r_old[i] : 0.0;
        zi[u] = 0.0;
        if (ok) a.p[i] = 0.0;
      }
      {
        // block rows: the first MBU entries through predicated, fully unrolled loads (all in flight at once; a loop with a run-time
        // trip count serialises one DRAM round trip per row), longer blocks finish in the loop
        double bv[UNR][MBU], rv[UNR][MBU];
#pragma unroll
        for (int u = 0; u < UNR; ++u)
#pragma unroll
          for (int j = 0; j < MBU; ++j) {
            const bool on = j < mb[u];
            bv[u][j] = on ? a.binv[ro[u] + j] : 0.0;
            rv[u][j] = on ? r_old[sb[u] + j] : 0.0;
          }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
#pragma unroll
          for (int j = 0; j < MBU; ++j)
            if (j < mb[u]) zi[u] = fma(bv[u][j], rv[u][j], zi[u]);
          for (int j = MBU; j < mb[u]; ++j) zi[u] = fma(a.binv[ro[u] + j], r_old[sb[u] + j], zi[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (!own[u]) continue;
        a.z[i0 + u * nth] = zi[u];
        lrz = fma(ri[u], zi[u], lrz);
        lrr0 = fma(ri[u], ri[u], lrr0);
      }
    }
    lrz = ebe_block_sum(lrz, sh);
    lrr0 = ebe_block_sum(lrr0, sh);
    if (threadIdx.x == 0) { P0[blockIdx.x] = lrz; P2[blockIdx.x] = lrr0; }
    grid.sync();
    if (multi) { dist_push(a.dist, epoch, a.z, tid, nth, n); grid.sync(); }
    rz = ebe_grid_total(P0, nbk, sh);
    rr = ebe_grid_total(P2, nbk, sh);
    if (multi) {
      double v[2] = {rz, rr};
      dist_allreduce<2>(grid, a.dist, epoch, v);
      rz = v[0]; rr = v[1];
      dist_unpack(a.dist, epoch, a.z, tid, nth, n);  // ghost z: from the exchange slot into the vector (the barrier below orders it)
    }
  }
  double beta = 0.0;
  unsigned long long tA = 0, tB = 0, tC = 0, tR = 0, tP = 0;
  grid.sync();  // P2 is reused by the first iteration
  const unsigned long long tp3 = ebe_now();
  const double tol2 = a.rtol * a.rtol * bb;
  if (bb > 0.0 && rr > tol2) {
    for (it = 1; it <= a.max_iter; ++it) {
      // -- A: element vectors of S p with p = z + beta p_old formed on the fly (z may have been written by a peer: L2 loads)
      unsigned long long t0 = ebe_now();
      // (tiling this phase like the condensation passes was measured and dropped: 42.8 -> 73.0 us per iteration on RT_2 / 256^2 --
      // the phase is bound by the L1 wavefronts of the p / z gathers, not by the S reads, and the tile loop adds two CTA barriers
      // per 192 cells, idles a quarter of the threads and lengthens the instruction stream of the iteration loop)
      for (int64_t c = tid; c < nc; c += nth) ebe_apply_cell<MB, CLS>(a, c, a.ev, [&](int32_t d) { return fma(beta, a.p[d], __ldcg(a.z + d)); });
      grid.sync();
      unsigned long long t1 = ebe_now();
      tA += t1 - t0;
      // -- B: p = z + beta p ; Ap = gather ; pAp
      double lpap = 0.0;
      for (int64_t i0 = tid; i0 < n; i0 += UNRB * nth) {
        int32_t lo[UNRB], hi[UNRB], s0[UNRB], s1[UNRB];
        double pv[UNRB], acc[UNRB];
        bool own[UNRB];
#pragma unroll
        for (int u = 0; u < UNRB; ++u) {
          const int64_t i = i0 + u * nth;
          const bool ok = i < n;
          // the ownership byte is loaded alongside the row data, not ahead of it (no extra dependent-load level);
          // ghost rows gather too (harmless, few) but are not stored
          own[u] = ok && (!owned || owned[i]);
          ebe_row_sources(a, i, ok, lo[u], hi[u], s0[u], s1[u]);
          pv[u] = ok ? fma(beta, a.p[i], __ldcg(a.z + i)) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < UNRB; ++u) {
          acc[u] = s0[u] >= 0 ? a.ev[s0[u]] : 0.0;
          if (s1[u] >= 0) acc[u] += a.ev[s1[u]];
        }
#pragma unroll
        for (int u = 0; u < UNRB; ++u) {
          for (int32_t k = lo[u] + 2; k < hi[u]; ++k) acc[u] += a.ev[a.vsrc[k]];
          const int64_t i = i0 + u * nth;
          if (i < n) a.p[i] = pv[u];
          if (own[u]) { a.Ap[i] = acc[u]; lpap = fma(pv[u], acc[u], lpap); }
        }
      }
      lpap = ebe_block_sum(lpap, sh);
      if (threadIdx.x == 0) P2[blockIdx.x] = lpap;
      grid.sync();
      t0 = ebe_now();
      tB += t0 - t1;
      double pap = ebe_grid_total(P2, nbk, sh);
      if (multi) {
        double v[1] = {pap};
        dist_allreduce<1>(grid, a.dist, epoch, v);
        pap = v[0];
        const unsigned long long tr = ebe_now();
        tR += tr - t0;
      }
      const double alpha = rz / pap;
      const unsigned long long tc0 = ebe_now();
      // -- C: x += alpha p ; r -= alpha Ap ; z = Binv r (block rows recomputed: no barrier) ; rz, rr  (owned rows only)
      double lrz2 = 0.0, lrr = 0.0;
      if (a.cblk) {
        // one thread per preconditioner block (<= 3 rows: the DOFs of a facet / vertex / edge): every operand of the block is
        // loaded once, all loads are independent of each other (block position by arithmetic) and in flight together -- the row
        // loop below needs two dependent DRAM round trips per round (index arrays, then the block rows)
        for (int64_t b = tid; b < a.n_blocks; b += nth) {
          int k = 0;
          if (a.bc.n > 1 && b >= a.bc.b0[1]) k = 1;
          if (a.bc.n > 2 && b >= a.bc.b0[2]) k = 2;
          const int m = a.bc.size[k], g = (int)b - a.bc.b0[k];
          const int64_t s0 = a.bc.start[k] + (int64_t)g * m;
          const double* __restrict__ B = a.binv + a.bc.off[k] + (int64_t)g * m * m;
          double Av[3], Rv[3], Xv[3], Pv[3], Bm[9];
          uint8_t ob[3];
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const bool on = j < m;
            Av[j] = on ? a.Ap[s0 + j] : 0.0; Rv[j] = on ? r_old[s0 + j] : 0.0;
            Xv[j] = on ? a.x[s0 + j] : 0.0; Pv[j] = on ? a.p[s0 + j] : 0.0;
            ob[j] = on ? (owned ? owned[s0 + j] : (uint8_t)1) : (uint8_t)0;
          }
#pragma unroll
          for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int l = 0; l < 3; ++l) Bm[j * 3 + l] = (j < m && l < m) ? B[j * m + l] : 0.0;
          double rn[3];
#pragma unroll
          for (int j = 0; j < 3; ++j) rn[j] = fma(-alpha, Av[j], Rv[j]);
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            if (!ob[j]) continue;
            double zj = 0.0;
#pragma unroll
            for (int l = 0; l < 3; ++l)
              if (l < m) zj = fma(Bm[j * 3 + l], rn[l], zj);
            a.x[s0 + j] = fma(alpha, Pv[j], Xv[j]);
            r_new[s0 + j] = rn[j]; a.z[s0 + j] = zj;
            lrz2 = fma(rn[j], zj, lrz2);
            lrr = fma(rn[j], rn[j], lrr);
            if (ob[j] == 2) dist_push_row(a.dist, epoch, s0 + j, zj);
          }
        }
      } else
      for (int64_t i0 = tid; i0 < n; i0 += UNR * nth) {
        int sb[UNR], mb[UNR], ro[UNR];
        double ri[UNR], zi[UNR], xi[UNR], pi[UNR];
        uint8_t ob[UNR];  // 0 ghost row, 1 owned, 2 owned and pushed to neighbours
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int64_t i = i0 + u * nth;
          const bool ok = i < n;
          ob[u] = ok ? (owned ? owned[i] : (uint8_t)1) : (uint8_t)0;
          sb[u] = 0; mb[u] = 0; ro[u] = 0;
          if (ok) ebe_block_of(a, i, sb[u], mb[u], ro[u]);
          ri[u] = ok ? fma(-alpha, a.Ap[i], r_old[i]) : 0.0;
          xi[u] = ok ? a.x[i] : 0.0; pi[u] = ok ? a.p[i] : 0.0;
          zi[u] = 0.0;
        }
#pragma unroll
        for (int h = 0; h < UNR; h += UNRH) {   // UNRH rows at a time: 9 loads per row in flight, within the register budget
          double bv[UNRH][MBU], av[UNRH][MBU], rv[UNRH][MBU];
#pragma unroll
          for (int v = 0; v < UNRH; ++v)
#pragma unroll
            for (int j = 0; j < MBU; ++j) {
              const int u = h + v;
              const bool on = j < mb[u];
              bv[v][j] = on ? a.binv[ro[u] + j] : 0.0;
              av[v][j] = on ? a.Ap[sb[u] + j] : 0.0;
              rv[v][j] = on ? r_old[sb[u] + j] : 0.0;
            }
#pragma unroll
          for (int v = 0; v < UNRH; ++v) {
            const int u = h + v;
#pragma unroll
            for (int j = 0; j < MBU; ++j)
              if (j < mb[u]) zi[u] = fma(bv[v][j], fma(-alpha, av[v][j], rv[v][j]), zi[u]);
            for (int j = MBU; j < mb[u]; ++j) zi[u] = fma(a.binv[ro[u] + j], fma(-alpha, a.Ap[sb[u] + j], r_old[sb[u] + j]), zi[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int64_t i = i0 + u * nth;
          if (ob[u]) {
            a.x[i] = fma(alpha, pi[u], xi[u]);
            r_new[i] = ri[u]; a.z[i] = zi[u];
            lrz2 = fma(ri[u], zi[u], lrz2);
            lrr = fma(ri[u], ri[u], lrr);
          }
        }
        // halo: the owner stores z straight into the neighbours' exchange slots (ordered before the flag by the grid
        // barrier and the cumulative system-scope fence of the posting thread, gg_dist.cuh)
#pragma unroll
        for (int u = 0; u < UNR; ++u)
          if (ob[u] == 2) dist_push_row(a.dist, epoch, i0 + u * nth, zi[u]);
      }
      lrz2 = ebe_block_sum(lrz2, sh);
      lrr = ebe_block_sum(lrr, sh);
      if (threadIdx.x == 0) { P0[blockIdx.x] = lrz2; P1[blockIdx.x] = lrr; }
      const unsigned long long tc1 = ebe_now();
      grid.sync();
      t1 = ebe_now();
      tC += t1 - t0;
      tP += tc1 - tc0;
      { double* t = r_old; r_old = r_new; r_new = t; }
      double rz_new = ebe_grid_total(P0, nbk, sh);
      rr = ebe_grid_total(P1, nbk, sh);
      if (multi) {
        double v[2] = {rz_new, rr};
        dist_allreduce<2>(grid, a.dist, epoch, v);
        rz_new = v[0]; rr = v[1];
        if (rr > tol2) { dist_unpack(a.dist, epoch, a.z, tid, nth, n); grid.sync(); }
        tR += ebe_now() - t1;
      }
      if (rr <= tol2) break;  // uniform across the grid and across the ranks
      beta = rz_new / rz;
      rz = rz_new;
    }
  }
  if (multi) {
    // owner -> ghost update of the solution (the exchange slot is free again: z is dead)
    grid.sync();
    dist_push(a.dist, epoch, a.x, tid, nth, n);
    grid.sync();
    double v[1] = {0.0};
    dist_allreduce<1>(grid, a.dist, epoch, v);
    dist_unpack(a.dist, epoch, a.x, tid, nth, n);
    grid.sync();
    if (tid == 0) a.dist.win[a.dist.rank]->epoch = epoch;
  }
  const unsigned long long tp4 = ebe_now();
  // ---- eliminated DOFs: x_i = W (s b_i) - T x_f ----
  auto recover = [&](int64_t c, auto ldT, auto ldW) {
    double bi[MIa], xf[MB];
    const int8_t* __restrict__ sg = CLS ? a.sg + c * a.mloc : nullptr;
#pragma unroll
    for (int t = 0; t < MI; ++t)
      bi[t] = a.scale_b * (a.be ? a.be[(int64_t)(MB + t) * nc + c] : a.b[n + c * MI + t]) * (sg ? (double)sg[MB + t] : 1.0);
#pragma unroll
    for (int j = 0; j < MB; ++j) {
      const int32_t e = __ldg(a.dofT + (int64_t)j * nc + c);
      xf[j] = e >= 0 ? a.x[e >> 1] : 0.0;
      if (CLS && e >= 0 && (e & 1)) xf[j] = -xf[j];
    }
#pragma unroll
    for (int t = 0; t < MI; ++t) {
      double acc = 0.0;
#pragma unroll
      for (int s = 0; s < MI; ++s) {
        const int lo = t < s ? t : s, hi = t < s ? s : t;
        acc = fma(ldW(ebe_sym(MI, lo, hi)), bi[s], acc);
      }
#pragma unroll
      for (int j = 0; j < MB; ++j) acc = fma(-ldT(t * MB + j), xf[j], acc);
      a.x[n + c * MI + t] = sg ? acc * (double)sg[MB + t] : acc;
    }
  };
  if (MI > 0) {
    if (tiles) {
      constexpr int NT = MI * MB, NW = MIa * (MIa + 1) / 2;
      const int64_t ntiles = (a.ns + EBE_TILE_CLS - 1) / EBE_TILE_CLS;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t k0 = tile * EBE_TILE_CLS;
        __syncthreads();
        ebe_stage_tile(ebe_tile, a.T, NT, a.ns, k0);
        ebe_stage_tile(ebe_tile + NT * EBE_TILE_CLS, a.W, NW, a.ns, k0);
        __syncthreads();
        const int kk = threadIdx.x % EBE_TILE_CLS;
        for (int r = threadIdx.x / EBE_TILE_CLS; r < EBE_TILE_REP; r += blockDim.x / EBE_TILE_CLS) {
          const int64_t c = k0 + kk < a.ns ? a.ccell[(k0 + kk) * EBE_TILE_REP + r] : -1;
          if (c < 0) continue;
          recover(c, [&](int ent) { return ebe_tile[ent * EBE_TILE_CLS + kk]; },
                  [&](int ent) { return ebe_tile[(NT + ent) * EBE_TILE_CLS + kk]; });
        }
      }
    } else {
      for (int64_t c = tid; c < nc; c += nth) {
        const int64_t k = CLS ? a.cls[c] : c, ns = CLS ? a.ns : nc;
        recover(c, [&](int ent) { return a.T[(int64_t)ent * ns + k]; }, [&](int ent) { return a.W[(int64_t)ent * ns + k]; });
      }
    }
  }
  if (tid == 0) {
    const int itc = it > a.max_iter ? a.max_iter : it;
    a.stats[0] = (double)itc;
    a.stats[1] = bb > 0.0 ? sqrt(rr / bb) : 0.0;
    a.stats[2] += (double)tA; a.stats[3] += (double)tB; a.stats[4] += (double)tC; a.stats[5] += (double)itc;
    a.stats[6] += (double)tR; a.stats[7] += (double)tP;
    // set-up and tear-down phases of the solve: cell pre-pass, row pass, first preconditioner pass, interior recovery; solves
    a.stats[8] += (double)(tp1 - tp0); a.stats[9] += (double)(tp2 - tp1); a.stats[10] += (double)(tp3 - tp2);
    a.stats[11] += (double)(ebe_now() - tp4); a.stats[12] += 1.0;
  }
}

static void ebe_build_precond(EbeSolver* s);
static void ebe_compress(EbeSolver* s);

// per-class copies of the per-cell operators of the representative cells, with the DOF signs taken out again
// (planes [n_planes][nc] -> [n_planes][ns]; sign_a, sign_b: local DOF of the row / column of every plane)
__global__ void k_ebe_compact(int64_t ns, int64_t nc, int n_planes, int mloc, const int32_t* __restrict__ rep,
                              const int8_t* __restrict__ sg, const uint8_t* __restrict__ plane_a, const uint8_t* __restrict__ plane_b,
                              const double* __restrict__ in, double* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ns * n_planes) return;
  const int64_t k = t % ns;
  const int pl = (int)(t / ns);
  const int64_t c = rep[k];
  out[(int64_t)pl * ns + k] = in[(int64_t)pl * nc + c] * (double)(sg[c * mloc + plane_a[pl]] * sg[c * mloc + plane_b[pl]]);
}

template <int MB, int MI>
static void ebe_launch_condense(gg_ctx* ctx, int64_t nc, const double* ebuf, double* S, double* T, double* W) {
  k_ebe_condense<MB, MI><<<nblk(nc, 64), 64, 0, ctx->stream>>>(nc, ebuf, S, T, W);
}

#define GG_EBE_CASES(X) X(4, 0) X(8, 4) X(12, 12) X(8, 1) X(12, 4)

static void* ebe_kernel(int MB, int MI, bool classes = false) {
#define X(MB_, MI_) if (MB == MB_ && MI == MI_) return classes ? (void*)k_pcg_ebe<MB_, MI_, true> : (void*)k_pcg_ebe<MB_, MI_, false>;
  GG_EBE_CASES(X)
#undef X
  return nullptr;
}

// (boundary, interior) local DOF split of a builtin space, or false when the condensed solver does not apply
bool ebe_supported(const gg_space* sp) {
  if (sp->blocks.empty() || sp->constraint != GG_CONSTRAINT_NONE) return false;
  int MB, MI;
  if (sp->kind == GG_SPACE_RT) { MB = 4 * (sp->order + 1); MI = 2 * sp->order * (sp->order + 1); }
  else if (sp->kind == GG_SPACE_CG) { MB = 4 * sp->order; MI = (sp->order - 1) * (sp->order - 1); }
  else return false;
  return ebe_kernel(MB, MI) != nullptr;
}

EbeSolver* ebe_create(gg_space* sp, double rtol, int max_iter) {
  GG_REQUIRE(ebe_supported(sp), GG_ERR_UNSUPPORTED, "the condensed element-by-element solver needs a builtin RT_0..2 / CG_1..3 space");
  gg_ctx* ctx = sp->mesh->ctx;
  cudaStream_t st = ctx->stream;
  std::unique_ptr<EbeSolver> s(new EbeSolver);
  s->ctx = ctx; s->sp = sp;
  s->rtol = rtol > 0 ? rtol : 1e-12;
  s->max_iter = max_iter > 0 ? max_iter : 1000;
  if (sp->kind == GG_SPACE_RT) { s->MB = 4 * (sp->order + 1); s->MI = 2 * sp->order * (sp->order + 1); }
  else { s->MB = 4 * sp->order; s->MI = (sp->order - 1) * (sp->order - 1); }
  const int64_t nc = sp->mesh->n_cells;
  s->nc = nc;
  // boundary DOFs come first in the builtin numberings, the interior ones are numbered cell by cell after them
  s->nb = sp->n_free - nc * s->MI;
  if (s->MI > 0) {
    const auto& last = sp->blocks.back();
    GG_REQUIRE(last.start == s->nb && last.count == nc && last.size == s->MI, GG_ERR_INVALID, "unexpected interior DOF numbering");
  }
  const int64_t n = s->nb;
  std::vector<int32_t> bstart(n), rowoff(n), blk_off;
  std::vector<uint8_t> bsize(n), blk_size;
  int64_t off = 0, covered = 0;
  for (const auto& bc : sp->blocks) {
    if (bc.start >= n) break;
    GG_REQUIRE(bc.start == covered && bc.size <= EBE_MAX_BLOCK, GG_ERR_INVALID, "inconsistent DOF block classes");
    for (int64_t g = 0; g < bc.count; ++g) {
      blk_off.push_back((int32_t)off); blk_size.push_back((uint8_t)bc.size);
      for (int i = 0; i < bc.size; ++i) {
        bstart[covered + i] = (int32_t)covered; bsize[covered + i] = (uint8_t)bc.size;
        rowoff[covered + i] = (int32_t)(off + (int64_t)i * bc.size);
      }
      off += (int64_t)bc.size * bc.size;
      covered += bc.size;
    }
  }
  GG_REQUIRE(covered == n, GG_ERR_INVALID, "DOF block classes do not cover the boundary DOFs");
  GG_REQUIRE(off < (int64_t)INT32_MAX, GG_ERR_UNSUPPORTED, "block preconditioner too large");
  s->n_blocks = (int64_t)blk_off.size();
  {
    // runs of equally sized blocks -> arithmetic block lookup in the solver kernel (at most 3 runs, else the index arrays are read)
    int nrun = 0;
    int64_t cov = 0, o = 0, nblk_seen = 0;
    bool fits = true, small = true;
    for (const auto& bc : sp->blocks) {
      if (bc.start >= n) break;
      if (bc.count == 0) continue;
      if (nrun == 3) { fits = false; break; }
      s->bc_start[nrun] = (int)cov; s->bc_size[nrun] = bc.size; s->bc_off[nrun] = (int)o; s->bc_b0[nrun] = (int)nblk_seen;
      if (bc.size > 3) small = false;
      nblk_seen += bc.count;
      cov += bc.count * bc.size; o += bc.count * (int64_t)bc.size * bc.size;
      ++nrun;
    }
    // measured on B200 (RT_2, 256^2 per panel): inside the ROW loops the arithmetic lookup is 1 us slower than the index arrays
    // (GG_EBE_BLOCK_CLASSES=1 switches it on there); the block-wise update phase below is built on it
    s->bc_n = fits ? nrun : 0;
    // block-wise update phase: needs the arithmetic block positions and blocks of at most 3 rows; GG_EBE_ROW_UPDATE=1 keeps the row loop
    s->cblk = fits && small && nrun > 0 && getenv("GG_EBE_ROW_UPDATE") == nullptr ? 1 : 0;
  }
  s->bstart.upload(bstart, st); s->rowoff.upload(rowoff, st); s->bsize.upload(bsize, st);
  s->blk_off.upload(blk_off, st); s->blk_size.upload(blk_size, st);
  s->binv.alloc(off);
  const int MB = s->MB, MI = s->MI;
  s->S.alloc((size_t)(MB * (MB + 1) / 2) * nc);
  s->T.alloc((size_t)MI * MB * nc);
  s->W.alloc((size_t)(MI * (MI + 1) / 2) * nc);
  s->ev.alloc((size_t)MB * std::max<int64_t>(nc, 1));
  s->ev2.alloc((size_t)MB * std::max<int64_t>(nc, 1));
  s->r.alloc(n); s->r2.alloc(n); s->z.alloc(n); s->p.alloc(n); s->Ap.alloc(n);
  GG_REQUIRE(sp->n_free < (int64_t)1 << 30, GG_ERR_UNSUPPORTED, "too many DOFs for the sign-tagged DOF table");
  s->dofT.alloc((size_t)MB * std::max<int64_t>(nc, 1));
  if (nc) {
    k_ebe_dof_table<<<nblk(nc * MB, 256), 256, 0, st>>>(nc, sp->ndofs_loc, MB, sp->d_cell_dofs.p,
                                                        sp->d_cell_signs.n ? sp->d_cell_signs.p : nullptr, s->dofT.p);
    ctx->launches++;
    GG_CUDA(cudaGetLastError());
  }
  build_vec_map(sp);
  {
    std::vector<int32_t> vp((size_t)n + 1);
    if (n) GG_CUDA(cudaMemcpyAsync(vp.data(), sp->d_vptr.p, (n + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    GG_CUDA(cudaStreamSynchronize(st));
    bool tail = getenv("GG_EBE_ROW_POINTERS") != nullptr;   // measurement: always walk the row pointers
    for (int64_t i = 0; !tail && i < n; ++i) tail = vp[i + 1] - vp[i] > 2;
    s->tail = tail ? 1 : 0;
    s->pair.alloc(std::max<int64_t>(n, 1));
    if (n) {
      k_ebe_source_pairs<<<nblk(n, 256), 256, 0, st>>>(n, sp->d_vptr.p, sp->d_vsrc.p, s->pair.p);
      ctx->launches++;
      GG_CUDA(cudaGetLastError());
    }
  }
  void* fn = ebe_kernel(MB, MI);
  int per_sm = 0;
  GG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, 0));
  GG_REQUIRE(per_sm >= 1, GG_ERR_CUDA, "condensed PCG kernel does not fit on an SM");
  s->grid = ctx->sm_count * per_sm;
  s->partials.alloc(3 * (size_t)s->grid);
  s->stats.alloc(16);
  GG_CUDA(cudaMemsetAsync(s->stats.p, 0, 16 * sizeof(double), st));
  GG_CUDA(cudaStreamSynchronize(st));
  return s.release();
}

void ebe_destroy(EbeSolver* s) { delete s; }

// numerical_setup!: condense the element matrices (entry-major full layout [(i*m+j)][cell]) and rebuild the preconditioner
void ebe_set_matrix(EbeSolver* s, const double* ebuf) {
  gg_ctx* ctx = s->ctx;
  gg_space* sp = s->sp;
  if (s->nc == 0) return;
  if (s->compressed) {
    // back to one operator per cell before condensing again
    const int MB = s->MB, MI = s->MI;
    s->S.alloc((size_t)(MB * (MB + 1) / 2) * s->nc);
    s->T.alloc((size_t)MI * MB * s->nc);
    s->W.alloc((size_t)(MI * (MI + 1) / 2) * s->nc);
    s->compressed = false; s->ns = s->nc;
  }
  {
    ProfScope ps(ctx, "k_ebe_condense");
#define X(MB_, MI_) if (s->MB == MB_ && s->MI == MI_) ebe_launch_condense<MB_, MI_>(ctx, s->nc, ebuf, s->S.p, s->T.p, s->W.p);
    GG_EBE_CASES(X)
#undef X
    ctx->launches++;
  }
  GG_CUDA(cudaGetLastError());
  s->ns = s->nc; s->compressed = false;
  ebe_build_precond(s);
  ebe_compress(s);
}

// Constant operators on a mesh whose cells repeat chart boxes (every (x-interval, y-interval) pair of the cubed sphere occurs on
// all 6 panels): the metric depends on the chart point only, so cells of one geometry class share S_c, T_c, M_ii^-1 up to the
// DOF signs.  Keep one copy per class (6x less operator traffic per PCG iteration; the class operators of a 256^2 panel stay in
// L2) and apply the signs on the fly.  Bitwise identical results: the copies are the representative cell's own numbers.
// (Tried and slower on B200: staging the S of 42 classes per CTA in shared memory so that the 6 cells of a class read it from
// there -- the load / barrier / compute sequence per class block loses the latency hiding of the free-running loop: cell phase
// 93 us against 49 us per iteration, with either thread mapping.)
static void ebe_compress(EbeSolver* s) {
  static const bool off = getenv("GG_EBE_CLASSES") && getenv("GG_EBE_CLASSES")[0] == '0';
  gg_space* sp = s->sp;
  gg_mesh* m = sp->mesh;
  gg_ctx* ctx = s->ctx;
  const int64_t nc = s->nc;
  if (off || sp->kind != GG_SPACE_RT || m->dim != 2 || (int64_t)m->h_ix.size() != nc || nc == 0) return;
  mesh_geometry_classes(m);
  const std::vector<int32_t>&cls = m->cls, &rep = m->cls_rep;
  const int64_t ns = (int64_t)rep.size();
  if (2 * ns > nc) return;   // nothing to gain
  cudaStream_t st = ctx->stream;
  const int MB = s->MB, MI = s->MI, mloc = sp->ndofs_loc;
  DevBuf<int32_t> drep;
  drep.upload(rep, st);
  auto compact = [&](DevBuf<double>& buf, const std::vector<uint8_t>& pa, const std::vector<uint8_t>& pb) {
    const int np = (int)pa.size();
    if (np == 0) return;
    DevBuf<uint8_t> da, db;
    da.upload(pa, st); db.upload(pb, st);
    DevBuf<double> out((size_t)np * ns);
    k_ebe_compact<<<nblk(ns * np, 256), 256, 0, st>>>(ns, nc, np, mloc, drep.p, sp->d_cell_signs.p, da.p, db.p, buf.p, out.p);
    ctx->launches++;
    GG_CUDA(cudaGetLastError());
    GG_CUDA(cudaStreamSynchronize(st));
    buf = std::move(out);
  };
  std::vector<uint8_t> pa, pb;
  for (int i = 0; i < MB; ++i) for (int j = i; j < MB; ++j) { pa.push_back((uint8_t)i); pb.push_back((uint8_t)j); }
  compact(s->S, pa, pb);
  pa.clear(); pb.clear();
  for (int t = 0; t < MI; ++t) for (int j = 0; j < MB; ++j) { pa.push_back((uint8_t)(MB + t)); pb.push_back((uint8_t)j); }
  compact(s->T, pa, pb);
  pa.clear(); pb.clear();
  for (int i = 0; i < MI; ++i) for (int j = i; j < MI; ++j) { pa.push_back((uint8_t)(MB + i)); pb.push_back((uint8_t)(MB + j)); }
  compact(s->W, pa, pb);
  s->cls.upload(cls, st);
  // class -> cells table of the tiled condensation passes (at most EBE_TILE_REP congruent cells per class, else the cell loops stay)
  {
    std::vector<int32_t> cc((size_t)ns * EBE_TILE_REP, -1), cnt((size_t)ns, 0);
    bool ok = MI > 0 && !(getenv("GG_EBE_TILES") && getenv("GG_EBE_TILES")[0] == '0');
    for (int64_t c = 0; ok && c < nc; ++c) {
      const int32_t k = cls[c];
      if (cnt[k] == EBE_TILE_REP) { ok = false; break; }
      cc[(size_t)k * EBE_TILE_REP + cnt[k]++] = (int32_t)c;
    }
    s->tile_smem = 0;
    if (ok) {
      const size_t bytes = (size_t)(MI * MB + MI * (MI + 1) / 2) * EBE_TILE_CLS * sizeof(double);
      void* fn = ebe_kernel(MB, MI, true);
      int per_sm = 0;
      if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess &&
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, bytes) == cudaSuccess && per_sm * ctx->sm_count >= s->grid) {
        s->ccell.upload(cc, st);
        s->tile_smem = bytes;
      } else {
        (void)cudaGetLastError();   // the tiles would cost residency: keep the cell loops
      }
    }
  }
  GG_CUDA(cudaStreamSynchronize(st));
  s->ns = ns; s->compressed = true;
}

static void ebe_build_precond(EbeSolver* s) {
  gg_ctx* ctx = s->ctx;
  gg_space* sp = s->sp;
  if (s->nb == 0) return;
  ProfScope ps(ctx, "k_ebe_precond");
  k_ebe_block_rows<<<nblk(s->nb, 128), 128, 0, ctx->stream>>>(s->nb, s->nc, sp->ndofs_loc, s->MB, sp->d_vptr.p, sp->d_vsrc.p,
                                                              sp->d_cell_dofs.p, s->bstart.p, s->bsize.p, s->rowoff.p, s->S.p,
                                                              s->binv.p);
  k_ebe_block_inverse<<<nblk(s->n_blocks, 128), 128, 0, ctx->stream>>>(s->n_blocks, s->blk_off.p, s->blk_size.p, s->binv.p);
  ctx->launches += 2;
  GG_CUDA(cudaGetLastError());
}

// numerical_setup! for the scalar mass int coef q w sqrtg on a CG_P space: element matrices, condensation and preconditioner
// in one pass, coef given at the quadrature points [n_cells][Q] (x-fastest inside the cell) or NULL (= 1)
void ebe_set_scalar_mass(EbeSolver* s, int degree, const double* coef) {
  gg_ctx* ctx = s->ctx;
  gg_space* sp = s->sp;
  GG_REQUIRE(sp->kind == GG_SPACE_CG && sp->order >= 1 && sp->order <= 3, GG_ERR_UNSUPPORTED, "fused scalar mass needs CG_1..3");
  GG_REQUIRE(coef, GG_ERR_INVALID, "coefficient at the quadrature points is required");
  if (s->nc == 0) return;
  const int nq = num_points_1d(degree), P = sp->order, NB = P + 1, NS = NB * (NB + 1) / 2;
  GG_REQUIRE(nq <= MAX_NQ, GG_ERR_UNSUPPORTED, "quadrature degree too high");
  const int64_t key = ((int64_t)P << 8) | nq;
  if (ctx->ebe_key != key) {
    std::vector<double> xi(nq), w(nq), nn((size_t)nq * NS), N(NB), D(NB);
    gauss_legendre_01(nq, xi.data(), w.data());
    for (int q = 0; q < nq; ++q) {
      lagrange_1d(P, xi[q], N.data(), D.data());
      int k = 0;
      for (int i = 0; i < NB; ++i)
        for (int j = i; j < NB; ++j) nn[(size_t)q * NS + k++] = N[i] * N[j];
    }
    GG_CUDA(cudaStreamSynchronize(ctx->stream));  // earlier launches may still read the old tables
    GG_CUDA(cudaMemcpyToSymbol(c_ebe_w, w.data(), nq * sizeof(double)));
    GG_CUDA(cudaMemcpyToSymbol(c_ebe_nn, nn.data(), nn.size() * sizeof(double)));
    ctx->ebe_key = key;
  }
  CellGeo g = cell_geo(sp->mesh, nq);
  {
    ProfScope ps(ctx, "k_ebe_scalar_mass");
    unsigned nb = nblk(s->nc, 64);
    switch (P) {
      case 1: k_ebe_scalar_mass<1><<<nb, 64, 0, ctx->stream>>>(s->nc, nq, g, coef, s->S.p, s->T.p, s->W.p); break;
      case 2: k_ebe_scalar_mass<2><<<nb, 64, 0, ctx->stream>>>(s->nc, nq, g, coef, s->S.p, s->T.p, s->W.p); break;
      case 3: k_ebe_scalar_mass<3><<<nb, 64, 0, ctx->stream>>>(s->nc, nq, g, coef, s->S.p, s->T.p, s->W.p); break;
    }
    ctx->launches++;
    GG_CUDA(cudaGetLastError());
  }
  ebe_build_precond(s);
}

// enqueue x = A^-1 (scale_b b), x and b over ALL free DOFs of the space (boundary first, then interior); no host sync
void ebe_solve_async(EbeSolver* s, const double* b, double scale_b, double* x, bool warm, const double* be) {
  gg_ctx* ctx = s->ctx;
  gg_space* sp = s->sp;
  if (sp->n_free == 0) return;
  EbeArgs a;
  a.nb = s->nb; a.nc = s->nc; a.mloc = sp->ndofs_loc; a.cell_dofs = sp->d_cell_dofs.p; a.dofT = s->dofT.p;
  a.S = s->S.p; a.T = s->T.p; a.W = s->W.p; a.ev = s->ev.p; a.ev2 = s->ev2.p;
  a.ns = s->compressed ? s->ns : s->nc;
  a.cls = s->compressed ? s->cls.p : nullptr;
  // tiles pay once the cell loops take more than one round of the persistent grid; on small meshes (the 48^2 wave test: 72 tiles for
  // 296 CTAs) the staging barriers cost 3 us per pass, so the cell loops stay
  static const bool force_tiles = getenv("GG_EBE_TILES") && getenv("GG_EBE_TILES")[0] == '2';   // tests: tiles on small meshes too
  a.ccell = s->compressed && s->tile_smem && (force_tiles || s->nc >= (int64_t)s->grid * 256) ? s->ccell.p : nullptr;
  a.sg = s->compressed ? sp->d_cell_signs.p : nullptr;
  a.vptr = sp->d_vptr.p; a.vsrc = sp->d_vsrc.p; a.pair = s->pair.p; a.tail = s->tail;
  a.bc.n = s->bc_n; a.n_blocks = s->n_blocks; a.cblk = s->cblk; a.bc_rows = getenv("GG_EBE_BLOCK_CLASSES") != nullptr; 
  // merged block-wise set-up pass: one grid barrier less, which pays on small meshes (48^2 wave test: 14.3 -> 11.9 us per solve); on
  // large ones its three dependent load levels lose to the two row passes (256^2 RT_2: 91 -> 106 us), so it is chosen by size
  // (GG_EBE_BLOCK_SETUP=1 / 0 force it on / off)
  {
    const char* e = getenv("GG_EBE_BLOCK_SETUP");
    a.row_setup = e ? (e[0] == '0') : (s->n_blocks > (int64_t)s->grid * 256);
  }
  for (int k = 0; k < 3; ++k) { a.bc.start[k] = s->bc_start[k]; a.bc.size[k] = s->bc_size[k]; a.bc.off[k] = s->bc_off[k]; a.bc.b0[k] = s->bc_b0[k]; }
  a.bstart = s->bstart.p; a.rowoff = s->rowoff.p; a.bsize = s->bsize.p; a.binv = s->binv.p;
  a.b = be ? nullptr : b; a.be = be; a.scale_b = scale_b; a.x = x; a.ra = s->r.p; a.rb = s->r2.p; a.z = s->z.p; a.p = s->p.p; a.Ap = s->Ap.p;
  a.partials = s->partials.p; a.rtol = s->rtol; a.max_iter = s->max_iter; a.stats = s->stats.p;
  a.warm = warm ? 1 : 0;
  a.dist = dist_dev(sp);
  if (getenv("GG_FORCE_MULTI") && a.dist.win) a.dist.force_multi = 1;  // measurement: run the partitioned code path on one rank
  void* fn = ebe_kernel(s->MB, s->MI, s->compressed);
  int grid = s->grid;
  const int64_t want = (std::max(s->nb, s->nc) + 255) / 256;
  if (want < grid) grid = (int)std::max<int64_t>(want, 1);
  if (const char* g = getenv("GG_EBE_GRID")) grid = std::max(1, std::min(grid, atoi(g)));  // measurement: barrier latency against work per thread
  void* args[] = {&a};
  ProfScope ps(ctx, "k_pcg_ebe");
  GG_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(256), args, a.ccell ? s->tile_smem : 0, ctx->stream));
  ctx->launches++;
}

}  // namespace gg
