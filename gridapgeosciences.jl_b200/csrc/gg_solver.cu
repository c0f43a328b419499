// CSR SpMV and a persistent, single-launch block-Jacobi-preconditioned conjugate-gradient solver.
//
// Stands in for `LUSolver()` / `CGSolver(JacobiLinearSolver(); rtol=1e-12)` at the
// `Algebra.solve!(x, ::LinearSolver, op, cache)` seam of the reference (src/ODEs/StageOperators.jl:55-88,
// src/ODEs/DAE.jl:241-274; the reference's own tutorial uses Jacobi-CG at rtol 1e-12,
// test/Tutorials/ShallowWater.jl:175-176).
//
// B200 design: the mass matrices of the named configurations are small enough to live in the 126 MB L2, so a
// solve is latency-bound, not bandwidth-bound.  The whole PCG iteration therefore runs inside ONE cooperative
// kernel (one CTA per SM slot, grid-wide barriers between phases) instead of ~6 launches per iteration;
// dot products are reduced through per-CTA partials that every CTA re-sums in the same fixed order, so the
// result is deterministic and needs neither atomics nor host round trips.
// Preconditioner: block Jacobi over the DOFs owned by one mesh entity (the k+1 moments of an RT facet, the
// interior moments of a cell, the p-1 DOFs of an edge ...).  For the moment-based RT_1 mass matrix this cuts
// the iteration count from ~90 (point Jacobi) to ~26, i.e. 3.5x fewer grid-wide barriers per solve.  The
// residual is double-buffered so that the block solve z = B^-1 r can be fused into the update sweep without
// an extra barrier.
#include <cooperative_groups.h>
#include "gg_common.cuh"
#include "gg_dist_host.cuh"

namespace cg = cooperative_groups;

namespace gg {

__global__ void k_spmv(int64_t n_rows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                       const double* __restrict__ vals, double alpha, const double* __restrict__ x, double beta,
                       double* __restrict__ y) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  double acc = 0.0;
  for (int32_t k = rowptr[r]; k < rowptr[r + 1]; ++k) acc = fma(vals[k], x[colind[k]], acc);
  y[r] = beta == 0.0 ? alpha * acc : fma(alpha, acc, beta * y[r]);
}

constexpr int MAX_BLOCK = 16;

// one thread per block: extract the diagonal block from the CSR rows, invert it (Gauss-Jordan; SPD blocks)
__global__ void k_block_inverse(int64_t n_blocks, const int32_t* __restrict__ blk_start, const uint8_t* __restrict__ blk_size,
                                const int32_t* __restrict__ blk_off, const int32_t* __restrict__ rowptr,
                                const int32_t* __restrict__ colind, const double* __restrict__ vals,
                                double* __restrict__ binv) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  const int s = blk_start[b], m = blk_size[b];
  double a[MAX_BLOCK][MAX_BLOCK], v[MAX_BLOCK][MAX_BLOCK];
  for (int i = 0; i < m; ++i) {
    for (int j = 0; j < m; ++j) { a[i][j] = 0.0; v[i][j] = i == j ? 1.0 : 0.0; }
    for (int32_t k = rowptr[s + i]; k < rowptr[s + i + 1]; ++k) {
      int c = colind[k] - s;
      if (c >= 0 && c < m) a[i][c] = vals[k];
    }
    if (a[i][i] == 0.0) a[i][i] = 1.0;  // empty row (e.g. a DOF without cells on a sub-mesh)
  }
  for (int k = 0; k < m; ++k) {
    double d = 1.0 / a[k][k];
    for (int j = 0; j < m; ++j) { a[k][j] *= d; v[k][j] *= d; }
    for (int i = 0; i < m; ++i) {
      if (i == k) continue;
      double f = a[i][k];
      for (int j = 0; j < m; ++j) { a[i][j] = fma(-f, a[k][j], a[i][j]); v[i][j] = fma(-f, v[k][j], v[i][j]); }
    }
  }
  double* out = binv + blk_off[b];
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) out[i * m + j] = v[i][j];
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
  // fixed-order reduction: warp shuffles then the first warp over the warp partials
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  }
  return v;  // valid in thread 0
}

__device__ __forceinline__ double grid_total(const double* partials, int nblocks, double* sh) {
  // every CTA sums all partials in the same order -> identical value everywhere
  double v = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) v += partials[i];
  v = block_sum(v, sh);
  __shared__ double bc;
  if (threadIdx.x == 0) bc = v;
  __syncthreads();
  return bc;
}

// x = A^-1 (scale_b * b) by block-Jacobi PCG starting from x = 0 (the RK stages need -r on the right-hand side).
// DIST: the matrix holds the rows of a rank's sub-model (owned rows complete thanks to the ghost cells); dot products run over
// the owned rows and are all-reduced through the peer windows, the owner -> ghost update of z rides on the same round (as in
// the condensed solver, gg_ebe.cu), so ghost p = z + beta p stays consistent without an exchange of its own: two rounds per
// iteration, no host, no NCCL.
template <bool DIST>
__global__ void __launch_bounds__(256)
k_pcg_persistent(int64_t n, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                 const double* __restrict__ vals, const int32_t* __restrict__ bstart, const uint8_t* __restrict__ bsize,
                 const int32_t* __restrict__ rowoff, const double* __restrict__ binv, const double* __restrict__ b,
                 double scale_b, double* __restrict__ x, double* ra, double* rb, double* __restrict__ z,
                 double* __restrict__ p, double* __restrict__ Ap, double* __restrict__ partials, double rtol,
                 int max_iter, double* __restrict__ stats, DistDev d) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sh[32];
  const int nb = gridDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  double* P0 = partials;           // [nb]
  double* P1 = partials + nb;      // [nb]
  double* P2 = partials + 2 * nb;  // [nb]
  double* r_old = ra;
  double* r_new = rb;
  unsigned long long epoch = DIST ? d.win[d.rank]->epoch : 0ull;

  // r = scale_b b, z = Binv r, p = z, x = 0; rz, bb
  double lrz = 0.0, lbb = 0.0;
  for (int64_t i = tid; i < n; i += nth) {
    if (DIST && !d.owned[i]) { x[i] = 0.0; r_old[i] = 0.0; r_new[i] = 0.0; z[i] = 0.0; p[i] = 0.0; Ap[i] = 0.0; continue; }
    const int s = bstart[i], m = bsize[i];
    const double* bi_row = binv + rowoff[i];
    double zi = 0.0;
    for (int j = 0; j < m; ++j) zi = fma(bi_row[j], scale_b * b[s + j], zi);
    const double ri = scale_b * b[i];
    x[i] = 0.0; r_old[i] = ri; z[i] = zi; p[i] = zi;
    lrz = fma(ri, zi, lrz);
    lbb = fma(ri, ri, lbb);
    if (DIST && d.owned[i] == 2) dist_push_row(d, epoch, i, zi);
  }
  lrz = block_sum(lrz, sh);
  lbb = block_sum(lbb, sh);
  if (threadIdx.x == 0) { P0[blockIdx.x] = lrz; P1[blockIdx.x] = lbb; }
  grid.sync();
  double rz = grid_total(P0, nb, sh);
  double bb = grid_total(P1, nb, sh);
  if (DIST) {
    double v[2] = {rz, bb};
    dist_allreduce<2>(grid, d, epoch, v);
    rz = v[0]; bb = v[1];
    dist_unpack(d, epoch, p, tid, nth, n);     // ghost p = ghost z
    grid.sync();
  }
  int it = 0;
  double rr = bb;
  if (bb > 0.0) {
    const double tol2 = rtol * rtol * bb;
    for (it = 1; it <= max_iter; ++it) {
      // Ap = A p ; pAp
      double lpap = 0.0;
      for (int64_t i = tid; i < n; i += nth) {
        if (DIST && !d.owned[i]) continue;
        double acc = 0.0;
        for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) acc = fma(vals[k], p[colind[k]], acc);
        Ap[i] = acc;
        lpap = fma(p[i], acc, lpap);
      }
      lpap = block_sum(lpap, sh);
      if (threadIdx.x == 0) P2[blockIdx.x] = lpap;
      grid.sync();
      double pap = grid_total(P2, nb, sh);
      if (DIST) {
        double v[1] = {pap};
        dist_allreduce<1>(grid, d, epoch, v);
        pap = v[0];
      }
      const double alpha = rz / pap;
      // x += alpha p ; r_new = r_old - alpha Ap ; z = Binv r_new (recomputing the block's residuals: no barrier needed)
      double lrz2 = 0.0, lrr = 0.0;
      for (int64_t i = tid; i < n; i += nth) {
        if (DIST && !d.owned[i]) continue;
        const int s = bstart[i], m = bsize[i];
        const double* bi_row = binv + rowoff[i];
        double zi = 0.0;
        for (int j = 0; j < m; ++j) zi = fma(bi_row[j], fma(-alpha, Ap[s + j], r_old[s + j]), zi);
        const double ri = fma(-alpha, Ap[i], r_old[i]);
        x[i] = fma(alpha, p[i], x[i]);
        r_new[i] = ri; z[i] = zi;
        lrz2 = fma(ri, zi, lrz2);
        lrr = fma(ri, ri, lrr);
        if (DIST && d.owned[i] == 2) dist_push_row(d, epoch, i, zi);
      }
      lrz2 = block_sum(lrz2, sh);
      lrr = block_sum(lrr, sh);
      if (threadIdx.x == 0) { P0[blockIdx.x] = lrz2; P1[blockIdx.x] = lrr; }
      grid.sync();
      { double* t = r_old; r_old = r_new; r_new = t; }
      double rz_new = grid_total(P0, nb, sh);
      rr = grid_total(P1, nb, sh);
      if (DIST) {
        double v[2] = {rz_new, rr};
        dist_allreduce<2>(grid, d, epoch, v);
        rz_new = v[0]; rr = v[1];
        if (rr > tol2) { dist_unpack(d, epoch, z, tid, nth, n); grid.sync(); }
      }
      if (rr <= tol2) break;  // uniform across the grid (and the ranks): every CTA computed the same rr
      const double beta = rz_new / rz;
      rz = rz_new;
      for (int64_t i = tid; i < n; i += nth) p[i] = fma(beta, p[i], z[i]);
      grid.sync();
    }
  }
  if (DIST) {
    // owner -> ghost update of the solution
    grid.sync();
    dist_push(d, epoch, x, tid, nth, n);
    grid.sync();
    double v[1] = {0.0};
    dist_allreduce<1>(grid, d, epoch, v);
    dist_unpack(d, epoch, x, tid, nth, n);
    grid.sync();
    if (tid == 0) d.win[d.rank]->epoch = epoch;
  }
  if (tid == 0) {
    stats[0] = (double)(it > max_iter ? max_iter : it);
    stats[1] = bb > 0.0 ? sqrt(rr / bb) : 0.0;
  }
}

void spmv(gg_csr* A, double alpha, const double* x, double beta, double* y) {
  if (A->n_rows == 0) return;
  k_spmv<<<(unsigned)((A->n_rows + 127) / 128), 128, 0, A->ctx->stream>>>(A->n_rows, A->d_rowptr.p, A->d_colind.p,
                                                                          A->d_vals.p, alpha, x, beta, y);
  A->ctx->launches++;
  GG_CUDA(cudaGetLastError());
}

// numerical_setup!: (re)compute the inverses of the diagonal blocks from the current matrix values
void solver_update(gg_solver* s) {
  gg_csr* A = s->A;
  if (s->n_blocks == 0) return;
  ProfScope ps(A->ctx, "k_block_inverse");
  k_block_inverse<<<(unsigned)((s->n_blocks + 63) / 64), 64, 0, A->ctx->stream>>>(
      s->n_blocks, s->d_blk_start.p, s->d_blk_size.p, s->d_blk_off.p, A->d_rowptr.p, A->d_colind.p, A->d_vals.p, s->binv.p);
  A->ctx->launches++;
  GG_CUDA(cudaGetLastError());
}

// enqueue one solve A x = scale_b * b (x overwritten, starts from 0); no host synchronisation
void solver_solve_async(gg_solver* s, const double* b, double scale_b, double* x) {
  gg_csr* A = s->A;
  int64_t n = A->n_rows;
  if (n == 0) return;
  if (s->mf_on) { solver_solve_mf_async(s, b, scale_b, x); return; }
  const int32_t* rp = A->d_rowptr.p;
  const int32_t* ci = A->d_colind.p;
  const double* va = A->d_vals.p;
  const int32_t* bs = s->bstart.p;
  const uint8_t* bz = s->bsize.p;
  const int32_t* ro = s->rowoff.p;
  const double* bi = s->binv.p;
  double *r = s->r.p, *r2 = s->r2.p, *z = s->z.p, *p = s->p.p, *Ap = s->Ap.p, *pa = s->partials.p, *stt = s->stats.p;
  double rtol = s->rtol;
  int mi = s->max_iter;
  // partitioned space with connected peer windows: the distributed variant (halo of z and all-reduces through peer memory)
  gg_space* sp = A->test;
  const bool multi = sp && sp->dist && sp->mesh->comm && sp->mesh->comm->world > 1;
  if (multi) GG_REQUIRE(A->test == A->trial, GG_ERR_UNSUPPORTED, "the distributed PCG needs a square operator on one partitioned space");
  DistDev d = multi ? dist_dev(sp) : DistDev();
  void* args[] = {&n, &rp, &ci, &va, &bs, &bz, &ro, &bi, &b, &scale_b, &x, &r, &r2, &z, &p, &Ap, &pa, &rtol, &mi, &stt, &d};
  // do not launch more CTAs than there is work for (fewer CTAs = cheaper grid barriers)
  int grid = s->grid;
  int64_t want = (n + s->block - 1) / s->block;
  if (want < grid) grid = (int)std::max<int64_t>(want, 1);
  ProfScope ps(A->ctx, "k_pcg_persistent");
  if (multi) GG_CUDA(cudaLaunchCooperativeKernel((void*)k_pcg_persistent<true>, dim3(grid), dim3(s->block), args, 0, A->ctx->stream));
  else GG_CUDA(cudaLaunchCooperativeKernel((void*)k_pcg_persistent<false>, dim3(grid), dim3(s->block), args, 0, A->ctx->stream));
  A->ctx->launches++;
}

}  // namespace gg

using namespace gg;

extern "C" {

int gg_csr_spmv(gg_csr* A, double alpha, gg_vec* x, double beta, gg_vec* y) {
  GG_API_BEGIN
  GG_REQUIRE(A && x && y, GG_ERR_INVALID, "null argument");
  GG_REQUIRE(gg_vec_size(x) == A->n_cols && gg_vec_size(y) == A->n_rows, GG_ERR_INVALID, "vector sizes do not match the matrix");
  GG_CUDA(cudaSetDevice(A->ctx->device));
  spmv(A, alpha, x->d.p, beta, y->d.p);
  finish_call(A->ctx);
  GG_API_END
}

int gg_solver_pcg_create(gg_csr* A, double rtol, int max_iter, gg_solver** out) {
  GG_API_BEGIN
  GG_REQUIRE(A && out, GG_ERR_INVALID, "null argument");
  GG_REQUIRE(A->n_rows == A->n_cols, GG_ERR_INVALID, "PCG needs a square matrix");
  gg_ctx* ctx = A->ctx;
  GG_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<gg_solver> s(new gg_solver);
  s->A = A;
  s->rtol = rtol > 0 ? rtol : 1e-12;
  s->max_iter = max_iter > 0 ? max_iter : 1000;
  int64_t n = A->n_rows;
  // entity blocks of the (test == trial) space; point Jacobi where the numbering is not ours
  std::vector<int32_t> bstart(n), rowoff(n), blk_start, blk_off;
  std::vector<uint8_t> bsize(n), blk_size;
  int64_t off = 0, covered = 0;
  auto add_block = [&](int64_t first, int m) {
    blk_start.push_back((int32_t)first); blk_size.push_back((uint8_t)m); blk_off.push_back((int32_t)off);
    for (int i = 0; i < m; ++i) { bstart[first + i] = (int32_t)first; bsize[first + i] = (uint8_t)m; rowoff[first + i] = (int32_t)(off + (int64_t)i * m); }
    off += (int64_t)m * m;
  };
  if (A->test == A->trial) {
    for (const auto& bc : A->test->blocks) {
      GG_REQUIRE(bc.start == covered, GG_ERR_INVALID, "inconsistent DOF block classes");
      for (int64_t g = 0; g < bc.count && covered < n; ++g) {
        int m = (int)std::min<int64_t>(bc.size, n - covered);  // a constrained space loses its last DOF
        if (m > MAX_BLOCK) {
          for (int i = 0; i < m; ++i) add_block(covered + i, 1);
        } else {
          add_block(covered, m);
        }
        covered += m;
      }
    }
  }
  for (; covered < n; ++covered) add_block(covered, 1);
  GG_REQUIRE(off < (int64_t)INT32_MAX, GG_ERR_UNSUPPORTED, "block preconditioner too large");
  cudaStream_t st = ctx->stream;
  s->n_blocks = (int64_t)blk_start.size();
  s->bstart.upload(bstart, st); s->rowoff.upload(rowoff, st); s->bsize.upload(bsize, st);
  s->d_blk_start.upload(blk_start, st); s->d_blk_off.upload(blk_off, st); s->d_blk_size.upload(blk_size, st);
  s->binv.alloc(off);
  s->r.alloc(n); s->r2.alloc(n); s->z.alloc(n); s->p.alloc(n); s->Ap.alloc(n);
  int per_sm = 0;
  GG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg_persistent<true>, s->block, 0));
  GG_REQUIRE(per_sm >= 1, GG_ERR_CUDA, "persistent PCG kernel does not fit on an SM");
  s->grid = ctx->sm_count * std::min(per_sm, 2);
  s->partials.alloc(3 * (size_t)s->grid);
  s->stats.alloc(8);
  GG_CUDA(cudaMemsetAsync(s->stats.p, 0, 8 * sizeof(double), st));
  solver_update(s.get());
  GG_CUDA(cudaStreamSynchronize(st));
  *out = s.release();
  GG_API_END
}

int gg_solver_update(gg_solver* s) {
  GG_API_BEGIN
  GG_REQUIRE(s, GG_ERR_INVALID, "null solver");
  GG_CUDA(cudaSetDevice(s->A->ctx->device));
  solver_update(s);
  finish_call(s->A->ctx);
  GG_API_END
}

int gg_solver_solve(gg_solver* s, gg_vec* b, gg_vec* x, int32_t* iters, double* rel_res) {
  GG_API_BEGIN
  GG_REQUIRE(s && b && x, GG_ERR_INVALID, "null argument");
  GG_REQUIRE(gg_vec_size(b) == s->A->n_rows && gg_vec_size(x) == s->A->n_rows, GG_ERR_INVALID, "vector sizes do not match the matrix");
  GG_CUDA(cudaSetDevice(s->A->ctx->device));
  solver_solve_async(s, b->d.p, 1.0, x->d.p);
  double st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (s->A->n_rows) s->stats.download(st, s->A->ctx->stream);
  if (iters) *iters = (int32_t)st[0];
  if (rel_res) *rel_res = st[1];
  if (s->A->test) comm_check(s->A->test->mesh->comm);   // a lost peer: GG_ERR_COMM, not a bogus "not converged"
  GG_REQUIRE(st[1] <= s->rtol * 1.0000001 || s->A->n_rows == 0, GG_ERR_NOCONV,
             "PCG did not converge: relative residual " + std::to_string(st[1]) + " after " + std::to_string((int)st[0]) + " iterations");
  GG_API_END
}

void gg_solver_destroy(gg_solver* s) { delete s; }

}  // extern "C"
