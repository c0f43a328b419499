// Matrix-free, persistent, single-launch block-Jacobi PCG for the mass solves of the explicit RK stages.
//
// Stands in, like gg_solver.cu, for `LUSolver()` / `CGSolver(JacobiLinearSolver(); rtol=1e-12)` at the
// `Algebra.solve!(x, ::LinearSolver, op, cache)` seam (src/ODEs/StageOperators.jl:55-88, src/ODEs/DAE.jl:241-274) for the
// two operators the transient drivers solve with at every stage:
//     RT_k mass      int (v.(g u)) / sqrtg                (test/Transient/TransientWaveEquation.jl:63, TransientShallowWater.jl:125)
//     Q_k mass       int c q w sqrtg, c at quadrature points (potential-vorticity block of jac_y, TransientShallowWater.jl:122)
//
// B200 design: on a 256x256-per-panel mesh the assembled RT_2 mass matrix is 3.4 GB of CSR, i.e. 0.5 ms of HBM traffic
// per CG iteration, while the operator itself is 5.7 kFMA per cell = 0.1 ms of FP64 pipe.  FP64 FLOPs are cheaper than
// HBM bytes here, so the solver never reads a matrix: every iteration re-integrates the form cell by cell (one thread per
// cell, all accumulators in registers, basis tabulations in __constant__ memory, tangents of the abscissae from the per-mesh
// tables, one MUFU.RSQ64H per quadrature point for the whole metric) and sums the element vectors per DOF in fixed cell
// order (atomic-free, deterministic).  The whole solve is ONE cooperative launch: 4 grid barriers per iteration, dot products
// through per-CTA partials that every CTA re-sums in the same order.  The preconditioner is the block-Jacobi of gg_solver.cu
// (blocks = DOFs of one mesh entity), built once from an assembled matrix (for the coefficient mass: unit coefficient; the
// scaling does not matter to CG).
#include <cooperative_groups.h>
#include "gg_common.cuh"
#include "gg_refel.cuh"

namespace cg = cooperative_groups;

namespace gg {

constexpr int MF_MAX_RT = 49 * 24 * 2;  // RT_2 at 7x7 points
constexpr int MF_MAX_SC = 49 * 16;      // Q_3 at 7x7 points
__constant__ double c_mf_w[MAX_NQ];
__constant__ double c_mf_rt[MF_MAX_RT];  // [Q][MU][2] reference RT basis
__constant__ double c_mf_sc[MF_MAX_SC];  // [Q][M] reference scalar basis (Gridap local order)

struct MfArgs {
  int64_t n, nc;
  int nq;
  CellGeo g;
  const int32_t* cell_dofs;
  const int8_t* signs;
  const double* qcoef;  // scalar mass: coefficient at the quadrature points, [c*Q + q]
  double* ev;           // [m][nc] element vectors
  const int32_t *vptr, *vsrc;
  const int32_t *bstart, *rowoff;
  const uint8_t* bsize;
  const double* binv;
  const double* b;
  double scale_b;
  double *x, *ra, *rb, *z, *p, *Ap, *partials;
  double rtol;
  int max_iter;
  double* stats;
};

__device__ __forceinline__ double mf_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(x, -(y * y), 1.0);
  return fma(fma(e, 0.375, 0.5), y * e, y);
}

// ev[:, c] = Mu_c p_c for RT_K: v_I.(g u)/sqrtg = (1/rho) v_I.[[sa, -ab], [-ab, sb]] u  (the radius cancels)
template <int K>
__device__ __forceinline__ void rt_mass_cell(const MfArgs& a, int64_t c) {
  constexpr int MU = 2 * (K + 1) * (K + 2);
  const int nq = a.nq;
  const double HX = a.g.hx[c], HY = a.g.hy[c];
  const double* __restrict__ ta = a.g.tx + (int64_t)a.g.ix[c] * nq;
  const double* __restrict__ tb = a.g.ty + (int64_t)a.g.iy[c] * nq;
  double ue[MU], acc[MU];
#pragma unroll
  for (int j = 0; j < MU; ++j) {
    ue[j] = (double)a.signs[c * MU + j] * a.p[a.cell_dofs[c * MU + j]];
    acc[j] = 0.0;
  }
  const double rxx = HX / HY, ryy = HY / HX;
#pragma unroll 1
  for (int q2 = 0; q2 < nq; ++q2) {
    const double b = tb[q2], sb = fma(b, b, 1.0), w2 = c_mf_w[q2];
#pragma unroll 1
    for (int q1 = 0; q1 < nq; ++q1) {
      const double* __restrict__ rv = c_mf_rt + (q1 + nq * q2) * (2 * MU);
      const double t = ta[q1], sa = fma(t, t, 1.0);
      const double wr = c_mf_w[q1] * w2 * mf_rsqrt(sa + b * b);
      // 4 + 4 independent accumulation chains (the FP64 pipe needs several DFMAs in flight per warp)
      double Uxp[4] = {0.0, 0.0, 0.0, 0.0}, Uyp[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int j = 0; j < MU; ++j) { Uxp[j & 3] = fma(rv[2 * j], ue[j], Uxp[j & 3]); Uyp[j & 3] = fma(rv[2 * j + 1], ue[j], Uyp[j & 3]); }
      const double Ux = (Uxp[0] + Uxp[1]) + (Uxp[2] + Uxp[3]), Uy = (Uyp[0] + Uyp[1]) + (Uyp[2] + Uyp[3]);
      // w HX HY (1/rho) [ (sa ux - ab uy) vx/HY + (sb uy - ab ux) vy/HX ],  ux = Ux/HY, uy = Uy/HX
      const double ab = t * b;
      const double cx = wr * (sa * rxx * Ux - ab * Uy), cy = wr * (sb * ryy * Uy - ab * Ux);
#pragma unroll
      for (int j = 0; j < MU; ++j) acc[j] = fma(cx, rv[2 * j], fma(cy, rv[2 * j + 1], acc[j]));
    }
  }
#pragma unroll
  for (int j = 0; j < MU; ++j) a.ev[(int64_t)j * a.nc + c] = (double)a.signs[c * MU + j] * acc[j];
}

// ev[:, c] = M_c(coef) p_c for scalar Q_K: coef phi_I phi_J sqrtg, sqrtg = r^2 sa sb / rho^3
template <int K>
__device__ __forceinline__ void scalar_mass_cell(const MfArgs& a, int64_t c) {
  constexpr int M = (K + 1) * (K + 1);
  const int nq = a.nq, Q = nq * nq;
  const double HX = a.g.hx[c], HY = a.g.hy[c];
  const double* __restrict__ ta = a.g.tx + (int64_t)a.g.ix[c] * nq;
  const double* __restrict__ tb = a.g.ty + (int64_t)a.g.iy[c] * nq;
  double ue[M], acc[M];
#pragma unroll
  for (int j = 0; j < M; ++j) {
    const int32_t d = a.cell_dofs[c * M + j];
    ue[j] = d >= 0 ? a.p[d] : 0.0;
    acc[j] = 0.0;
  }
  const double scale = a.g.radius * a.g.radius * HX * HY;
#pragma unroll 1
  for (int q2 = 0; q2 < nq; ++q2) {
    const double b = tb[q2], sb = fma(b, b, 1.0), w2 = c_mf_w[q2] * scale;
#pragma unroll 1
    for (int q1 = 0; q1 < nq; ++q1) {
      const int q = q1 + nq * q2;
      const double* __restrict__ sv = c_mf_sc + q * M;
      const double t = ta[q1], sa = fma(t, t, 1.0);
      const double ri = mf_rsqrt(sa + b * b);
      double wq = c_mf_w[q1] * w2 * sa * sb * ri * ri * ri;
      if (a.qcoef) wq *= a.qcoef[c * Q + q];
      double up[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int j = 0; j < M; ++j) up[j & 3] = fma(sv[j], ue[j], up[j & 3]);
      const double uh = wq * ((up[0] + up[1]) + (up[2] + up[3]));
#pragma unroll
      for (int j = 0; j < M; ++j) acc[j] = fma(uh, sv[j], acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < M; ++j) a.ev[(int64_t)j * a.nc + c] = acc[j];
}

__device__ __forceinline__ unsigned long long mf_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ double mf_block_sum(double v, double* sh) {
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

__device__ __forceinline__ double mf_grid_total(const double* partials, int nblocks, double* sh) {
  double v = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) v += partials[i];
  v = mf_block_sum(v, sh);
  __shared__ double bc;
  if (threadIdx.x == 0) bc = v;
  __syncthreads();
  return bc;
}

// x = A^-1 (scale_b b) from x = 0.  OP 0: RT_K mass, OP 1: scalar Q_K mass with coefficient.
template <int OP, int K>
__global__ void __launch_bounds__(128)
k_pcg_mf(const MfArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sh[32];
  const int nb = gridDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  const int64_t n = a.n;
  constexpr int UNR = 4;
  double* P0 = a.partials;
  double* P1 = a.partials + nb;
  double* P2 = a.partials + 2 * nb;
  double* r_old = a.ra;
  double* r_new = a.rb;
  double lrz = 0.0, lbb = 0.0;
  for (int64_t i = tid; i < n; i += nth) {
    const int s = a.bstart[i], m = a.bsize[i];
    const double* bi_row = a.binv + a.rowoff[i];
    double zi = 0.0;
    for (int j = 0; j < m; ++j) zi = fma(bi_row[j], a.scale_b * a.b[s + j], zi);
    const double ri = a.scale_b * a.b[i];
    a.x[i] = 0.0; r_old[i] = ri; a.z[i] = zi; a.p[i] = zi;
    lrz = fma(ri, zi, lrz);
    lbb = fma(ri, ri, lbb);
  }
  lrz = mf_block_sum(lrz, sh);
  lbb = mf_block_sum(lbb, sh);
  if (threadIdx.x == 0) { P0[blockIdx.x] = lrz; P1[blockIdx.x] = lbb; }
  grid.sync();
  double rz = mf_grid_total(P0, nb, sh);
  const double bb = mf_grid_total(P1, nb, sh);
  int it = 0;
  double rr = bb;
  unsigned long long tA = 0, tB = 0, tC = 0;
  if (bb > 0.0) {
    const double tol2 = a.rtol * a.rtol * bb;
    for (it = 1; it <= a.max_iter; ++it) {
      // element vectors of A p (cell-wise quadrature, no matrix)
      unsigned long long t0 = mf_now();
      for (int64_t c = tid; c < a.nc; c += nth) {
        if (OP == 0) rt_mass_cell<K>(a, c); else scalar_mass_cell<K>(a, c);
      }
      grid.sync();
      unsigned long long t1 = mf_now();
      tA += t1 - t0;
      // Ap = sum of the element contributions in cell order ; pAp
      // (UNR rows per thread in flight: the persistent grid has few warps per SM, memory parallelism must come from ILP)
      double lpap = 0.0;
      for (int64_t i0 = tid; i0 < n; i0 += UNR * nth) {
        int32_t lo[UNR], hi[UNR], s0[UNR], s1[UNR];
        double pv[UNR], acc[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int64_t i = i0 + u * nth;
          const bool ok = i < n;
          lo[u] = ok ? a.vptr[i] : 0; hi[u] = ok ? a.vptr[i + 1] : 0; pv[u] = ok ? a.p[i] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) { s0[u] = lo[u] < hi[u] ? a.vsrc[lo[u]] : -1; s1[u] = lo[u] + 1 < hi[u] ? a.vsrc[lo[u] + 1] : -1; }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          acc[u] = s0[u] >= 0 ? a.ev[s0[u]] : 0.0;
          if (s1[u] >= 0) acc[u] += a.ev[s1[u]];
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          for (int32_t k = lo[u] + 2; k < hi[u]; ++k) acc[u] += a.ev[a.vsrc[k]];
          const int64_t i = i0 + u * nth;
          if (i < n) { a.Ap[i] = acc[u]; lpap = fma(pv[u], acc[u], lpap); }
        }
      }
      lpap = mf_block_sum(lpap, sh);
      if (threadIdx.x == 0) P2[blockIdx.x] = lpap;
      grid.sync();
      t0 = mf_now();
      tB += t0 - t1;
      const double alpha = rz / mf_grid_total(P2, nb, sh);
      double lrz2 = 0.0, lrr = 0.0;
      for (int64_t i0 = tid; i0 < n; i0 += UNR * nth) {
        int sb[UNR], mb[UNR], ro[UNR];
        double ri[UNR], zi[UNR], xi[UNR], pi[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int64_t i = i0 + u * nth;
          const bool ok = i < n;
          sb[u] = ok ? a.bstart[i] : 0; mb[u] = ok ? a.bsize[i] : 0; ro[u] = ok ? a.rowoff[i] : 0;
          ri[u] = ok ? fma(-alpha, a.Ap[i], r_old[i]) : 0.0;
          xi[u] = ok ? a.x[i] : 0.0; pi[u] = ok ? a.p[i] : 0.0;
          zi[u] = 0.0;
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const double* bi_row = a.binv + ro[u];
          const int s = sb[u];
          for (int j = 0; j < mb[u]; ++j) zi[u] = fma(bi_row[j], fma(-alpha, a.Ap[s + j], r_old[s + j]), zi[u]);
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int64_t i = i0 + u * nth;
          if (i < n) {
            a.x[i] = fma(alpha, pi[u], xi[u]);
            r_new[i] = ri[u]; a.z[i] = zi[u];
            lrz2 = fma(ri[u], zi[u], lrz2);
            lrr = fma(ri[u], ri[u], lrr);
          }
        }
      }
      lrz2 = mf_block_sum(lrz2, sh);
      lrr = mf_block_sum(lrr, sh);
      if (threadIdx.x == 0) { P0[blockIdx.x] = lrz2; P1[blockIdx.x] = lrr; }
      grid.sync();
      t1 = mf_now();
      tC += t1 - t0;
      { double* t = r_old; r_old = r_new; r_new = t; }
      const double rz_new = mf_grid_total(P0, nb, sh);
      rr = mf_grid_total(P1, nb, sh);
      if (rr <= tol2) break;
      const double beta = rz_new / rz;
      rz = rz_new;
      for (int64_t i0 = tid; i0 < n; i0 += UNR * nth) {
        double pv[UNR], zv[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int64_t i = i0 + u * nth;
          pv[u] = i < n ? a.p[i] : 0.0; zv[u] = i < n ? a.z[i] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int64_t i = i0 + u * nth;
          if (i < n) a.p[i] = fma(beta, pv[u], zv[u]);
        }
      }
      grid.sync();
    }
  }
  if (tid == 0) {
    a.stats[0] = (double)(it > a.max_iter ? a.max_iter : it);
    a.stats[1] = bb > 0.0 ? sqrt(rr / bb) : 0.0;
    // accumulated ns spent in the cell phase, the DOF gather + dot, and the update + preconditioner phases
    a.stats[2] += (double)tA; a.stats[3] += (double)tB; a.stats[4] += (double)tC; a.stats[5] += (double)(it > a.max_iter ? a.max_iter : it);
  }
}

template <int OP, int K>
static void* mf_kernel() { return (void*)k_pcg_mf<OP, K>; }

static void* mf_select(int op, int order) {
  if (op == 0) {
    switch (order) { case 0: return mf_kernel<0, 0>(); case 1: return mf_kernel<0, 1>(); case 2: return mf_kernel<0, 2>(); }
  } else {
    switch (order) { case 0: return mf_kernel<1, 0>(); case 1: return mf_kernel<1, 1>(); case 2: return mf_kernel<1, 2>(); case 3: return mf_kernel<1, 3>(); }
  }
  return nullptr;
}

// whether (space kind, order, nq) has a matrix-free operator (tables must fit the __constant__ blocks)
bool mf_supported(int kind, int order, int nq) {
  int Q = nq * nq;
  if (nq > MAX_NQ) return false;
  if (kind == GG_SPACE_RT) return order >= 0 && order <= 2 && Q * 4 * (order + 1) * (order + 2) <= MF_MAX_RT;
  return order >= 0 && order <= 3 && Q * (order + 1) * (order + 1) <= MF_MAX_SC;
}

// make the tables of (kind, order, nq) resident in __constant__ memory (stream-ordered; no-op when already there)
static void mf_load_tables(gg_ctx* ctx, int kind, int order, int nq) {
  const bool rt = kind == GG_SPACE_RT;
  int64_t key = ((int64_t)order << 8) | nq;
  int64_t& cur = rt ? ctx->mf_key_rt : ctx->mf_key_sc;
  cudaStream_t st = ctx->stream;
  if (ctx->mf_key_w != nq) {
    std::vector<double> xi(nq), w(nq);
    gauss_legendre_01(nq, xi.data(), w.data());
    GG_CUDA(cudaStreamSynchronize(st));  // earlier launches may still read the old tables
    GG_CUDA(cudaMemcpyToSymbol(c_mf_w, w.data(), nq * sizeof(double)));
    ctx->mf_key_w = nq;
    ctx->mf_key_rt = ctx->mf_key_sc = -1;  // tabulations depend on the abscissae
  }
  if (cur == key) return;
  auto T = get_tab(ctx, kind, order, nq);
  GG_CUDA(cudaStreamSynchronize(st));
  if (rt) GG_CUDA(cudaMemcpyToSymbolAsync(c_mf_rt, T->val.p, T->val.n * sizeof(double), 0, cudaMemcpyDeviceToDevice, ctx->stream));
  else GG_CUDA(cudaMemcpyToSymbolAsync(c_mf_sc, T->val.p, T->val.n * sizeof(double), 0, cudaMemcpyDeviceToDevice, ctx->stream));
  cur = key;
}

// switch solver `s` (created on an assembled matrix of the same space, which fixes the preconditioner) to the matrix-free
// operator of its space; qcoef (scalar mass only) may be NULL
void solver_set_matrix_free(gg_solver* s, int degree, const double* qcoef) {
  gg_space* sp = s->A->test;
  gg_ctx* ctx = s->A->ctx;
  int nq = num_points_1d(degree);
  GG_REQUIRE(s->A->test == s->A->trial, GG_ERR_INVALID, "matrix-free PCG needs test == trial");
  GG_REQUIRE(mf_supported(sp->kind, sp->order, nq), GG_ERR_UNSUPPORTED, "no matrix-free operator for this space / quadrature");
  build_vec_map(sp);
  s->mf_on = true; s->mf_nq = nq; s->mf_qcoef = qcoef;
  s->mf_ev.alloc((size_t)sp->ndofs_loc * std::max<int64_t>(sp->mesh->n_cells, 1));
  void* fn = mf_select(sp->kind == GG_SPACE_RT ? 0 : 1, sp->order);
  GG_REQUIRE(fn, GG_ERR_UNSUPPORTED, "no matrix-free kernel for this order");
  int per_sm = 0;
  GG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 128, 0));
  GG_REQUIRE(per_sm >= 1, GG_ERR_CUDA, "matrix-free PCG kernel does not fit on an SM");
  s->mf_grid = ctx->sm_count * per_sm;
  if ((size_t)3 * s->mf_grid > s->partials.n) s->partials.alloc(3 * (size_t)s->mf_grid);
}

void solver_solve_mf_async(gg_solver* s, const double* b, double scale_b, double* x) {
  gg_csr* A = s->A;
  gg_space* sp = A->test;
  gg_ctx* ctx = A->ctx;
  if (A->n_rows == 0) return;
  mf_load_tables(ctx, sp->kind, sp->order, s->mf_nq);
  MfArgs a;
  a.n = A->n_rows; a.nc = sp->mesh->n_cells; a.nq = s->mf_nq;
  a.g = cell_geo(sp->mesh, s->mf_nq);
  a.cell_dofs = sp->d_cell_dofs.p; a.signs = sp->d_cell_signs.p; a.qcoef = s->mf_qcoef; a.ev = s->mf_ev.p;
  a.vptr = sp->d_vptr.p; a.vsrc = sp->d_vsrc.p;
  a.bstart = s->bstart.p; a.rowoff = s->rowoff.p; a.bsize = s->bsize.p; a.binv = s->binv.p;
  a.b = b; a.scale_b = scale_b; a.x = x; a.ra = s->r.p; a.rb = s->r2.p; a.z = s->z.p; a.p = s->p.p; a.Ap = s->Ap.p;
  a.partials = s->partials.p; a.rtol = s->rtol; a.max_iter = s->max_iter; a.stats = s->stats.p;
  void* fn = mf_select(sp->kind == GG_SPACE_RT ? 0 : 1, sp->order);
  int grid = s->mf_grid;
  int64_t want = (std::max(a.n, a.nc) + 127) / 128;
  if (want < grid) grid = (int)std::max<int64_t>(want, 1);
  void* args[] = {&a};
  ProfScope ps(ctx, "k_pcg_mf");
  GG_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(128), args, 0, ctx->stream));
  ctx->launches++;
}

}  // namespace gg
