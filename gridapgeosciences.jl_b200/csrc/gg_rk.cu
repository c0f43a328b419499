// Explicit Runge-Kutta stepping of the linear wave equation on the device.
//
// Replaces, for RT_p x DG_p on the cubed sphere (test/Transient/TransientWaveEquation.jl:28-77), the stock
// Gridap.ODEs `ode_march!` of EXRungeKutta with a constant mass matrix:
//     for each stage i:  x_i = x_0 + dt sum_j A_ij k_j ;  M k_i = -res(x_i)
//     x_F = x_0 + dt sum_i b_i k_i                         (same stage algebra as src/ODEs/RungeKuttaEX.jl:78-122)
// with  mass((du,dp),(v,q)) = int v.(g du)/sqrtg + int q dp sqrtg   and   res((u,p),(v,q)) = int q div(u) - int p div(v).
//
// B200 design
//   * state, slopes and all operators stay resident in HBM/L2; a step enqueues a fixed sequence of launches with no
//     host synchronisation (the PCG iteration count is decided inside the persistent kernel);
//   * the stage residual is matrix-free and owner-computes: ONE launch, one thread per global DOF, no scatter and
//     no atomics.  For the contravariant Piola map on an axis-aligned cell the Jacobian determinant cancels in
//     int q div(u), so the cell-wise quadrature collapses to the reference matrix Bhat[I][J] = sum_q w_q q_I divhat_J
//     (same quadrature rule as the reference, evaluated once);
//   * the DG mass block is block diagonal: its cell blocks are inverted once (batched Gauss-Jordan, one thread per
//     cell) and applied as a dense mat-vec; the RT block is solved by the persistent Jacobi-PCG (gg_solver.cu).
#include <cstring>
#include <cstdlib>
#include "gg_common.cuh"
#include "gg_refel.cuh"

#include "gg_rk.cuh"
#include "gg_dist_host.cuh"

namespace gg {

static inline unsigned nblk(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// one thread per cell: invert the m x m DG mass block held entry-major in ebuf[(i*m+j)*nc + c]
template <int M>
__global__ void k_invert_blocks(int64_t nc, const double* __restrict__ ebuf, double* __restrict__ inv) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  double a[M][M], b[M][M];
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) {
      a[i][j] = ebuf[(int64_t)(i * M + j) * nc + c];
      b[i][j] = i == j ? 1.0 : 0.0;
    }
  // SPD block: Gauss-Jordan without pivoting is stable
#pragma unroll
  for (int k = 0; k < M; ++k) {
    double d = 1.0 / a[k][k];
#pragma unroll
    for (int j = 0; j < M; ++j) { a[k][j] *= d; b[k][j] *= d; }
#pragma unroll
    for (int i = 0; i < M; ++i) {
      if (i == k) continue;
      double f = a[i][k];
#pragma unroll
      for (int j = 0; j < M; ++j) { a[i][j] = fma(-f, a[k][j], a[i][j]); b[i][j] = fma(-f, b[k][j], b[i][j]); }
    }
  }
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) inv[(int64_t)(i * M + j) * nc + c] = b[i][j];
}

// matrix-free wave residual, one thread per global DOF of [u; p]
//   u-dof d: r = - sum_{(c,J) touching d} s_cJ sum_I Bhat[I][J] p[c*mq+I]
//   p-dof (c,I): r = sum_J Bhat[I][J] s_cJ u[dof(c,J)]
__global__ void __launch_bounds__(128)
k_wave_residual(int64_t nu, int64_t np, int64_t nc, int mq, int mu, const int32_t* __restrict__ vptr,
                const int32_t* __restrict__ vsrc, const int32_t* __restrict__ udofs, const int8_t* __restrict__ signs,
                const double* __restrict__ Bhat, const double* __restrict__ x, double* __restrict__ r) {
  int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= nu + np) return;
  if (d < nu) {
    double acc = 0.0;
    for (int32_t e = vptr[d]; e < vptr[d + 1]; ++e) {
      int32_t src = vsrc[e];
      int J = (int)(src / nc);
      int64_t c = src - (int64_t)J * nc;
      double s = 0.0;
      for (int I = 0; I < mq; ++I) s = fma(Bhat[I * mu + J], x[nu + c * mq + I], s);
      acc = fma(-(double)signs[c * mu + J], s, acc);
    }
    r[d] = acc;
  } else {
    int64_t e = d - nu;
    int64_t c = e / mq;
    int I = (int)(e - c * mq);
    double acc = 0.0;
    for (int J = 0; J < mu; ++J) acc = fma(Bhat[I * mu + J] * (double)signs[c * mu + J], x[udofs[c * mu + J]], acc);
    r[d] = acc;
  }
}

// k_p = -MpInv r_p (one thread per DG dof)
__global__ void k_apply_dg_inverse(int64_t nc, int mq, const double* __restrict__ inv, const double* __restrict__ rp,
                                   double* __restrict__ kp) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * mq) return;
  int64_t c = t / mq;
  int i = (int)(t - c * mq);
  double acc = 0.0;
  for (int j = 0; j < mq; ++j) acc = fma(inv[(int64_t)(i * mq + j) * nc + c], rp[c * mq + j], acc);
  kp[t] = -acc;
}

// out = x0 + sum_j coef[j] * k_j  (up to 4 slopes)
__global__ void k_combine(int64_t n, const double* x0, int ns, double c0, const double* __restrict__ k0,
                          double c1, const double* __restrict__ k1, double c2, const double* __restrict__ k2, double c3,
                          const double* __restrict__ k3, double* out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = x0[i];
  if (ns > 0) v = fma(c0, k0[i], v);
  if (ns > 1) v = fma(c1, k1[i], v);
  if (ns > 2) v = fma(c2, k2[i], v);
  if (ns > 3) v = fma(c3, k3[i], v);
  out[i] = v;
}

__global__ void k_accumulate_stats(const double* __restrict__ solver_stats, double rtol, double* __restrict__ accum) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    accum[0] += solver_stats[0]; accum[1] += 1.0;
    if (!(solver_stats[1] <= rtol * 1.0000001)) accum[2] += 1.0;   // stopped at max_iter (or NaN): reported by gg_rk_step
  }
}

static void rk_residual_async(gg_rk* rk, const double* x, double* r) {
  gg_ctx* ctx = rk->ctx;
  int64_t n = rk_nx(rk);
  if (n == 0) return;
  k_wave_residual<<<nblk(n, 128), 128, 0, ctx->stream>>>(rk->nu, rk->np, rk->mesh->n_cells, rk->mq, rk->mu, rk->V->d_vptr.p,
                                                        rk->V->d_vsrc.p, rk->V->d_cell_dofs.p, rk->V->d_cell_signs.p,
                                                        rk->Bhat.p, x, r);
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
}

void rk_accumulate_stats(gg_rk* rk, const double* solver_stats, double rtol) {
  k_accumulate_stats<<<1, 32, 0, rk->ctx->stream>>>(solver_stats, rtol, rk->iter_accum.p);
  rk->ctx->launches++;
}

// x = polynomial extrapolation of the last L solutions of this solve to the current step (the same-stage solutions of
// consecutive steps are equally spaced in time): sum_j (-1)^(j+1) C(L, j) h_j, i.e. h1, 2 h1 - h2, 3 h1 - 3 h2 + h3, ...
struct HistPtrs {
  const double* h[GG_HIST_LEVELS];
  double c[GG_HIST_LEVELS];
};

__global__ void k_extrapolate(int64_t n, int levels, HistPtrs H, double* __restrict__ x) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = 0.0;
  for (int j = levels - 1; j >= 0; --j) v = fma(H.c[j], H.h[j][i], v);
  x[i] = v;
}

void rk_mass_solve(gg_rk* rk, int which, const double* b, double scale, double* x, int slot, const double* be) {
  EbeSolver* e = which == 0 ? rk->ebe_u : rk->ebe_q;
  if (e) {
    gg_rk_hist* H = nullptr;
    const int64_t n = which == 0 ? rk->nu : rk->nr;
    cudaStream_t st = rk->ctx->stream;
    if (slot >= 0 && rk->warm_start && rk->extrap_order > 0 && n > 0) {
      H = &rk->hist[slot];
      // all levels of a slot are allocated at its first solve: no cudaMalloc (a device-wide synchronisation of unpredictable
      // cost) ever happens in the steady state of a march
      const int R = rk->extrap_order;   // ring size = levels in use
      if ((int64_t)H->h[0].n != n)
        for (int l = 0; l < R; ++l) H->h[l].alloc(n);
      const int levels = std::min(H->count, R);
      if (levels >= 1) {
        HistPtrs P;
        double binom = 1.0;
        for (int j = 0; j < levels; ++j) {
          binom = binom * (levels - j) / (j + 1);                                   // C(levels, j + 1)
          P.h[j] = H->h[(H->head + R - j) % R].p;
          P.c[j] = (j % 2 == 0 ? 1.0 : -1.0) * binom;
        }
        k_extrapolate<<<nblk(n, 256), 256, 0, st>>>(n, levels, P, x);
        rk->ctx->launches++;
      }
    }
    ebe_solve_async(e, b, scale, x, rk->warm_start, be);
    rk_accumulate_stats(rk, e->stats.p, e->rtol);
    if (H) {
      // newest slot moves forward; the oldest level is overwritten
      const int R = rk->extrap_order;
      H->head = (H->head + 1) % R;
      GG_CUDA(cudaMemcpyAsync(H->h[H->head].p, x, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
      if (H->count < R) H->count++;
    }
  } else {
    gg_solver* s = which == 0 ? rk->solver : rk->solver_q;
    GG_REQUIRE(!be, GG_ERR_INVALID, "element-form right-hand sides need the element-by-element solver");
    solver_solve_async(s, b, scale, x);
    rk_accumulate_stats(rk, s->stats.p, s->rtol);
  }
}

void rk_combine(gg_rk* rk, int ns, const double* co, double* out) {
  int64_t n = rk_nx(rk);
  if (n == 0) return;
  auto kp = [&](int j) { return j < (int)rk->k.size() ? rk->k[j].p : nullptr; };
  k_combine<<<nblk(n, 256), 256, 0, rk->ctx->stream>>>(n, rk->x.p, ns, co[0], kp(0), co[1], kp(1), co[2], kp(2), co[3], kp(3), out);
  rk->ctx->launches++;
}

std::vector<double> build_bhat(int order, int degree) {
  int nq = num_points_1d(degree);
  GG_REQUIRE(nq <= MAX_NQ, GG_ERR_UNSUPPORTED, "quadrature degree too high");
  std::vector<double> xi(nq), w(nq);
  gauss_legendre_01(nq, xi.data(), w.data());
  RTRef R = build_rt_ref(order);
  std::vector<double> pts((size_t)nq * nq * 2), val, div;
  for (int q2 = 0; q2 < nq; ++q2)
    for (int q1 = 0; q1 < nq; ++q1) { pts[2 * (q1 + nq * q2)] = xi[q1]; pts[2 * (q1 + nq * q2) + 1] = xi[q2]; }
  R.tabulate(nq * nq, pts.data(), val, div);
  int nb1 = order + 1, mq = nb1 * nb1, mu = R.ndofs;
  std::vector<int32_t> l2x = quad_local_to_lex(order);
  std::vector<double> N(nq * nb1), D(nq * nb1);
  for (int q = 0; q < nq; ++q) lagrange_1d(order, xi[q], &N[q * nb1], &D[q * nb1]);
  std::vector<double> B((size_t)mq * mu, 0.0);
  for (int q2 = 0; q2 < nq; ++q2)
    for (int q1 = 0; q1 < nq; ++q1) {
      int q = q1 + nq * q2;
      for (int I = 0; I < mq; ++I) {
        double qI = N[q1 * nb1 + l2x[I] % nb1] * N[q2 * nb1 + l2x[I] / nb1];
        for (int J = 0; J < mu; ++J) B[(size_t)I * mu + J] += w[q1] * w[q2] * qI * div[(size_t)q * mu + J];
      }
    }
  return B;
}

void build_dg_mass_inverse(gg_rk* rk) {
  gg_ctx* ctx = rk->ctx;
  int64_t nc = rk->mesh->n_cells;
  gg_form_args fa;
  std::memset(&fa, 0, sizeof(fa));
  fa.test = rk->Q; fa.trial = rk->Q; fa.degree = rk->degree;
  int layout = 0;
  const double* mp = element_matrices_device(GG_FORM_MASS_SCALAR, &fa, &layout);
  GG_REQUIRE(layout == 0, GG_ERR_INVALID, "unexpected element buffer layout");
  rk->MpInv.alloc((size_t)rk->mq * rk->mq * nc);
  if (nc) {
    unsigned nb = nblk(nc, 64);
    switch (rk->mq) {
      case 1: k_invert_blocks<1><<<nb, 64, 0, ctx->stream>>>(nc, mp, rk->MpInv.p); break;
      case 4: k_invert_blocks<4><<<nb, 64, 0, ctx->stream>>>(nc, mp, rk->MpInv.p); break;
      case 9: k_invert_blocks<9><<<nb, 64, 0, ctx->stream>>>(nc, mp, rk->MpInv.p); break;
      default: throw Error(GG_ERR_UNSUPPORTED, "DG order > 2 in the RK steppers");
    }
    ctx->launches++;
    GG_CUDA(cudaGetLastError());
  }
}

static void rk_step_async(gg_rk* rk) {
  if (rk->problem == GG_PROBLEM_SHALLOW_WATER) { swe_step_async(rk); return; }
  if (rk->problem == GG_PROBLEM_ADVECTION) { adv_step_async(rk); return; }
  if (rk->problem == GG_PROBLEM_BOUSSINESQ) { bq_step_async(rk); return; }
  gg_ctx* ctx = rk->ctx;
  cudaStream_t st = ctx->stream;
  int64_t n = rk_nx(rk);
  int s = rk->n_stages;
  auto kp = [&](int j) { return j < s ? rk->k[j].p : nullptr; };
  for (int i = 0; i < s; ++i) {
    const double* xs = rk->x.p;
    if (i > 0) {
      double co[4] = {0, 0, 0, 0};
      for (int j = 0; j < i && j < 4; ++j) co[j] = rk->dt * rk->A[i * s + j];
      k_combine<<<nblk(n, 256), 256, 0, st>>>(n, rk->x.p, std::min(i, 4), co[0], kp(0), co[1], kp(1), co[2], kp(2), co[3], kp(3), rk->xi.p);
      ctx->launches++;
      xs = rk->xi.p;
    }
    rk_residual_async(rk, xs, rk->r.p);
    // slope: M k = -r
    rk_mass_solve(rk, 0, rk->r.p, -1.0, rk->k[i].p, i);
    if (rk->np) {
      k_apply_dg_inverse<<<nblk(rk->np, 128), 128, 0, st>>>(rk->mesh->n_cells, rk->mq, rk->MpInv.p, rk->r.p + rk->nu, rk->k[i].p + rk->nu);
      ctx->launches++;
    }
  }
  double co[4] = {0, 0, 0, 0};
  for (int j = 0; j < s && j < 4; ++j) co[j] = rk->dt * rk->b[j];
  k_combine<<<nblk(n, 256), 256, 0, st>>>(n, rk->x.p, std::min(s, 4), co[0], kp(0), co[1], kp(1), co[2], kp(2), co[3], kp(3), rk->x.p);
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
}

}  // namespace gg

using namespace gg;

extern "C" {

void gg_rk_destroy(gg_rk* rk) {
  if (!rk) return;
  swe_destroy(rk);
  ebe_destroy(rk->ebe_u);
  gg_solver_destroy(rk->solver);
  gg_csr_destroy(rk->Mu);
  gg_space_destroy(rk->V);
  gg_space_destroy(rk->Q);
  delete rk;
}

int gg_rk_create(gg_mesh* m, const gg_rk_args* a, gg_rk** out) {
  gg_rk* raw = nullptr;
  GG_API_BEGIN
  GG_REQUIRE(m && a && out, GG_ERR_INVALID, "null argument");
  GG_REQUIRE(a->problem == GG_PROBLEM_WAVE || a->problem == GG_PROBLEM_SHALLOW_WATER || a->problem == GG_PROBLEM_ADVECTION ||
                 a->problem == GG_PROBLEM_THERMAL_SHALLOW_WATER || a->problem == GG_PROBLEM_BOUSSINESQ, GG_ERR_UNSUPPORTED,
             "unknown problem id (GG_PROBLEM_WAVE, _SHALLOW_WATER, _ADVECTION, _THERMAL_SHALLOW_WATER and _BOUSSINESQ are implemented)");
  GG_REQUIRE(a->n_stages >= 1 && a->n_stages <= 4 && a->A && a->b && a->c, GG_ERR_INVALID, "tableau must have 1..4 stages");
  GG_REQUIRE(a->order >= 0 && a->order <= 2, GG_ERR_UNSUPPORTED, "the RK steppers support RT_p x DG_p with p <= 2");
  GG_REQUIRE(a->dt > 0, GG_ERR_INVALID, "dt must be positive");
  GG_REQUIRE(m->dim == (a->problem == GG_PROBLEM_BOUSSINESQ ? 3 : 2), GG_ERR_UNSUPPORTED,
             "the wave / shallow-water / advection steppers run on the 2-D cubed sphere, the Boussinesq stepper on the 3-D shell");
  gg_ctx* ctx = m->ctx;
  GG_CUDA(cudaSetDevice(ctx->device));
  raw = new gg_rk;
  gg_rk* rk = raw;
  rk->ctx = ctx; rk->mesh = m; rk->problem = a->problem; rk->order = a->order; rk->degree = a->degree;
  rk->n_stages = a->n_stages; rk->dt = a->dt;
  int s = a->n_stages;
  rk->A.assign(a->A, a->A + s * s); rk->b.assign(a->b, a->b + s); rk->c.assign(a->c, a->c + s);
  for (int i = 0; i < s; ++i)
    for (int j = i; j < s; ++j) GG_REQUIRE(rk->A[i * s + j] == 0.0, GG_ERR_INVALID, "tableau is not explicit");
  auto chk = [](int code) { if (code != GG_OK) throw Error(code, gg_last_error(nullptr)); };
  if (a->problem == GG_PROBLEM_BOUSSINESQ) {
    bq_create(rk, a);
    *out = raw;
    return GG_OK;
  }
  if (a->problem == GG_PROBLEM_ADVECTION) {
    // scalar DG state only: no RT space, no global solve
    chk(gg_space_builtin(m, GG_SPACE_DG, a->order, GG_CONSTRAINT_NONE, &rk->Q));
    rk->nu = 0; rk->np = rk->Q->n_free; rk->mq = rk->Q->ndofs_loc; rk->mu = 0;
    build_dg_mass_inverse(rk);
    int64_t n = rk->np;
    rk->x.alloc(n); rk->xi.alloc(n); rk->r.alloc(n);
    rk->k.resize(s);
    for (int i = 0; i < s; ++i) rk->k[i].alloc(n);
    rk->iter_accum.alloc(3);
    GG_CUDA(cudaMemsetAsync(rk->iter_accum.p, 0, 3 * sizeof(double), ctx->stream));
    if (n) GG_CUDA(cudaMemsetAsync(rk->x.p, 0, n * sizeof(double), ctx->stream));
    adv_create(rk, a);
    GG_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = raw;
    return GG_OK;
  }
  chk(gg_space_builtin(m, GG_SPACE_RT, a->order, GG_CONSTRAINT_NONE, &rk->V));
  chk(gg_space_builtin(m, GG_SPACE_DG, a->order, GG_CONSTRAINT_NONE, &rk->Q));
  rk->nu = rk->V->n_free; rk->np = rk->Q->n_free; rk->mq = rk->Q->ndofs_loc; rk->mu = rk->V->ndofs_loc;
  // mass solver: GG_PCG = ebe (default: statically condensed element-by-element), mf (matrix-free re-integration), csr
  const char* mode_env = getenv("GG_PCG");
  const std::string mode = mode_env ? mode_env : "ebe";
  // every mass solve starts from the previous solution held by its output vector (the same stage of the previous step,
  // or the previous diagnostics) unless GG_WARM_START=0
  const char* ws = getenv("GG_WARM_START");
  rk->warm_start = !(ws && ws[0] == '0');
  if (const char* ex = getenv("GG_EXTRAPOLATE")) rk->extrap_order = std::max(0, std::min(GG_HIST_LEVELS, atoi(ex)));
  gg_form_args fa;
  std::memset(&fa, 0, sizeof(fa));
  fa.test = rk->V; fa.trial = rk->V; fa.degree = a->degree;
  const bool partitioned = m->plan != nullptr;
  if (mode == "ebe" && ebe_supported(rk->V)) {
    // RT mass: element matrices once, condensed and kept per cell (never assembled)
    rk->ebe_u = ebe_create(rk->V, a->rtol > 0 ? a->rtol : 1e-12, 2000);
    int layout = 0;
    const double* eb = element_matrices_device(GG_FORM_MASS_RT, &fa, &layout);
    GG_REQUIRE(layout == 0, GG_ERR_INVALID, "unexpected element buffer layout");
    ebe_set_matrix(rk->ebe_u, eb);
  } else {
    GG_REQUIRE(!partitioned, GG_ERR_UNSUPPORTED, "partitioned models need the element-by-element mass solver (GG_PCG=ebe)");
    // assembled RT mass matrix and block-Jacobi PCG on it
    chk(gg_pattern(rk->V, rk->V, &rk->Mu));
    chk(gg_assemble_matrix(GG_FORM_MASS_RT, &fa, rk->Mu, 0));
    chk(gg_solver_pcg_create(rk->Mu, a->rtol > 0 ? a->rtol : 1e-12, 2000, &rk->solver));
  }
  // DG mass blocks -> inverses
  build_dg_mass_inverse(rk);
  // reference div matrix, same rule as Measure(Omega, degree)
  rk->Bhat.upload(build_bhat(a->order, a->degree), ctx->stream);
  build_vec_map(rk->V);
  const bool thermal = a->problem == GG_PROBLEM_THERMAL_SHALLOW_WATER;
  if (thermal) { rk->problem = GG_PROBLEM_SHALLOW_WATER; rk->thermal = true; rk->nB = rk->np; }
  int64_t n = rk_nx(rk);
  rk->x.alloc(n); rk->xi.alloc(n); rk->r.alloc(n);
  rk->k.resize(s);
  for (int i = 0; i < s; ++i) {
    rk->k[i].alloc(n);
    // the slopes double as initial guesses of the warm-started mass solves: they must start out consistent across ranks
    if (n) GG_CUDA(cudaMemsetAsync(rk->k[i].p, 0, n * sizeof(double), ctx->stream));
  }
  if (n) {
    GG_CUDA(cudaMemsetAsync(rk->xi.p, 0, n * sizeof(double), ctx->stream));
    GG_CUDA(cudaMemsetAsync(rk->r.p, 0, n * sizeof(double), ctx->stream));
  }
  rk->iter_accum.alloc(3);
  GG_CUDA(cudaMemsetAsync(rk->iter_accum.p, 0, 3 * sizeof(double), ctx->stream));
  if (n) GG_CUDA(cudaMemsetAsync(rk->x.p, 0, n * sizeof(double), ctx->stream));
  if (a->problem == GG_PROBLEM_SHALLOW_WATER || thermal) {
    swe_create(rk, a);
    if (thermal) tswe_create(rk);
  } else if (mode == "mf" && mf_supported(GG_SPACE_RT, a->order, num_points_1d(a->degree))) {
    solver_set_matrix_free(rk->solver, a->degree, nullptr);
  }
  GG_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = raw;
  raw = nullptr;
  }
  catch (...) {
    gg_rk_destroy(raw);
    try { throw; }
    catch (const gg::Error& e) { gg::set_last_error(e.what()); return e.code; }
    catch (const std::exception& e) { gg::set_last_error(e.what()); return GG_ERR_INVALID; }
  }
  return GG_OK;
}

int gg_rk_info(gg_rk* rk, int64_t* n_u, int64_t* n_p) {
  GG_API_BEGIN
  GG_REQUIRE(rk, GG_ERR_INVALID, "null stepper");
  if (n_u) *n_u = rk->nu;
  if (n_p) *n_p = rk->np + rk->nB;   // thermal shallow water: [p; B]
  GG_API_END
}

int gg_rk_diag_info(gg_rk* rk, int64_t* n_q, int64_t* n_F, int64_t* n_Phi) {
  GG_API_BEGIN
  GG_REQUIRE(rk, GG_ERR_INVALID, "null stepper");
  bool swe = rk->problem == GG_PROBLEM_SHALLOW_WATER;
  if (n_q) *n_q = swe ? rk->nr : 0;
  if (n_F) *n_F = swe ? rk->nu : 0;
  if (n_Phi) *n_Phi = swe ? rk->np + rk->nB : 0;   // thermal shallow water: [Phi; b]
  GG_API_END
}

int gg_rk_set_state(gg_rk* rk, const double* x) {
  GG_API_BEGIN
  GG_REQUIRE(rk && x, GG_ERR_INVALID, "null argument");
  GG_CUDA(cudaSetDevice(rk->ctx->device));
  int64_t n = rk_nx(rk);
  if (n) GG_CUDA(cudaMemcpyAsync(rk->x.p, x, n * sizeof(double), cudaMemcpyHostToDevice, rk->ctx->stream));
  rk->y_valid = false;
  for (auto& h : rk->hist) h.second.count = 0;   // a new state: the solution histories no longer extrapolate to it
  GG_CUDA(cudaStreamSynchronize(rk->ctx->stream));
  GG_API_END
}

int gg_rk_get_state(gg_rk* rk, double* x) {
  GG_API_BEGIN
  GG_REQUIRE(rk && x, GG_ERR_INVALID, "null argument");
  GG_CUDA(cudaSetDevice(rk->ctx->device));
  rk->x.download(x, rk->ctx->stream);
  GG_API_END
}

double* gg_rk_state_device_ptr(gg_rk* rk) { return rk ? rk->x.p : nullptr; }

int gg_rk_step(gg_rk* rk, int32_t n_steps) {
  GG_API_BEGIN
  GG_REQUIRE(rk && n_steps >= 0, GG_ERR_INVALID, "bad argument");
  GG_CUDA(cudaSetDevice(rk->ctx->device));
  for (int32_t i = 0; i < n_steps; ++i) rk_step_async(rk);
  GG_CUDA(cudaStreamSynchronize(rk->ctx->stream));
  // a march that lost a peer or a mass solve that stopped at max_iter must not return GG_OK with a garbage state
  comm_check(rk->mesh->comm);
  if (rk->iter_accum.n >= 3) {
    double h[3];
    rk->iter_accum.download(h, rk->ctx->stream);
    if (h[2] > rk->noconv_seen) {
      const int64_t fresh = (int64_t)(h[2] - rk->noconv_seen);
      rk->noconv_seen = h[2];
      throw Error(GG_ERR_NOCONV, std::to_string(fresh) + " mass / diagnostic solve(s) of this march stopped at max_iter without reaching rtol");
    }
  }
  GG_API_END
}

int gg_rk_residual(gg_rk* rk, const double* x, double* r) {
  GG_API_BEGIN
  GG_REQUIRE(rk && x && r, GG_ERR_INVALID, "null argument");
  GG_CUDA(cudaSetDevice(rk->ctx->device));
  int64_t n = rk_nx(rk);
  if (n) GG_CUDA(cudaMemcpyAsync(rk->xi.p, x, n * sizeof(double), cudaMemcpyHostToDevice, rk->ctx->stream));
  if (rk->problem == GG_PROBLEM_SHALLOW_WATER) {
    rk->y_valid = false;
    swe_diagnostics_async(rk, rk->xi.p, rk->y.p);
    swe_residual_async(rk, rk->xi.p, rk->y.p, rk->y.p, rk->r.p, nullptr);
  } else if (rk->problem == GG_PROBLEM_BOUSSINESQ) {
    bq_residual_async(rk, rk->xi.p, rk->r.p, nullptr);
  } else if (rk->problem == GG_PROBLEM_ADVECTION) {
    adv_slope_async(rk, rk->xi.p, nullptr, rk->r.p);
  } else {
    rk_residual_async(rk, rk->xi.p, rk->r.p);
  }
  rk->r.download(r, rk->ctx->stream);
  GG_API_END
}

int gg_rk_get_diagnostics(gg_rk* rk, double* y) {
  GG_API_BEGIN
  GG_REQUIRE(rk && y, GG_ERR_INVALID, "null argument");
  GG_REQUIRE(rk->problem == GG_PROBLEM_SHALLOW_WATER, GG_ERR_INVALID, "this stepper has no diagnostic variables");
  GG_CUDA(cudaSetDevice(rk->ctx->device));
  rk->y.download(y, rk->ctx->stream);
  GG_API_END
}

int gg_rk_diagnostics(gg_rk* rk, const double* x, double* y) {
  GG_API_BEGIN
  GG_REQUIRE(rk && x && y, GG_ERR_INVALID, "null argument");
  GG_REQUIRE(rk->problem == GG_PROBLEM_SHALLOW_WATER, GG_ERR_INVALID, "this stepper has no diagnostic variables");
  GG_CUDA(cudaSetDevice(rk->ctx->device));
  int64_t n = rk_nx(rk);
  if (n) GG_CUDA(cudaMemcpyAsync(rk->xi.p, x, n * sizeof(double), cudaMemcpyHostToDevice, rk->ctx->stream));
  rk->y_valid = false;
  swe_diagnostics_async(rk, rk->xi.p, rk->y.p);
  rk->y.download(y, rk->ctx->stream);
  GG_API_END
}

int gg_rk_swe_residual(gg_rk* rk, const double* x, const double* y, const double* q0, double* r) {
  GG_API_BEGIN
  GG_REQUIRE(rk && x && y && q0 && r, GG_ERR_INVALID, "null argument");
  GG_REQUIRE(rk->problem == GG_PROBLEM_SHALLOW_WATER, GG_ERR_INVALID, "not a shallow-water stepper");
  GG_CUDA(cudaSetDevice(rk->ctx->device));
  cudaStream_t st = rk->ctx->stream;
  int64_t n = rk_nx(rk), ny = rk->nr + rk_nx(rk);
  if (n) GG_CUDA(cudaMemcpyAsync(rk->xi.p, x, n * sizeof(double), cudaMemcpyHostToDevice, st));
  rk->y_valid = false;
  if (ny) GG_CUDA(cudaMemcpyAsync(rk->y.p, y, ny * sizeof(double), cudaMemcpyHostToDevice, st));
  if (rk->nr) GG_CUDA(cudaMemcpyAsync(rk->q0.p, q0, rk->nr * sizeof(double), cudaMemcpyHostToDevice, st));
  swe_residual_async(rk, rk->xi.p, rk->y.p, rk->q0.p, rk->r.p, nullptr);
  rk->r.download(r, st);
  GG_API_END
}

int gg_rk_solver_phase_ns(gg_rk* rk, int which, double* out) {
  GG_API_BEGIN
  GG_REQUIRE(rk && out, GG_ERR_INVALID, "null argument");
  // which = 0 / 1: iteration phases of the RT / potential-vorticity solver; which = 10 / 11: set-up phases of the condensed
  // solver (cell pre-pass, row pass, first preconditioner pass, interior recovery), each followed by the count in out[3] / ...
  const bool setup = which >= 10;
  if (setup) which -= 10;
  gg_solver* s = which == 0 ? rk->solver : rk->solver_q;
  EbeSolver* e = which == 0 ? rk->ebe_u : rk->ebe_q;
  for (int i = 0; i < 4; ++i) out[i] = 0.0;
  if (!s && !e) return GG_OK;
  GG_CUDA(cudaSetDevice(rk->ctx->device));
  double h[16] = {0};
  if (e) e->stats.download(h, rk->ctx->stream);
  else { double h8[8]; s->stats.download(h8, rk->ctx->stream); for (int i = 0; i < 8; ++i) h[i] = h8[i]; }
  if (setup) {
    const double ns = h[12] > 0 ? h[12] : 1.0;
    for (int i = 0; i < 4; ++i) out[i] = h[8 + i] / ns;   // ns per solve
    return GG_OK;
  }
  for (int i = 0; i < 4; ++i) out[i] = h[2 + i];
  if (getenv("GG_COMM_STATS") && h[5] > 0)
    fprintf(stderr, "[gg_rk solver %d] per iteration: cells %.1f  gather+dot %.1f  update+precond %.1f us, of which all-reduce rounds %.1f "
                    "and thread 0's own phase-C loop %.1f us\n", which, 1e-3 * h[2] / h[5], 1e-3 * h[3] / h[5], 1e-3 * h[4] / h[5], 1e-3 * h[6] / h[5],
            1e-3 * h[7] / h[5]);
  GG_API_END
}

int gg_rk_stats(gg_rk* rk, int64_t* pcg_iters_total, int64_t* solves_total) {
  GG_API_BEGIN
  GG_REQUIRE(rk, GG_ERR_INVALID, "null stepper");
  GG_CUDA(cudaSetDevice(rk->ctx->device));
  double h[3] = {0, 0, 0};
  rk->iter_accum.download(h, rk->ctx->stream);
  if (pcg_iters_total) *pcg_iters_total = (int64_t)h[0];
  if (solves_total) *solves_total = (int64_t)h[1];
  GG_API_END
}

}  // extern "C"
