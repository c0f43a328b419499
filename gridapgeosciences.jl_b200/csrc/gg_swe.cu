// Rotating shallow water as a DAE with explicit Runge-Kutta stepping, on the device.
//
// Replaces, for X = RT_k x DG_k (prognostic u, p) and Y = CG_{k+1} x RT_k x DG_k (diagnostic q, F, Phi) on the cubed
// sphere, the reference's DAE-aware `ode_march!` (src/ODEs/RungeKuttaEX.jl:48-134) applied to the forms of
// test/Transient/TransientShallowWater.jl:113-139:
//   diagnostics (src/ODEs/DAE.jl:241-274: x = 0, b = res_y, A = jac_y, one linear solve A y = -b; jac_y is block
//   diagonal, so the three blocks are solved separately):
//     int q p w sqrtg            = int f w sqrtg + int ((R u).ginv).grad(w) sqrtg          (CG_{k+1}, matrix depends on p)
//     int F.(g v)/sqrtg          = int p u.(g v)/sqrtg                                    (RT_k mass, constant)
//     int Phi psi sqrtg          = int g_r (p + b) psi sqrtg + int 1/2 u.(g u) psi/sqrtg  (DG_k mass, block diagonal)
//   stage slope (src/ODEs/StageOperators.jl:55-88: x = 0, b = res_x + mass(0), A = mass, M k = -b):
//     r_u = int [q - tau (q - q0)/dt - tau (u.grad q)/sqrtg] (R F).v - int Phi div(v),   r_p = int r div(F)
// `q0` follows SURVEY.md F11: GG_Q0_ALIAS reproduces the reference (q0 and q share one array, the (q - q0) term
// vanishes), GG_Q0_START_OF_STEP keeps the potential vorticity of the beginning of the step.
//
// B200 design
//   * state, slopes, diagnostics and every operator stay resident in HBM/L2; a step enqueues a fixed sequence of
//     launches with no host synchronisation (PCG iteration counts are decided inside the persistent solver kernel);
//   * the nonlinear cell integrals are sum-factorised with 4 or 8 lanes of a warp per cell, one row of the tensor rule per lane
//     (gg_swe_lanes.cuh; the rules the drivers use by default) -- or, for any other rule, one thread per cell with all accumulators
//     in FP64 registers (gg_swe_sf.cuh and the direct kernels below); tangents of the quadrature abscissae come from the per-mesh
//     tables (no tan() in the kernels), 1-D basis tables are immediate / warp-uniform constant operands;
//   * terms that are linear with constant coefficients need no metric (the Jacobian determinant cancels in int q div(u) under the
//     contravariant Piola map): the one-thread kernels apply the reference matrix Bhat, the lane kernels integrate them on the
//     same rows with plain reference weights;
//   * the Bernoulli potential and the depth slope are cell-local (DG): the inverse mass block is applied inside the
//     kernel that computed the right-hand side; RT and CG right-hand sides go through the atomic-free DOF gather.
#include <cstring>
#include <cstdlib>
#include "gg_rk.cuh"
#include "gg_refel.cuh"
#include "gg_geom.cuh"
#include "gg_dist_host.cuh"

namespace gg {

static inline unsigned nblk(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// Basis tabulations of the resident (order, nq) pair in __constant__ memory (warp-uniform operands: no L1 traffic).
// Sized for RT_2 / DG_2 / CG_3 at 7x7 points (41 KB); larger rules fall back to the same tables in global memory.
constexpr int SWE_MAXQ = 49;
#ifndef SWE_MIN_BLOCKS
#define SWE_MIN_BLOCKS 6
#endif
__constant__ double c_swe_w[MAX_NQ];
__constant__ double c_swe_rt[SWE_MAXQ * 24 * 2];  // [Q][MU][2] reference RT basis
__constant__ double c_swe_dg[SWE_MAXQ * 9];       // [Q][MP]
__constant__ double c_swe_cg[SWE_MAXQ * 16];      // [Q][MR]
__constant__ double c_swe_cgd[SWE_MAXQ * 16 * 2]; // [Q][MR][2] reference gradients

struct SweTabs {
  const double* w;       // 1-D quadrature weights [nq]
  const double* rt_val;  // [Q][MU][2] reference RT basis
  const double* dg_val;  // [Q][MP]
  const double* cg_val;  // [Q][MR]
  const double* cg_der;  // [Q][MR][2] reference gradients
};

// table accessors: CONST selects the __constant__ copies
template <bool CONST> __device__ __forceinline__ const double* tab_w(const SweTabs& T) { return CONST ? c_swe_w : T.w; }
template <bool CONST> __device__ __forceinline__ const double* tab_rt(const SweTabs& T) { return CONST ? c_swe_rt : T.rt_val; }
template <bool CONST> __device__ __forceinline__ const double* tab_dg(const SweTabs& T) { return CONST ? c_swe_dg : T.dg_val; }
template <bool CONST> __device__ __forceinline__ const double* tab_cg(const SweTabs& T) { return CONST ? c_swe_cg : T.cg_val; }
template <bool CONST> __device__ __forceinline__ const double* tab_cgd(const SweTabs& T) { return CONST ? c_swe_cgd : T.cg_der; }

// 1/sqrt(x), x in [1,3]: hardware seed + one cubic step (< 1 ulp)
__device__ __forceinline__ double swe_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(x, -(y * y), 1.0);
  return fma(fma(e, 0.375, 0.5), y * e, y);
}

// sqrt(det g) = r^2 sa sb / rho^3 (src/CellData/CellFields.jl:80-82 evaluated in closed form)
__device__ __forceinline__ double csphere_sqrtg(double a, double b, double r) {
  const double sa = fma(a, a, 1.0), sb = fma(b, b, 1.0);
  const double rho2 = sa + b * b;
  return r * r * sa * sb / (rho2 * sqrt(rho2));
}

// Both cell kernels run one thread per cell in two passes over the quadrature points so that the live register state
// stays below the 255-register limit for RT_2 / CG_3: pass 1 interpolates the fields and stages two coefficients per
// point (the chart-space vector that multiplies the RT test functions) in shared memory, laid out [point][component]
// [thread] (conflict-free); pass 2 only accumulates the RT element vector.  Rules with more than 7x7 points stage the
// coefficients in a global scratch instead ([point][component][cell], coalesced).
struct Stage2 {
  double* base;    // this thread's column
  int64_t stride;  // distance between consecutive (point, component) slots
  __device__ __forceinline__ double& at(int q, int comp) { return base[(int64_t)(2 * q + comp) * stride]; }
};
__device__ __forceinline__ Stage2 stage2_of(double* gscratch, int64_t nc, int64_t c) {
  extern __shared__ double swe_smem[];
  Stage2 s;
  if (gscratch) { s.base = gscratch + c; s.stride = nc; }
  else { s.base = swe_smem + threadIdx.x; s.stride = blockDim.x; }
  return s;
}

}  // namespace gg
#include "gg_swe_sf.cuh"
#include "gg_swe_lanes.cuh"
namespace gg {

// Right-hand sides of the three diagnostic equations, one thread per cell.
//   pq_out [c*Q+q]      depth at the quadrature points (coefficient of the q mass matrix)
//   eq [MR][nc], eF [MU][nc]  entry-major element vectors of the q and F equations
//   Phi [c*MP+i]        Bernoulli potential (DG mass block inverted in place)
template <int K, bool CONST>
__global__ void __launch_bounds__(64, SWE_MIN_BLOCKS)
k_swe_diag(int64_t nc, int nq, CellGeo g, SweTabs T, const int32_t* __restrict__ vdofs, const int8_t* __restrict__ vsigns,
           const double* __restrict__ u, const double* __restrict__ p, const double* __restrict__ fq,
           const double* __restrict__ bq, double gravity, const double* __restrict__ MpInv, double* __restrict__ pq_out,
           double* __restrict__ eq, double* __restrict__ eF, double* __restrict__ Phi, double* __restrict__ gscratch) {
  constexpr int MU = 2 * (K + 1) * (K + 2), MP = (K + 1) * (K + 1), MR = (K + 2) * (K + 2);
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  Stage2 st = stage2_of(gscratch, nc, c);
  const double HX = g.hx[c], HY = g.hy[c], iHX = 1.0 / HX, iHY = 1.0 / HY;
  const double* __restrict__ ta = g.tx + (int64_t)g.ix[c] * nq;
  const double* __restrict__ tb = g.ty + (int64_t)g.iy[c] * nq;
  const int Q = nq * nq;
  const double r2 = g.radius * g.radius;
  const double* __restrict__ tw = tab_w<CONST>(T);
  {
    double ue[MU], pe[MP], aq[MR], aP[MP];
#pragma unroll
    for (int j = 0; j < MU; ++j) ue[j] = (double)vsigns[c * MU + j] * u[vdofs[c * MU + j]];
#pragma unroll
    for (int i = 0; i < MP; ++i) { pe[i] = p[c * MP + i]; aP[i] = 0.0; }
#pragma unroll
    for (int i = 0; i < MR; ++i) aq[i] = 0.0;
#pragma unroll 1
    for (int q2 = 0; q2 < nq; ++q2) {
      const double b = tb[q2], sb = fma(b, b, 1.0);
#pragma unroll 1
      for (int q1 = 0; q1 < nq; ++q1) {
        const int q = q1 + nq * q2;
        // metric from one rsqrt: sqrtg = r^2 sa sb / rho^3, g / sqrtg = [[sa, -ab], [-ab, sb]] / rho,
        // ginv sqrtg = [[sb, ab], [ab, sa]] / rho
        const double t = ta[q1], sa = fma(t, t, 1.0), ab = t * b;
        const double ri = swe_rsqrt(sa + b * b);
        const double sqrtg = r2 * sa * sb * ri * ri * ri;
        const double* __restrict__ rv = tab_rt<CONST>(T) + (size_t)q * MU * 2;
        double Uxp[2] = {0.0, 0.0}, Uyp[2] = {0.0, 0.0};
#pragma unroll
        for (int j = 0; j < MU; ++j) { Uxp[j & 1] = fma(rv[2 * j], ue[j], Uxp[j & 1]); Uyp[j & 1] = fma(rv[2 * j + 1], ue[j], Uyp[j & 1]); }
        const double ux = (Uxp[0] + Uxp[1]) * iHY, uy = (Uyp[0] + Uyp[1]) * iHX;  // contravariant Piola, axis-aligned cell
        const double* __restrict__ dv = tab_dg<CONST>(T) + (size_t)q * MP;
        double ph = 0.0;
#pragma unroll
        for (int i = 0; i < MP; ++i) ph = fma(dv[i], pe[i], ph);
        pq_out[c * Q + q] = ph;
        const double wq = tw[q1] * tw[q2] * HX * HY;
        const double wr = wq * ri;
        // vorticity equation: f w sqrtg + ((R u).ginv).grad(w) sqrtg, R u = (-u_y, u_x)
        {
          const double f0 = wq * sqrtg * fq[c * Q + q];
          const double fx = wr * (ux * ab - uy * sb) * iHX, fy = wr * (ux * sa - uy * ab) * iHY;
          const double* __restrict__ cv = tab_cg<CONST>(T) + (size_t)q * MR;
          const double* __restrict__ cd = tab_cgd<CONST>(T) + (size_t)q * MR * 2;
#pragma unroll
          for (int i = 0; i < MR; ++i) aq[i] = fma(f0, cv[i], fma(fx, cd[2 * i], fma(fy, cd[2 * i + 1], aq[i])));
        }
        // mass flux p u.(g v)/sqrtg, (g u)/sqrtg = (sa ux - ab uy, sb uy - ab ux) / rho: staged for pass 2
        const double gux = sa * ux - ab * uy, guy = sb * uy - ab * ux;
        st.at(q, 0) = wr * ph * gux * iHY;
        st.at(q, 1) = wr * ph * guy * iHX;
        // Bernoulli potential: g_r (p + b) sqrtg + 1/2 u.(g u)/sqrtg
        {
          const double bt = bq ? bq[c * Q + q] : 0.0;
          const double cP = wq * (gravity * (ph + bt) * sqrtg + 0.5 * ri * (ux * gux + uy * guy));
#pragma unroll
          for (int i = 0; i < MP; ++i) aP[i] = fma(cP, dv[i], aP[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < MR; ++i) eq[(int64_t)i * nc + c] = aq[i];
#pragma unroll
    for (int i = 0; i < MP; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < MP; ++j) acc = fma(MpInv[(int64_t)(i * MP + j) * nc + c], aP[j], acc);
      Phi[c * MP + i] = acc;
    }
  }
  // pass 2: RT element vector of the mass-flux equation
  double aF[MU];
#pragma unroll
  for (int j = 0; j < MU; ++j) aF[j] = 0.0;
#pragma unroll 1
  for (int q = 0; q < Q; ++q) {
    const double* __restrict__ rv = tab_rt<CONST>(T) + (size_t)q * MU * 2;
    const double fx = st.at(q, 0), fy = st.at(q, 1);
#pragma unroll
    for (int j = 0; j < MU; ++j) aF[j] = fma(fx, rv[2 * j], fma(fy, rv[2 * j + 1], aF[j]));
  }
#pragma unroll
  for (int j = 0; j < MU; ++j) eF[(int64_t)j * nc + c] = (double)vsigns[c * MU + j] * aF[j];
}

// Stage residual, one thread per cell.  eru [MU][nc] entry-major element vector of r_u; rp [c*MP+i] = r_p (DG, cell
// local); kp (optional) = -Mp^-1 r_p, the depth slope.
template <int K, bool ALIAS, bool CONST>
__global__ void __launch_bounds__(64, SWE_MIN_BLOCKS)
k_swe_stage(int64_t nc, int nq, CellGeo g, SweTabs T, const int32_t* __restrict__ vdofs, const int8_t* __restrict__ vsigns,
            const int32_t* __restrict__ rdofs, const double* __restrict__ u, const double* __restrict__ qv,
            const double* __restrict__ q0v, const double* __restrict__ Fv, const double* __restrict__ Phi,
            const double* __restrict__ Bhat, const double* __restrict__ MpInv, double tau, double tau_over_dt,
            double* __restrict__ eru, double* __restrict__ rp, double* __restrict__ kp, double* __restrict__ gscratch) {
  constexpr int MU = 2 * (K + 1) * (K + 2), MP = (K + 1) * (K + 1), MR = (K + 2) * (K + 2);
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  Stage2 st = stage2_of(gscratch, nc, c);
  const double HX = g.hx[c], HY = g.hy[c], iHX = 1.0 / HX, iHY = 1.0 / HY;
  const double* __restrict__ ta = g.tx + (int64_t)g.ix[c] * nq;
  const double* __restrict__ tb = g.ty + (int64_t)g.iy[c] * nq;
  const int Q = nq * nq;
  double Fe[MU];
#pragma unroll
  for (int j = 0; j < MU; ++j) Fe[j] = (double)vsigns[c * MU + j] * Fv[vdofs[c * MU + j]];
  // pass 1: coefficient of (R F).v at every point
  {
    double ue[MU], qe[MR], qc[ALIAS ? 1 : MR];
#pragma unroll
    for (int j = 0; j < MU; ++j) ue[j] = (double)vsigns[c * MU + j] * u[vdofs[c * MU + j]];
#pragma unroll
    for (int i = 0; i < MR; ++i) {
      const int32_t d = rdofs[c * MR + i];
      qe[i] = qv[d];
      // q - tau (q - q0)/dt = (1 - tau/dt) q + (tau/dt) q0
      if (!ALIAS) qc[i] = fma(tau_over_dt, q0v[d] - qe[i], qe[i]);
    }
    const double ir2 = 1.0 / (g.radius * g.radius);
    const double* __restrict__ tw = tab_w<CONST>(T);
#pragma unroll 1
    for (int q2 = 0; q2 < nq; ++q2) {
      const double b = tb[q2], isb = 1.0 / fma(b, b, 1.0);
#pragma unroll 1
      for (int q1 = 0; q1 < nq; ++q1) {
        const int q = q1 + nq * q2;
        // 1 / sqrtg = rho^3 / (r^2 sa sb)
        const double t = ta[q1], sa = fma(t, t, 1.0), rho2 = sa + b * b;
        const double isqrtg = ir2 * isb * rho2 * rho2 * swe_rsqrt(rho2) / sa;
        const double* __restrict__ rv = tab_rt<CONST>(T) + (size_t)q * MU * 2;
        double Ux = 0.0, Uy = 0.0, Fx = 0.0, Fy = 0.0;
#pragma unroll
        for (int j = 0; j < MU; ++j) {
          Ux = fma(rv[2 * j], ue[j], Ux); Uy = fma(rv[2 * j + 1], ue[j], Uy);
          Fx = fma(rv[2 * j], Fe[j], Fx); Fy = fma(rv[2 * j + 1], Fe[j], Fy);
        }
        const double* __restrict__ cv = tab_cg<CONST>(T) + (size_t)q * MR;
        const double* __restrict__ cd = tab_cgd<CONST>(T) + (size_t)q * MR * 2;
        double qq = 0.0, qx = 0.0, qy = 0.0;
#pragma unroll
        for (int i = 0; i < MR; ++i) {
          qq = fma(cv[i], ALIAS ? qe[i] : qc[i], qq);
          qx = fma(cd[2 * i], qe[i], qx);
          qy = fma(cd[2 * i + 1], qe[i], qy);
        }
        // u.grad(q) in chart coordinates: (Ux/HY)(qx/HX) + (Uy/HX)(qy/HY)
        const double ugq = (Ux * qx + Uy * qy) * iHX * iHY;
        const double coef = tw[q1] * tw[q2] * HX * HY * (qq - tau * ugq * isqrtg);
        // (R F).v_I with F = (Fx/HY, Fy/HX), v_I = (vx/HY, vy/HX): (-Fy/HX)(vx/HY) + (Fx/HY)(vy/HX)
        st.at(q, 0) = -coef * Fy * iHX * iHY;
        st.at(q, 1) = coef * Fx * iHX * iHY;
      }
    }
  }
  // linear, constant-coefficient parts through the reference divergence matrix
  double ru[MU];
  {
    double pe[MP], r[MP];
#pragma unroll
    for (int i = 0; i < MP; ++i) { pe[i] = Phi[c * MP + i]; r[i] = 0.0; }
#pragma unroll
    for (int j = 0; j < MU; ++j) {
      double acc = 0.0;
#pragma unroll
      for (int i = 0; i < MP; ++i) {
        const double bij = Bhat[i * MU + j];
        acc = fma(bij, pe[i], acc);
        r[i] = fma(bij, Fe[j], r[i]);
      }
      ru[j] = -acc;
    }
#pragma unroll
    for (int i = 0; i < MP; ++i) rp[c * MP + i] = r[i];
    if (kp) {
#pragma unroll
      for (int i = 0; i < MP; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < MP; ++j) acc = fma(MpInv[(int64_t)(i * MP + j) * nc + c], r[j], acc);
        kp[c * MP + i] = -acc;
      }
    }
  }
  // pass 2: RT element vector
#pragma unroll 1
  for (int q = 0; q < Q; ++q) {
    const double* __restrict__ rv = tab_rt<CONST>(T) + (size_t)q * MU * 2;
    const double fx = st.at(q, 0), fy = st.at(q, 1);
#pragma unroll
    for (int j = 0; j < MU; ++j) ru[j] = fma(fx, rv[2 * j], fma(fy, rv[2 * j + 1], ru[j]));
  }
#pragma unroll
  for (int j = 0; j < MU; ++j) eru[(int64_t)j * nc + c] = (double)vsigns[c * MU + j] * ru[j];
}

__global__ void k_copy(int64_t n, const double* __restrict__ a, double* __restrict__ b) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) b[i] = a[i];
}

// dynamic shared memory of the two-pass cell kernels: 2 coefficients per point and thread; 0 = stage through global memory
static size_t swe_stage_smem(gg_rk* rk, int threads) {
  const size_t bytes = (size_t)2 * rk->nq1 * rk->nq1 * threads * sizeof(double);
  // Staging through shared memory only below GG_SWE_STAGE_KB per CTA (default 36): at 7x7 points the buffer is 50 KB per
  // 64-thread CTA, which caps the residency at 4 CTAs per SM while the registers allow 6; staging through a global scratch
  // ([point][component][cell], coalesced, 1.5 KB more traffic per cell on kernels that use 17 % of the bandwidth) measured
  // 2.35 -> 2.14 ms (k_swe_diag) and 2.47 -> 2.09 ms (k_swe_stage) per step at 256^2, order 2.
  static const size_t limit = [] { const char* e = getenv("GG_SWE_STAGE_KB"); return (size_t)(e ? atoi(e) : 36) * 1024; }();
  if (bytes > limit) {
    if (rk->stage2.n == 0) rk->stage2.alloc((size_t)2 * rk->nq1 * rk->nq1 * std::max<int64_t>(rk->mesh->n_cells, 1));
    return 0;
  }
  return bytes;
}

template <typename Kern, typename... Args>
static void swe_launch(Kern kern, unsigned nb, size_t smem, cudaStream_t st, Args... args) {
  if (smem > 48 * 1024) GG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<nb, 64, smem, st>>>(args...);
}

// global-memory tables; *use_const: the same tables are (made) resident in the __constant__ block
static SweTabs swe_tabs(gg_rk* rk, bool* use_const) {
  gg_ctx* ctx = rk->ctx;
  auto trt = get_tab(ctx, GG_SPACE_RT, rk->order, rk->nq1);
  auto tdg = get_tab(ctx, GG_SPACE_DG, rk->order, rk->nq1);
  auto tcg = get_tab(ctx, GG_SPACE_CG, rk->order + 1, rk->nq1);
  SweTabs T;
  T.w = tdg->w.p; T.rt_val = trt->val.p; T.dg_val = tdg->val.p; T.cg_val = tcg->val.p; T.cg_der = tcg->der.p;
  *use_const = rk->nq1 * rk->nq1 <= SWE_MAXQ;
  const int64_t key = ((int64_t)rk->order << 8) | rk->nq1;
  if (*use_const && ctx->swe_key != key) {
    GG_CUDA(cudaStreamSynchronize(ctx->stream));  // earlier launches may still read the old tables
    GG_CUDA(cudaMemcpyToSymbolAsync(c_swe_w, tdg->w.p, tdg->w.n * sizeof(double), 0, cudaMemcpyDeviceToDevice, ctx->stream));
    GG_CUDA(cudaMemcpyToSymbolAsync(c_swe_rt, trt->val.p, trt->val.n * sizeof(double), 0, cudaMemcpyDeviceToDevice, ctx->stream));
    GG_CUDA(cudaMemcpyToSymbolAsync(c_swe_dg, tdg->val.p, tdg->val.n * sizeof(double), 0, cudaMemcpyDeviceToDevice, ctx->stream));
    GG_CUDA(cudaMemcpyToSymbolAsync(c_swe_cg, tcg->val.p, tcg->val.n * sizeof(double), 0, cudaMemcpyDeviceToDevice, ctx->stream));
    GG_CUDA(cudaMemcpyToSymbolAsync(c_swe_cgd, tcg->der.p, tcg->der.n * sizeof(double), 0, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->swe_key = key;
  }
  return T;
}

// tables of the sum-factorised kernels: node values of the reference RT basis, 1-D Lagrange bases of orders K+1 and K
static void swe_sf_tabs(gg_rk* rk) {
  gg_ctx* ctx = rk->ctx;
  const int K = rk->order, nq = rk->nq1;
  const int64_t key = ((int64_t)K << 8) | nq;
  if (ctx->swe_sf_key == key) return;
  GG_REQUIRE(nq <= SF_MAXNQ && K <= 2, GG_ERR_UNSUPPORTED, "sum-factorised shallow-water kernels: order <= 2, at most 16 points per direction");
  const int NA = K + 2, NB = K + 1, NX = NA * NB;
  RTRef R = build_rt_ref(K);
  const int MU = R.ndofs;
  auto node = [](int p, int i) { return p > 0 ? (double)i / p : 0.5; };
  std::vector<double> pts, val, div, phix((size_t)NX * MU), phiy((size_t)NX * MU);
  // component x: nodes (x_a of order K+1, y_b of order K), index b * NA + a;  component y: (x_a of order K, y_b of order K+1), b * NB + a
  for (int b = 0; b < NB; ++b) for (int a = 0; a < NA; ++a) { pts.push_back(node(K + 1, a)); pts.push_back(node(K, b)); }
  for (int b = 0; b < NA; ++b) for (int a = 0; a < NB; ++a) { pts.push_back(node(K, a)); pts.push_back(node(K + 1, b)); }
  R.tabulate(2 * NX, pts.data(), val, div);
  for (int i = 0; i < NX; ++i)
    for (int j = 0; j < MU; ++j) {
      phix[(size_t)i * MU + j] = val[((size_t)i * MU + j) * 2 + 0];
      phiy[(size_t)i * MU + j] = val[((size_t)(NX + i) * MU + j) * 2 + 1];
    }
  std::vector<double> xi(nq), w(nq), la((size_t)nq * NA), da((size_t)nq * NA), lb((size_t)nq * NB), db((size_t)nq * NB);
  gauss_legendre_01(nq, xi.data(), w.data());
  for (int q = 0; q < nq; ++q) {
    lagrange_1d(K + 1, xi[q], &la[(size_t)q * NA], &da[(size_t)q * NA]);
    lagrange_1d(K, xi[q], &lb[(size_t)q * NB], &db[(size_t)q * NB]);
  }
  GG_CUDA(cudaStreamSynchronize(ctx->stream));  // earlier launches may still read the old tables
  GG_CUDA(cudaMemcpyToSymbol(c_sf_phix, phix.data(), phix.size() * sizeof(double)));
  GG_CUDA(cudaMemcpyToSymbol(c_sf_phiy, phiy.data(), phiy.size() * sizeof(double)));
  GG_CUDA(cudaMemcpyToSymbol(c_sf_la, la.data(), la.size() * sizeof(double)));
  GG_CUDA(cudaMemcpyToSymbol(c_sf_da, da.data(), da.size() * sizeof(double)));
  GG_CUDA(cudaMemcpyToSymbol(c_sf_lb, lb.data(), lb.size() * sizeof(double)));
  GG_CUDA(cudaMemcpyToSymbol(c_sf_w, w.data(), w.size() * sizeof(double)));
  ctx->swe_sf_key = key;
  // ---- kernels with several lanes per cell (gg_swe_lanes.cuh): the RT basis as a tensor product of 1-D dual bases ----
  // A_alpha in P_{K+1} dual to Lambda_0 f = -f(0), Lambda_1 f = f(1), Lambda_{2+i} f = int x^i f;  B_m in P_K dual to int s^m f
  auto dual_basis = [](int n, const std::vector<long double>& L) {   // L[functional][monomial] -> coefficients [monomial][function]
    std::vector<long double> M = L, I((size_t)n * n, 0.0L);
    for (int i = 0; i < n; ++i) I[(size_t)i * n + i] = 1.0L;
    for (int c = 0; c < n; ++c) {
      int piv = c;
      for (int r = c + 1; r < n; ++r) if (fabsl(M[(size_t)r * n + c]) > fabsl(M[(size_t)piv * n + c])) piv = r;
      for (int k2 = 0; k2 < n; ++k2) { std::swap(M[(size_t)c * n + k2], M[(size_t)piv * n + k2]); std::swap(I[(size_t)c * n + k2], I[(size_t)piv * n + k2]); }
      const long double d = 1.0L / M[(size_t)c * n + c];
      for (int k2 = 0; k2 < n; ++k2) { M[(size_t)c * n + k2] *= d; I[(size_t)c * n + k2] *= d; }
      for (int r = 0; r < n; ++r) {
        if (r == c) continue;
        const long double f = M[(size_t)r * n + c];
        for (int k2 = 0; k2 < n; ++k2) { M[(size_t)r * n + k2] -= f * M[(size_t)c * n + k2]; I[(size_t)r * n + k2] -= f * I[(size_t)c * n + k2]; }
      }
    }
    return I;   // L^-1: [monomial e][function alpha]
  };
  std::vector<long double> LA((size_t)NA * NA), LB((size_t)NB * NB);
  for (int e = 0; e < NA; ++e) {
    LA[(size_t)0 * NA + e] = e == 0 ? -1.0L : 0.0L;
    LA[(size_t)1 * NA + e] = 1.0L;
    for (int i = 0; i + 2 < NA; ++i) LA[(size_t)(2 + i) * NA + e] = 1.0L / (i + e + 1);
  }
  for (int m = 0; m < NB; ++m) for (int e = 0; e < NB; ++e) LB[(size_t)m * NB + e] = 1.0L / (m + e + 1);
  const std::vector<long double> CA = dual_basis(NA, LA), CB = dual_basis(NB, LB);
  std::vector<double> ra((size_t)nq * NA), rda((size_t)nq * NA), rb((size_t)nq * NB);
  for (int q = 0; q < nq; ++q) {
    const long double x = xi[q];
    for (int al = 0; al < NA; ++al) {
      long double v = 0.0L, d = 0.0L, xp = 1.0L, xm = 0.0L;   // xp = x^e, xm = e x^(e-1)
      for (int e = 0; e < NA; ++e) { v += CA[(size_t)e * NA + al] * xp; d += CA[(size_t)e * NA + al] * xm; xm = (e + 1) * xp; xp *= x; }
      ra[(size_t)q * NA + al] = (double)v; rda[(size_t)q * NA + al] = (double)d;
    }
    for (int m = 0; m < NB; ++m) {
      long double v = 0.0L, xp = 1.0L;
      for (int e = 0; e < NB; ++e) { v += CB[(size_t)e * NB + m] * xp; xp *= x; }
      rb[(size_t)q * NB + m] = (double)v;
    }
  }
  {
    // check the factorisation against the tabulated basis at the tensor Gauss points
    auto jx = [&](int al, int m) { return al == 0 ? 2 * NB + m : al == 1 ? 3 * NB + m : 4 * NB + m * K + (al - 2); };
    auto jy = [&](int m, int be) { return be == 0 ? m : be == 1 ? NB + m : 4 * NB + K * NB + (be - 2) * NB + m; };
    std::vector<double> tp, tv, td;
    for (int q2 = 0; q2 < nq; ++q2) for (int q1 = 0; q1 < nq; ++q1) { tp.push_back(xi[q1]); tp.push_back(xi[q2]); }
    R.tabulate(nq * nq, tp.data(), tv, td);
    std::vector<double> ex(tv.size(), 0.0), exd(td.size(), 0.0);
    for (int q2 = 0; q2 < nq; ++q2)
      for (int q1 = 0; q1 < nq; ++q1) {
        const size_t q = (size_t)q1 + (size_t)nq * q2;
        for (int al = 0; al < NA; ++al)
          for (int m = 0; m < NB; ++m) {
            ex[(q * MU + jx(al, m)) * 2 + 0] = ra[(size_t)q1 * NA + al] * rb[(size_t)q2 * NB + m];
            exd[q * MU + jx(al, m)] = rda[(size_t)q1 * NA + al] * rb[(size_t)q2 * NB + m];
            ex[(q * MU + jy(m, al)) * 2 + 1] = rb[(size_t)q1 * NB + m] * ra[(size_t)q2 * NA + al];
            exd[q * MU + jy(m, al)] = rb[(size_t)q1 * NB + m] * rda[(size_t)q2 * NA + al];
          }
      }
    // (the tabulation goes through the monomial prebasis: 1e-12 relative for K = 2, values up to 1e3; the 1-D tables are exact
    // to the last bit, so only the order of magnitude of a mismatch matters here)
    double err = 0.0, scale = 1.0;
    for (size_t i = 0; i < tv.size(); ++i) { err = std::max(err, std::fabs(tv[i] - ex[i])); scale = std::max(scale, std::fabs(tv[i])); }
    for (size_t i = 0; i < td.size(); ++i) { err = std::max(err, std::fabs(td[i] - exd[i])); scale = std::max(scale, std::fabs(td[i])); }
    GG_REQUIRE(err < 1e-9 * scale, GG_ERR_INVALID, "Raviart-Thomas basis does not factorise into the 1-D dual bases");
  }
  GG_CUDA(cudaMemcpyToSymbol(c_sf_ra, ra.data(), ra.size() * sizeof(double)));
  GG_CUDA(cudaMemcpyToSymbol(c_sf_rda, rda.data(), rda.size() * sizeof(double)));
  GG_CUDA(cudaMemcpyToSymbol(c_sf_rb, rb.data(), rb.size() * sizeof(double)));
  std::vector<double> blk;   // SweLanes<K, NQ>::T_*
  blk.insert(blk.end(), la.begin(), la.end());
  blk.insert(blk.end(), da.begin(), da.end());
  blk.insert(blk.end(), lb.begin(), lb.end());
  blk.insert(blk.end(), ra.begin(), ra.end());
  blk.insert(blk.end(), rda.begin(), rda.end());
  blk.insert(blk.end(), rb.begin(), rb.end());
  blk.insert(blk.end(), w.begin(), w.end());
  blk.resize((blk.size() + 1) & ~(size_t)1, 0.0);
  ctx->swe_lane_tab.upload(blk, ctx->stream);
  GG_CUDA(cudaStreamSynchronize(ctx->stream));
}

// (order, points per direction) pairs the lane-parallel kernels are instantiated for: the default rules of the drivers
// (degree 4 (K + 1): 3, 5, 7 points; thermal shallow water 4 K: 3 points for K = 1) and degree 8 for K = 2.
// ONE list: the predicate and the two dispatches expand it, so they cannot disagree.
#define GG_LANE_PAIRS(X) X(0, 3) X(1, 3) X(1, 5) X(2, 5) X(2, 7)
static bool swe_lanes_enabled(int K, int nq) {
  static const bool on = !(getenv("GG_SWE_LANES") && getenv("GG_SWE_LANES")[0] == '0');
  bool have = false;
#define GG_PAIR(K_, NQ_) have = have || (K == K_ && nq == NQ_);
  GG_LANE_PAIRS(GG_PAIR)
#undef GG_PAIR
  return on && have;
}

template <typename Kern, typename... Args>
static void swe_launch_lanes(gg_ctx* ctx, Kern kern, int64_t nc, int cpb, size_t smem, Args... args) {
  static std::map<const void*, int> per_sm;   // resident CTAs per SM of each instantiation
  auto it = per_sm.find((const void*)kern);
  if (it == per_sm.end()) {
    if (smem > 48 * 1024) GG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int n = 0;
    GG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, 128, smem));
    it = per_sm.emplace((const void*)kern, std::max(n, 1)).first;
  }
  const int64_t nblk = (nc + cpb - 1) / cpb;
  const unsigned grid = (unsigned)std::min<int64_t>(nblk, (int64_t)ctx->sm_count * it->second);
  kern<<<grid, 128, smem, ctx->stream>>>(nc, args...);
}

static bool swe_sf_enabled() {
  static const bool on = !(getenv("GG_SWE_SF") && getenv("GG_SWE_SF")[0] == '0');
  return on;
}

// y = [q; F; Phi] from the state x = [u; p]
void swe_diagnostics_async(gg_rk* rk, const double* x, double* y) {
  gg_ctx* ctx = rk->ctx;
  cudaStream_t st = ctx->stream;
  const int64_t nc = rk->mesh->n_cells;
  if (nc == 0) return;
  CellGeo g = cell_geo(rk->mesh, rk->nq1);
  bool uc = false;
  SweTabs T = swe_tabs(rk, &uc);
  double *yq = y, *yF = y + rk->nr, *yP = y + rk->nr + rk->nu;
  const double *u = x, *p = x + rk->nu;
  unsigned nb = nblk(nc, 64);
  if (swe_lanes_enabled(rk->order, rk->nq1)) {
    ProfScope ps(ctx, "k_swe_diag");
    swe_sf_tabs(rk);
#define GG_LANES(K_, NQ_)                                                                                                  \
  if (rk->order == K_ && rk->nq1 == NQ_)                                                                                   \
    swe_launch_lanes(ctx, k_swe_diag_lanes<K_, NQ_>, nc, SweLanes<K_, NQ_>::CPB, SweLanes<K_, NQ_>::diag_smem(), g,        \
                     ctx->swe_lane_tab.p, rk->V->d_cell_dofs.p, rk->V->d_cell_signs.p, u, p, rk->fq.p,                     \
                     rk->bq.n ? rk->bq.p : nullptr, rk->gravity, rk->MpInv.p, rk->pq->d.p, rk->evec_q.p, rk->evec_u.p, yP);
    GG_LANE_PAIRS(GG_LANES)
#undef GG_LANES
    ctx->launches++;
    GG_CUDA(cudaGetLastError());
  } else {
    ProfScope ps(ctx, "k_swe_diag");
    const size_t smem = swe_stage_smem(rk, 64);
    double* gs = smem ? nullptr : rk->stage2.p;
#define GG_ARGS nc, rk->nq1, g, T, rk->V->d_cell_dofs.p, rk->V->d_cell_signs.p, u, p, rk->fq.p, rk->bq.n ? rk->bq.p : nullptr, \
                rk->gravity, rk->MpInv.p, rk->pq->d.p, rk->evec_q.p, rk->evec_u.p, yP, gs
#define GG_CASE(K_)                                                 \
  case K_:                                                          \
    if (uc) swe_launch(k_swe_diag<K_, true>, nb, smem, st, GG_ARGS);       \
    else swe_launch(k_swe_diag<K_, false>, nb, smem, st, GG_ARGS);         \
    break;
    switch (rk->order) {
      GG_CASE(0) GG_CASE(1) GG_CASE(2)
      default: throw Error(GG_ERR_UNSUPPORTED, "shallow-water order must be <= 2");
    }
#undef GG_CASE
#undef GG_ARGS
    ctx->launches++;
    GG_CUDA(cudaGetLastError());
  }
  // potential vorticity: assemble int q p w sqrtg (the Jacobian block that depends on the state), refresh the
  // preconditioner (numerical_setup!), solve
  {
    // the Jacobian block that depends on the state, int q p w sqrtg, then numerical_setup!
    if (rk->ebe_q) {
      // element matrices, static condensation and preconditioner in one fused pass (gg_ebe.cu)
      ebe_set_scalar_mass(rk->ebe_q, rk->degree, rk->pq->d.p);
    } else if (!rk->solver_q->mf_on) {
      // assembled path: element matrices, atomic-free CSR fill, block inverses
      gg_form_args fa;
      std::memset(&fa, 0, sizeof(fa));
      fa.test = rk->R; fa.trial = rk->R; fa.degree = rk->degree; fa.qdata = rk->pq;
      int layout = 0;
      const double* eb = element_matrices_device(GG_FORM_MASS_SCALAR, &fa, &layout);
      gather_matrix(rk->Mq, layout, eb, 1.0, 0);
      solver_update(rk->solver_q);
    }
    // the element-by-element solver gathers the element vectors itself (one launch and one pass over the vector less per solve)
    if (!rk->ebe_q) gather_vector(rk->R, rk->evec_q.p, rk->rhs_q.p, 1.0, 0);
    rk_mass_solve(rk, 1, rk->rhs_q.p, 1.0, yq, rk->diag_slot >= 0 ? 100 + rk->diag_slot : -1, rk->ebe_q ? rk->evec_q.p : nullptr);
  }
  // mass flux
  if (!rk->ebe_u) gather_vector(rk->V, rk->evec_u.p, rk->rhs_F.p, 1.0, 0);
  rk_mass_solve(rk, 0, rk->rhs_F.p, 1.0, yF, rk->diag_slot >= 0 ? 200 + rk->diag_slot : -1, rk->ebe_u ? rk->evec_u.p : nullptr);
  if (rk->thermal) tswe_diagnostics_async(rk, x, y);   // Phi += 0.5 B, b
  rk->diag_solves++;
}

// r = [r_u; r_p]; kp (optional) receives the depth slope -Mp^-1 r_p
void swe_residual_async(gg_rk* rk, const double* x, const double* y, const double* q0, double* r, double* kp, bool gather_u) {
  gg_ctx* ctx = rk->ctx;
  cudaStream_t st = ctx->stream;
  const int64_t nc = rk->mesh->n_cells;
  if (nc == 0) return;
  CellGeo g = cell_geo(rk->mesh, rk->nq1);
  bool uc = false;
  SweTabs T = swe_tabs(rk, &uc);
  const double *yq = y, *yF = y + rk->nr, *yP = y + rk->nr + rk->nu;
  const bool alias = (q0 == yq);
  unsigned nb = nblk(nc, 64);
  if (swe_lanes_enabled(rk->order, rk->nq1)) {
    ProfScope ps(ctx, "k_swe_stage");
    swe_sf_tabs(rk);
#define GG_LANES_A(K_, NQ_, A_)                                                                                            \
  swe_launch_lanes(ctx, k_swe_stage_lanes<K_, A_, NQ_>, nc, SweLanes<K_, NQ_>::CPB, SweLanes<K_, NQ_>::stage_smem(), g,    \
                   ctx->swe_lane_tab.p, rk->V->d_cell_dofs.p, rk->V->d_cell_signs.p, rk->R->d_cell_dofs.p, x, yq, q0, yF,  \
                   yP, rk->MpInv.p, rk->tau, rk->tau / rk->dt, rk->evec_u.p, r + rk->nu, kp);
#define GG_LANES(K_, NQ_)                                                                                                  \
  if (rk->order == K_ && rk->nq1 == NQ_) {                                                                                 \
    if (alias) { GG_LANES_A(K_, NQ_, true) } else { GG_LANES_A(K_, NQ_, false) }                                           \
  }
    GG_LANE_PAIRS(GG_LANES)
#undef GG_LANES
#undef GG_LANES_A
    ctx->launches++;
    GG_CUDA(cudaGetLastError());
  } else if (swe_sf_enabled()) {
    ProfScope ps(ctx, "k_swe_stage");
    swe_sf_tabs(rk);
    // the sum-factorised kernels are register-limited to 4 CTAs per SM, so the 50 KB staging buffer fits in shared memory
    const size_t bytes = (size_t)2 * rk->nq1 * rk->nq1 * 64 * sizeof(double);
    const size_t smem = bytes <= 56 * 1024 ? bytes : 0;
    if (!smem && rk->stage2.n == 0) rk->stage2.alloc((size_t)2 * rk->nq1 * rk->nq1 * nc);
    double* gs = smem ? nullptr : rk->stage2.p;
    auto tdg = get_tab(ctx, GG_SPACE_DG, rk->order, rk->nq1);
#define GG_ARGS nc, rk->nq1, g, tdg->w.p, rk->V->d_cell_dofs.p, rk->V->d_cell_signs.p, rk->R->d_cell_dofs.p, x, yq, q0, yF, yP, \
                rk->Bhat.p, rk->MpInv.p, rk->tau, rk->tau / rk->dt, rk->evec_u.p, r + rk->nu, kp, gs
// (compile-time rule sizes with unrolled x-loops were measured too: 1.79 against 1.74 ms per step -- more spills, no gain)
#define GG_NQ(K_, A_) swe_launch(k_swe_stage_sf<K_, A_, 0>, nb, smem, st, GG_ARGS);
#define GG_CASE(K_)                                                                          \
  case K_:                                                                                   \
    if (alias) { GG_NQ(K_, true) } else { GG_NQ(K_, false) }                                 \
    break;
    switch (rk->order) {
      GG_CASE(0) GG_CASE(1) GG_CASE(2)
      default: throw Error(GG_ERR_UNSUPPORTED, "shallow-water order must be <= 2");
    }
#undef GG_CASE
#undef GG_NQ
#undef GG_ARGS
    ctx->launches++;
    GG_CUDA(cudaGetLastError());
  } else {
    ProfScope ps(ctx, "k_swe_stage");
    const size_t smem = swe_stage_smem(rk, 64);
    double* gs = smem ? nullptr : rk->stage2.p;
#define GG_ARGS nc, rk->nq1, g, T, rk->V->d_cell_dofs.p, rk->V->d_cell_signs.p, rk->R->d_cell_dofs.p, x, yq, q0, yF, yP, \
                rk->Bhat.p, rk->MpInv.p, rk->tau, rk->tau / rk->dt, rk->evec_u.p, r + rk->nu, kp, gs
#define GG_CASE(K_)                                                                   \
  case K_:                                                                            \
    if (alias && uc) swe_launch(k_swe_stage<K_, true, true>, nb, smem, st, GG_ARGS);         \
    else if (alias) swe_launch(k_swe_stage<K_, true, false>, nb, smem, st, GG_ARGS);         \
    else if (uc) swe_launch(k_swe_stage<K_, false, true>, nb, smem, st, GG_ARGS);            \
    else swe_launch(k_swe_stage<K_, false, false>, nb, smem, st, GG_ARGS);                   \
    break;
    switch (rk->order) {
      GG_CASE(0) GG_CASE(1) GG_CASE(2)
      default: throw Error(GG_ERR_UNSUPPORTED, "shallow-water order must be <= 2");
    }
#undef GG_CASE
#undef GG_ARGS
    ctx->launches++;
    GG_CUDA(cudaGetLastError());
  }
  if (rk->thermal) tswe_residual_async(rk, x, y, r + rk->nu + rk->np, kp ? kp + rk->np : nullptr);
  if (gather_u) gather_vector(rk->V, rk->evec_u.p, r, 1.0, 0);
}

// one ode_march! (src/ODEs/RungeKuttaEX.jl:48-134).  The reference solves the diagnostics five times per SSPRK3 step; two of
// those solves repeat an earlier one on the same state (the first stage starts from u0, :85-93 after :70-75, and the final
// solve of a step, :125-127, is the initial one of the next), so their result is reused: identical numbers, 3 solves per step.
void swe_step_async(gg_rk* rk) {
  gg_ctx* ctx = rk->ctx;
  cudaStream_t st = ctx->stream;
  const int s = rk->n_stages;
  // :70-75 initial diagnostics of the step
  if (!rk->y_valid) { swe_diagnostics_async(rk, rk->x.p, rk->y.p); rk->y_valid = true; }
  if (rk->q0_mode == GG_Q0_START_OF_STEP && rk->nr) {
    k_copy<<<nblk(rk->nr, 256), 256, 0, st>>>(rk->nr, rk->y.p, rk->q0.p);
    ctx->launches++;
  }
  for (int i = 0; i < s; ++i) {
    const double* xs = rk->x.p;
    bool same_state = true;
    double co[4] = {0, 0, 0, 0};
    for (int j = 0; j < i && j < 4; ++j) { co[j] = rk->dt * rk->A[i * s + j]; same_state = same_state && co[j] == 0.0; }
    if (i > 0) {
      // :85-88 stage state
      rk_combine(rk, std::min(i, 4), co, rk->xi.p);
      if (rk->thermal) vec_consistent_async(rk->Q, rk->xi.p + rk->nu + rk->np);   // B on the ghost cells (partitioned models)
      xs = rk->xi.p;
    }
    // :91-93 stage diagnostics (y still holds diagnostics(x) when the stage state is x itself)
    if (!(same_state && rk->y_valid)) {
      rk->diag_slot = i;
      swe_diagnostics_async(rk, xs, rk->y.p);
      rk->diag_slot = -1;
      rk->y_valid = false;
    }
    // :100-117 slope: M k = -(res_x + mass(0))
    const double* q0 = rk->q0_mode == GG_Q0_START_OF_STEP ? rk->q0.p : rk->y.p;
    swe_residual_async(rk, xs, rk->y.p, q0, rk->r.p, rk->k[i].p + rk->nu, rk->ebe_u == nullptr);
    rk_mass_solve(rk, 0, rk->r.p, -1.0, rk->k[i].p, i, rk->ebe_u ? rk->evec_u.p : nullptr);
  }
  // :121-122 update, :125-127 final diagnostics
  double co[4] = {0, 0, 0, 0};
  for (int j = 0; j < s && j < 4; ++j) co[j] = rk->dt * rk->b[j];
  rk_combine(rk, std::min(s, 4), co, rk->x.p);
  if (rk->thermal) vec_consistent_async(rk->Q, rk->x.p + rk->nu + rk->np);
  rk->diag_slot = s;
  swe_diagnostics_async(rk, rk->x.p, rk->y.p);
  rk->diag_slot = -1;
  rk->y_valid = true;
  GG_CUDA(cudaGetLastError());
}

void swe_create(gg_rk* rk, const gg_rk_args* a) {
  gg_ctx* ctx = rk->ctx;
  cudaStream_t st = ctx->stream;
  auto chk = [](int code) { if (code != GG_OK) throw Error(code, gg_last_error(nullptr)); };
  GG_REQUIRE(a->coriolis, GG_ERR_INVALID, "shallow water needs the Coriolis parameter at the quadrature points");
  GG_REQUIRE(a->q0_mode == GG_Q0_ALIAS || a->q0_mode == GG_Q0_START_OF_STEP, GG_ERR_INVALID, "unknown q0_mode");
  const int64_t nc = rk->mesh->n_cells;
  rk->nq1 = num_points_1d(a->degree);
  GG_REQUIRE(rk->nq1 <= MAX_NQ, GG_ERR_UNSUPPORTED, "quadrature degree too high");
  const int64_t Q = (int64_t)rk->nq1 * rk->nq1;
  rk->gravity = rk->thermal ? 0.0 : a->gravity;   // thermal: the potential term is 0.5 B, added by gg_tswe.cu
  rk->tau = a->tau < 0 ? 0.5 * a->dt : a->tau;  // TransientShallowWater.jl:150
  rk->q0_mode = a->q0_mode;
  chk(gg_space_builtin(rk->mesh, GG_SPACE_CG, a->order + 1, GG_CONSTRAINT_NONE, &rk->R));
  rk->nr = rk->R->n_free; rk->mr = rk->R->ndofs_loc;
  chk(gg_vec_create(ctx, nc * Q, &rk->pq));
  chk(gg_vec_fill(rk->pq, 1.0));
  build_vec_map(rk->R);
  const char* mode_env = getenv("GG_PCG");
  const std::string mode = mode_env ? mode_env : "ebe";
  if (rk->ebe_u && ebe_supported(rk->R)) {
    rk->ebe_q = ebe_create(rk->R, a->rtol > 0 ? a->rtol : 1e-12, 2000);
  } else {
    GG_REQUIRE(!rk->mesh->plan, GG_ERR_UNSUPPORTED, "partitioned models need the element-by-element mass solver (GG_PCG=ebe)");
    GG_REQUIRE(rk->solver, GG_ERR_INVALID, "no RT mass solver");
    chk(gg_pattern(rk->R, rk->R, &rk->Mq));
    // a first assembly with unit depth so that the solver is created on a positive definite matrix
    gg_form_args fa;
    std::memset(&fa, 0, sizeof(fa));
    fa.test = rk->R; fa.trial = rk->R; fa.degree = a->degree; fa.qdata = rk->pq;
    chk(gg_assemble_matrix(GG_FORM_MASS_SCALAR, &fa, rk->Mq, 0));
    chk(gg_solver_pcg_create(rk->Mq, a->rtol > 0 ? a->rtol : 1e-12, 2000, &rk->solver_q));
    if (mode == "mf" && mf_supported(GG_SPACE_RT, a->order, rk->nq1) && mf_supported(GG_SPACE_CG, a->order + 1, rk->nq1)) {
      solver_set_matrix_free(rk->solver, a->degree, nullptr);
      solver_set_matrix_free(rk->solver_q, a->degree, rk->pq->d.p);
    }
  }
  rk->fq.upload(a->coriolis, (size_t)(nc * Q), st);
  if (a->topography && !rk->thermal) rk->bq.upload(a->topography, (size_t)(nc * Q), st);
  rk->y.alloc(rk->nr + rk->nu + rk->np + rk->nB);
  rk->q0.alloc(rk->nr);
  rk->rhs_q.alloc(rk->nr); rk->rhs_F.alloc(rk->nu);
  rk->evec_q.alloc((size_t)rk->mr * nc); rk->evec_u.alloc((size_t)rk->mu * nc);
  if (rk->y.n) GG_CUDA(cudaMemsetAsync(rk->y.p, 0, rk->y.n * sizeof(double), st));
  if (rk->q0.n) GG_CUDA(cudaMemsetAsync(rk->q0.p, 0, rk->q0.n * sizeof(double), st));
  GG_CUDA(cudaStreamSynchronize(st));
}

void swe_destroy(gg_rk* rk) {
  ebe_destroy(rk->ebe_q);
  rk->ebe_q = nullptr;
  if (rk->solver_q) gg_solver_destroy(rk->solver_q);
  if (rk->Mq) gg_csr_destroy(rk->Mq);
  if (rk->pq) gg_vec_destroy(rk->pq);
  if (rk->R) gg_space_destroy(rk->R);
  rk->solver_q = nullptr; rk->Mq = nullptr; rk->pq = nullptr; rk->R = nullptr;
}

}  // namespace gg
