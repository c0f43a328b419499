// Thermal shallow water (test/Tutorials/ThermalShallowWater.jl): the terms it adds to the rotating shallow-water DAE.
//
//   prognostic x = [u (RT_k); p (DG_k); B (DG_k)]      diagnostic y = [q (CG_{k+1}); F (RT_k); Phi (DG_k); b (DG_k)]
// The q and F equations, the q-dependent part of res_u, -int Phi div(v), res_p and the mass solves are the shallow-water
// kernels of gg_swe.cu run unchanged (gravity 0, no topography); this file adds
//   * diagnostics (:173-174): Phi += 0.5 B  (the L2 projection of a DG_k function onto DG_k is the function itself) and
//     b from the cell-local system  int b p r sqrtg = int B r sqrtg;
//   * stage residual (:216-245): the buoyancy cell terms of res_u, the whole res_B, and their skeleton terms u_s1, u_s2,
//     B_s1, B_s2 -- one thread per cell, each cell integrating its own side of its four facets (owner computes, no atomics):
//     plus = lower-numbered cell, my_mean(x) = (x.plus - x.minus)/2, upwinding_sign = -1/2, 0, +1/2 with threshold 1e-4.
// RT functions carry sqrtg and are contravariant-Piola mapped, so the skeleton terms need no metric: on an axis-aligned
// chart cell the flux density times the facet measure is (reference normal component) x (1-D weight).  The normal trace of
// F (and of the RT test functions) is continuous, so (F.n).minus = -(F.n).plus is taken from the cell's own DOFs.
#include <cstring>
#include "gg_rk.cuh"
#include "gg_refel.cuh"
#include "gg_geom.cuh"

namespace gg {

static inline unsigned nblk(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

static const int TQEDGE_V[4][2] = {{0, 1}, {2, 3}, {0, 2}, {1, 3}};
static const double TQVERT[4][2] = {{0, 0}, {1, 0}, {0, 1}, {1, 1}};
static const double TQNORM[4][2] = {{0, -1}, {0, 1}, {-1, 0}, {1, 0}};

__device__ __forceinline__ double tswe_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(x, -(y * y), 1.0);
  return fma(fma(e, 0.375, 0.5), y * e, y);
}

// b from int b p r sqrtg = int B r sqrtg (cell-local, SPD), and Phi += 0.5 B
template <int K>
__global__ void __launch_bounds__(64)
k_tswe_buoyancy(int64_t nc, int nq, CellGeo g, const double* __restrict__ w1, const double* __restrict__ dgv,
                const double* __restrict__ p, const double* __restrict__ B, double* __restrict__ Phi, double* __restrict__ bout) {
  constexpr int MP = (K + 1) * (K + 1);
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const double HX = g.hx[c], HY = g.hy[c];
  const double* __restrict__ ta = g.tx + (int64_t)g.ix[c] * nq;
  const double* __restrict__ tb = g.ty + (int64_t)g.iy[c] * nq;
  const double r2 = g.radius * g.radius;
  double pe[MP], Be[MP], A[MP][MP], rhs[MP];
#pragma unroll
  for (int i = 0; i < MP; ++i) {
    pe[i] = p[c * MP + i]; Be[i] = B[c * MP + i]; rhs[i] = 0.0;
    Phi[c * MP + i] += 0.5 * Be[i];
#pragma unroll
    for (int j = 0; j < MP; ++j) A[i][j] = 0.0;
  }
#pragma unroll 1
  for (int q2 = 0; q2 < nq; ++q2) {
    const double b = tb[q2], sb = fma(b, b, 1.0);
#pragma unroll 1
    for (int q1 = 0; q1 < nq; ++q1) {
      const double t = ta[q1], sa = fma(t, t, 1.0);
      const double ri = tswe_rsqrt(sa + b * b);
      const double wsg = w1[q1] * w1[q2] * HX * HY * r2 * sa * sb * ri * ri * ri;
      const double* __restrict__ dv = dgv + (size_t)(q1 + nq * q2) * MP;
      double ph = 0.0, Bh = 0.0;
#pragma unroll
      for (int i = 0; i < MP; ++i) { ph = fma(dv[i], pe[i], ph); Bh = fma(dv[i], Be[i], Bh); }
#pragma unroll
      for (int i = 0; i < MP; ++i) {
        const double wi = wsg * dv[i];
        rhs[i] = fma(wi, Bh, rhs[i]);
#pragma unroll
        for (int j = i; j < MP; ++j) A[i][j] = fma(wi * ph, dv[j], A[i][j]);
      }
    }
  }
  // Cholesky-free elimination on the symmetric positive definite block (upper triangle held)
#pragma unroll
  for (int i = 0; i < MP; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) A[i][j] = A[j][i];
#pragma unroll
  for (int k = 0; k < MP; ++k) {
    const double d = 1.0 / A[k][k];
#pragma unroll
    for (int i = k + 1; i < MP; ++i) {
      const double f = A[i][k] * d;
#pragma unroll
      for (int j = k + 1; j < MP; ++j) A[i][j] = fma(-f, A[k][j], A[i][j]);
      rhs[i] = fma(-f, rhs[k], rhs[i]);
    }
  }
#pragma unroll
  for (int i = MP - 1; i >= 0; --i) {
    double v = rhs[i];
#pragma unroll
    for (int j = i + 1; j < MP; ++j) v = fma(-A[i][j], rhs[j], v);
    rhs[i] = v / A[i][i];
  }
#pragma unroll
  for (int i = 0; i < MP; ++i) bout[c * MP + i] = rhs[i];
}

struct TsweTabs {
  const double* w1;     // [nq] 1-D weights
  const double* rtv;    // [Q][MU][2] reference RT basis
  const double* rtd;    // [Q][MU] reference divergence
  const double* dgv;    // [Q][MP]
  const double* dgd;    // [Q][MP][2] reference gradients
  const double* fval;   // [4][nqf][MP] DG basis at the facet points
  const double* fvn;    // [4][nqf][MU] reference RT normal component at the facet points
  const double* wf;     // [nqf]
};

// adds the buoyancy terms to the RT element vector eru [MU][nc] of res_u; writes r_B (cell-local) and, optionally, k_B = -Mp^-1 r_B
template <int K>
__global__ void __launch_bounds__(64)
k_tswe_stage(int64_t nc, int nq, TsweTabs T, const double* __restrict__ hx, const double* __restrict__ hy,
             const int32_t* __restrict__ vdofs, const int8_t* __restrict__ vsigns, const int32_t* __restrict__ nbr_cell,
             const int8_t* __restrict__ nbr_edge, const int8_t* __restrict__ is_plus, const double* __restrict__ p,
             const double* __restrict__ bv, const double* __restrict__ Fv, const double* __restrict__ MpInv,
             double* __restrict__ eru, double* __restrict__ rB, double* __restrict__ kB) {
  constexpr int MU = 2 * (K + 1) * (K + 2), MP = (K + 1) * (K + 1);
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const double HX = hx[c], HY = hy[c], iHX = 1.0 / HX, iHY = 1.0 / HY;
  double pe[MP], be[MP], Fe[MU], ru[MU], rb[MP];
#pragma unroll
  for (int i = 0; i < MP; ++i) { pe[i] = p[c * MP + i]; be[i] = bv[c * MP + i]; rb[i] = 0.0; }
#pragma unroll
  for (int j = 0; j < MU; ++j) { Fe[j] = (double)vsigns[c * MU + j] * Fv[vdofs[c * MU + j]]; ru[j] = 0.0; }
  // ---- cell terms (:216-224, :234-238) -------------------------------------------------------------------------------------
  const int Q = nq * nq;
#pragma unroll 1
  for (int q = 0; q < Q; ++q) {
    const double wq = T.w1[q % nq] * T.w1[q / nq] * HX * HY;
    const double* __restrict__ dv = T.dgv + (size_t)q * MP;
    const double* __restrict__ dd = T.dgd + (size_t)q * MP * 2;
    const double* __restrict__ rv = T.rtv + (size_t)q * MU * 2;
    const double* __restrict__ rd = T.rtd + (size_t)q * MU;
    double th = 0.0, tx = 0.0, ty = 0.0, bh = 0.0, bx = 0.0, by = 0.0;
#pragma unroll
    for (int i = 0; i < MP; ++i) {
      th = fma(dv[i], pe[i], th); tx = fma(dd[2 * i], pe[i], tx); ty = fma(dd[2 * i + 1], pe[i], ty);
      bh = fma(dv[i], be[i], bh); bx = fma(dd[2 * i], be[i], bx); by = fma(dd[2 * i + 1], be[i], by);
    }
    th *= 0.5; tx *= 0.5 * iHX; ty *= 0.5 * iHY; bx *= iHX; by *= iHY;          // theta = p / 2, chart gradients
    double Fx = 0.0, Fy = 0.0, dF = 0.0;
#pragma unroll
    for (int j = 0; j < MU; ++j) { Fx = fma(rv[2 * j], Fe[j], Fx); Fy = fma(rv[2 * j + 1], Fe[j], Fy); dF = fma(rd[j], Fe[j], dF); }
    Fx *= iHY; Fy *= iHX; dF *= iHX * iHY;                                      // contravariant Piola, axis-aligned cell
    // res_u: 0.5 b (grad(th).v) - 0.5 th (grad(b).v) - 0.5 b th div(v)
    const double cx = wq * 0.5 * (bh * tx - th * bx) * iHY, cy = wq * 0.5 * (bh * ty - th * by) * iHX;
    const double cd = -wq * 0.5 * bh * th * iHX * iHY;
#pragma unroll
    for (int j = 0; j < MU; ++j) ru[j] = fma(cx, rv[2 * j], fma(cy, rv[2 * j + 1], fma(cd, rd[j], ru[j])));
    // res_B: -0.5 b (grad(w).F) + 0.5 b w div(F) + 0.5 w (grad(b).F)
    const double gx = -wq * 0.5 * bh * Fx * iHX, gy = -wq * 0.5 * bh * Fy * iHY;
    const double c0 = wq * 0.5 * (bh * dF + bx * Fx + by * Fy);
#pragma unroll
    for (int i = 0; i < MP; ++i) rb[i] = fma(gx, dd[2 * i], fma(gy, dd[2 * i + 1], fma(c0, dv[i], rb[i])));
  }
  // ---- skeleton terms: this cell's side of its four facets (:226-231, :240-245) ---------------------------------------------
#pragma unroll 1
  for (int e = 0; e < 4; ++e) {
    const int64_t cn = nbr_cell[c * 4 + e];
    const int en = nbr_edge[c * 4 + e];
    const bool plus = is_plus[c * 4 + e];
    const double pf = e < 2 ? iHX : iHY;       // Piola factor of the normal component on this edge (times the edge length = 1)
    double pn[MP], bn[MP];
#pragma unroll
    for (int i = 0; i < MP; ++i) { pn[i] = p[cn * MP + i]; bn[i] = bv[cn * MP + i]; }
#pragma unroll 1
    for (int q = 0; q < nq; ++q) {
      const double* __restrict__ fo = T.fval + (size_t)(e * nq + q) * MP;
      const double* __restrict__ fn = T.fval + (size_t)(en * nq + q) * MP;
      const double* __restrict__ vn = T.fvn + (size_t)(e * nq + q) * MU;
      double bo = 0.0, to = 0.0, bt = 0.0, tt = 0.0, Fr = 0.0;
#pragma unroll
      for (int i = 0; i < MP; ++i) {
        bo = fma(fo[i], be[i], bo); to = fma(fo[i], pe[i], to);
        bt = fma(fn[i], bn[i], bt); tt = fma(fn[i], pn[i], tt);
      }
      to *= 0.5; tt *= 0.5;
#pragma unroll
      for (int j = 0; j < MU; ++j) Fr = fma(vn[j], Fe[j], Fr);
      const double Fn_own = Fr * pf;                       // (F.n) of this side
      const double Fnp = plus ? Fn_own : -Fn_own;          // (F.n).plus
      const double sg = Fnp < -1e-4 ? -0.5 : (Fnp > 1e-4 ? 0.5 : 0.0);
      const double jb = plus ? bo - bt : bt - bo, jt = plus ? to - tt : tt - to;
      const double wf = T.wf[q];                           // = w * edge length * Piola factor of the normal component
      // u equation: own-side part of u_s1, and u_s2 on the plus side
      const double cu = plus ? (-0.25 * bo * jt + 0.25 * to * jb - 0.5 * sg * jb * jt) : (0.25 * bo * jt - 0.25 * to * jb);
#pragma unroll
      for (int j = 0; j < MU; ++j) ru[j] = fma(wf * cu, vn[j], ru[j]);
      // B equation: (F.n).minus = -(F.n).plus
      const double mFb = 0.5 * Fnp * (bo + bt);
      const double up = 0.5 * sg * Fnp * jb;
      const double cB = plus ? (0.5 * mFb - 0.25 * Fnp * jb + up) : (-0.5 * mFb - 0.25 * Fnp * jb - up);
      const double wl = wf / pf;                           // w * edge length
#pragma unroll
      for (int i = 0; i < MP; ++i) rb[i] = fma(wl * cB, fo[i], rb[i]);
    }
  }
#pragma unroll
  for (int j = 0; j < MU; ++j) eru[(int64_t)j * nc + c] += (double)vsigns[c * MU + j] * ru[j];
#pragma unroll
  for (int i = 0; i < MP; ++i) rB[c * MP + i] = rb[i];
  if (kB) {
#pragma unroll
    for (int i = 0; i < MP; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < MP; ++j) acc = fma(MpInv[(int64_t)(i * MP + j) * nc + c], rb[j], acc);
      kB[c * MP + i] = -acc;
    }
  }
}

void tswe_create(gg_rk* rk) {
  gg_ctx* ctx = rk->ctx;
  // Partitioned models: everything is cell-local on consistent data except r_B of a ghost cell (its far neighbours are not
  // local), so the B block of every stage state is made consistent through the peer windows (swe_step_async); the facet
  // tables give outer edges of the ghost layer the cell itself as neighbour (their rows are never used).
  const int nq = rk->nq1;
  facet_tables(rk, nq);
  // reference RT normal component at the facet points of the four local edges
  RTRef R = build_rt_ref(rk->order);
  std::vector<double> s(nq), wf(nq), pts((size_t)4 * nq * 2), val, div;
  gauss_legendre_01(nq, s.data(), wf.data());
  for (int e = 0; e < 4; ++e)
    for (int q = 0; q < nq; ++q)
      for (int d = 0; d < 2; ++d)
        pts[((size_t)e * nq + q) * 2 + d] = TQVERT[TQEDGE_V[e][0]][d] + s[q] * (TQVERT[TQEDGE_V[e][1]][d] - TQVERT[TQEDGE_V[e][0]][d]);
  R.tabulate(4 * nq, pts.data(), val, div);
  const int MU = R.ndofs;
  std::vector<double> fvn((size_t)4 * nq * MU);
  for (int e = 0; e < 4; ++e)
    for (int q = 0; q < nq; ++q)
      for (int j = 0; j < MU; ++j) {
        const size_t k = ((size_t)e * nq + q) * MU + j;
        fvn[k] = val[k * 2] * TQNORM[e][0] + val[k * 2 + 1] * TQNORM[e][1];
      }
  rk->tsw_fvn.upload(fvn, ctx->stream);
  rk->tsw_wf.upload(wf, ctx->stream);
  GG_CUDA(cudaStreamSynchronize(ctx->stream));
}

static TsweTabs tswe_tabs(gg_rk* rk) {
  gg_ctx* ctx = rk->ctx;
  auto trt = get_tab(ctx, GG_SPACE_RT, rk->order, rk->nq1);
  auto tdg = get_tab(ctx, GG_SPACE_DG, rk->order, rk->nq1);
  return {tdg->w.p, trt->val.p, trt->der.p, tdg->val.p, tdg->der.p, rk->adv_fval.p, rk->tsw_fvn.p, rk->tsw_wf.p};
}

// after the shallow-water diagnostics: Phi += 0.5 B, b = (p-weighted mass)^-1 int B r sqrtg
void tswe_diagnostics_async(gg_rk* rk, const double* x, double* y) {
  gg_ctx* ctx = rk->ctx;
  const int64_t nc = rk->mesh->n_cells;
  if (nc == 0) return;
  CellGeo g = cell_geo(rk->mesh, rk->nq1);
  TsweTabs T = tswe_tabs(rk);
  const double *p = x + rk->nu, *B = x + rk->nu + rk->np;
  double *Phi = y + rk->nr + rk->nu, *b = Phi + rk->np;
  ProfScope ps(ctx, "k_tswe_buoyancy");
  unsigned nb = nblk(nc, 64);
  switch (rk->order) {
    case 0: k_tswe_buoyancy<0><<<nb, 64, 0, ctx->stream>>>(nc, rk->nq1, g, T.w1, T.dgv, p, B, Phi, b); break;
    case 1: k_tswe_buoyancy<1><<<nb, 64, 0, ctx->stream>>>(nc, rk->nq1, g, T.w1, T.dgv, p, B, Phi, b); break;
    case 2: k_tswe_buoyancy<2><<<nb, 64, 0, ctx->stream>>>(nc, rk->nq1, g, T.w1, T.dgv, p, B, Phi, b); break;
    default: throw Error(GG_ERR_UNSUPPORTED, "thermal shallow-water order must be <= 2");
  }
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
}

// between the shallow-water stage kernel and the DOF gather of r_u: buoyancy terms into evec_u, r_B and (optionally) k_B
void tswe_residual_async(gg_rk* rk, const double* x, const double* y, double* rB, double* kB) {
  gg_ctx* ctx = rk->ctx;
  const int64_t nc = rk->mesh->n_cells;
  if (nc == 0) return;
  TsweTabs T = tswe_tabs(rk);
  const double* p = x + rk->nu;
  const double *F = y + rk->nr, *b = y + rk->nr + rk->nu + rk->np;
  gg_mesh* m = rk->mesh;
  ProfScope ps(ctx, "k_tswe_stage");
  unsigned nb = nblk(nc, 64);
#define GG_ARGS nc, rk->nq1, T, m->d_hx.p, m->d_hy.p, rk->V->d_cell_dofs.p, rk->V->d_cell_signs.p, rk->adv_nbr_cell.p, \
                rk->adv_nbr_edge.p, rk->adv_is_plus.p, p, b, F, rk->MpInv.p, rk->evec_u.p, rB, kB
  switch (rk->order) {
    case 0: k_tswe_stage<0><<<nb, 64, 0, ctx->stream>>>(GG_ARGS); break;
    case 1: k_tswe_stage<1><<<nb, 64, 0, ctx->stream>>>(GG_ARGS); break;
    case 2: k_tswe_stage<2><<<nb, 64, 0, ctx->stream>>>(GG_ARGS); break;
    default: throw Error(GG_ERR_UNSUPPORTED, "thermal shallow-water order must be <= 2");
  }
#undef GG_ARGS
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
}

}  // namespace gg
