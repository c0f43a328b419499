// Linearised Boussinesq equations on the 3-D cubed-sphere shell: RT_k x DG_k x DG_k on hexahedra, explicit Runge-Kutta.
//
// Replaces, for the explicit march of the hot path, what test/Tutorials/LinearBoussinesq.jl assembles through Gridap:
//   mass :112-114   int v.(g u)/meas + int q p meas + int r b meas           (constant: constant_mass = true, :134)
//   resu :116-120   int omega (n x (g u / meas)).(g v)/meas - int p div(v) - int b (n.v)
//   resp :124       int c^2 q div(u)
//   resb :128       int N^2 r (n.u)
// with the shell metric of src/Fields/CubedSphereWithThicknessMetric.jl:24-35 (g = diag(t^2, s^2 g_surf), s = 1 + t gamma / r,
// meas = t s^2 sqrt(det g_surf)), the unit radial normal normal_3D_cf = (1,0,0)/sqrt(ginv_11) = (t, 0, 0) (:103-105) and the
// contravariant Piola map of the hexahedral Raviart-Thomas basis on the axis-aligned chart boxes.
//
// Design: one thread per hexahedron, reference tabulations read through L1 (the same for every cell), geometry in FP64
// registers from two tangents per surface abscissa.  The stage residual kernel leaves the RT element vectors in the
// entry-major scratch (atomic-free DOF gather, gg_pattern.cu) and applies the inverse DG mass blocks in place, so a stage is:
// residual kernel, DOF gather, one PCG solve with the assembled RT mass.  (The tutorial itself marches with
// ThetaMethod + LU; implicit solvers are outside the hot path)
#include <cstring>
#include "gg_common.cuh"
#include "gg_refel.cuh"
#include "gg_rk.cuh"
#include "gg_dist_host.cuh"
#include "gg_geom.cuh"

namespace gg {

struct BqGeo {
  int64_t nc;
  int n, L, nq;
  double r, t;
  const int32_t *col, *layer;
  const double *xi, *w;   // 1-D abscissae / weights
};

static inline unsigned nblk(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// geometry of one cell: chart box and the tangents of its surface abscissae
struct BqCell {
  double hg, ha, hb, g0, det;
  double ta[8], tb[8];
};

__device__ __forceinline__ void bq_cell(const BqGeo& G, int64_t c, BqCell& C) {
  const double H = 0.78539816339744830962;
  const int cl = G.col[c] % (G.n * G.n), ia = cl % G.n, jb = cl / G.n;
  C.hg = 1.0 / G.L; C.ha = 2.0 * H / G.n; C.hb = 2.0 * H / G.n;
  C.g0 = (double)G.layer[c] / G.L;
  C.det = C.hg * C.ha * C.hb;
  const double a0 = -H + ia * C.ha, b0 = -H + jb * C.hb;
  for (int q = 0; q < G.nq; ++q) { C.ta[q] = tan(fma(C.ha, G.xi[q], a0)); C.tb[q] = tan(fma(C.hb, G.xi[q], b0)); }
}

// metric at the point (q0, q1, q2): g = diag(g00, [g11 g12; g12 g22]), meas, quadrature weight dV = w det
struct BqPoint { double g00, g11, g12, g22, meas, dV; };

__device__ __forceinline__ BqPoint bq_point(const BqGeo& G, const BqCell& C, int q0, int q1, int q2) {
  const double a = C.ta[q1], b = C.tb[q2];
  const double sa = fma(a, a, 1.0), sb = fma(b, b, 1.0), rho2 = sa + b * b, rho = sqrt(rho2);
  const double cs = G.r * G.r * sa * sb / (rho2 * rho2);               // CubedSphereMetric.jl:16-22
  const double s = 1.0 + G.t * fma(C.hg, G.xi[q0], C.g0) / G.r, s2 = s * s;
  BqPoint P;
  P.g00 = G.t * G.t;
  P.g11 = s2 * cs * sa; P.g12 = -s2 * cs * a * b; P.g22 = s2 * cs * sb;
  P.meas = G.t * s2 * G.r * G.r * sa * sb / (rho2 * rho);
  P.dV = G.w[q0] * G.w[q1] * G.w[q2] * C.det;
  return P;
}

// RT mass element matrices, full entry-major layout ebuf[(a * MU + b) * nc + c]; one thread per (cell, row a)
template <int MU>
__global__ void __launch_bounds__(128)
k_bq_rt_mass(BqGeo G, const double* __restrict__ rt_val, const int8_t* __restrict__ signs, double* __restrict__ ebuf) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= G.nc * MU) return;
  const int64_t c = t / MU;
  const int a = (int)(t - c * MU);
  BqCell C;
  bq_cell(G, c, C);
  double row[MU];
#pragma unroll
  for (int b = 0; b < MU; ++b) row[b] = 0.0;
  const double hd[3] = {C.hg / C.det, C.ha / C.det, C.hb / C.det};     // Piola: u_d = h_d uhat_d / det
  const int nq = G.nq;
  for (int q2 = 0; q2 < nq; ++q2)
    for (int q1 = 0; q1 < nq; ++q1)
      for (int q0 = 0; q0 < nq; ++q0) {
        const BqPoint P = bq_point(G, C, q0, q1, q2);
        const double* __restrict__ v = rt_val + (size_t)(q0 + nq * (q1 + nq * q2)) * MU * 3;
        const double f = P.dV / P.meas;
        const double u0 = hd[0] * v[a * 3], u1 = hd[1] * v[a * 3 + 1], u2 = hd[2] * v[a * 3 + 2];
        const double gu0 = f * P.g00 * u0, gu1 = f * (P.g11 * u1 + P.g12 * u2), gu2 = f * (P.g12 * u1 + P.g22 * u2);
#pragma unroll
        for (int b = 0; b < MU; ++b)
          row[b] = fma(gu0, hd[0] * v[b * 3], fma(gu1, hd[1] * v[b * 3 + 1], fma(gu2, hd[2] * v[b * 3 + 2], row[b])));
      }
  const double sa_ = (double)signs[c * MU + a];
#pragma unroll
  for (int b = 0; b < MU; ++b) ebuf[(int64_t)(a * MU + b) * G.nc + c] = sa_ * (double)signs[c * MU + b] * row[b];
}

// inverse of the DG mass blocks int psi_a psi_b meas, entry-major out[(a * MQ + b) * nc + c]; one thread per cell
template <int MQ>
__global__ void __launch_bounds__(64)
k_bq_dg_mass_inverse(BqGeo G, const double* __restrict__ dg_val, double* __restrict__ out) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= G.nc) return;
  BqCell C;
  bq_cell(G, c, C);
  double A[MQ][2 * MQ];
#pragma unroll
  for (int a = 0; a < MQ; ++a)
#pragma unroll
    for (int b = 0; b < 2 * MQ; ++b) A[a][b] = b == MQ + a ? 1.0 : 0.0;
  const int nq = G.nq;
  for (int q2 = 0; q2 < nq; ++q2)
    for (int q1 = 0; q1 < nq; ++q1)
      for (int q0 = 0; q0 < nq; ++q0) {
        const BqPoint P = bq_point(G, C, q0, q1, q2);
        const double* __restrict__ v = dg_val + (size_t)(q0 + nq * (q1 + nq * q2)) * MQ;
        const double f = P.dV * P.meas;
#pragma unroll
        for (int a = 0; a < MQ; ++a)
#pragma unroll
          for (int b = 0; b < MQ; ++b) A[a][b] = fma(f * v[a], v[b], A[a][b]);
      }
  // Gauss-Jordan without pivoting (symmetric positive definite)
#pragma unroll
  for (int k = 0; k < MQ; ++k) {
    const double d = 1.0 / A[k][k];
#pragma unroll
    for (int b = 0; b < 2 * MQ; ++b) A[k][b] *= d;
#pragma unroll
    for (int a = 0; a < MQ; ++a) {
      if (a == k) continue;
      const double f = A[a][k];
#pragma unroll
      for (int b = 0; b < 2 * MQ; ++b) A[a][b] = fma(-f, A[k][b], A[a][b]);
    }
  }
#pragma unroll
  for (int a = 0; a < MQ; ++a)
#pragma unroll
    for (int b = 0; b < MQ; ++b) out[(int64_t)(a * MQ + b) * G.nc + c] = A[a][MQ + b];
}

// stage residual: element vectors of resu into vbuf[m * nc + c]; r_p, r_b (cell-local DG DOFs) into r_pb[0 .. 2 nc MQ), and
// the slopes k = -Mp^-1 r of the two DG fields into k_pb if given.  x = [u (free RT DOFs); p; b]
template <int MU, int MQ>
__global__ void __launch_bounds__(64)
k_bq_residual(BqGeo G, const double* __restrict__ rt_val, const double* __restrict__ rt_div, const double* __restrict__ dg_val,
              const int32_t* __restrict__ rt_dofs, const int8_t* __restrict__ signs, const double* __restrict__ omega,
              double c2, double N2, int64_t nu, const double* __restrict__ x, const double* __restrict__ mpinv,
              double* __restrict__ vbuf, double* __restrict__ r_pb, double* __restrict__ k_pb) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= G.nc) return;
  BqCell C;
  bq_cell(G, c, C);
  const int64_t np = G.nc * MQ;
  double ue[MU], pe[MQ], be[MQ], ru[MU], rp[MQ], rb[MQ];
#pragma unroll
  for (int m = 0; m < MU; ++m) {
    const int32_t d = rt_dofs[c * MU + m];
    ue[m] = d >= 0 ? (double)signs[c * MU + m] * x[d] : 0.0;        // sign folded into the coefficient; Dirichlet value 0
    ru[m] = 0.0;
  }
#pragma unroll
  for (int a = 0; a < MQ; ++a) { pe[a] = x[nu + c * MQ + a]; be[a] = x[nu + np + c * MQ + a]; rp[a] = 0.0; rb[a] = 0.0; }
  const double hd[3] = {C.hg / C.det, C.ha / C.det, C.hb / C.det};
  const double idet = 1.0 / C.det;
  const int nq = G.nq;
  for (int q2 = 0; q2 < nq; ++q2)
    for (int q1 = 0; q1 < nq; ++q1)
      for (int q0 = 0; q0 < nq; ++q0) {
        const int q = q0 + nq * (q1 + nq * q2);
        const BqPoint P = bq_point(G, C, q0, q1, q2);
        const double* __restrict__ v = rt_val + (size_t)q * MU * 3;
        const double* __restrict__ dv = rt_div + (size_t)q * MU;
        const double* __restrict__ ps = dg_val + (size_t)q * MQ;
        double u0 = 0.0, u1 = 0.0, u2 = 0.0, divu = 0.0, ph = 0.0, bh = 0.0;
#pragma unroll
        for (int m = 0; m < MU; ++m) {
          u0 = fma(v[m * 3], ue[m], u0); u1 = fma(v[m * 3 + 1], ue[m], u1); u2 = fma(v[m * 3 + 2], ue[m], u2);
          divu = fma(dv[m], ue[m], divu);
        }
        u0 *= hd[0]; u1 *= hd[1]; u2 *= hd[2]; divu *= idet;
#pragma unroll
        for (int a = 0; a < MQ; ++a) { ph = fma(ps[a], pe[a], ph); bh = fma(ps[a], be[a], bh); }
        // g u / meas, then n x (.) with n = (t, 0, 0): (0, -t a2, t a1)
        const double im = 1.0 / P.meas;
        const double a1 = (P.g11 * u1 + P.g12 * u2) * im, a2 = (P.g12 * u1 + P.g22 * u2) * im;
        const double cor1 = -G.t * a2, cor2 = G.t * a1;
        const double fc = P.dV * omega[c * (int64_t)(nq * nq * nq) + q] * im;
        // test side: cor . (g v) = cor1 (g11 v1 + g12 v2) + cor2 (g12 v1 + g22 v2)
        const double w1 = fc * (cor1 * P.g11 + cor2 * P.g12) * hd[1], w2 = fc * (cor1 * P.g12 + cor2 * P.g22) * hd[2];
        const double wp = -P.dV * ph * idet, wb = -P.dV * bh * G.t * hd[0];
#pragma unroll
        for (int m = 0; m < MU; ++m)
          ru[m] = fma(w1, v[m * 3 + 1], fma(w2, v[m * 3 + 2], fma(wp, dv[m], fma(wb, v[m * 3], ru[m]))));
        const double fp = c2 * P.dV * divu, fb = N2 * P.dV * G.t * u0;
#pragma unroll
        for (int a = 0; a < MQ; ++a) { rp[a] = fma(fp, ps[a], rp[a]); rb[a] = fma(fb, ps[a], rb[a]); }
      }
#pragma unroll
  for (int m = 0; m < MU; ++m) vbuf[(int64_t)m * G.nc + c] = (double)signs[c * MU + m] * ru[m];
#pragma unroll
  for (int a = 0; a < MQ; ++a) { r_pb[c * MQ + a] = rp[a]; r_pb[np + c * MQ + a] = rb[a]; }
  if (k_pb) {
#pragma unroll
    for (int a = 0; a < MQ; ++a) {
      double sp = 0.0, sb = 0.0;
#pragma unroll
      for (int b = 0; b < MQ; ++b) {
        const double mi = mpinv[(int64_t)(a * MQ + b) * G.nc + c];
        sp = fma(mi, rp[b], sp); sb = fma(mi, rb[b], sb);
      }
      k_pb[c * MQ + a] = -sp; k_pb[np + c * MQ + a] = -sb;
    }
  }
}

// ---- interpolation of an ambient vector field into the hexahedral RT space (LinearBoussinesq.jl:107-108, 115):
//      u_cf = meas * (pinvJ(covariant basis) . u0(X)), then the moment functionals of the reference element ---------------------
__global__ void k_bq_interp_points(BqGeo G, const int32_t* __restrict__ chart_id, int N, const double* __restrict__ ref,
                                   double* __restrict__ chart, double* __restrict__ xyz) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= G.nc * N) return;
  const int64_t c = t / N;
  const int q = (int)(t - c * N);
  const double H = 0.78539816339744830962;
  const int cl = G.col[c] % (G.n * G.n), ia = cl % G.n, jb = cl / G.n;
  const double gam = ((double)G.layer[c] + ref[q * 3]) / G.L;
  const double al = -H + (ia + ref[q * 3 + 1]) * 2.0 * H / G.n, be = -H + (jb + ref[q * 3 + 2]) * 2.0 * H / G.n;
  if (chart) { chart[t * 3] = gam; chart[t * 3 + 1] = al; chart[t * 3 + 2] = be; }
  if (xyz) {
    double X[3];
    csphere_point(chart_id[c], tan(al), tan(be), G.r, X);
    const double s = 1.0 + G.t * gam / G.r;
    xyz[t * 3] = s * X[0]; xyz[t * 3 + 1] = s * X[1]; xyz[t * 3 + 2] = s * X[2];
  }
}

// one thread per (cell, local DOF): dof = sign * sum_n W[m][n][:] . uhat(pt n), uhat = det J_ref^-1 (meas J^-T v); the owner of a
// DOF (sign +1) writes it
template <int MU>
__global__ void __launch_bounds__(128)
k_bq_interpolate(BqGeo G, const int32_t* __restrict__ chart_id, int N, const double* __restrict__ ref, const double* __restrict__ W,
                 const int32_t* __restrict__ dofs, const int8_t* __restrict__ signs, const double* __restrict__ vals,
                 double* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= G.nc * MU) return;
  const int64_t c = t / MU;
  const int m = (int)(t - c * MU);
  const int32_t d = dofs[c * MU + m];
  if (d < 0 || signs[c * MU + m] < 0) return;
  const double H = 0.78539816339744830962;
  const int cl = G.col[c] % (G.n * G.n), ia = cl % G.n, jb = cl / G.n;
  const double h[3] = {1.0 / G.L, 2.0 * H / G.n, 2.0 * H / G.n}, det = h[0] * h[1] * h[2];
  double acc = 0.0;
  for (int q = 0; q < N; ++q) {
    const double* __restrict__ w = W + ((size_t)m * N + q) * 3;
    if (w[0] == 0.0 && w[1] == 0.0 && w[2] == 0.0) continue;
    const double gam = ((double)G.layer[c] + ref[q * 3]) / G.L;
    const double al = -H + (ia + ref[q * 3 + 1]) * 2.0 * H / G.n, be = -H + (jb + ref[q * 3 + 2]) * 2.0 * H / G.n;
    const double a = tan(al), b = tan(be);
    double X[3], Js[2][3];
    csphere_point(chart_id[c], a, b, G.r, X);
    csphere_jacobian(chart_id[c], a, b, G.r, Js);
    const double s = 1.0 + G.t * gam / G.r;
    // rows of the 3-D Jacobian J[i][k] = d X_k / d x_i (CubedSphereWithThicknessMap.jl:31-44)
    double J[3][3];
    for (int k = 0; k < 3; ++k) { J[0][k] = G.t / G.r * X[k]; J[1][k] = s * Js[0][k]; J[2][k] = s * Js[1][k]; }
    const double* v = vals + ((size_t)c * N + q) * 3;
    // contravariant components: sum_i cc_i J[i][:] = v  (Cramer on J^T)
    const double D = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                     J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    double cc[3];
    for (int i = 0; i < 3; ++i) {
      double Mx[3][3];
      for (int r2 = 0; r2 < 3; ++r2)
        for (int k = 0; k < 3; ++k) Mx[r2][k] = r2 == i ? v[k] : J[r2][k];
      cc[i] = (Mx[0][0] * (Mx[1][1] * Mx[2][2] - Mx[1][2] * Mx[2][1]) - Mx[0][1] * (Mx[1][0] * Mx[2][2] - Mx[1][2] * Mx[2][0]) +
               Mx[0][2] * (Mx[1][0] * Mx[2][1] - Mx[1][1] * Mx[2][0])) / D;
    }
    const double sa = fma(a, a, 1.0), sb = fma(b, b, 1.0), rho2 = sa + b * b;
    const double meas = G.t * s * s * G.r * G.r * sa * sb / (rho2 * sqrt(rho2));
    for (int k = 0; k < 3; ++k) acc = fma(w[k], det * meas * cc[k] / h[k], acc);
  }
  out[d] = acc;
}

static BqGeo mesh_geo(gg_mesh* m) {
  BqGeo G;
  G.nc = m->n_cells; G.n = m->n; G.L = m->n_layers; G.nq = 0; G.r = m->radius; G.t = m->thickness;
  G.col = m->d_col.p; G.layer = m->d_layer.p; G.xi = nullptr; G.w = nullptr;
  return G;
}

int bq_interp_points(gg_space* s, double* chart, double* xyz) {
  gg_mesh* m = s->mesh;
  gg_ctx* ctx = m->ctx;
  std::vector<double> ref, W;
  rt_hex_moment_functionals(s->order, ref, W);
  const int N = (int)(ref.size() / 3);
  const int64_t tot = m->n_cells * N;
  if ((!chart && !xyz) || tot == 0) return N;
  DevBuf<double> dref, dc(chart ? 3 * tot : 0), dx(xyz ? 3 * tot : 0);
  dref.upload(ref, ctx->stream);
  k_bq_interp_points<<<nblk(tot, 256), 256, 0, ctx->stream>>>(mesh_geo(m), m->d_chart.p, N, dref.p, chart ? dc.p : nullptr, xyz ? dx.p : nullptr);
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
  if (chart) dc.download(chart, ctx->stream);
  if (xyz) dx.download(xyz, ctx->stream);
  GG_CUDA(cudaStreamSynchronize(ctx->stream));
  return N;
}

void bq_interpolate(gg_space* s, const double* vals, double* out) {
  gg_mesh* m = s->mesh;
  gg_ctx* ctx = m->ctx;
  std::vector<double> ref, W;
  rt_hex_moment_functionals(s->order, ref, W);
  const int N = (int)(ref.size() / 3);
  const int64_t nc = m->n_cells;
  if (nc == 0) return;
  DevBuf<double> dref, dW, dv;
  dref.upload(ref, ctx->stream); dW.upload(W, ctx->stream); dv.upload(vals, (size_t)nc * N * 3, ctx->stream);
  if (s->ndofs_loc == 6)
    k_bq_interpolate<6><<<nblk(nc * 6, 128), 128, 0, ctx->stream>>>(mesh_geo(m), m->d_chart.p, N, dref.p, dW.p, s->d_cell_dofs.p, s->d_cell_signs.p, dv.p, out);
  else
    k_bq_interpolate<36><<<nblk(nc * 36, 128), 128, 0, ctx->stream>>>(mesh_geo(m), m->d_chart.p, N, dref.p, dW.p, s->d_cell_dofs.p, s->d_cell_signs.p, dv.p, out);
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
  GG_CUDA(cudaStreamSynchronize(ctx->stream));
}

static BqGeo bq_geo(gg_rk* rk) {
  gg_mesh* m = rk->mesh;
  BqGeo G;
  G.nc = m->n_cells; G.n = m->n; G.L = m->n_layers; G.nq = rk->nq1; G.r = m->radius; G.t = m->thickness;
  G.col = m->d_col.p; G.layer = m->d_layer.p; G.xi = rk->bq_xi.p; G.w = rk->bq_w.p;
  return G;
}

void bq_residual_async(gg_rk* rk, const double* x, double* r, double* k_pb) {
  gg_ctx* ctx = rk->ctx;
  const int64_t nc = rk->mesh->n_cells;
  if (nc == 0) return;
  BqGeo G = bq_geo(rk);
  double* vb = rk->evec_u.p;
  if (rk->mu == 6)
    k_bq_residual<6, 1><<<nblk(nc, 64), 64, 0, ctx->stream>>>(G, rk->bq_rt_val.p, rk->bq_rt_div.p, rk->bq_dg_val.p, rk->V->d_cell_dofs.p,
                                                              rk->V->d_cell_signs.p, rk->fq.p, rk->bq_c2, rk->bq_N2, rk->nu, x, rk->MpInv.p,
                                                              vb, r + rk->nu, k_pb);
  else
    k_bq_residual<36, 8><<<nblk(nc, 64), 64, 0, ctx->stream>>>(G, rk->bq_rt_val.p, rk->bq_rt_div.p, rk->bq_dg_val.p, rk->V->d_cell_dofs.p,
                                                               rk->V->d_cell_signs.p, rk->fq.p, rk->bq_c2, rk->bq_N2, rk->nu, x, rk->MpInv.p,
                                                               vb, r + rk->nu, k_pb);
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
  gather_vector(rk->V, vb, r, 1.0, 0);
}

void bq_step_async(gg_rk* rk) {
  const int64_t n = rk_nx(rk);
  const int s = rk->n_stages;
  for (int i = 0; i < s; ++i) {
    const double* xs = rk->x.p;
    if (i > 0) {
      double co[4] = {0, 0, 0, 0};
      for (int j = 0; j < i && j < 4; ++j) co[j] = rk->dt * rk->A[i * s + j];
      rk_combine(rk, std::min(i, 4), co, rk->xi.p);
      xs = rk->xi.p;
    }
    bq_residual_async(rk, xs, rk->r.p, rk->k[i].p + rk->nu);
    rk_mass_solve(rk, 0, rk->r.p, -1.0, rk->k[i].p, i);          // M_u k_u = -r_u
  }
  double co[4] = {0, 0, 0, 0};
  for (int j = 0; j < s && j < 4; ++j) co[j] = rk->dt * rk->b[j];
  rk_combine(rk, std::min(s, 4), co, rk->x.p);
  (void)n;
}

void bq_create(gg_rk* rk, const gg_rk_args* a) {
  gg_mesh* m = rk->mesh;
  gg_ctx* ctx = rk->ctx;
  cudaStream_t st = ctx->stream;
  const int k = a->order;
  GG_REQUIRE(m->dim == 3, GG_ERR_INVALID, "the Boussinesq stepper runs on the 3-D shell");
  GG_REQUIRE(k == 0 || k == 1, GG_ERR_UNSUPPORTED, "the Boussinesq stepper supports RT_k x DG_k x DG_k with k <= 1 on hexahedra");
  GG_REQUIRE(m->n > 0 && (int64_t)m->cell_col.size() == m->n_cells, GG_ERR_UNSUPPORTED, "the Boussinesq stepper needs a builtin shell mesh");
  GG_REQUIRE(a->coriolis, GG_ERR_INVALID, "gg_rk_args.coriolis: the rotation rate omega at the quadrature points is required");
  auto chk = [](int code) { if (code != GG_OK) throw Error(code, gg_last_error(nullptr)); };
  chk(gg_space_builtin(m, GG_SPACE_RT, k, GG_CONSTRAINT_NOFLUX, &rk->V));
  chk(gg_space_builtin(m, GG_SPACE_DG, k, GG_CONSTRAINT_NONE, &rk->Q));
  rk->nu = rk->V->n_free; rk->np = rk->Q->n_free; rk->nB = rk->np;
  rk->mu = rk->V->ndofs_loc; rk->mq = rk->Q->ndofs_loc;
  rk->nq1 = num_points_1d(a->degree);
  GG_REQUIRE(rk->nq1 <= 8, GG_ERR_UNSUPPORTED, "quadrature degree too high for the Boussinesq kernels (max 15)");
  rk->bq_c2 = a->sound_speed > 0 ? a->sound_speed * a->sound_speed : 1.0;
  rk->bq_N2 = a->buoyancy_frequency > 0 ? a->buoyancy_frequency * a->buoyancy_frequency : 1.0;
  // reference tabulations at the tensor Gauss points (gamma fastest)
  const int nq = rk->nq1, Q = nq * nq * nq;
  std::vector<double> xi(nq), w(nq), pts((size_t)Q * 3), val, div, dg((size_t)Q * rk->mq);
  gauss_legendre_01(nq, xi.data(), w.data());
  for (int q2 = 0; q2 < nq; ++q2)
    for (int q1 = 0; q1 < nq; ++q1)
      for (int q0 = 0; q0 < nq; ++q0) {
        double* p = &pts[(size_t)(q0 + nq * (q1 + nq * q2)) * 3];
        p[0] = xi[q0]; p[1] = xi[q1]; p[2] = xi[q2];
      }
  RTRefHex R = build_rt_hex_ref(k);
  GG_REQUIRE(R.ndofs == rk->mu, GG_ERR_INVALID, "hexahedral RT size mismatch");
  R.tabulate(Q, pts.data(), val, div);
  // DG_k on the hexahedron: Gridap local node order -> lexicographic tensor index (gamma fastest)
  std::vector<int32_t> l2x = hex_local_to_lex(k);
  const int nb = k + 1;
  for (int q = 0; q < Q; ++q) {
    double N[3][MAX_NB], D[MAX_NB];
    for (int d = 0; d < 3; ++d) lagrange_1d(k, pts[(size_t)q * 3 + d], N[d], D);
    for (int a2 = 0; a2 < rk->mq; ++a2) {
      const int lx = l2x[a2], i0 = lx % nb, i1 = (lx / nb) % nb, i2 = lx / (nb * nb);
      dg[(size_t)q * rk->mq + a2] = N[0][i0] * N[1][i1] * N[2][i2];
    }
  }
  rk->bq_xi.upload(xi, st); rk->bq_w.upload(w, st);
  rk->bq_rt_val.upload(val, st); rk->bq_rt_div.upload(div, st); rk->bq_dg_val.upload(dg, st);
  rk->fq.upload(a->coriolis, (size_t)m->n_cells * Q, st);
  const int64_t nc = m->n_cells;
  BqGeo G = bq_geo(rk);
  // RT mass: assembled once (constant), block-Jacobi PCG on it
  chk(gg_pattern(rk->V, rk->V, &rk->Mu));
  if (nc) {
    if (ctx->ebuf.n < (size_t)rk->mu * rk->mu * nc) ctx->ebuf.alloc((size_t)rk->mu * rk->mu * nc);
    double* eb = ctx->ebuf.p;
    if (rk->mu == 6) k_bq_rt_mass<6><<<nblk(nc * 6, 128), 128, 0, st>>>(G, rk->bq_rt_val.p, rk->V->d_cell_signs.p, eb);
    else k_bq_rt_mass<36><<<nblk(nc * 36, 128), 128, 0, st>>>(G, rk->bq_rt_val.p, rk->V->d_cell_signs.p, eb);
    ctx->launches++;
    GG_CUDA(cudaGetLastError());
    gather_matrix(rk->Mu, 0, eb, 1.0, 0);
    rk->MpInv.alloc((size_t)rk->mq * rk->mq * nc);
    if (rk->mq == 1) k_bq_dg_mass_inverse<1><<<nblk(nc, 64), 64, 0, st>>>(G, rk->bq_dg_val.p, rk->MpInv.p);
    else k_bq_dg_mass_inverse<8><<<nblk(nc, 64), 64, 0, st>>>(G, rk->bq_dg_val.p, rk->MpInv.p);
    ctx->launches++;
    GG_CUDA(cudaGetLastError());
  }
  chk(gg_solver_pcg_create(rk->Mu, a->rtol > 0 ? a->rtol : 1e-12, 2000, &rk->solver));
  build_vec_map(rk->V);
  rk->evec_u.alloc((size_t)rk->mu * std::max<int64_t>(nc, 1));
  const int64_t n = rk_nx(rk);
  const int s = rk->n_stages;
  rk->x.alloc(n); rk->xi.alloc(n); rk->r.alloc(n);
  rk->k.resize(s);
  for (int i = 0; i < s; ++i) {
    rk->k[i].alloc(n);
    if (n) GG_CUDA(cudaMemsetAsync(rk->k[i].p, 0, n * sizeof(double), st));
  }
  rk->iter_accum.alloc(3);
  GG_CUDA(cudaMemsetAsync(rk->iter_accum.p, 0, 3 * sizeof(double), st));
  if (n) {
    GG_CUDA(cudaMemsetAsync(rk->x.p, 0, n * sizeof(double), st));
    GG_CUDA(cudaMemsetAsync(rk->r.p, 0, n * sizeof(double), st));
  }
  GG_CUDA(cudaStreamSynchronize(st));
}

}  // namespace gg
