// Vector-valued H1 Lagrangian space on the intrinsic cubed sphere (GG_SPACE_CGV): panel-interface change of basis, the
// mass / source forms of the L2 projection driver, interpolation and the error norm.
//
// Reference: src/FESpaces/GradConformingFESpaces.jl:1-113 (master cells, GenerateChangeOfBasisMatrixMap) and
// test/Projection/L2_projection_Lagrangian_vector.jl:17-58.  A DOF holds a contravariant component in the chart of the master
// cell of its vertex / edge; a cell on another panel evaluates  u(x) = sum_a N_a(x) C_a U_a  with the 2x2 block
// C_a = pinvJ(J_cell)(x_a) . J_master^T(x_a)  (identity away from the 12 panel interfaces), so element matrices are
// C^T K C block by block: no dense change-of-basis matrix is ever formed.
#include <cub/cub.cuh>
#include "gg_common.cuh"
#include "gg_refel.cuh"
#include "gg_geom.cuh"
#include "gg_vec.cuh"

namespace gg {

static inline unsigned nblk(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// one thread per (cell, node): the 2x2 block, row = component in this cell's chart, column = component in the master's chart
__global__ void k_vec_cob(int64_t nc, int nn, const double* __restrict__ ref, const double* __restrict__ x0,
                          const double* __restrict__ y0, const double* __restrict__ hx, const double* __restrict__ hy,
                          const int32_t* __restrict__ chart, double radius, const int32_t* __restrict__ mpanel,
                          const double* __restrict__ mpoint, double* __restrict__ cob) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * nn) return;
  const int64_t c = t / nn;
  const int a = (int)(t - c * nn);
  double* o = cob + 4 * t;
  const int32_t mp = mpanel[t];
  if (mp < 0) { o[0] = 1.0; o[1] = 0.0; o[2] = 0.0; o[3] = 1.0; return; }
  double JS[2][3], JM[2][3];
  const double as = tan(fma(hx[c], ref[2 * a], x0[c])), bs = tan(fma(hy[c], ref[2 * a + 1], y0[c]));
  csphere_jacobian(chart[c], as, bs, radius, JS);
  csphere_jacobian(mp, tan(mpoint[2 * t]), tan(mpoint[2 * t + 1]), radius, JM);
  const Metric2 m = csphere_metric_terms(as, bs, radius);
  // pinvJ = ginv . J  (src/Helpers/Operators.jl:16-24 for the 2x3 covariant basis)
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const double d0 = JS[0][0] * JM[j][0] + JS[0][1] * JM[j][1] + JS[0][2] * JM[j][2];
    const double d1 = JS[1][0] * JM[j][0] + JS[1][1] * JM[j][1] + JS[1][2] * JM[j][2];
    o[0 * 2 + j] = m.i11 * d0 + m.i12 * d1;
    o[1 * 2 + j] = m.i12 * d0 + m.i22 * d1;
  }
}

void vec_space_finish(gg_space* s) {
  gg_mesh* m = s->mesh;
  if (!m->ctx || s->kind != GG_SPACE_CGV) return;
  gg_ctx* ctx = m->ctx;
  cudaStream_t st = ctx->stream;
  const int nn = s->ndofs_loc / 2;
  const int64_t nc = m->n_cells;
  std::vector<int32_t> l2x = quad_local_to_lex(s->order);
  const int nb = s->order + 1;
  std::vector<double> ref((size_t)nn * 2);
  for (int a = 0; a < nn; ++a) { ref[2 * a] = (double)(l2x[a] % nb) / s->order; ref[2 * a + 1] = (double)(l2x[a] / nb) / s->order; }
  DevBuf<double> dref;
  DevBuf<int32_t> dmc;
  DevBuf<double> dmn;
  if (s->cob_master_panel.size() != (size_t)nc * nn) {   // caller-numbered space without master data: identity blocks
    s->cob_master_panel.assign((size_t)nc * nn, -1);
    s->cob_master_point.assign((size_t)nc * nn * 2, 0.0);
  }
  dref.upload(ref, st); dmc.upload(s->cob_master_panel, st); dmn.upload(s->cob_master_point, st);
  s->d_cob.alloc((size_t)nc * nn * 4);
  if (nc) {
    k_vec_cob<<<nblk(nc * nn, 128), 128, 0, st>>>(nc, nn, dref.p, m->d_x0.p, m->d_y0.p, m->d_hx.p, m->d_hy.p, m->d_chart.p, m->radius,
                                                 dmc.p, dmn.p, s->d_cob.p);
    ctx->launches++;
    GG_CUDA(cudaGetLastError());
  }
  GG_CUDA(cudaStreamSynchronize(st));
}

// element matrix of int (u.(g v)) sqrtg, one thread per (cell, node pair a, b): G[k][l] = int N_a N_b g_kl sqrtg, then the
// block C_a^T G C_b goes to the full entry-major buffer ebuf[(a + nn i) * m + (b + nn j)][cell]
__global__ void k_vec_mass(int64_t nc, int nq, int nn, const double* __restrict__ xi, const double* __restrict__ w,
                           const double* __restrict__ val, const double* __restrict__ x0, const double* __restrict__ y0,
                           const double* __restrict__ hx, const double* __restrict__ hy, double radius,
                           const double* __restrict__ cob, double* __restrict__ ebuf) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * nn * nn) return;
  const int64_t c = t % nc;               // consecutive threads -> consecutive cells: coalesced stores
  const int ab = (int)(t / nc), a = ab / nn, b = ab - a * nn;
  const double HX = hx[c], HY = hy[c];
  double G11 = 0.0, G12 = 0.0, G22 = 0.0;
  for (int q2 = 0; q2 < nq; ++q2) {
    const double tb = tan(fma(HY, xi[q2], y0[c]));
    for (int q1 = 0; q1 < nq; ++q1) {
      const int q = q1 + nq * q2;
      const Metric2 mt = csphere_metric_terms(tan(fma(HX, xi[q1], x0[c])), tb, radius);
      const double f = w[q1] * w[q2] * HX * HY * mt.sqrtg * val[q * nn + a] * val[q * nn + b];
      G11 = fma(f, mt.g11, G11); G12 = fma(f, mt.g12, G12); G22 = fma(f, mt.g22, G22);
    }
  }
  const double* Ca = cob + 4 * (c * nn + a);
  const double* Cb = cob + 4 * (c * nn + b);
  const double G[2][2] = {{G11, G12}, {G12, G22}};
  const int m = 2 * nn;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int l = 0; l < 2; ++l) v = fma(Ca[k * 2 + i] * G[k][l], Cb[l * 2 + j], v);
      ebuf[(int64_t)((a + nn * i) * m + (b + nn * j)) * nc + c] = v;
    }
}

// element vector of int (f.(g v)) sqrtg, one thread per (cell, node): b[k] = int N_a (g f)_k sqrtg, then C_a^T b
__global__ void k_vec_source(int64_t nc, int nq, int nn, const double* __restrict__ xi, const double* __restrict__ w,
                             const double* __restrict__ val, const double* __restrict__ x0, const double* __restrict__ y0,
                             const double* __restrict__ hx, const double* __restrict__ hy, double radius,
                             const double* __restrict__ cob, const double* __restrict__ fq, double* __restrict__ vbuf) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * nn) return;
  const int64_t c = t % nc;
  const int a = (int)(t / nc);
  const double HX = hx[c], HY = hy[c];
  const int Q = nq * nq;
  double b0 = 0.0, b1 = 0.0;
  for (int q2 = 0; q2 < nq; ++q2) {
    const double tb = tan(fma(HY, xi[q2], y0[c]));
    for (int q1 = 0; q1 < nq; ++q1) {
      const int q = q1 + nq * q2;
      const Metric2 mt = csphere_metric_terms(tan(fma(HX, xi[q1], x0[c])), tb, radius);
      const double f = w[q1] * w[q2] * HX * HY * mt.sqrtg * val[q * nn + a];
      const double f0 = fq[(c * Q + q) * 2], f1 = fq[(c * Q + q) * 2 + 1];
      b0 = fma(f, mt.g11 * f0 + mt.g12 * f1, b0);
      b1 = fma(f, mt.g12 * f0 + mt.g22 * f1, b1);
    }
  }
  const double* Ca = cob + 4 * (c * nn + a);
  vbuf[(int64_t)a * nc + c] = Ca[0] * b0 + Ca[2] * b1;
  vbuf[(int64_t)(a + nn) * nc + c] = Ca[1] * b0 + Ca[3] * b1;
}

void vec_mass_element_matrices(gg_space* s, int nq, double* ebuf) {
  gg_mesh* m = s->mesh;
  gg_ctx* ctx = m->ctx;
  const int nn = s->ndofs_loc / 2;
  const int64_t nc = m->n_cells;
  auto T = get_tab(ctx, GG_SPACE_CG, s->order, nq);
  k_vec_mass<<<nblk(nc * nn * nn, 128), 128, 0, ctx->stream>>>(nc, nq, nn, T->xi.p, T->w.p, T->val.p, m->d_x0.p, m->d_y0.p, m->d_hx.p,
                                                               m->d_hy.p, m->radius, s->d_cob.p, ebuf);
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
}

void vec_source_element_vectors(gg_space* s, int nq, const double* fq, double* vbuf) {
  gg_mesh* m = s->mesh;
  gg_ctx* ctx = m->ctx;
  const int nn = s->ndofs_loc / 2;
  const int64_t nc = m->n_cells;
  auto T = get_tab(ctx, GG_SPACE_CG, s->order, nq);
  k_vec_source<<<nblk(nc * nn, 128), 128, 0, ctx->stream>>>(nc, nq, nn, T->xi.p, T->w.p, T->val.p, m->d_x0.p, m->d_y0.p, m->d_hx.p,
                                                           m->d_hy.p, m->radius, s->d_cob.p, fq, vbuf);
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
}

// ---- vector Laplacian / surface Stokes blocks (test/SurfaceStokes/VectorLaplacian2D.jl:38-80, SurfaceStokes_TaylorHood.jl) --------
// For the untransformed shape function N_a e_i the two first-order operators of the reference's weak form are
//   D1 = div(sqrtg u)          = sqrtg d_i N_a + N_a d_i sqrtg                                   ("divg_product_rule")
//   D2 = -curl(g u)            = -(g_2i d_1 N_a - g_1i d_2 N_a + N_a (d_1 g_2i - d_2 g_1i))      ("divergenceRgu")
// so per quadrature point 8 numbers describe both: D1 = p_i . (dN_i, N), D2 = r_i . (dN_1, dN_2, N).
struct VecLapPoint {
  double sqrtg, gs0, gs1;        // sqrtg and its chart gradient, 1/(2 sqrtg) cof(g) : d_k g   (deriv_sqrt, deriv_det, cpAB)
  double g11, g12, g22, c0, c1;  // metric and the curl of its columns
};

__device__ __forceinline__ VecLapPoint vec_lap_point(double ta, double tb, double radius) {
  const Metric2 mt = csphere_metric_terms(ta, tb, radius);
  double dg[2][3], dgi[2][3];
  csphere_metric_gradients(ta, tb, radius, dg, dgi);
  VecLapPoint P;
  P.sqrtg = mt.sqrtg;
  const double h = 0.5 / mt.sqrtg;
  P.gs0 = h * (mt.g22 * dg[0][0] + mt.g11 * dg[0][2] - 2.0 * mt.g12 * dg[0][1]);
  P.gs1 = h * (mt.g22 * dg[1][0] + mt.g11 * dg[1][2] - 2.0 * mt.g12 * dg[1][1]);
  P.g11 = mt.g11; P.g12 = mt.g12; P.g22 = mt.g22;
  P.c0 = dg[0][1] - dg[1][0];
  P.c1 = dg[0][2] - dg[1][1];
  return P;
}

// element matrix of int (D1(u) D1(v) - D2(u) D2(v)) / sqrtg, one thread per (cell, node pair a, b): the 2x2 block of the
// untransformed functions, then C_a^T G C_b into the full entry-major buffer (same layout as k_vec_mass)
__global__ void k_vec_laplacian(int64_t nc, int nq, int nn, const double* __restrict__ xi, const double* __restrict__ w,
                                const double* __restrict__ val, const double* __restrict__ der, const double* __restrict__ x0,
                                const double* __restrict__ y0, const double* __restrict__ hx, const double* __restrict__ hy,
                                double radius, const double* __restrict__ cob, double* __restrict__ ebuf) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * nn * nn) return;
  const int64_t c = t % nc;
  const int ab = (int)(t / nc), a = ab / nn, b = ab - a * nn;
  const double HX = hx[c], HY = hy[c];
  double G[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  for (int q2 = 0; q2 < nq; ++q2) {
    const double tb = tan(fma(HY, xi[q2], y0[c]));
    for (int q1 = 0; q1 < nq; ++q1) {
      const int q = q1 + nq * q2;
      const VecLapPoint P = vec_lap_point(tan(fma(HX, xi[q1], x0[c])), tb, radius);
      const double f = w[q1] * w[q2] * HX * HY / P.sqrtg;
      const double Na = val[q * nn + a], ax = der[(q * nn + a) * 2] / HX, ay = der[(q * nn + a) * 2 + 1] / HY;
      const double Nb = val[q * nn + b], bx = der[(q * nn + b) * 2] / HX, by = der[(q * nn + b) * 2 + 1] / HY;
      const double d1a[2] = {fma(P.sqrtg, ax, Na * P.gs0), fma(P.sqrtg, ay, Na * P.gs1)};
      const double d1b[2] = {fma(P.sqrtg, bx, Nb * P.gs0), fma(P.sqrtg, by, Nb * P.gs1)};
      const double d2a[2] = {-(P.g12 * ax - P.g11 * ay + Na * P.c0), -(P.g22 * ax - P.g12 * ay + Na * P.c1)};
      const double d2b[2] = {-(P.g12 * bx - P.g11 * by + Nb * P.c0), -(P.g22 * bx - P.g12 * by + Nb * P.c1)};
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) G[i][j] = fma(f, d1a[i] * d1b[j] - d2a[i] * d2b[j], G[i][j]);
    }
  }
  const double* Ca = cob + 4 * (c * nn + a);
  const double* Cb = cob + 4 * (c * nn + b);
  const int m = 2 * nn;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int l = 0; l < 2; ++l) v = fma(Ca[k * 2 + i] * G[k][l], Cb[l * 2 + j], v);
      ebuf[(int64_t)((a + nn * i) * m + (b + nn * j)) * nc + c] = v;
    }
}

// element matrix of int q D1(u) (rows: scalar test functions, columns: vector trial functions), one thread per
// (cell, scalar function I, node b): B_k = int q_I D1(N_b e_k), then B . C_b
__global__ void k_vec_div(int64_t nc, int nq, int mt, int nn, const double* __restrict__ xi, const double* __restrict__ w,
                          const double* __restrict__ tval, const double* __restrict__ val, const double* __restrict__ der,
                          const double* __restrict__ x0, const double* __restrict__ y0, const double* __restrict__ hx,
                          const double* __restrict__ hy, double radius, const double* __restrict__ cob, double* __restrict__ ebuf) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * mt * nn) return;
  const int64_t c = t % nc;
  const int Ib = (int)(t / nc), I = Ib / nn, b = Ib - I * nn;
  const double HX = hx[c], HY = hy[c];
  double B0 = 0.0, B1 = 0.0;
  for (int q2 = 0; q2 < nq; ++q2) {
    const double tb = tan(fma(HY, xi[q2], y0[c]));
    for (int q1 = 0; q1 < nq; ++q1) {
      const int q = q1 + nq * q2;
      const VecLapPoint P = vec_lap_point(tan(fma(HX, xi[q1], x0[c])), tb, radius);
      const double f = w[q1] * w[q2] * HX * HY * tval[q * mt + I];
      const double Nb = val[q * nn + b], bx = der[(q * nn + b) * 2] / HX, by = der[(q * nn + b) * 2 + 1] / HY;
      B0 = fma(f, fma(P.sqrtg, bx, Nb * P.gs0), B0);
      B1 = fma(f, fma(P.sqrtg, by, Nb * P.gs1), B1);
    }
  }
  const double* Cb = cob + 4 * (c * nn + b);
  const int m = 2 * nn;
  ebuf[(int64_t)(I * m + b) * nc + c] = B0 * Cb[0] + B1 * Cb[2];
  ebuf[(int64_t)(I * m + b + nn) * nc + c] = B0 * Cb[1] + B1 * Cb[3];
}

void vec_laplacian_element_matrices(gg_space* s, int nq, double* ebuf) {
  gg_mesh* m = s->mesh;
  gg_ctx* ctx = m->ctx;
  const int nn = s->ndofs_loc / 2;
  const int64_t nc = m->n_cells;
  auto T = get_tab(ctx, GG_SPACE_CG, s->order, nq);
  k_vec_laplacian<<<nblk(nc * nn * nn, 128), 128, 0, ctx->stream>>>(nc, nq, nn, T->xi.p, T->w.p, T->val.p, T->der.p, m->d_x0.p, m->d_y0.p,
                                                                    m->d_hx.p, m->d_hy.p, m->radius, s->d_cob.p, ebuf);
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
}

void vec_div_element_matrices(gg_space* test, gg_space* trial, int nq, double* ebuf) {
  gg_mesh* m = trial->mesh;
  gg_ctx* ctx = m->ctx;
  const int nn = trial->ndofs_loc / 2, mt = test->ndofs_loc;
  const int64_t nc = m->n_cells;
  auto T = get_tab(ctx, GG_SPACE_CG, trial->order, nq);
  auto Tt = get_tab(ctx, test->kind, test->order, nq);
  k_vec_div<<<nblk(nc * mt * nn, 128), 128, 0, ctx->stream>>>(nc, nq, mt, nn, T->xi.p, T->w.p, Tt->val.p, T->val.p, T->der.p, m->d_x0.p,
                                                              m->d_y0.p, m->d_hx.p, m->d_hy.p, m->radius, trial->d_cob.p, ebuf);
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
}

// nodal values of an ambient vector field -> master-frame contravariant components per (cell, local DOF): C_a^-1 pinvJ v
__global__ void k_vec_interp(int64_t nc, int nn, const double* __restrict__ ref, const double* __restrict__ x0,
                             const double* __restrict__ y0, const double* __restrict__ hx, const double* __restrict__ hy,
                             const int32_t* __restrict__ chart, double radius, const double* __restrict__ cob,
                             const double* __restrict__ v, double* __restrict__ vals) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * nn) return;
  const int64_t c = t / nn;
  const int a = (int)(t - c * nn);
  const double ta = tan(fma(hx[c], ref[2 * a], x0[c])), tb = tan(fma(hy[c], ref[2 * a + 1], y0[c]));
  double J[2][3];
  csphere_jacobian(chart[c], ta, tb, radius, J);
  const Metric2 m = csphere_metric_terms(ta, tb, radius);
  const double Jv0 = J[0][0] * v[3 * t] + J[0][1] * v[3 * t + 1] + J[0][2] * v[3 * t + 2];
  const double Jv1 = J[1][0] * v[3 * t] + J[1][1] * v[3 * t + 1] + J[1][2] * v[3 * t + 2];
  const double u0 = m.i11 * Jv0 + m.i12 * Jv1, u1 = m.i12 * Jv0 + m.i22 * Jv1;
  const double* C = cob + 4 * t;
  const double det = C[0] * C[3] - C[1] * C[2];
  vals[c * 2 * nn + a] = (C[3] * u0 - C[1] * u1) / det;
  vals[c * 2 * nn + a + nn] = (-C[2] * u0 + C[0] * u1) / det;
}

void vec_interp_values(gg_space* s, const double* ref_dev, const double* v_dev, double* vals_dev) {
  gg_mesh* m = s->mesh;
  gg_ctx* ctx = m->ctx;
  const int nn = s->ndofs_loc / 2;
  const int64_t nc = m->n_cells;
  k_vec_interp<<<nblk(nc * nn, 128), 128, 0, ctx->stream>>>(nc, nn, ref_dev, m->d_x0.p, m->d_y0.p, m->d_hx.p, m->d_hy.p, m->d_chart.p,
                                                           m->radius, s->d_cob.p, v_dev, vals_dev);
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
}

// per-cell int (u_h - f).g.(u_h - f) sqrtg, u_h = sum_a N_a C_a U_a, f contravariant at the quadrature points (may be NULL)
__global__ void k_vec_error_cells(int64_t nc, int nq, int nn, const double* __restrict__ xi, const double* __restrict__ w,
                                  const double* __restrict__ val, const double* __restrict__ x0, const double* __restrict__ y0,
                                  const double* __restrict__ hx, const double* __restrict__ hy, double radius,
                                  const double* __restrict__ cob, const int32_t* __restrict__ dofs, const double* __restrict__ U,
                                  const double* __restrict__ fq, const uint8_t* __restrict__ cell_owned, double* __restrict__ sums) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  if (cell_owned && !cell_owned[c]) { sums[c] = 0.0; return; }   // partitioned model: ghost cells are counted by their owner
  const double HX = hx[c], HY = hy[c];
  const int Q = nq * nq;
  double acc = 0.0;
  for (int q2 = 0; q2 < nq; ++q2) {
    const double tb = tan(fma(HY, xi[q2], y0[c]));
    for (int q1 = 0; q1 < nq; ++q1) {
      const int q = q1 + nq * q2;
      const Metric2 mt = csphere_metric_terms(tan(fma(HX, xi[q1], x0[c])), tb, radius);
      double e0 = 0.0, e1 = 0.0;
      for (int a = 0; a < nn; ++a) {
        const double* C = cob + 4 * (c * nn + a);
        const double U0 = U[dofs[c * 2 * nn + a]], U1 = U[dofs[c * 2 * nn + a + nn]];
        const double N = val[q * nn + a];
        e0 = fma(N, C[0] * U0 + C[1] * U1, e0);
        e1 = fma(N, C[2] * U0 + C[3] * U1, e1);
      }
      if (fq) { e0 -= fq[(c * Q + q) * 2]; e1 -= fq[(c * Q + q) * 2 + 1]; }
      acc = fma(w[q1] * w[q2] * HX * HY * mt.sqrtg, e0 * (mt.g11 * e0 + mt.g12 * e1) + e1 * (mt.g12 * e0 + mt.g22 * e1), acc);
    }
  }
  sums[c] = acc;
}

double vec_error_sq(gg_space* s, const double* U, const double* fq, int nq) {
  gg_mesh* m = s->mesh;
  gg_ctx* ctx = m->ctx;
  cudaStream_t st = ctx->stream;
  const int nn = s->ndofs_loc / 2;
  const int64_t nc = m->n_cells;
  if (nc == 0) return 0.0;
  auto T = get_tab(ctx, GG_SPACE_CG, s->order, nq);
  DevBuf<double> sums(nc), total(1);
  k_vec_error_cells<<<nblk(nc, 128), 128, 0, st>>>(nc, nq, nn, T->xi.p, T->w.p, T->val.p, m->d_x0.p, m->d_y0.p, m->d_hx.p, m->d_hy.p,
                                                   m->radius, s->d_cob.p, s->d_cell_dofs.p, U, fq, m->d_cell_owned.n ? m->d_cell_owned.p : nullptr, sums.p);
  size_t tb = 0;
  GG_CUDA(cub::DeviceReduce::Sum(nullptr, tb, sums.p, total.p, (int)nc, st));
  DevBuf<uint8_t> tmp(tb);
  GG_CUDA(cub::DeviceReduce::Sum(tmp.p, tb, sums.p, total.p, (int)nc, st));
  ctx->launches += 3;
  GG_CUDA(cudaGetLastError());
  double out = 0.0;
  total.download(&out, st);
  return out;
}

// per-cell int grad(e) : (ginv . grad(e)) sqrtg, grad(e)[k][i] = d_k e^i of the FE function e = sum_a N_a C_a E_a
// (the gradient part of eu_h1, test/SurfaceStokes/SurfaceStokes_TaylorHood.jl:122)
__global__ void k_vec_h1_cells(int64_t nc, int nq, int nn, const double* __restrict__ xi, const double* __restrict__ w,
                               const double* __restrict__ der, const double* __restrict__ x0, const double* __restrict__ y0,
                               const double* __restrict__ hx, const double* __restrict__ hy, double radius,
                               const double* __restrict__ cob, const int32_t* __restrict__ dofs, const double* __restrict__ U,
                               const uint8_t* __restrict__ cell_owned, double* __restrict__ sums) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  if (cell_owned && !cell_owned[c]) { sums[c] = 0.0; return; }
  const double HX = hx[c], HY = hy[c];
  double acc = 0.0;
  for (int q2 = 0; q2 < nq; ++q2) {
    const double tb = tan(fma(HY, xi[q2], y0[c]));
    for (int q1 = 0; q1 < nq; ++q1) {
      const int q = q1 + nq * q2;
      const Metric2 mt = csphere_metric_terms(tan(fma(HX, xi[q1], x0[c])), tb, radius);
      double G[2][2] = {{0.0, 0.0}, {0.0, 0.0}};   // G[k][i] = d_k e^i
      for (int a = 0; a < nn; ++a) {
        const double* C = cob + 4 * (c * nn + a);
        const double U0 = U[dofs[c * 2 * nn + a]], U1 = U[dofs[c * 2 * nn + a + nn]];
        const double e0 = C[0] * U0 + C[1] * U1, e1 = C[2] * U0 + C[3] * U1;
        const double dx = der[(q * nn + a) * 2] / HX, dy = der[(q * nn + a) * 2 + 1] / HY;
        G[0][0] = fma(dx, e0, G[0][0]); G[0][1] = fma(dx, e1, G[0][1]);
        G[1][0] = fma(dy, e0, G[1][0]); G[1][1] = fma(dy, e1, G[1][1]);
      }
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < 2; ++i)
        s += G[0][i] * (mt.i11 * G[0][i] + mt.i12 * G[1][i]) + G[1][i] * (mt.i12 * G[0][i] + mt.i22 * G[1][i]);
      acc = fma(w[q1] * w[q2] * HX * HY * mt.sqrtg, s, acc);
    }
  }
  sums[c] = acc;
}

double vec_h1_seminorm_sq(gg_space* s, const double* U, int nq) {
  gg_mesh* m = s->mesh;
  gg_ctx* ctx = m->ctx;
  cudaStream_t st = ctx->stream;
  const int nn = s->ndofs_loc / 2;
  const int64_t nc = m->n_cells;
  if (nc == 0) return 0.0;
  auto T = get_tab(ctx, GG_SPACE_CG, s->order, nq);
  DevBuf<double> sums(nc), total(1);
  k_vec_h1_cells<<<nblk(nc, 128), 128, 0, st>>>(nc, nq, nn, T->xi.p, T->w.p, T->der.p, m->d_x0.p, m->d_y0.p, m->d_hx.p, m->d_hy.p,
                                                m->radius, s->d_cob.p, s->d_cell_dofs.p, U, m->d_cell_owned.n ? m->d_cell_owned.p : nullptr, sums.p);
  size_t tb = 0;
  GG_CUDA(cub::DeviceReduce::Sum(nullptr, tb, sums.p, total.p, (int)nc, st));
  DevBuf<uint8_t> tmp(tb);
  GG_CUDA(cub::DeviceReduce::Sum(tmp.p, tb, sums.p, total.p, (int)nc, st));
  ctx->launches += 3;
  GG_CUDA(cudaGetLastError());
  double out = 0.0;
  total.download(&out, st);
  return out;
}

}  // namespace gg

using namespace gg;

extern "C" {

int gg_space_change_of_basis(gg_space* s, double* out) {
  GG_API_BEGIN
  GG_REQUIRE(s && out, GG_ERR_INVALID, "null argument");
  GG_REQUIRE(s->kind == GG_SPACE_CGV, GG_ERR_INVALID, "only vector H1 spaces carry a change of basis");
  GG_CUDA(cudaSetDevice(s->mesh->ctx->device));
  if (s->d_cob.n) s->d_cob.download(out, s->mesh->ctx->stream);
  GG_API_END
}

int gg_h1_seminorm_sq(gg_space* s, gg_vec* e, int32_t degree, double* out) {
  GG_API_BEGIN
  GG_REQUIRE(s && e && out, GG_ERR_INVALID, "null argument");
  GG_REQUIRE(s->kind == GG_SPACE_CGV && s->mesh->dim == 2, GG_ERR_UNSUPPORTED, "the H1 seminorm is implemented for vector H1 spaces on the 2-D sphere");
  GG_REQUIRE(gg_vec_size(e) == s->n_free, GG_ERR_INVALID, "vector has the wrong length");
  const int nq = num_points_1d(degree);
  GG_REQUIRE(degree >= 0 && nq <= MAX_NQ_NORMS, GG_ERR_UNSUPPORTED, "quadrature degree out of range");
  GG_CUDA(cudaSetDevice(s->mesh->ctx->device));
  *out = vec_h1_seminorm_sq(s, e->d.p, nq);
  GG_API_END
}

}  // extern "C"
