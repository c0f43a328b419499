// Interpolation of ambient functions into the FE spaces and the norms of the transient drivers' error measures.
//
// Replaces `interpolate(f o ambient_map_cf, P)` (Lagrangian) and
// `interpolate(meas_cf*((pinvJ o transpose o grad(ambient_map_cf)) . (vX o ambient_map_cf)), U)` (Raviart-Thomas) of
// test/Transient/TransientShallowWater.jl:88-95 (pinvJ: src/Helpers/Operators.jl:16-24), and the error integrals
// sqrt(sum(int e.(g e)/sqrtg)) / sqrt(sum(int e e sqrtg)) of test/Transient/TransientWaveEquation.jl:112-117.
// The caller only evaluates its own ambient function at the points handed out by gg_space_interp_points(); geometry,
// pull-back, inverse Piola map and moments run on the device.
#include <cub/cub.cuh>
#include "gg_common.cuh"
#include "gg_refel.cuh"
#include "gg_geom.cuh"
#include "gg_vec.cuh"
#include "gg_hex_ops.cuh"
#include "gg_rk.cuh"

namespace gg {

static inline unsigned nblk(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// reference interpolation points of the space (host)
static std::vector<double> interp_ref_points(const gg_space* s, std::vector<double>* W) {
  std::vector<double> pts;
  if (s->kind == GG_SPACE_RT) {
    std::vector<double> w;
    rt_moment_functionals(s->order, pts, w);
    if (W) *W = std::move(w);
  } else {
    // Lagrangian nodes (the vector H1 space has one point per node, both components)
    int p = s->order, nb = p + 1;
    std::vector<int32_t> l2x = quad_local_to_lex(p);
    pts.resize((size_t)nb * nb * 2);
    for (int a = 0; a < nb * nb; ++a) {
      pts[2 * a] = p == 0 ? 0.5 : (double)(l2x[a] % nb) / p;
      pts[2 * a + 1] = p == 0 ? 0.5 : (double)(l2x[a] / nb) / p;
    }
  }
  return pts;
}

__global__ void k_interp_points(int64_t nc, int N, const double* __restrict__ ref, const double* __restrict__ x0,
                                const double* __restrict__ y0, const double* __restrict__ hx, const double* __restrict__ hy,
                                const int32_t* __restrict__ chart, double radius, double* __restrict__ out_chart,
                                double* __restrict__ out_xyz) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * N) return;
  int64_t c = t / N;
  int n = (int)(t - c * N);
  double x = fma(hx[c], ref[2 * n], x0[c]), y = fma(hy[c], ref[2 * n + 1], y0[c]);
  if (out_chart) { out_chart[2 * t] = x; out_chart[2 * t + 1] = y; }
  if (out_xyz) csphere_point(chart[c], tan(x), tan(y), radius, &out_xyz[3 * t]);
}

// RT: reference-space field uhat = det(Jc) Jc^-1 u at every moment point, u = sqrtg (J J^T)^-1 J v  (sqrtg * pinvJ . v)
__global__ void k_rt_pullback(int64_t nc, int N, const double* __restrict__ ref, const double* __restrict__ x0,
                              const double* __restrict__ y0, const double* __restrict__ hx, const double* __restrict__ hy,
                              const int32_t* __restrict__ chart, double radius, const double* __restrict__ v,
                              double* __restrict__ uhat) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * N) return;
  int64_t c = t / N;
  int n = (int)(t - c * N);
  double a = tan(fma(hx[c], ref[2 * n], x0[c])), b = tan(fma(hy[c], ref[2 * n + 1], y0[c]));
  double J[2][3];
  csphere_jacobian(chart[c], a, b, radius, J);
  Metric2 m = csphere_metric_terms(a, b, radius);
  double Jv0 = J[0][0] * v[3 * t] + J[0][1] * v[3 * t + 1] + J[0][2] * v[3 * t + 2];
  double Jv1 = J[1][0] * v[3 * t] + J[1][1] * v[3 * t + 1] + J[1][2] * v[3 * t + 2];
  double ux = m.sqrtg * (m.i11 * Jv0 + m.i12 * Jv1), uy = m.sqrtg * (m.i12 * Jv0 + m.i22 * Jv1);
  uhat[2 * t] = hy[c] * ux;
  uhat[2 * t + 1] = hx[c] * uy;
}

// one thread per free DOF: take the value from the first (cell, local) pair that touches it
__global__ void k_interp_pick_scalar(int64_t n_free, int64_t nc, int m, const int32_t* __restrict__ vptr,
                                     const int32_t* __restrict__ vsrc, const double* __restrict__ vals,
                                     double* __restrict__ out) {
  int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n_free) return;
  if (vptr[d + 1] == vptr[d]) { out[d] = 0.0; return; }
  int32_t src = vsrc[vptr[d]];
  int i = (int)(src / nc);
  int64_t c = src - (int64_t)i * nc;
  out[d] = vals[c * m + i];
}

__global__ void k_interp_pick_rt(int64_t n_free, int64_t nc, int m, int N, const int32_t* __restrict__ vptr,
                                 const int32_t* __restrict__ vsrc, const int8_t* __restrict__ signs,
                                 const double* __restrict__ W, const double* __restrict__ uhat, double* __restrict__ out) {
  int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n_free) return;
  if (vptr[d + 1] == vptr[d]) { out[d] = 0.0; return; }
  int32_t src = vsrc[vptr[d]];
  int i = (int)(src / nc);
  int64_t c = src - (int64_t)i * nc;
  const double* w = W + (size_t)i * N * 2;
  const double* uh = uhat + (size_t)c * N * 2;
  double acc = 0.0;
  for (int n = 0; n < 2 * N; ++n) acc = fma(w[n], uh[n], acc);
  out[d] = (double)signs[c * m + i] * acc;
}

// per-cell contribution to the squared norm
__global__ void k_norm_cells(int64_t nc, int nq, int m, int is_rt, const double* __restrict__ x0, const double* __restrict__ y0,
                             const double* __restrict__ hx, const double* __restrict__ hy, double radius,
                             const double* __restrict__ xi, const double* __restrict__ w, const double* __restrict__ val,
                             const int32_t* __restrict__ dofs, const int8_t* __restrict__ signs, const double* __restrict__ e,
                             const double* __restrict__ fq, const uint8_t* __restrict__ cell_owned, double* __restrict__ sums) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  if (cell_owned && !cell_owned[c]) { sums[c] = 0.0; return; }  // partitioned model: ghost cells are counted by their owner
  const double HX = hx[c], HY = hy[c];
  double acc = 0.0;
  for (int q2 = 0; q2 < nq; ++q2) {
    const double b = tan(fma(HY, xi[q2], y0[c]));
    for (int q1 = 0; q1 < nq; ++q1) {
      const int q = q1 + nq * q2;
      Metric2 mt = csphere_metric_terms(tan(fma(HX, xi[q1], x0[c])), b, radius);
      const double wq = w[q1] * w[q2] * HX * HY;
      if (is_rt) {
        double Ux = 0.0, Uy = 0.0;
        for (int j = 0; j < m; ++j) {
          int32_t d = dofs[c * m + j];
          double ej = d >= 0 ? (double)signs[c * m + j] * e[d] : 0.0;
          Ux = fma(val[(q * m + j) * 2], ej, Ux);
          Uy = fma(val[(q * m + j) * 2 + 1], ej, Uy);
        }
        // contravariant components u_h / sqrtg, minus the exact contravariant field when one is given
        double ux = Ux / (HY * mt.sqrtg), uy = Uy / (HX * mt.sqrtg);
        if (fq) { ux -= fq[(c * nq * nq + q) * 2]; uy -= fq[(c * nq * nq + q) * 2 + 1]; }
        acc = fma(wq * mt.sqrtg, ux * (mt.g11 * ux + mt.g12 * uy) + uy * (mt.g12 * ux + mt.g22 * uy), acc);
      } else {
        double eh = 0.0;
        for (int j = 0; j < m; ++j) {
          int32_t d = dofs[c * m + j];
          eh = fma(val[q * m + j], d >= 0 ? e[d] : 0.0, eh);
        }
        acc = fma(wq * mt.sqrtg, eh * eh, acc);
      }
    }
  }
  sums[c] = acc;
}

// Closed-form Laplace-Beltrami operator of an ambient function at the quadrature points (src/CellData/SurfaceDiffOps.jl:26-65):
//   Delta_s f = 1/sqrtg [ grad(sqrtg).(ginv grad_f) + sqrtg div(ginv).grad_f + sqrtg ginv : hess_f ],
//   grad_f = J grad_X f,   hess_f = J (hess_X f) J^T + grad_X f . d2phi,   grad(sqrtg)_k = cof(g) : d_k g / (2 sqrtg)
__global__ void k_surface_laplacian(int64_t nc, int nq, const double* __restrict__ xi, const double* __restrict__ x0,
                                    const double* __restrict__ y0, const double* __restrict__ hx, const double* __restrict__ hy,
                                    const int32_t* __restrict__ chart, double radius, const double* __restrict__ gf,
                                    const double* __restrict__ hf, double* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int Q = nq * nq;
  if (t >= nc * Q) return;
  const int64_t c = t / Q;
  const int q = (int)(t - c * Q), q1 = q % nq, q2 = q / nq;
  const double a = tan(fma(hx[c], xi[q1], x0[c])), b = tan(fma(hy[c], xi[q2], y0[c]));
  const int panel = chart[c];
  double J[2][3], H[2][2][3], dg[2][3], dgi[2][3];
  csphere_jacobian(panel, a, b, radius, J);
  csphere_hessian(panel, a, b, radius, H);
  csphere_metric_gradients(a, b, radius, dg, dgi);
  const Metric2 m = csphere_metric_terms(a, b, radius);
  const double* g1 = gf + 3 * t;
  const double* h2 = hf + 9 * t;
  double gradf[2], hess[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i) gradf[i] = J[i][0] * g1[0] + J[i][1] * g1[1] + J[i][2] * g1[2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int l = 0; l < 3; ++l) v = fma(J[i][k] * h2[3 * k + l], J[j][l], v);
        v = fma(g1[k], H[i][j][k], v);
      }
      hess[i][j] = v;
    }
  const double gi[2][2] = {{m.i11, m.i12}, {m.i12, m.i22}};
  // cof(g) : d_k g = g22 d g11 - 2 g12 d g12 + g11 d g22
  double gm[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) gm[k] = 0.5 / m.sqrtg * (m.g22 * dg[k][0] - 2.0 * m.g12 * dg[k][1] + m.g11 * dg[k][2]);
  // (div ginv)_j = d_i ginv_ij
  const double dv[2] = {dgi[0][0] + dgi[1][1], dgi[0][1] + dgi[1][2]};
  double t1 = 0.0, t3 = 0.0;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) { t1 = fma(gm[i] * gi[i][j], gradf[j], t1); t3 = fma(gi[i][j], hess[i][j], t3); }
  const double t2 = dv[0] * gradf[0] + dv[1] * gradf[1];
  out[t] = t1 / m.sqrtg + t2 + t3;
}

// per-cell functionals of u_h.  MODE 0: int (u_h - f)^2 sqrtg with f given at the quadrature points (f may be NULL);
// MODE 1: int u_h in the CHART measure (no sqrtg); MODE 2: chart area of the cell
template <int MODE>
__global__ void k_l2_error_cells(int64_t nc, int nq, int m, const double* __restrict__ x0, const double* __restrict__ y0,
                                 const double* __restrict__ hx, const double* __restrict__ hy, double radius,
                                 const double* __restrict__ xi, const double* __restrict__ w, const double* __restrict__ val,
                                 const int32_t* __restrict__ dofs, const double* __restrict__ u, double dirichlet,
                                 const double* __restrict__ fq, const uint8_t* __restrict__ cell_owned, double* __restrict__ sums) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  if (cell_owned && !cell_owned[c]) { sums[c] = 0.0; return; }
  const double HX = hx[c], HY = hy[c];
  const int Q = nq * nq;
  double acc = 0.0;
  for (int q2 = 0; q2 < nq; ++q2) {
    const double b = tan(fma(HY, xi[q2], y0[c]));
    for (int q1 = 0; q1 < nq; ++q1) {
      const int q = q1 + nq * q2;
      const Metric2 mt = csphere_metric_terms(tan(fma(HX, xi[q1], x0[c])), b, radius);
      double uh = 0.0;
      for (int j = 0; j < m; ++j) {
        const int32_t d = dofs[c * m + j];
        uh = fma(val[q * m + j], d >= 0 ? u[d] : dirichlet, uh);
      }
      if (MODE == 0) {
        const double e = uh - (fq ? fq[c * Q + q] : 0.0);
        acc = fma(w[q1] * w[q2] * HX * HY * mt.sqrtg, e * e, acc);
      } else {
        acc = fma(w[q1] * w[q2] * HX * HY, MODE == 1 ? uh : 1.0, acc);
      }
    }
  }
  sums[c] = acc;
}

// MODE 0: contravariant surface gradient ginv (J grad_X f) -> out[t][2]           (src/CellData/SurfaceDiffOps.jl:107-115)
// MODE 1: surface divergence 1/sqrtg div(sqrtg ginv J F) of an ambient field -> out[t]   (:228-259)
template <int MODE>
__global__ void k_surface_grad_div(int64_t nc, int nq, const double* __restrict__ xi, const double* __restrict__ x0,
                                   const double* __restrict__ y0, const double* __restrict__ hx, const double* __restrict__ hy,
                                   const int32_t* __restrict__ chart, double radius, const double* __restrict__ F,
                                   const double* __restrict__ dF, double* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int Q = nq * nq;
  if (t >= nc * Q) return;
  const int64_t c = t / Q;
  const int q = (int)(t - c * Q), q1 = q % nq, q2 = q / nq;
  const double a = tan(fma(hx[c], xi[q1], x0[c])), b = tan(fma(hy[c], xi[q2], y0[c]));
  const int panel = chart[c];
  double J[2][3];
  csphere_jacobian(panel, a, b, radius, J);
  const Metric2 m = csphere_metric_terms(a, b, radius);
  const double gi[2][2] = {{m.i11, m.i12}, {m.i12, m.i22}};
  const double* f = F + 3 * t;
  double JF[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) JF[i] = J[i][0] * f[0] + J[i][1] * f[1] + J[i][2] * f[2];
  if (MODE == 0) {
    out[2 * t] = gi[0][0] * JF[0] + gi[0][1] * JF[1];
    out[2 * t + 1] = gi[1][0] * JF[0] + gi[1][1] * JF[1];
    return;
  }
  double H[2][2][3], dg[2][3], dgi[2][3];
  csphere_hessian(panel, a, b, radius, H);
  csphere_metric_gradients(a, b, radius, dg, dgi);
  const double* d = dF + 9 * t;
  double gm[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) gm[k] = 0.5 / m.sqrtg * (m.g22 * dg[k][0] - 2.0 * m.g12 * dg[k][1] + m.g11 * dg[k][2]);
  const double dv[2] = {dgi[0][0] + dgi[1][1], dgi[0][1] + dgi[1][2]};
  double acc = dv[0] * JF[0] + dv[1] * JF[1];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        v = fma(H[i][j][k], f[k], v);                                                   // (d_i J_jk) F_k
        v = fma(J[j][k], J[i][0] * d[k] + J[i][1] * d[3 + k] + J[i][2] * d[6 + k], v);  // J_jk J_il dF_lk
      }
      acc = fma(gi[i][j], v, acc);
      acc = fma(gm[i] * gi[i][j], JF[j] / m.sqrtg, acc);
    }
  out[t] = acc;
}

}  // namespace gg

using namespace gg;

extern "C" {

int gg_space_interp_points(gg_space* s, int32_t* n_pts, double* chart, double* xyz) {
  GG_API_BEGIN
  GG_REQUIRE(s, GG_ERR_INVALID, "null space");
  gg_mesh* m = s->mesh;
  gg_ctx* ctx = m->ctx;
  GG_CUDA(cudaSetDevice(ctx->device));
  if (m->dim == 3) {
    if (s->kind == GG_SPACE_RT) {
      // moment points of the hexahedral Raviart-Thomas element (gg_bq.cu)
      const int N = bq_interp_points(s, chart, xyz);
      if (n_pts) *n_pts = N;
      return GG_OK;
    }
    // Lagrange nodes of the hexahedra in the LOCAL DOF order of the space
    GG_REQUIRE(s->kind == GG_SPACE_CG || s->kind == GG_SPACE_DG, GG_ERR_UNSUPPORTED, "interpolation on the shell needs a Lagrangian or RT space");
    const int ml = s->ndofs_loc;
    if (n_pts) *n_pts = ml;
    const int64_t tot = m->n_cells * ml;
    if ((!chart && !xyz) || tot == 0) return GG_OK;
    std::vector<double> cl(chart ? 3 * tot : 0), xl(xyz ? 3 * tot : 0);
    hex_points(m, s->order, 0, chart ? cl.data() : nullptr, xyz ? xl.data() : nullptr);
    for (int64_t c = 0; c < m->n_cells; ++c)
      for (int a = 0; a < ml; ++a)
        for (int d = 0; d < 3; ++d) {
          const size_t src = ((size_t)c * ml + s->loc2lex[a]) * 3 + d, dst = ((size_t)c * ml + a) * 3 + d;
          if (chart) chart[dst] = cl[src];
          if (xyz) xyz[dst] = xl[src];
        }
    return GG_OK;
  }
  std::vector<double> ref = interp_ref_points(s, nullptr);
  int N = (int)(ref.size() / 2);
  if (n_pts) *n_pts = N;
  int64_t tot = m->n_cells * N;
  if ((!chart && !xyz) || tot == 0) return GG_OK;
  cudaStream_t st = ctx->stream;
  DevBuf<double> dref; dref.upload(ref, st);
  DevBuf<double> dc(chart ? 2 * tot : 0), dx(xyz ? 3 * tot : 0);
  k_interp_points<<<nblk(tot, 256), 256, 0, st>>>(m->n_cells, N, dref.p, m->d_x0.p, m->d_y0.p, m->d_hx.p, m->d_hy.p,
                                                  m->d_chart.p, m->radius, dc.p, dx.p);
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
  if (chart) dc.download(chart, st);
  if (xyz) dx.download(xyz, st);
  GG_CUDA(cudaStreamSynchronize(st));
  GG_API_END
}

int gg_space_interpolate(gg_space* s, const double* vals, gg_vec* out) {
  GG_API_BEGIN
  GG_REQUIRE(s && vals && out, GG_ERR_INVALID, "null argument");
  GG_REQUIRE(gg_vec_size(out) == s->n_free, GG_ERR_INVALID, "output vector has the wrong length");
  gg_mesh* m = s->mesh;
  gg_ctx* ctx = m->ctx;
  GG_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  int64_t nc = m->n_cells;
  if (s->n_free == 0) return GG_OK;
  build_vec_map(s);
  if (m->dim == 3 && s->kind == GG_SPACE_RT) {
    GG_CUDA(cudaMemsetAsync(out->d.p, 0, s->n_free * sizeof(double), st));
    bq_interpolate(s, vals, out->d.p);
    return GG_OK;
  }
  if (m->dim == 3) {
    GG_REQUIRE(s->kind == GG_SPACE_CG || s->kind == GG_SPACE_DG, GG_ERR_UNSUPPORTED, "interpolation on the shell needs a Lagrangian or RT space");
    DevBuf<double> dv; dv.upload(vals, (size_t)nc * s->ndofs_loc, st);
    k_interp_pick_scalar<<<nblk(s->n_free, 256), 256, 0, st>>>(s->n_free, nc, s->ndofs_loc, s->d_vptr.p, s->d_vsrc.p, dv.p, out->d.p);
    ctx->launches++;
    GG_CUDA(cudaGetLastError());
    GG_CUDA(cudaStreamSynchronize(st));
    return GG_OK;
  }
  std::vector<double> W;
  std::vector<double> ref = interp_ref_points(s, &W);
  int N = (int)(ref.size() / 2);
  int ml = s->ndofs_loc;
  if (s->kind == GG_SPACE_RT) {
    DevBuf<double> dref, dW, dv, duh((size_t)nc * N * 2);
    dref.upload(ref, st); dW.upload(W, st); dv.upload(vals, (size_t)nc * N * 3, st);
    if (nc) k_rt_pullback<<<nblk(nc * N, 128), 128, 0, st>>>(nc, N, dref.p, m->d_x0.p, m->d_y0.p, m->d_hx.p, m->d_hy.p,
                                                             m->d_chart.p, m->radius, dv.p, duh.p);
    k_interp_pick_rt<<<nblk(s->n_free, 128), 128, 0, st>>>(s->n_free, nc, ml, N, s->d_vptr.p, s->d_vsrc.p, s->d_cell_signs.p,
                                                           dW.p, duh.p, out->d.p);
    ctx->launches += 2;
    GG_CUDA(cudaGetLastError());
    GG_CUDA(cudaStreamSynchronize(st));
  } else if (s->kind == GG_SPACE_CGV) {
    // ambient 3-vectors at the nodes -> contravariant components in the cell's chart -> master frame (C^-1) -> DOFs
    DevBuf<double> dref, dv, dl((size_t)nc * ml);
    dref.upload(ref, st); dv.upload(vals, (size_t)nc * N * 3, st);
    if (nc) vec_interp_values(s, dref.p, dv.p, dl.p);
    k_interp_pick_scalar<<<nblk(s->n_free, 256), 256, 0, st>>>(s->n_free, nc, ml, s->d_vptr.p, s->d_vsrc.p, dl.p, out->d.p);
    ctx->launches++;
    GG_CUDA(cudaGetLastError());
    GG_CUDA(cudaStreamSynchronize(st));
  } else {
    DevBuf<double> dv; dv.upload(vals, (size_t)nc * ml, st);
    k_interp_pick_scalar<<<nblk(s->n_free, 256), 256, 0, st>>>(s->n_free, nc, ml, s->d_vptr.p, s->d_vsrc.p, dv.p, out->d.p);
    ctx->launches++;
    GG_CUDA(cudaGetLastError());
    GG_CUDA(cudaStreamSynchronize(st));
  }
  GG_API_END
}

int gg_norm_sq(gg_space* s, gg_vec* e, int32_t degree, double* out) {
  GG_API_BEGIN
  GG_REQUIRE(s && e && out, GG_ERR_INVALID, "null argument");
  GG_REQUIRE(gg_vec_size(e) == s->n_free, GG_ERR_INVALID, "vector has the wrong length");
  gg_mesh* m = s->mesh;
  gg_ctx* ctx = m->ctx;
  GG_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  int nq = num_points_1d(degree);
  GG_REQUIRE(degree >= 0 && nq <= (m->dim == 3 ? MAX_NQ : MAX_NQ_NORMS), GG_ERR_UNSUPPORTED, "quadrature degree out of range");
  if (m->dim == 3) {
    GG_REQUIRE(s->kind == GG_SPACE_CG, GG_ERR_UNSUPPORTED, "norms on the shell are implemented for CG spaces");
    *out = hex_cell_functional(s, 0, nq, e->d.p, 0.0, nullptr);
    return GG_OK;
  }
  int64_t nc = m->n_cells;
  *out = 0.0;
  if (nc == 0) return GG_OK;
  if (s->kind == GG_SPACE_CGV) { *out = vec_error_sq(s, e->d.p, nullptr, nq); return GG_OK; }
  auto T = get_tab(ctx, s->kind, s->order, nq);
  DevBuf<double> sums(nc), total(1);
  int is_rt = s->kind == GG_SPACE_RT ? 1 : 0;
  k_norm_cells<<<nblk(nc, 128), 128, 0, st>>>(nc, nq, s->ndofs_loc, is_rt, m->d_x0.p, m->d_y0.p, m->d_hx.p, m->d_hy.p, m->radius,
                                              T->xi.p, T->w.p, T->val.p, s->d_cell_dofs.p, is_rt ? s->d_cell_signs.p : nullptr,
                                              e->d.p, nullptr, m->d_cell_owned.n ? m->d_cell_owned.p : nullptr, sums.p);
  size_t tb = 0;
  GG_CUDA(cub::DeviceReduce::Sum(nullptr, tb, sums.p, total.p, (int)nc, st));
  DevBuf<uint8_t> tmp(tb);
  GG_CUDA(cub::DeviceReduce::Sum(tmp.p, tb, sums.p, total.p, (int)nc, st));
  ctx->launches += 3;
  GG_CUDA(cudaGetLastError());
  total.download(out, st);
  GG_API_END
}

int gg_surface_laplacian(gg_mesh* m, int32_t degree, const double* grad_f, const double* hess_f, gg_vec* out) {
  GG_API_BEGIN
  GG_REQUIRE(m && grad_f && hess_f && out, GG_ERR_INVALID, "null argument");
  GG_REQUIRE(m->dim == 2, GG_ERR_UNSUPPORTED, "the surface Laplacian of an ambient function is implemented on the 2-D surface");
  gg_ctx* ctx = m->ctx;
  GG_CUDA(cudaSetDevice(ctx->device));
  const int nq = num_points_1d(degree);
  GG_REQUIRE(degree >= 0 && nq <= (m->dim == 3 ? MAX_NQ : MAX_NQ_NORMS), GG_ERR_UNSUPPORTED, "quadrature degree out of range");
  const int64_t n = m->n_cells * nq * nq;
  GG_REQUIRE(gg_vec_size(out) == n, GG_ERR_INVALID, "output vector must have n_cells * Q entries");
  if (n == 0) return GG_OK;
  cudaStream_t st = ctx->stream;
  auto T = get_tab(ctx, GG_SPACE_DG, 0, nq);
  DevBuf<double> dg, dh;
  dg.upload(grad_f, (size_t)3 * n, st); dh.upload(hess_f, (size_t)9 * n, st);
  k_surface_laplacian<<<nblk(n, 128), 128, 0, st>>>(m->n_cells, nq, T->xi.p, m->d_x0.p, m->d_y0.p, m->d_hx.p, m->d_hy.p, m->d_chart.p,
                                                   m->radius, dg.p, dh.p, out->d.p);
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
  GG_CUDA(cudaStreamSynchronize(st));
  GG_API_END
}

int gg_l2_error_sq(gg_space* s, gg_vec* u, double dirichlet, gg_vec* fq, int32_t degree, double* out) {
  GG_API_BEGIN
  GG_REQUIRE(s && u && out, GG_ERR_INVALID, "null argument");
  GG_REQUIRE(gg_vec_size(u) == s->n_free, GG_ERR_INVALID, "vector has the wrong length");
  if (s->mesh->dim == 3) {
    GG_REQUIRE(s->kind == GG_SPACE_CG, GG_ERR_UNSUPPORTED, "L2 errors on the shell are implemented for CG spaces");
    const int nq3 = num_points_1d(degree);
    GG_REQUIRE(degree >= 0 && nq3 <= MAX_NQ, GG_ERR_UNSUPPORTED, "quadrature degree out of range");
    if (fq) GG_REQUIRE(gg_vec_size(fq) == s->mesh->n_cells * nq3 * nq3 * nq3, GG_ERR_INVALID, "f must be given at the n_cells * nq^3 quadrature points");
    GG_CUDA(cudaSetDevice(s->mesh->ctx->device));
    *out = hex_cell_functional(s, 0, nq3, u->d.p, dirichlet, fq ? fq->d.p : nullptr);
    return GG_OK;
  }
  gg_mesh* m = s->mesh;
  gg_ctx* ctx = m->ctx;
  GG_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int nq = num_points_1d(degree);
  GG_REQUIRE(degree >= 0 && nq <= (m->dim == 3 ? MAX_NQ : MAX_NQ_NORMS), GG_ERR_UNSUPPORTED, "quadrature degree out of range");
  const int64_t nc = m->n_cells;
  const bool is_rt = s->kind == GG_SPACE_RT, is_vec = s->kind == GG_SPACE_CGV;
  if (fq) GG_REQUIRE(gg_vec_size(fq) == nc * nq * nq * (is_rt || is_vec ? 2 : 1), GG_ERR_INVALID,
                     "f must be given at the n_cells * Q quadrature points (two contravariant components for vector-valued spaces)");
  *out = 0.0;
  if (nc == 0) return GG_OK;
  if (is_vec) { *out = vec_error_sq(s, u->d.p, fq ? fq->d.p : nullptr, nq); return GG_OK; }
  auto T = get_tab(ctx, s->kind, s->order, nq);
  DevBuf<double> sums(nc), total(1);
  if (is_rt)
    k_norm_cells<<<nblk(nc, 128), 128, 0, st>>>(nc, nq, s->ndofs_loc, 1, m->d_x0.p, m->d_y0.p, m->d_hx.p, m->d_hy.p, m->radius, T->xi.p,
                                                T->w.p, T->val.p, s->d_cell_dofs.p, s->d_cell_signs.p, u->d.p, fq ? fq->d.p : nullptr,
                                                m->d_cell_owned.n ? m->d_cell_owned.p : nullptr, sums.p);
  else
  k_l2_error_cells<0><<<nblk(nc, 128), 128, 0, st>>>(nc, nq, s->ndofs_loc, m->d_x0.p, m->d_y0.p, m->d_hx.p, m->d_hy.p, m->radius, T->xi.p,
                                                  T->w.p, T->val.p, s->d_cell_dofs.p, u->d.p, dirichlet, fq ? fq->d.p : nullptr,
                                                  m->d_cell_owned.n ? m->d_cell_owned.p : nullptr, sums.p);
  size_t tb = 0;
  GG_CUDA(cub::DeviceReduce::Sum(nullptr, tb, sums.p, total.p, (int)nc, st));
  DevBuf<uint8_t> tmp(tb);
  GG_CUDA(cub::DeviceReduce::Sum(tmp.p, tb, sums.p, total.p, (int)nc, st));
  ctx->launches += 3;
  GG_CUDA(cudaGetLastError());
  total.download(out, st);
  GG_API_END
}

int gg_chart_mean(gg_space* s, gg_vec* u, double dirichlet, double* out) {
  GG_API_BEGIN
  GG_REQUIRE(s && u && out, GG_ERR_INVALID, "null argument");
  GG_REQUIRE(s->kind != GG_SPACE_RT && s->kind != GG_SPACE_CGV, GG_ERR_UNSUPPORTED, "chart means are implemented for scalar Lagrangian spaces");
  GG_REQUIRE(gg_vec_size(u) == s->n_free, GG_ERR_INVALID, "vector has the wrong length");
  if (s->mesh->dim == 3) {
    GG_REQUIRE(s->kind == GG_SPACE_CG, GG_ERR_UNSUPPORTED, "chart means on the shell are implemented for CG spaces");
    GG_CUDA(cudaSetDevice(s->mesh->ctx->device));
    const int nq3 = num_points_1d(2 * s->order);
    *out = hex_cell_functional(s, 1, nq3, u->d.p, dirichlet, nullptr) / hex_cell_functional(s, 2, nq3, u->d.p, dirichlet, nullptr);
    return GG_OK;
  }
  gg_mesh* m = s->mesh;
  gg_ctx* ctx = m->ctx;
  GG_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int nq = num_points_1d(2 * (s->order > 0 ? s->order : 1));
  const int64_t nc = m->n_cells;
  *out = 0.0;
  if (nc == 0) return GG_OK;
  auto T = get_tab(ctx, s->kind, s->order, nq);
  DevBuf<double> sums(2 * nc), total(2);
  const uint8_t* own = m->d_cell_owned.n ? m->d_cell_owned.p : nullptr;
  k_l2_error_cells<1><<<nblk(nc, 128), 128, 0, st>>>(nc, nq, s->ndofs_loc, m->d_x0.p, m->d_y0.p, m->d_hx.p, m->d_hy.p, m->radius, T->xi.p,
                                                     T->w.p, T->val.p, s->d_cell_dofs.p, u->d.p, dirichlet, nullptr, own, sums.p);
  k_l2_error_cells<2><<<nblk(nc, 128), 128, 0, st>>>(nc, nq, s->ndofs_loc, m->d_x0.p, m->d_y0.p, m->d_hx.p, m->d_hy.p, m->radius, T->xi.p,
                                                     T->w.p, T->val.p, s->d_cell_dofs.p, u->d.p, dirichlet, nullptr, own, sums.p + nc);
  size_t tb = 0;
  GG_CUDA(cub::DeviceReduce::Sum(nullptr, tb, sums.p, total.p, (int)nc, st));
  DevBuf<uint8_t> tmp(tb);
  GG_CUDA(cub::DeviceReduce::Sum(tmp.p, tb, sums.p, total.p, (int)nc, st));
  GG_CUDA(cub::DeviceReduce::Sum(tmp.p, tb, sums.p + nc, total.p + 1, (int)nc, st));
  ctx->launches += 6;
  GG_CUDA(cudaGetLastError());
  double h[2];
  total.download(h, st);
  *out = h[0] / h[1];
  GG_API_END
}

}  // extern "C"

static int surface_grad_div(gg_mesh* m, int32_t degree, const double* F, const double* dF, gg_vec* out, int mode) {
  GG_API_BEGIN
  GG_REQUIRE(m && F && out && (mode == 0 || dF), GG_ERR_INVALID, "null argument");
  GG_REQUIRE(m->dim == 2, GG_ERR_UNSUPPORTED, "surface differential operators of ambient functions are implemented on the 2-D surface");
  gg_ctx* ctx = m->ctx;
  GG_CUDA(cudaSetDevice(ctx->device));
  const int nq = num_points_1d(degree);
  GG_REQUIRE(degree >= 0 && nq <= (m->dim == 3 ? MAX_NQ : MAX_NQ_NORMS), GG_ERR_UNSUPPORTED, "quadrature degree out of range");
  const int64_t n = m->n_cells * nq * nq;
  GG_REQUIRE(gg_vec_size(out) == (mode == 0 ? 2 * n : n), GG_ERR_INVALID, "output vector has the wrong length");
  if (n == 0) return GG_OK;
  cudaStream_t st = ctx->stream;
  auto T = get_tab(ctx, GG_SPACE_DG, 0, nq);
  DevBuf<double> dg, dh;
  dg.upload(F, (size_t)3 * n, st);
  if (mode == 1) dh.upload(dF, (size_t)9 * n, st);
  if (mode == 0)
    k_surface_grad_div<0><<<nblk(n, 128), 128, 0, st>>>(m->n_cells, nq, T->xi.p, m->d_x0.p, m->d_y0.p, m->d_hx.p, m->d_hy.p, m->d_chart.p,
                                                        m->radius, dg.p, nullptr, out->d.p);
  else
    k_surface_grad_div<1><<<nblk(n, 128), 128, 0, st>>>(m->n_cells, nq, T->xi.p, m->d_x0.p, m->d_y0.p, m->d_hx.p, m->d_hy.p, m->d_chart.p,
                                                        m->radius, dg.p, dh.p, out->d.p);
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
  GG_CUDA(cudaStreamSynchronize(st));
  GG_API_END
}

extern "C" {

int gg_surface_gradient(gg_mesh* m, int32_t degree, const double* grad_f, gg_vec* out) {
  return surface_grad_div(m, degree, grad_f, nullptr, out, 0);
}

int gg_surface_divergence(gg_mesh* m, int32_t degree, const double* F, const double* grad_F, gg_vec* out) {
  return surface_grad_div(m, degree, F, grad_F, out, 1);
}

}  // extern "C"
