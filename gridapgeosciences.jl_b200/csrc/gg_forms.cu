// Numeric phase: cell-wise quadrature -> element matrices / vectors -> atomic-free gather into CSR / DOF vectors.
//
// Replaces Gridap's collect_cell_matrix / collect_cell_vector + assemble_* on the forms of the reference's
// drivers (src/ODEs/ODEs.jl:33-39; forms: test/Laplacian/LaplaceBeltrami.jl:47,53, test/Laplacian/Helmholtz.jl:43,
// test/AmbientModel/AmbientLaplaceBeltrami.jl:44, test/Transient/TransientWaveEquation.jl:63-66).
// Two kernel families:
//   * gg_lb_fast.cuh: the hot sum-factorised one-thread-per-cell kernels for scalar Q1/Q2 stiffness forms;
//   * the generic kernels below (one thread per (cell, test dof), tabulated bases) for every other supported
//     (space, order, degree, form) combination.  Both are CUDA; nothing falls back to the host.
#include <cub/cub.cuh>
#include <cstring>
#include <cstdlib>
#include "gg_common.cuh"
#include "gg_vec.cuh"
#include "gg_hex_ops.cuh"
#include "gg_refel.cuh"
#include "gg_geom.cuh"
#include "gg_lb_fast.cuh"
#include "gg_lb_tile.cuh"
#include "gg_hex.cuh"

namespace gg {

void gather_matrix(gg_csr* A, int layout, const double* ebuf, double scale, int add);
void gather_vector(gg_space* s, const double* ebuf, double* out, double scale, int add);
void gather_matrix_and_vector(gg_csr* A, int layout, const double* ebuf, const double* vbuf, double* vec_out, double scale, int add);

static inline unsigned nblk(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// ---- device tabulations ---------------------------------------------------------------------------------
std::shared_ptr<FormTab> get_tab(gg_ctx* ctx, int kind, int order, int nq) {
  int k = kind == GG_SPACE_RT ? 1 : 0;
  uint64_t key = ((uint64_t)k << 32) | ((uint64_t)order << 16) | (uint64_t)nq;
  auto it = ctx->tabs.find(key);
  if (it != ctx->tabs.end()) return it->second;
  GG_REQUIRE(nq >= 1 && nq <= MAX_NQ_NORMS, GG_ERR_UNSUPPORTED, "quadrature degree too high (max 63)");
  auto T = std::make_shared<FormTab>();
  T->kind = k; T->order = order; T->nq = nq; T->Q = nq * nq;
  std::vector<double> xi(nq), w(nq);
  gauss_legendre_01(nq, xi.data(), w.data());
  std::vector<double> val, der;
  if (k == 0) {
    int nb = order + 1, m = nb * nb;
    T->m = m;
    std::vector<int32_t> l2x = quad_local_to_lex(order);
    val.resize((size_t)T->Q * m);
    der.resize((size_t)T->Q * m * 2);
    std::vector<double> N(nq * nb), D(nq * nb);
    for (int q = 0; q < nq; ++q) lagrange_1d(order, xi[q], &N[q * nb], &D[q * nb]);
    for (int q2 = 0; q2 < nq; ++q2)
      for (int q1 = 0; q1 < nq; ++q1) {
        int q = q1 + nq * q2;
        for (int a = 0; a < m; ++a) {
          int ix = l2x[a] % nb, iy = l2x[a] / nb;
          val[(size_t)q * m + a] = N[q1 * nb + ix] * N[q2 * nb + iy];
          der[((size_t)q * m + a) * 2 + 0] = D[q1 * nb + ix] * N[q2 * nb + iy];
          der[((size_t)q * m + a) * 2 + 1] = N[q1 * nb + ix] * D[q2 * nb + iy];
        }
      }
  } else {
    RTRef R = build_rt_ref(order);
    T->m = R.ndofs;
    std::vector<double> pts((size_t)T->Q * 2);
    for (int q2 = 0; q2 < nq; ++q2)
      for (int q1 = 0; q1 < nq; ++q1) {
        pts[2 * (q1 + nq * q2)] = xi[q1];
        pts[2 * (q1 + nq * q2) + 1] = xi[q2];
      }
    R.tabulate(T->Q, pts.data(), val, der);
  }
  cudaStream_t st = ctx->stream;
  T->xi.upload(xi, st); T->w.upload(w, st); T->val.upload(val, st); T->der.upload(der, st);
  GG_CUDA(cudaStreamSynchronize(st));
  ctx->tabs[key] = T;
  return T;
}

struct Geo {
  int64_t nc;
  const double *x0, *y0, *hx, *hy;
  const int32_t* chart;
  double radius;
};
static Geo geo_of(gg_mesh* m) { return {m->n_cells, m->d_x0.p, m->d_y0.p, m->d_hx.p, m->d_hy.p, m->d_chart.p, m->radius}; }

// ---- generic kernels: one thread per (cell, test dof) -----------------------------------------------------

// Ke[I][J] = sum_q w hx hy [ cM * qcoef * v_I u_J sqrtg  +  cK * grad v_I . ginv grad u_J sqrtg ]
template <int MU, bool EXTRINSIC>
__global__ void __launch_bounds__(128)
k_scalar_matrix(Geo g, int nq, const double* __restrict__ xi, const double* __restrict__ w, int mt,
                const double* __restrict__ tval, const double* __restrict__ tder, const double* __restrict__ uval,
                const double* __restrict__ uder, double cM, double cK, const double* __restrict__ qcoef,
                double* __restrict__ ebuf) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.nc * mt) return;
  int I = (int)(t / g.nc);
  int64_t c = t - (int64_t)I * g.nc;
  const double X0 = g.x0[c], Y0 = g.y0[c], HX = g.hx[c], HY = g.hy[c];
  const int panel = g.chart[c];
  double acc[MU];
#pragma unroll
  for (int j = 0; j < MU; ++j) acc[j] = 0.0;
  double ta[MAX_NQ];
  for (int q = 0; q < nq; ++q) ta[q] = tan(fma(HX, xi[q], X0));
  const int Q = nq * nq;
  for (int q2 = 0; q2 < nq; ++q2) {
    const double b = tan(fma(HY, xi[q2], Y0));
    for (int q1 = 0; q1 < nq; ++q1) {
      const int q = q1 + nq * q2;
      const double a = ta[q1];
      Metric2 m = EXTRINSIC ? extrinsic_metric_terms(panel, a, b, g.radius) : csphere_metric_terms(a, b, g.radius);
      const double dV = w[q1] * w[q2] * HX * HY * m.sqrtg;
      const double vI = tval[q * mt + I];
      const double gx = tder[(q * mt + I) * 2] / HX, gy = tder[(q * mt + I) * 2 + 1] / HY;
      const double fx = cK * dV * (m.i11 * gx + m.i12 * gy), fy = cK * dV * (m.i12 * gx + m.i22 * gy);
      double fm = cM * dV * vI;
      if (qcoef) fm *= qcoef[c * Q + q];
#pragma unroll
      for (int j = 0; j < MU; ++j) {
        double v = acc[j];
        v = fma(fm, uval[q * MU + j], v);
        v = fma(fx / HX, uder[(q * MU + j) * 2], v);
        v = fma(fy / HY, uder[(q * MU + j) * 2 + 1], v);
        acc[j] = v;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < MU; ++j) ebuf[(int64_t)(I * MU + j) * g.nc + c] = acc[j];
}

// RT mass: Ke[I][J] = sum_q w hx hy (v_I . g u_J) / sqrtg, contravariant Piola u = (uhat_x/hy, uhat_y/hx) * sign
template <int MU>
__global__ void __launch_bounds__(128)
k_rt_mass_matrix(Geo g, int nq, const double* __restrict__ xi, const double* __restrict__ w,
                 const double* __restrict__ val, const int8_t* __restrict__ signs, double* __restrict__ ebuf) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.nc * MU) return;
  int I = (int)(t / g.nc);
  int64_t c = t - (int64_t)I * g.nc;
  const double X0 = g.x0[c], Y0 = g.y0[c], HX = g.hx[c], HY = g.hy[c];
  double acc[MU];
#pragma unroll
  for (int j = 0; j < MU; ++j) acc[j] = 0.0;
  double ta[MAX_NQ];
  for (int q = 0; q < nq; ++q) ta[q] = tan(fma(HX, xi[q], X0));
  const double sI = signs[c * MU + I];
  for (int q2 = 0; q2 < nq; ++q2) {
    const double b = tan(fma(HY, xi[q2], Y0));
    for (int q1 = 0; q1 < nq; ++q1) {
      const int q = q1 + nq * q2;
      Metric2 m = csphere_metric_terms(ta[q1], b, g.radius);
      const double dV = w[q1] * w[q2] * HX * HY / m.sqrtg;
      const double vx = sI * val[(q * MU + I) * 2] / HY, vy = sI * val[(q * MU + I) * 2 + 1] / HX;
      const double fx = dV * (m.g11 * vx + m.g12 * vy) / HY, fy = dV * (m.g12 * vx + m.g22 * vy) / HX;
#pragma unroll
      for (int j = 0; j < MU; ++j)
        acc[j] = fma(fx, val[(q * MU + j) * 2], fma(fy, val[(q * MU + j) * 2 + 1], acc[j]));
    }
  }
#pragma unroll
  for (int j = 0; j < MU; ++j) ebuf[(int64_t)(I * MU + j) * g.nc + c] = acc[j] * (double)signs[c * MU + j];
}

// div: Ke[I][J] = sum_q w q_I divhat(uhat_J) sign_J   (the Jacobian determinant cancels)
template <int MU>
__global__ void __launch_bounds__(128)
k_div_matrix(int64_t nc, int nq, const double* __restrict__ w, int mt, const double* __restrict__ tval,
             const double* __restrict__ udiv, const int8_t* __restrict__ signs, double* __restrict__ ebuf) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * mt) return;
  int I = (int)(t / nc);
  int64_t c = t - (int64_t)I * nc;
  double acc[MU];
#pragma unroll
  for (int j = 0; j < MU; ++j) acc[j] = 0.0;
  for (int q2 = 0; q2 < nq; ++q2)
    for (int q1 = 0; q1 < nq; ++q1) {
      const int q = q1 + nq * q2;
      const double f = w[q1] * w[q2] * tval[q * mt + I];
#pragma unroll
      for (int j = 0; j < MU; ++j) acc[j] = fma(f, udiv[q * MU + j], acc[j]);
    }
#pragma unroll
  for (int j = 0; j < MU; ++j) ebuf[(int64_t)(I * MU + j) * nc + c] = acc[j] * (double)signs[c * MU + j];
}

// fe[I] = sum_q w hx hy sqrtg f_q v_I
__global__ void __launch_bounds__(128)
k_source_vector(Geo g, int nq, const double* __restrict__ xi, const double* __restrict__ w, int mt,
                const double* __restrict__ tval, const double* __restrict__ qdata, double* __restrict__ vbuf) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.nc * mt) return;
  int I = (int)(t / g.nc);
  int64_t c = t - (int64_t)I * g.nc;
  const double X0 = g.x0[c], Y0 = g.y0[c], HX = g.hx[c], HY = g.hy[c];
  double ta[MAX_NQ];
  for (int q = 0; q < nq; ++q) ta[q] = tan(fma(HX, xi[q], X0));
  const int Q = nq * nq;
  double acc = 0.0;
  for (int q2 = 0; q2 < nq; ++q2) {
    const double b = tan(fma(HY, xi[q2], Y0));
    for (int q1 = 0; q1 < nq; ++q1) {
      const int q = q1 + nq * q2;
      Metric2 m = csphere_metric_terms(ta[q1], b, g.radius);
      acc = fma(w[q1] * w[q2] * HX * HY * m.sqrtg * qdata[c * Q + q], tval[q * mt + I], acc);
    }
  }
  vbuf[(int64_t)I * g.nc + c] = acc;
}

// generic action of the scalar forms: fe[I] = sum_q dV [ cM v_I u_h + cK grad v_I . ginv grad u_h ]
template <int MU, bool EXTRINSIC>
__global__ void __launch_bounds__(128)
k_scalar_action(Geo g, int nq, const double* __restrict__ xi, const double* __restrict__ w, int mt,
                const double* __restrict__ tval, const double* __restrict__ tder, const double* __restrict__ uval,
                const double* __restrict__ uder, double cM, double cK, const int32_t* __restrict__ udofs,
                const double* __restrict__ u, double dirichlet, double* __restrict__ vbuf) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.nc * mt) return;
  int I = (int)(t / g.nc);
  int64_t c = t - (int64_t)I * g.nc;
  const double X0 = g.x0[c], Y0 = g.y0[c], HX = g.hx[c], HY = g.hy[c];
  const int panel = g.chart[c];
  double ue[MU];
#pragma unroll
  for (int j = 0; j < MU; ++j) {
    int32_t d = udofs[c * MU + j];
    ue[j] = d >= 0 ? u[d] : dirichlet;
  }
  double ta[MAX_NQ];
  for (int q = 0; q < nq; ++q) ta[q] = tan(fma(HX, xi[q], X0));
  double acc = 0.0;
  for (int q2 = 0; q2 < nq; ++q2) {
    const double b = tan(fma(HY, xi[q2], Y0));
    for (int q1 = 0; q1 < nq; ++q1) {
      const int q = q1 + nq * q2;
      Metric2 m = EXTRINSIC ? extrinsic_metric_terms(panel, ta[q1], b, g.radius) : csphere_metric_terms(ta[q1], b, g.radius);
      double uh = 0.0, ux = 0.0, uy = 0.0;
#pragma unroll
      for (int j = 0; j < MU; ++j) {
        uh = fma(uval[q * MU + j], ue[j], uh);
        ux = fma(uder[(q * MU + j) * 2], ue[j], ux);
        uy = fma(uder[(q * MU + j) * 2 + 1], ue[j], uy);
      }
      ux /= HX; uy /= HY;
      const double dV = w[q1] * w[q2] * HX * HY * m.sqrtg;
      const double gx = tder[(q * mt + I) * 2] / HX, gy = tder[(q * mt + I) * 2 + 1] / HY;
      acc += dV * (cM * tval[q * mt + I] * uh + cK * (gx * (m.i11 * ux + m.i12 * uy) + gy * (m.i12 * ux + m.i22 * uy)));
    }
  }
  vbuf[(int64_t)I * g.nc + c] = acc;
}

__global__ void k_qpoints(Geo g, int nq, const double* __restrict__ xi, double* __restrict__ chart_out, double* __restrict__ xyz_out) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int Q = nq * nq;
  if (t >= g.nc * Q) return;
  int64_t c = t / Q;
  int q = (int)(t - c * Q);
  int q1 = q % nq, q2 = q / nq;
  double x = fma(g.hx[c], xi[q1], g.x0[c]), y = fma(g.hy[c], xi[q2], g.y0[c]);
  if (chart_out) { chart_out[2 * t] = x; chart_out[2 * t + 1] = y; }
  if (xyz_out) csphere_point(g.chart[c], tan(x), tan(y), g.radius, &xyz_out[3 * t]);
}

__global__ void k_integrate_cells(Geo g, int nq, const double* __restrict__ xi, const double* __restrict__ w,
                                  const double* __restrict__ qdata, const uint8_t* __restrict__ cell_owned,
                                  double* __restrict__ cell_sums) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= g.nc) return;
  if (cell_owned && !cell_owned[c]) { cell_sums[c] = 0.0; return; }  // partitioned model: counted by the owner
  const double X0 = g.x0[c], Y0 = g.y0[c], HX = g.hx[c], HY = g.hy[c];
  int Q = nq * nq;
  double acc = 0.0;
  for (int q2 = 0; q2 < nq; ++q2) {
    const double b = tan(fma(HY, xi[q2], Y0));
    for (int q1 = 0; q1 < nq; ++q1) {
      Metric2 m = csphere_metric_terms(tan(fma(HX, xi[q1], X0)), b, g.radius);
      double f = qdata ? qdata[c * Q + q1 + nq * q2] : 1.0;
      acc = fma(w[q1] * w[q2] * HX * HY * m.sqrtg, f, acc);
    }
  }
  cell_sums[c] = acc;
}

__global__ void k_eval_geometry(int panel, double r, int64_t n, const double* __restrict__ pts, double* xyz, double* jac,
                                double* metric, double* inv_metric, double* sqrtg) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  double a = tan(pts[2 * t]), b = tan(pts[2 * t + 1]);
  if (xyz) csphere_point(panel, a, b, r, &xyz[3 * t]);
  if (jac) {
    double J[2][3];
    csphere_jacobian(panel, a, b, r, J);
    for (int i = 0; i < 2; ++i)
      for (int k = 0; k < 3; ++k) jac[6 * t + 3 * i + k] = J[i][k];
  }
  Metric2 m = csphere_metric_terms(a, b, r);
  if (metric) { metric[3 * t] = m.g11; metric[3 * t + 1] = m.g12; metric[3 * t + 2] = m.g22; }
  if (inv_metric) { inv_metric[3 * t] = m.i11; inv_metric[3 * t + 1] = m.i12; inv_metric[3 * t + 2] = m.i22; }
  if (sqrtg) sqrtg[t] = m.sqrtg;
}

// ---- dispatch -------------------------------------------------------------------------------------------

static bool is_scalar(const gg_space* s) { return s->kind == GG_SPACE_CG || s->kind == GG_SPACE_DG; }

static void load_const_tables(gg_ctx* ctx, int order, int nq) {
  if (ctx->const_order == order && ctx->const_nq == nq) return;
  ScalarTables t;
  build_scalar_tables(order, nq, &t);
  int nb = order + 1, M = nb * nb;
  std::vector<int32_t> l2x = quad_local_to_lex(order);
  std::vector<int32_t> x2l(M);
  for (int a = 0; a < M; ++a) x2l[l2x[a]] = a;
  std::vector<int32_t> plane;
  for (int I = 0; I < M; ++I)
    for (int J = I; J < M; ++J) {
      int a = std::min(x2l[I], x2l[J]), b = std::max(x2l[I], x2l[J]);
      plane.push_back(a * M - (a * (a - 1)) / 2 + (b - a));
    }
  // compact pair tables in use order: row q = [NN sym | DD sym | ND], quadrature weight folded in
  int NS = nb * (nb + 1) / 2, ROW = 2 * NS + M;
  std::vector<double> lb((size_t)nq * ROW);
  for (int q = 0; q < nq; ++q) {
    int k = 0;
    for (int i = 0; i < nb; ++i)
      for (int j = i; j < nb; ++j) {
        lb[(size_t)q * ROW + k] = t.NN[i * nb + j][q];
        lb[(size_t)q * ROW + NS + k] = t.DD[i * nb + j][q];
        ++k;
      }
    for (int k2 = 0; k2 < M; ++k2) lb[(size_t)q * ROW + 2 * NS + k2] = t.ND[k2][q];
  }
  // synchronous w.r.t. the stream: earlier launches reading the old tables must have drained
  GG_CUDA(cudaStreamSynchronize(ctx->stream));
  GG_CUDA(cudaMemcpyToSymbol(c_tab, &t, sizeof(t)));
  GG_CUDA(cudaMemcpyToSymbol(c_lb, lb.data(), lb.size() * sizeof(double)));
  GG_CUDA(cudaMemcpyToSymbol(c_plane, plane.data(), plane.size() * sizeof(int32_t)));
  GG_CUDA(cudaMemcpyToSymbol(c_lex2loc, x2l.data(), x2l.size() * sizeof(int32_t)));
  ctx->const_order = order;
  ctx->const_nq = nq;
}

// per-(mesh, nq) tangent tables of the distinct chart intervals, built on the device at first use
CellGeo cell_geo(gg_mesh* m, int nq) {
  gg_ctx* ctx = m->ctx;
  auto it = m->tan_tabs.find(nq);
  if (it == m->tan_tabs.end()) {
    auto T = std::make_shared<gg_mesh::TanTab>();
    auto tab = get_tab(ctx, GG_SPACE_DG, 0, nq);  // 1-D abscissae on the device
    T->tx.alloc((size_t)m->n_xint * nq);
    T->ty.alloc((size_t)m->n_yint * nq);
    if (m->n_xint) k_tan_table<<<nblk(m->n_xint * nq, 128), 128, 0, ctx->stream>>>(m->n_xint, nq, m->d_xlo.p, m->d_xh.p, tab->xi.p, T->tx.p);
    if (m->n_yint) k_tan_table<<<nblk(m->n_yint * nq, 128), 128, 0, ctx->stream>>>(m->n_yint, nq, m->d_ylo.p, m->d_yh.p, tab->xi.p, T->ty.p);
    ctx->launches += 2;
    GG_CUDA(cudaGetLastError());
    it = m->tan_tabs.emplace(nq, T).first;
  }
  CellGeo g;
  g.hx = m->d_hx.p; g.hy = m->d_hy.p; g.ix = m->d_ix.p; g.iy = m->d_iy.p;
  g.tx = it->second->tx.p; g.ty = it->second->ty.p;
  g.chart = m->d_chart.p; g.radius = m->radius;
  return g;
}

template <int NB, int NQ>
static void launch_lb_fast(gg_ctx* ctx, int64_t nc, int64_t c0, int64_t c1, const CellGeo& g, bool extrinsic,
                           const int32_t* dofs, const double* u, double dir, double* ebuf, double* vbuf, bool leave_room) {
  if (c1 <= c0) return;
  unsigned nb = nblk(c1 - c0, 128);
  // leave_room: cap residency at 2 CTAs per SM (an unused 80 KB dynamic shared-memory reservation), so that a third of the
  // register file and 67 KB of shared memory stay free for the concurrently running CSR-fill CTAs of the second stream
  size_t pad = leave_room ? 80 * 1024 : 0;
  static bool attr_set = false;
  if (!attr_set) {
    GG_CUDA(cudaFuncSetAttribute(k_lb_fast<NB, NQ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
    GG_CUDA(cudaFuncSetAttribute(k_lb_fast<NB, NQ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
    attr_set = true;
  }
  ProfScope ps(ctx, "k_lb_fast");
  if (extrinsic)
    k_lb_fast<NB, NQ, true><<<nb, 128, pad, ctx->stream>>>(nc, c0, c1, g, dofs, u, dir, ebuf, vbuf);
  else
    k_lb_fast<NB, NQ, false><<<nb, 128, pad, ctx->stream>>>(nc, c0, c1, g, dofs, u, dir, ebuf, vbuf);
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
}

// geometry-class mode: one element matrix per class (representative cells), residual from the class matrices
template <int NB, int NQ>
static void launch_lb_classes(gg_ctx* ctx, gg_mesh* m, const CellGeo& g, const int32_t* dofs, const double* u, double dir,
                              double* ebuf, double* vbuf) {
  const int64_t ncls = m->n_classes, nc = m->n_cells;
  {
    ProfScope ps(ctx, "k_lb_fast");
    k_lb_fast<NB, NQ, false, true><<<nblk(ncls, 128), 128, 0, ctx->stream>>>(ncls, 0, ncls, g, nullptr, nullptr, 0.0, ebuf, nullptr,
                                                                            m->d_cls_rep.p);
    ctx->launches++;
  }
  if (vbuf) {
    ProfScope ps(ctx, "k_lb_residual_classes");
    k_lb_residual_classes<NB><<<nblk(nc, 128), 128, 0, ctx->stream>>>(nc, ncls, m->d_cls.p, ebuf, dofs, u, dir, vbuf);
    ctx->launches++;
  }
  GG_CUDA(cudaGetLastError());
}

// When set, the CSR fill is pipelined behind the element kernel: the cells are cut into gg_csr::N_BATCH batches; as soon
// as batch b has been computed (event on the main stream) the chunks it completes are gathered on the second stream
// while the FP64-bound element kernel proceeds with batch b+1 (the two kernels use complementary resources).
struct FillPipe {
  gg_csr* A;
  double scale;
  int add;
  double* vec_out = nullptr;   // fused residual goes straight into this DOF vector (tile path); set by gg_assemble_matrix_and_vector
  bool vec_done = false;       // out: the residual has been written (no element-vector gather needed)
};

// fast-path availability: scalar Q1/Q2, test == trial numbering family, the quadrature sizes the reference uses
static bool fast_lb_available(int order, int nq) {
  if (order == 1) return nq == 2 || nq == 3 || nq == 4 || nq == 7;
  if (order == 2) return nq == 3 || nq == 4 || nq == 5 || nq == 7 || nq == 10;
  return false;
}

static void run_lb_fast_range(gg_mesh* m, int order, int nq, bool extrinsic, int64_t c0, int64_t c1, const CellGeo& g,
                              const int32_t* dofs, const double* u, double dir, double* ebuf, double* vbuf,
                              bool leave_room = false) {
  gg_ctx* ctx = m->ctx;
  int64_t nc = m->n_cells;
#define GG_CASE(NB_, NQ_) \
  if (order + 1 == NB_ && nq == NQ_) { launch_lb_fast<NB_, NQ_>(ctx, nc, c0, c1, g, extrinsic, dofs, u, dir, ebuf, vbuf, leave_room); return; }
  GG_CASE(2, 2) GG_CASE(2, 3) GG_CASE(2, 4) GG_CASE(2, 7)
  GG_CASE(3, 3) GG_CASE(3, 4) GG_CASE(3, 5) GG_CASE(3, 7) GG_CASE(3, 10)
#undef GG_CASE
  throw Error(GG_ERR_UNSUPPORTED, "no fast Laplace-Beltrami kernel for this (order, degree)");
}

// fused tile path (gg_tile.cuh, gg_lb_tile.cuh): one kernel computes the element matrices and writes the CSR values (and the
// residual); on by default, GG_LB_TILES=0 selects the two-kernel path (element buffer + k_gather_chunks)
template <int NB, int NQ, bool EXT, bool MULTI>
static void launch_lb_tile_k(gg_ctx* ctx, const TilePlan& plan, const TileDev& T, const TileArgs& a, const CellGeo& g) {
  constexpr int M = NB * NB;
  const size_t smem = tile_smem_bytes(M, a.n_acc, a.n_vacc, a.n_rows);
  static size_t attr_set = 0;
  if (attr_set < smem) {
    GG_CUDA(cudaFuncSetAttribute(k_lb_tile<NB, NQ, EXT, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = smem;
  }
  ProfScope ps(ctx, "k_lb_tile");
  k_lb_tile<NB, NQ, EXT, MULTI><<<(unsigned)plan.n_super, TILE_CELLS, smem, ctx->stream>>>(T, a, g);
  ctx->launches++;
}

template <int NB, int NQ>
static void launch_lb_tile(gg_ctx* ctx, const TilePlan& plan, const TileDev& T, const TileArgs& a, const CellGeo& g, bool extrinsic) {
  constexpr int M = NB * NB;
  if (extrinsic) launch_lb_tile_k<NB, NQ, true, false>(ctx, plan, T, a, g);     // geometry classes are an intrinsic-form feature
  else if (plan.classes) launch_lb_tile_k<NB, NQ, false, true>(ctx, plan, T, a, g);
  else launch_lb_tile_k<NB, NQ, false, false>(ctx, plan, T, a, g);
  if (plan.group_slots) {
    ProfScope ps(ctx, "k_lb_tile_groups");
    k_lb_tile_groups<M><<<nblk(plan.group_slots, 2 * 256), 256, 0, ctx->stream>>>(T, a, plan.gcls);
    ctx->launches++;
  }
  GG_CUDA(cudaGetLastError());
}

static bool run_lb_tile(gg_mesh* m, int order, int nq, bool extrinsic, const int32_t* dofs, const double* u, double dir,
                        FillPipe* pipe) {
  // GG_LB_TILES = 0 / 1 forces the two-kernel / the tile path; default: by size.  A tile CTA lives ~35 us (its warps are
  // latency-bound), so the tile kernel needs a few waves of CTAs to pay off: measured on B200 with the L2 flushed between
  // steps (tools/tile_sizes.py), 3072 tiles (393 216 cells, 6.9 waves of 444 CTAs) 0.263 ms against 0.286 ms for the
  // two-kernel path, 1536 tiles (a 2-way slice) 0.151 / 0.159, 1055 tiles 0.112 / 0.117, 768 tiles and fewer within 1 %
  // (and slower with a warm L2) -- so slices below 2 waves keep the two-kernel path
  static const char* force = getenv("GG_LB_TILES");
  gg_ctx* ctx = m->ctx;
  gg_csr* A = pipe->A;
  if ((force && force[0] == '0') || A->test != A->trial) return false;
  if (!(force && force[0] == '1') && (m->n_cells + TILE_CELLS - 1) / TILE_CELLS < 2LL * 3 * ctx->sm_count) return false;
  const int mode = (ctx->geometry_classes && !extrinsic) ? 1 : 0;
  // geometry classes: a CTA would serve the six congruent tiles of a class from one set of element matrices, but 512 long CTAs
  // fill the 444 CTA slots of a B200 badly (measured 0.268 ms against 0.172 ms of the class path through k_gather_chunks), so
  // that combination stays opt-in
  static const bool tile_classes = getenv("GG_LB_TILES_CLASSES") && getenv("GG_LB_TILES_CLASSES")[0] == '1';
  if (mode == 1 && !tile_classes) return false;
  if (!A->tiles_tried[mode]) {
    A->tiles_tried[mode] = true;
    A->tiles[mode] = build_tile_plan(A, mode == 1);
  }
  if (!A->tiles[mode]) return false;
  const TilePlan& plan = *A->tiles[mode];
  load_const_tables(ctx, order, nq);
  CellGeo g = cell_geo(m, nq);
  TileDev T = plan.dev();
  auto even = [](int x) { return (x + 1) & ~1; };
  TileArgs a{dofs, pipe->vec_out ? u : nullptr, dir, pipe->scale, pipe->add, A->d_vals.p, pipe->vec_out,
             even(plan.max_slots + 32), even(plan.max_rows + 32), even(plan.max_rows)};
#define GG_CASE(NB_, NQ_) \
  if (order + 1 == NB_ && nq == NQ_) { launch_lb_tile<NB_, NQ_>(ctx, plan, T, a, g, extrinsic); pipe->vec_done = pipe->vec_out != nullptr; return true; }
  GG_CASE(2, 2) GG_CASE(2, 3) GG_CASE(2, 4) GG_CASE(2, 7)
  GG_CASE(3, 3) GG_CASE(3, 4) GG_CASE(3, 5) GG_CASE(3, 7) GG_CASE(3, 10)
#undef GG_CASE
  return false;
}

static void run_lb_fast(gg_mesh* m, int order, int nq, bool extrinsic, const int32_t* dofs, const double* u,
                        double dir, double* ebuf, double* vbuf, FillPipe* pipe) {
  gg_ctx* ctx = m->ctx;
  load_const_tables(ctx, order, nq);
  CellGeo g = cell_geo(m, nq);
  int64_t nc = m->n_cells;
  const int NBt = pipeline_batches();
  // Measured on B200 (round 1, 393 216 cells, tools/pipeline_sweep.sh): 0.282 ms sequential vs 0.306-0.41 ms pipelined for
  // 2, 3, 4 and 8 batches, with and without the shared-memory reservation that leaves room for the fill CTAs, fill stream at
  // the highest priority -- the co-resident fill CTAs slow the FP64-bound element kernel more than the overlap gains, so the
  // pipeline stays off unless GG_PIPELINE=1 is exported (GG_PIPELINE_BATCHES, GG_PIPELINE_ROOM select the variant).
  static const bool want_pipeline = getenv("GG_PIPELINE") && getenv("GG_PIPELINE")[0] == '1';
  if (pipe && ctx->geometry_classes && !extrinsic) {
    // Opt-in (gg_set_geometry_classes): cells with the same chart box share their intrinsic element matrix -- on the cubed
    // sphere every (column, row) pair occurs on all 6 panels -- so the FP64-bound element kernel runs on one representative per
    // class and the CSR fill reads the (L2-resident) class matrices.  Same numbers bit for bit, recomputed on every call.
    mesh_geometry_classes(m);
    if (2 * m->n_classes <= nc) {
#define GG_CASE(NB_, NQ_) \
  if (order + 1 == NB_ && nq == NQ_) { launch_lb_classes<NB_, NQ_>(ctx, m, g, dofs, u, dir, ebuf, vbuf); } else
      GG_CASE(2, 2) GG_CASE(2, 3) GG_CASE(2, 4) GG_CASE(2, 7)
      GG_CASE(3, 3) GG_CASE(3, 4) GG_CASE(3, 5) GG_CASE(3, 7) GG_CASE(3, 10)
      throw Error(GG_ERR_UNSUPPORTED, "no fast Laplace-Beltrami kernel for this (order, degree)");
#undef GG_CASE
      gather_matrix(pipe->A, 2, ebuf, pipe->scale, pipe->add);
      return;
    }
  }
  if (!pipe || !want_pipeline || nc < 64 * 1024 || ctx->profile) {
    // default, small meshes (launch-bound) and per-kernel profiling runs: one element launch, one fill launch
    run_lb_fast_range(m, order, nq, extrinsic, 0, nc, g, dofs, u, dir, ebuf, vbuf);
    if (pipe && pipe->vec_out && vbuf && !pipe->vec_done) {
      // matrix and residual of one call: the DOF gather of the residual rides on the CSR fill launch
      gather_matrix_and_vector(pipe->A, 1, ebuf, vbuf, pipe->vec_out, pipe->scale, pipe->add);
      pipe->vec_done = true;
    } else if (pipe) {
      gather_matrix(pipe->A, 1, ebuf, pipe->scale, pipe->add);
    }
    return;
  }
  prepare_gather(pipe->A, 1);
  for (int b = 0; b < NBt; ++b) {
    // batch boundaries rounded to whole CTAs; batch b owns cells with (cell * NBt) / nc == b
    int64_t c0 = (b * nc + NBt - 1) / NBt, c1 = ((b + 1) * nc + NBt - 1) / NBt;
    static const bool room = !(getenv("GG_PIPELINE_ROOM") && getenv("GG_PIPELINE_ROOM")[0] == '0');
    run_lb_fast_range(m, order, nq, extrinsic, c0, c1, g, dofs, u, dir, ebuf, vbuf, room);
    GG_CUDA(cudaEventRecord(ctx->ev_batch[b], ctx->stream));
    GG_CUDA(cudaStreamWaitEvent(ctx->stream2, ctx->ev_batch[b], 0));
    gather_matrix_batch(pipe->A, b, ebuf, pipe->scale, pipe->add, ctx->stream2);
  }
  GG_CUDA(cudaEventRecord(ctx->ev_join, ctx->stream2));
  GG_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
}

template <int NB, int NQ>
static void launch_action_fast(gg_ctx* ctx, int64_t nc, const CellGeo& g, const int32_t* dofs, const double* u, double dir,
                               double* vbuf) {
  ProfScope ps(ctx, "k_lb_action_fast");
  k_lb_action_fast<NB, NQ><<<nblk(nc, 128), 128, 0, ctx->stream>>>(nc, g, dofs, u, dir, vbuf);
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
}

static void run_action_fast(gg_mesh* m, int order, int nq, const int32_t* dofs, const double* u, double dir, double* vbuf) {
  gg_ctx* ctx = m->ctx;
  load_const_tables(ctx, order, nq);
  CellGeo g = cell_geo(m, nq);
  int64_t nc = m->n_cells;
#define GG_CASE(NB_, NQ_) \
  if (order + 1 == NB_ && nq == NQ_) { launch_action_fast<NB_, NQ_>(ctx, nc, g, dofs, u, dir, vbuf); return; }
  GG_CASE(2, 2) GG_CASE(2, 3) GG_CASE(2, 4) GG_CASE(2, 7)
  GG_CASE(3, 3) GG_CASE(3, 4) GG_CASE(3, 5) GG_CASE(3, 7) GG_CASE(3, 10)
#undef GG_CASE
  throw Error(GG_ERR_UNSUPPORTED, "no fast Laplace-Beltrami action kernel for this (order, degree)");
}

static double* scratch(DevBuf<double>& b, size_t n) {
  if (b.n < n) b.alloc(n);
  return b.p;
}

struct Plan {
  gg_ctx* ctx;
  gg_mesh* mesh;
  gg_space *test, *trial;
  int nq, mt, mu;
  double scale;
};

static Plan make_plan(gg_form form, const gg_form_args* a, bool need_trial) {
  GG_REQUIRE(a && a->test, GG_ERR_INVALID, "form args need a test space");
  Plan P;
  P.test = a->test;
  P.trial = a->trial;
  if (need_trial) GG_REQUIRE(P.trial, GG_ERR_INVALID, "this form needs a trial space");
  if (P.trial) GG_REQUIRE(P.trial->mesh == P.test->mesh, GG_ERR_INVALID, "test and trial spaces live on different meshes");
  P.mesh = P.test->mesh;
  P.ctx = P.mesh->ctx;
  GG_REQUIRE(a->degree >= 0, GG_ERR_INVALID, "negative quadrature degree");
  P.nq = num_points_1d(a->degree);
  GG_REQUIRE(P.nq <= MAX_NQ, GG_ERR_UNSUPPORTED, "quadrature degree too high (max 31)");
  P.mt = P.test->ndofs_loc;
  P.mu = P.trial ? P.trial->ndofs_loc : 0;
  P.scale = a->scale == 0.0 ? 1.0 : a->scale;
  (void)form;
  GG_CUDA(cudaSetDevice(P.ctx->device));
  return P;
}

static void form_coefficients(gg_form form, double* cM, double* cK, bool* extrinsic) {
  *extrinsic = false;
  switch (form) {
    case GG_FORM_LAPLACE_BELTRAMI: *cM = 0.0; *cK = 1.0; break;
    case GG_FORM_LAPLACE_BELTRAMI_EXTRINSIC: *cM = 0.0; *cK = 1.0; *extrinsic = true; break;
    case GG_FORM_HELMHOLTZ: *cM = 1.0; *cK = -1.0; break;
    case GG_FORM_MASS_SCALAR: *cM = 1.0; *cK = 0.0; break;
    default: throw Error(GG_ERR_INVALID, "not a scalar bilinear form");
  }
}

// ---- 3-D shell: Kronecker path (gg_hex.cuh) ----------------------------------------------------------------------------------
struct HexWork {
  const double *K2, *M2, *Dg, *Ng;
  int64_t ncol;
};

static void load_hex_tables(gg_ctx* ctx, int order) {
  if (ctx->hex_order == order) return;
  const int nb = order + 1, m2 = nb * nb, m3 = m2 * nb, G = nb * nb;
  std::vector<int32_t> l2 = quad_local_to_lex(order), l3 = hex_local_to_lex(order);
  std::vector<int32_t> loc3(m3);
  for (int I = 0; I < m3; ++I) loc3[l3[I]] = I;
  std::vector<int32_t> t3(3 * 9, 0), plane(45 * 9, -1);
  for (int i0 = 0; i0 < nb; ++i0)
    for (int A = 0; A < m2; ++A) t3[i0 * m2 + A] = loc3[i0 + nb * l2[A]];
  int p2 = 0;
  for (int a = 0; a < m2; ++a)
    for (int b = a; b < m2; ++b) {
      for (int i0 = 0; i0 < nb; ++i0)
        for (int j0 = 0; j0 < nb; ++j0) {
          if (a == b && i0 > j0) continue;  // the unordered pair is produced by (j0, i0)
          int I = t3[i0 * m2 + a], J = t3[j0 * m2 + b];
          int lo = std::min(I, J), hi = std::max(I, J);
          plane[p2 * G + i0 * nb + j0] = lo * m3 - (lo * (lo - 1)) / 2 + (hi - lo);
        }
      ++p2;
    }
  GG_CUDA(cudaStreamSynchronize(ctx->stream));  // earlier launches may still read the old tables
  GG_CUDA(cudaMemcpyToSymbol(c_hex_t3, t3.data(), t3.size() * sizeof(int32_t)));
  GG_CUDA(cudaMemcpyToSymbol(c_hex_plane, plane.data(), plane.size() * sizeof(int32_t)));
  ctx->hex_order = order;
}

// per-column surface matrices (re-integrated on every call) and per-layer radial matrices (cached) of a shell mesh
// The form cM * int phi_I phi_J sqrtg + cK * int grad(phi_I).(ginv grad(phi_J)) sqrtg keeps the Kronecker structure: with
// sqrtg = t s^2 sqrtg_surf the mass term is t NSg (x) M2, NSg = int N_i0 N_j0 s^2 h_gamma, so the kernels run unchanged on the
// per-layer tables Dg' = cK Dg + cM t^2 NSg (coefficient of M2 after the division by t) and Ng' = cK Ng.
static int hex_form_code(double cM, double cK) { return cM == 0.0 ? 0 : (cK == 0.0 ? 2 : 1); }

static HexWork hex_prepare(gg_mesh* m, int order, int nq, double cM = 0.0, double cK = 1.0) {
  gg_ctx* ctx = m->ctx;
  gg_mesh* ms = m->surface.get();
  GG_REQUIRE(order == 1 || order == 2, GG_ERR_UNSUPPORTED, "the 3-D shell path implements hexahedral Q1 / Q2");
  GG_REQUIRE(fast_lb_available(order, nq), GG_ERR_UNSUPPORTED, "no surface kernel for this (order, quadrature degree) on the shell");
  const int nb = order + 1, m2 = nb * nb, G = nb * nb;
  const int64_t ncol = ms->n_cells;
  const int key = (hex_form_code(cM, cK) * 8 + order) * 64 + nq;
  auto it = m->gamma_tabs.find(key);
  if (it == m->gamma_tabs.end()) {
    std::vector<double> xi(nq), w(nq), N(nb), D(nb), Dg((size_t)m->n_layers * G, 0.0), Ng((size_t)m->n_layers * G, 0.0);
    const double t2 = m->thickness * m->thickness;
    gauss_legendre_01(nq, xi.data(), w.data());
    const double hg = 1.0 / m->n_layers;
    for (int l = 0; l < m->n_layers; ++l)
      for (int q = 0; q < nq; ++q) {
        lagrange_1d(order, xi[q], N.data(), D.data());
        const double s = 1.0 + m->thickness * (l + xi[q]) * hg / m->radius;
        for (int i = 0; i < nb; ++i)
          for (int j = 0; j < nb; ++j) {
            Dg[(size_t)l * G + i * nb + j] += cK * w[q] * D[i] * D[j] * s * s / hg + cM * t2 * w[q] * N[i] * N[j] * s * s * hg;
            Ng[(size_t)l * G + i * nb + j] += cK * w[q] * N[i] * N[j] * hg;
          }
      }
    auto T = std::make_shared<gg_mesh::GammaTab>();
    T->Dg.upload(Dg, ctx->stream); T->Ng.upload(Ng, ctx->stream);
    T->K2.alloc((size_t)(m2 * (m2 + 1) / 2) * ncol); T->M2.alloc((size_t)m2 * m2 * ncol);
    GG_CUDA(cudaStreamSynchronize(ctx->stream));
    it = m->gamma_tabs.emplace(key, T).first;
  }
  auto& T = *it->second;
  // K2: the 2-D hot kernel on the surface mesh (packed symmetric, entry-major)
  load_const_tables(ctx, order, nq);
  CellGeo g2 = cell_geo(ms, nq);
  run_lb_fast_range(ms, order, nq, false, 0, ncol, g2, nullptr, nullptr, 0.0, T.K2.p, nullptr);
  // M2: scalar mass with sqrtg_surf (full, entry-major)
  {
    auto tt = get_tab(ctx, GG_SPACE_CG, order, nq);
    Geo g = geo_of(ms);
    unsigned nbk = nblk(ncol * m2, 128);
    if (m2 == 4) k_scalar_matrix<4, false><<<nbk, 128, 0, ctx->stream>>>(g, nq, tt->xi.p, tt->w.p, m2, tt->val.p, tt->der.p, tt->val.p, tt->der.p, 1.0, 0.0, nullptr, T.M2.p);
    else k_scalar_matrix<9, false><<<nbk, 128, 0, ctx->stream>>>(g, nq, tt->xi.p, tt->w.p, m2, tt->val.p, tt->der.p, tt->val.p, tt->der.p, 1.0, 0.0, nullptr, T.M2.p);
    ctx->launches++;
    GG_CUDA(cudaGetLastError());
  }
  load_hex_tables(ctx, order);
  return {T.K2.p, T.M2.p, T.Dg.p, T.Ng.p, ncol};
}

struct HexTensor {
  gg_space* Vs = nullptr;          // surface CG-Q_p space
  gg_csr *AK = nullptr, *AM = nullptr;  // assembled surface stiffness / mass (same pattern)
  int NR = 0;
  DevBuf<int32_t> slot2d;          // [nnz] slot of the surface pattern
  DevBuf<uint16_t> rr;             // [nnz] radial pair rho_i * NR + rho_j
  DevBuf<double2> KM;              // [nnz of the surface pattern] (mass, stiffness) interleaved
  std::map<int, std::shared_ptr<DevBuf<double2>>> radial;  // (order, nq) -> (DG / t, t NG) interleaved
  ~HexTensor() { gg_csr_destroy(AK); gg_csr_destroy(AM); gg_space_destroy(Vs); }
};

// whether the matrix can take the tensor-structured numeric phase: a complete shell (every column with all its layers)
static bool hex_tensor_applicable(gg_csr* A) {
  static const bool off = getenv("GG_SHELL_ASSEMBLY") && std::string(getenv("GG_SHELL_ASSEMBLY")) == "cellwise";
  gg_mesh* m = A->test->mesh;
  return !off && A->test == A->trial && m->dim == 3 && m->n_cells == 6LL * m->n * m->n * m->n_layers &&
         A->test->constraint == GG_CONSTRAINT_NONE && (int64_t)(A->test->order * m->n_layers + 1) * (A->test->order * m->n_layers + 1) <= 65535;
}

static HexWork hex_tensor_fill(gg_csr* A, int order, int nq, double scale, int add, double cM = 0.0, double cK = 1.0) {
  gg_space* V = A->test;
  gg_mesh* m = V->mesh;
  gg_ctx* ctx = m->ctx;
  cudaStream_t st = ctx->stream;
  auto chk = [](int code) { if (code != GG_OK) throw Error(code, gg_last_error(nullptr)); };
  const int nb = order + 1, m2 = nb * nb, m3 = m2 * nb, G = nb * nb;
  const int NR = order * m->n_layers + 1;
  if (!A->tensor) {
    auto T = std::make_shared<HexTensor>();
    T->NR = NR;
    chk(gg_space_builtin(m->surface.get(), GG_SPACE_CG, order, GG_CONSTRAINT_NONE, &T->Vs));
    chk(gg_pattern(T->Vs, T->Vs, &T->AK));
    chk(gg_pattern(T->Vs, T->Vs, &T->AM));
    // (radial node, surface DOF) of every shell DOF
    std::vector<int32_t> l2 = quad_local_to_lex(order), l3 = hex_local_to_lex(order), loc3(m3);
    for (int I = 0; I < m3; ++I) loc3[l3[I]] = I;
    std::vector<int32_t> rho(V->n_free, 0), sig(V->n_free, 0);
    for (int64_t c = 0; c < m->n_cells; ++c)
      for (int i0 = 0; i0 < nb; ++i0)
        for (int a2 = 0; a2 < m2; ++a2) {
          const int32_t d = V->cell_dofs[c * m3 + loc3[i0 + nb * l2[a2]]];
          if (d < 0) continue;
          rho[d] = order * m->cell_layer[c] + i0;
          sig[d] = T->Vs->cell_dofs[(int64_t)m->cell_col[c] * m2 + a2];
        }
    DevBuf<int32_t> drho, dsig;
    drho.upload(rho, st); dsig.upload(sig, st);
    T->slot2d.alloc(A->nnz); T->rr.alloc(A->nnz);
    DevBuf<int> bad(1);
    GG_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), st));
    k_hex_tensor_maps<<<nblk(A->n_rows, 128), 128, 0, st>>>(A->n_rows, NR, A->d_rowptr.p, A->d_colind.p, drho.p, dsig.p,
                                                           T->AK->d_rowptr.p, T->AK->d_colind.p, T->slot2d.p, T->rr.p, bad.p);
    ctx->launches++;
    GG_CUDA(cudaGetLastError());
    int nbad = 0;
    bad.download(&nbad, st);
    GG_REQUIRE(nbad == 0, GG_ERR_INVALID, "shell matrix pattern does not have the surface x radial tensor structure");
    A->tensor = T;
  }
  HexTensor& T = *A->tensor;
  // surface matrices: cell-wise quadrature on the columns + atomic-free CSR fill, every call
  HexWork W = hex_prepare(m, order, nq, cM, cK);
  gather_matrix(T.AK, 1, W.K2, 1.0, 0);
  gather_matrix(T.AM, 0, W.M2, 1.0, 0);
  // assembled radial matrices (constant per mesh, rule and form)
  const int key = (hex_form_code(cM, cK) * 8 + order) * 64 + nq;
  auto it = T.radial.find(key);
  if (it == T.radial.end()) {
    std::vector<double> Dg((size_t)m->n_layers * G), Ng((size_t)m->n_layers * G), DG((size_t)NR * NR, 0.0), NG((size_t)NR * NR, 0.0);
    GG_CUDA(cudaMemcpyAsync(Dg.data(), W.Dg, Dg.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    GG_CUDA(cudaMemcpyAsync(Ng.data(), W.Ng, Ng.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    GG_CUDA(cudaStreamSynchronize(st));
    for (int l = 0; l < m->n_layers; ++l)
      for (int i = 0; i < nb; ++i)
        for (int j = 0; j < nb; ++j) {
          DG[(size_t)(order * l + i) * NR + order * l + j] += Dg[(size_t)l * G + i * nb + j];
          NG[(size_t)(order * l + i) * NR + order * l + j] += Ng[(size_t)l * G + i * nb + j];
        }
    std::vector<double2> DN((size_t)NR * NR);
    for (size_t k = 0; k < DN.size(); ++k) DN[k] = make_double2(DG[k] / m->thickness, m->thickness * NG[k]);
    // (a row only addresses DN[rho_i * NR + rho_j] with |rho_j - rho_i| <= order and both inside 0 .. NR-1)
    auto dDN = std::make_shared<DevBuf<double2>>();
    dDN->upload(DN, st);
    GG_CUDA(cudaStreamSynchronize(st));
    it = T.radial.emplace(key, dDN).first;
  }
  {
    if (T.KM.n != (size_t)T.AK->nnz) T.KM.alloc(T.AK->nnz);
    k_hex_interleave<<<nblk(T.AK->nnz, 256), 256, 0, st>>>(T.AK->nnz, T.AM->d_vals.p, T.AK->d_vals.p, T.KM.p);
    ProfScope ps(ctx, "k_hex_tensor_fill");
    k_hex_tensor_fill<<<nblk(A->nnz, 256), 256, 0, st>>>(A->nnz, T.slot2d.p, T.rr.p, it->second->p, T.KM.p, scale, add, A->d_vals.p);
    ctx->launches += 2;
    GG_CUDA(cudaGetLastError());
  }
  return W;
}

static void run_hex_lb(gg_mesh* m, int order, int nq, const int32_t* dofs, const double* u, double dir, double* ebuf, double* vbuf,
                       const HexWork* prepared = nullptr, double cM = 0.0, double cK = 1.0) {
  gg_ctx* ctx = m->ctx;
  HexWork W = prepared ? *prepared : hex_prepare(m, order, nq, cM, cK);
  const int64_t nc = m->n_cells;
  unsigned nbk = nblk(nc, 128);
  if (ebuf) {
    ProfScope ps(ctx, "k_hex_kron");
    if (order == 1) k_hex_kron<2><<<nbk, 128, 0, ctx->stream>>>(nc, W.ncol, m->d_col.p, m->d_layer.p, W.K2, W.M2, W.Dg, W.Ng, m->thickness, ebuf);
    else k_hex_kron<3><<<nbk, 128, 0, ctx->stream>>>(nc, W.ncol, m->d_col.p, m->d_layer.p, W.K2, W.M2, W.Dg, W.Ng, m->thickness, ebuf);
    ctx->launches++;
  }
  if (vbuf) {
    ProfScope ps(ctx, "k_hex_residual");
    if (order == 1) k_hex_residual<2><<<nbk, 128, 0, ctx->stream>>>(nc, W.ncol, m->d_col.p, m->d_layer.p, W.K2, W.M2, W.Dg, W.Ng, m->thickness, dofs, u, dir, vbuf);
    else k_hex_residual<3><<<nbk, 128, 0, ctx->stream>>>(nc, W.ncol, m->d_col.p, m->d_layer.p, W.K2, W.M2, W.Dg, W.Ng, m->thickness, dofs, u, dir, vbuf);
    ctx->launches++;
  }
  GG_CUDA(cudaGetLastError());
}

// element matrices into ctx->ebuf; returns the buffer layout (0 full entry-major, 1 symmetric packed entry-major)
static int compute_element_matrices(gg_form form, const gg_form_args* a, const Plan& P, bool want_fused_vector, double** vbuf_out,
                                    FillPipe* pipe = nullptr) {
  gg_ctx* ctx = P.ctx;
  Geo g = geo_of(P.mesh);
  cudaStream_t st = ctx->stream;
  int64_t nc = g.nc;
  if (vbuf_out) *vbuf_out = nullptr;
  if (nc == 0) return 0;
  if (P.mesh->dim == 3) {
    GG_REQUIRE((form == GG_FORM_LAPLACE_BELTRAMI || form == GG_FORM_HELMHOLTZ || form == GG_FORM_MASS_SCALAR) && !a->qdata,
               GG_ERR_UNSUPPORTED, "on the 3-D shell the Laplace-Beltrami, Helmholtz and scalar mass forms (without coefficient) are implemented");
    GG_REQUIRE(P.test->kind == GG_SPACE_CG && P.trial->kind == GG_SPACE_CG && P.test->order == P.trial->order, GG_ERR_INVALID,
               "the shell forms need CG test/trial spaces of the same order");
    double cM3, cK3; bool ext3;
    form_coefficients(form, &cM3, &cK3, &ext3);
    double* vb = nullptr;
    if (want_fused_vector) {
      GG_REQUIRE(a->u && gg_vec_size(a->u) == P.trial->n_free, GG_ERR_INVALID, "fused residual needs args->u of the trial size");
      vb = scratch(ctx->vbuf, (size_t)P.mt * nc);
      *vbuf_out = vb;
    }
    if (pipe && hex_tensor_applicable(pipe->A)) {
      // assembled matrix straight from the tensor structure (no element matrices); the residual stays cell-wise
      HexWork W = hex_tensor_fill(pipe->A, P.test->order, P.nq, pipe->scale, pipe->add, cM3, cK3);
      if (vb) run_hex_lb(P.mesh, P.test->order, P.nq, P.trial->d_cell_dofs.p, a->u->d.p, a->dirichlet, nullptr, vb, &W);
      return -1;
    }
    size_t np = (size_t)P.mt * (P.mt + 1) / 2;
    double* eb = scratch(ctx->ebuf, np * nc);
    run_hex_lb(P.mesh, P.test->order, P.nq, P.trial->d_cell_dofs.p, a->u ? a->u->d.p : nullptr, a->dirichlet, eb, vb, nullptr, cM3, cK3);
    return 1;
  }
  const bool stiffness = form == GG_FORM_LAPLACE_BELTRAMI || form == GG_FORM_LAPLACE_BELTRAMI_EXTRINSIC;
  if (form == GG_FORM_LAPLACE_BELTRAMI || form == GG_FORM_LAPLACE_BELTRAMI_EXTRINSIC || form == GG_FORM_HELMHOLTZ ||
      form == GG_FORM_MASS_SCALAR) {
    GG_REQUIRE(is_scalar(P.test) && is_scalar(P.trial), GG_ERR_INVALID, "scalar form needs Lagrangian test/trial spaces");
    double cM, cK; bool ext;
    form_coefficients(form, &cM, &cK, &ext);
    if (stiffness && !a->qdata && P.test->order == P.trial->order && fast_lb_available(P.test->order, P.nq)) {
      if (want_fused_vector) {
        GG_REQUIRE(a->u, GG_ERR_INVALID, "fused residual needs args->u");
        GG_REQUIRE(gg_vec_size(a->u) == P.trial->n_free, GG_ERR_INVALID, "args->u has the wrong length");
      }
      if (pipe && (!want_fused_vector || pipe->vec_out) &&
          run_lb_tile(P.mesh, P.test->order, P.nq, ext, P.trial->d_cell_dofs.p, a->u ? a->u->d.p : nullptr, a->dirichlet, pipe))
        return -1;  // CSR values (and the residual, pipe->vec_done) written by the fused tile kernel
      size_t np = (size_t)P.mt * (P.mt + 1) / 2;
      double* eb = scratch(ctx->ebuf, np * nc);
      double* vb = nullptr;
      if (want_fused_vector) {
        vb = scratch(ctx->vbuf, (size_t)P.mt * nc);
        *vbuf_out = vb;
      }
      run_lb_fast(P.mesh, P.test->order, P.nq, ext, P.trial->d_cell_dofs.p, a->u ? a->u->d.p : nullptr, a->dirichlet, eb, vb, pipe);
      return pipe ? -1 : 1;  // -1: the CSR fill has already been enqueued
    }
    GG_REQUIRE(!want_fused_vector, GG_ERR_UNSUPPORTED, "fused matrix+vector assembly needs the fast Q1/Q2 stiffness kernel");
    if (a->qdata) GG_REQUIRE(gg_vec_size(a->qdata) == nc * P.nq * P.nq, GG_ERR_INVALID, "qdata has the wrong length");
    auto tt = get_tab(ctx, P.test->kind, P.test->order, P.nq);
    auto tu = get_tab(ctx, P.trial->kind, P.trial->order, P.nq);
    double* eb = scratch(ctx->ebuf, (size_t)P.mt * P.mu * nc);
    const double* qc = a->qdata ? a->qdata->d.p : nullptr;
    unsigned nb = nblk(nc * P.mt, 128);
    ProfScope ps(ctx, "k_scalar_matrix");
#define GG_CASE(MU_)                                                                                                      \
  if (P.mu == MU_) {                                                                                                      \
    if (ext) k_scalar_matrix<MU_, true><<<nb, 128, 0, st>>>(g, P.nq, tt->xi.p, tt->w.p, P.mt, tt->val.p, tt->der.p, tu->val.p, tu->der.p, cM, cK, qc, eb); \
    else k_scalar_matrix<MU_, false><<<nb, 128, 0, st>>>(g, P.nq, tt->xi.p, tt->w.p, P.mt, tt->val.p, tt->der.p, tu->val.p, tu->der.p, cM, cK, qc, eb); \
  } else
    GG_CASE(1) GG_CASE(4) GG_CASE(9) GG_CASE(16) GG_CASE(25) throw Error(GG_ERR_UNSUPPORTED, "unsupported trial order");
#undef GG_CASE
    ctx->launches++;
    GG_CUDA(cudaGetLastError());
    return 0;
  }
  GG_REQUIRE(!want_fused_vector, GG_ERR_UNSUPPORTED, "fused matrix+vector assembly is only implemented for the stiffness forms");
  if (form == GG_FORM_MASS_RT) {
    GG_REQUIRE(P.test->kind == GG_SPACE_RT && P.trial->kind == GG_SPACE_RT && P.test->order == P.trial->order, GG_ERR_INVALID,
               "RT mass needs RT test and trial spaces of the same order");
    auto tu = get_tab(ctx, GG_SPACE_RT, P.trial->order, P.nq);
    double* eb = scratch(ctx->ebuf, (size_t)P.mt * P.mu * nc);
    unsigned nb = nblk(nc * P.mt, 128);
#define GG_CASE(MU_) \
  if (P.mu == MU_) k_rt_mass_matrix<MU_><<<nb, 128, 0, st>>>(g, P.nq, tu->xi.p, tu->w.p, tu->val.p, P.trial->d_cell_signs.p, eb); else
    GG_CASE(4) GG_CASE(12) GG_CASE(24) throw Error(GG_ERR_UNSUPPORTED, "RT order must be <= 2 for matrix assembly");
#undef GG_CASE
    ctx->launches++;
    GG_CUDA(cudaGetLastError());
    return 0;
  }
  if (form == GG_FORM_MASS_VECTOR) {
    GG_REQUIRE(P.test->kind == GG_SPACE_CGV && P.trial == P.test, GG_ERR_INVALID, "the vector mass form needs one vector H1 space as test and trial");
    double* eb = scratch(ctx->ebuf, (size_t)P.mt * P.mu * nc);
    vec_mass_element_matrices(P.test, P.nq, eb);
    return 0;
  }
  if (form == GG_FORM_VECTOR_LAPLACIAN) {
    GG_REQUIRE(P.test->kind == GG_SPACE_CGV && P.trial == P.test, GG_ERR_INVALID, "the vector Laplacian form needs one vector H1 space as test and trial");
    double* eb = scratch(ctx->ebuf, (size_t)P.mt * P.mu * nc);
    vec_laplacian_element_matrices(P.test, P.nq, eb);
    return 0;
  }
  if (form == GG_FORM_VECTOR_DIV) {
    GG_REQUIRE(is_scalar(P.test) && P.trial->kind == GG_SPACE_CGV, GG_ERR_INVALID, "the vector divergence form needs a Lagrangian test and a vector H1 trial space");
    double* eb = scratch(ctx->ebuf, (size_t)P.mt * P.mu * nc);
    vec_div_element_matrices(P.test, P.trial, P.nq, eb);
    return 0;
  }
  if (form == GG_FORM_DIV) {
    GG_REQUIRE(is_scalar(P.test) && P.trial->kind == GG_SPACE_RT, GG_ERR_INVALID, "div form needs a scalar test and an RT trial space");
    auto tt = get_tab(ctx, P.test->kind, P.test->order, P.nq);
    auto tu = get_tab(ctx, GG_SPACE_RT, P.trial->order, P.nq);
    double* eb = scratch(ctx->ebuf, (size_t)P.mt * P.mu * nc);
    unsigned nb = nblk(nc * P.mt, 128);
#define GG_CASE(MU_) \
  if (P.mu == MU_) k_div_matrix<MU_><<<nb, 128, 0, st>>>(nc, P.nq, tt->w.p, P.mt, tt->val.p, tu->der.p, P.trial->d_cell_signs.p, eb); else
    GG_CASE(4) GG_CASE(12) GG_CASE(24) throw Error(GG_ERR_UNSUPPORTED, "RT order must be <= 2 for matrix assembly");
#undef GG_CASE
    ctx->launches++;
    GG_CUDA(cudaGetLastError());
    return 0;
  }
  throw Error(GG_ERR_INVALID, "not a bilinear form id");
}

// element vectors into ctx->vbuf (entry-major [I][cell])
static double* compute_element_vectors(gg_form form, const gg_form_args* a, const Plan& P) {
  gg_ctx* ctx = P.ctx;
  Geo g = geo_of(P.mesh);
  cudaStream_t st = ctx->stream;
  int64_t nc = g.nc;
  double* vb = scratch(ctx->vbuf, (size_t)P.mt * std::max<int64_t>(nc, 1));
  if (nc == 0) return vb;
  if (P.mesh->dim == 3 && form == GG_FORM_SOURCE) {
    GG_REQUIRE(P.test->kind == GG_SPACE_CG, GG_ERR_UNSUPPORTED, "the shell source form needs a CG test space");
    GG_REQUIRE(a->qdata && gg_vec_size(a->qdata) == nc * P.nq * P.nq * P.nq, GG_ERR_INVALID, "source form needs qdata of length n_cells*nq^3");
    hex_source_vectors(P.test, P.nq, a->qdata->d.p, vb);
    return vb;
  }
  if (P.mesh->dim == 3 && form == GG_FORM_SHELL_BOUNDARY_FLUX) {
    GG_REQUIRE(P.test->kind == GG_SPACE_CG, GG_ERR_UNSUPPORTED, "the shell boundary form needs a CG test space");
    GG_REQUIRE(a->u && gg_vec_size(a->u) == P.test->n_free, GG_ERR_INVALID, "args->u missing or of the wrong length");
    hex_boundary_flux_vectors(P.test, P.nq, a->u->d.p, a->dirichlet, vb);
    return vb;
  }
  GG_REQUIRE(form != GG_FORM_SHELL_BOUNDARY_FLUX, GG_ERR_UNSUPPORTED, "the boundary-flux form exists on the 3-D shell only");
  if (P.mesh->dim == 3) {
    GG_REQUIRE((form == GG_FORM_LAPLACE_BELTRAMI || form == GG_FORM_HELMHOLTZ || form == GG_FORM_MASS_SCALAR) && !a->qdata && P.trial &&
                   P.test->kind == GG_SPACE_CG && P.trial->kind == GG_SPACE_CG && P.test->order == P.trial->order, GG_ERR_UNSUPPORTED,
               "on the 3-D shell the actions of the Laplace-Beltrami, Helmholtz and scalar mass forms are implemented");
    GG_REQUIRE(a->u && gg_vec_size(a->u) == P.trial->n_free, GG_ERR_INVALID, "args->u missing or of the wrong length");
    double cM3, cK3; bool ext3;
    form_coefficients(form, &cM3, &cK3, &ext3);
    run_hex_lb(P.mesh, P.test->order, P.nq, P.trial->d_cell_dofs.p, a->u->d.p, a->dirichlet, nullptr, vb, nullptr, cM3, cK3);
    return vb;
  }
  if (form == GG_FORM_SOURCE_VECTOR) {
    GG_REQUIRE(P.test->kind == GG_SPACE_CGV, GG_ERR_INVALID, "the vector source form needs a vector H1 test space");
    GG_REQUIRE(a->qdata && gg_vec_size(a->qdata) == 2 * nc * P.nq * P.nq, GG_ERR_INVALID, "vector source form needs qdata of length n_cells*Q*2");
    vec_source_element_vectors(P.test, P.nq, a->qdata->d.p, vb);
    return vb;
  }
  if (form == GG_FORM_SOURCE) {
    GG_REQUIRE(is_scalar(P.test), GG_ERR_INVALID, "source form needs a Lagrangian test space");
    GG_REQUIRE(a->qdata && gg_vec_size(a->qdata) == nc * P.nq * P.nq, GG_ERR_INVALID, "source form needs qdata of length n_cells*Q");
    auto tt = get_tab(ctx, P.test->kind, P.test->order, P.nq);
    k_source_vector<<<nblk(nc * P.mt, 128), 128, 0, st>>>(g, P.nq, tt->xi.p, tt->w.p, P.mt, tt->val.p, a->qdata->d.p, vb);
    ctx->launches++;
    GG_CUDA(cudaGetLastError());
    return vb;
  }
  // action of a scalar bilinear form on u_h
  GG_REQUIRE(P.trial, GG_ERR_INVALID, "the action of a bilinear form needs the trial space");
  GG_REQUIRE(is_scalar(P.test) && is_scalar(P.trial), GG_ERR_INVALID, "scalar form needs Lagrangian test/trial spaces");
  GG_REQUIRE(a->u && gg_vec_size(a->u) == P.trial->n_free, GG_ERR_INVALID, "args->u missing or of the wrong length");
  double cM, cK; bool ext;
  form_coefficients(form, &cM, &cK, &ext);
  if (form == GG_FORM_LAPLACE_BELTRAMI && P.test->order == P.trial->order && fast_lb_available(P.test->order, P.nq)) {
    run_action_fast(P.mesh, P.test->order, P.nq, P.trial->d_cell_dofs.p, a->u->d.p, a->dirichlet, vb);
    return vb;
  }
  auto tt = get_tab(ctx, P.test->kind, P.test->order, P.nq);
  auto tu = get_tab(ctx, P.trial->kind, P.trial->order, P.nq);
  unsigned nb = nblk(nc * P.mt, 128);
#define GG_CASE(MU_)                                                                                                       \
  if (P.mu == MU_) {                                                                                                       \
    if (ext) k_scalar_action<MU_, true><<<nb, 128, 0, st>>>(g, P.nq, tt->xi.p, tt->w.p, P.mt, tt->val.p, tt->der.p, tu->val.p, tu->der.p, cM, cK, P.trial->d_cell_dofs.p, a->u->d.p, a->dirichlet, vb); \
    else k_scalar_action<MU_, false><<<nb, 128, 0, st>>>(g, P.nq, tt->xi.p, tt->w.p, P.mt, tt->val.p, tt->der.p, tu->val.p, tu->der.p, cM, cK, P.trial->d_cell_dofs.p, a->u->d.p, a->dirichlet, vb); \
  } else
  GG_CASE(1) GG_CASE(4) GG_CASE(9) GG_CASE(16) GG_CASE(25) throw Error(GG_ERR_UNSUPPORTED, "unsupported trial order");
#undef GG_CASE
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
  return vb;
}

// element matrices of `form` left on the device (ctx scratch), for other translation units
const double* element_matrices_device(gg_form form, const gg_form_args* args, int* layout) {
  Plan P = make_plan(form, args, true);
  *layout = compute_element_matrices(form, args, P, false, nullptr);
  return P.ctx->ebuf.p;
}

}  // namespace gg

using namespace gg;

extern "C" {

int gg_assemble_matrix(gg_form form, const gg_form_args* args, gg_csr* A, int add) {
  GG_API_BEGIN
  GG_REQUIRE(A, GG_ERR_INVALID, "null matrix");
  Plan P = make_plan(form, args, true);
  GG_REQUIRE(A->test == P.test && A->trial == P.trial, GG_ERR_INVALID, "matrix pattern was built for different spaces");
  GG_REQUIRE(!A->skeleton, GG_ERR_INVALID, "skeleton patterns are filled by gg_assemble_advection");
  FillPipe pipe{A, P.scale, add};
  int layout = compute_element_matrices(form, args, P, false, nullptr, &pipe);
  if (layout >= 0) gather_matrix(A, layout, P.ctx->ebuf.p, P.scale, add);
  finish_call(P.ctx);
  GG_API_END
}

int gg_assemble_vector(gg_form form, const gg_form_args* args, gg_vec* b, int add) {
  GG_API_BEGIN
  GG_REQUIRE(b, GG_ERR_INVALID, "null vector");
  Plan P = make_plan(form, args, false);
  GG_REQUIRE(gg_vec_size(b) == P.test->n_free, GG_ERR_INVALID, "output vector has the wrong length");
  double* vb = compute_element_vectors(form, args, P);
  gather_vector(P.test, vb, b->d.p, P.scale, add);
  finish_call(P.ctx);
  GG_API_END
}

int gg_assemble_matrix_and_vector(gg_form form, const gg_form_args* args, gg_csr* A, gg_vec* b) {
  GG_API_BEGIN
  GG_REQUIRE(A && b, GG_ERR_INVALID, "null output");
  Plan P = make_plan(form, args, true);
  GG_REQUIRE(A->test == P.test && A->trial == P.trial, GG_ERR_INVALID, "matrix pattern was built for different spaces");
  GG_REQUIRE(gg_vec_size(b) == P.test->n_free, GG_ERR_INVALID, "output vector has the wrong length");
  double* vb = nullptr;
  FillPipe pipe{A, P.scale, 0, b->d.p};
  int layout = compute_element_matrices(form, args, P, true, &vb, &pipe);
  if (layout >= 0) gather_matrix(A, layout, P.ctx->ebuf.p, P.scale, 0);
  if (!pipe.vec_done) gather_vector(P.test, vb, b->d.p, P.scale, 0);
  finish_call(P.ctx);
  GG_API_END
}

int gg_element_matrices(gg_form form, const gg_form_args* args, int64_t first_cell, int64_t n, double* out) {
  GG_API_BEGIN
  GG_REQUIRE(out, GG_ERR_INVALID, "null output");
  Plan P = make_plan(form, args, true);
  int64_t nc = P.mesh->n_cells;
  GG_REQUIRE(first_cell >= 0 && n >= 0 && first_cell + n <= nc, GG_ERR_INVALID, "cell range out of bounds");
  int layout = compute_element_matrices(form, args, P, false, nullptr);
  size_t planes = layout == 1 ? (size_t)P.mt * (P.mt + 1) / 2 : (size_t)P.mt * P.mu;
  std::vector<double> h(planes * nc);
  if (nc) GG_CUDA(cudaMemcpyAsync(h.data(), P.ctx->ebuf.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, P.ctx->stream));
  GG_CUDA(cudaStreamSynchronize(P.ctx->stream));
  for (int64_t c = 0; c < n; ++c)
    for (int i = 0; i < P.mt; ++i)
      for (int j = 0; j < P.mu; ++j) {
        size_t pl;
        if (layout == 1) {
          int a = std::min(i, j), b = std::max(i, j);
          pl = (size_t)a * P.mu - (size_t)(a * (a - 1)) / 2 + (b - a);
        } else {
          pl = (size_t)i * P.mu + j;
        }
        out[((size_t)c * P.mt + i) * P.mu + j] = P.scale * h[pl * nc + first_cell + c];
      }
  GG_API_END
}

int gg_element_vectors(gg_form form, const gg_form_args* args, int64_t first_cell, int64_t n, double* out) {
  GG_API_BEGIN
  GG_REQUIRE(out, GG_ERR_INVALID, "null output");
  Plan P = make_plan(form, args, false);
  int64_t nc = P.mesh->n_cells;
  GG_REQUIRE(first_cell >= 0 && n >= 0 && first_cell + n <= nc, GG_ERR_INVALID, "cell range out of bounds");
  double* vb = compute_element_vectors(form, args, P);
  std::vector<double> h((size_t)P.mt * nc);
  if (nc) GG_CUDA(cudaMemcpyAsync(h.data(), vb, h.size() * sizeof(double), cudaMemcpyDeviceToHost, P.ctx->stream));
  GG_CUDA(cudaStreamSynchronize(P.ctx->stream));
  for (int64_t c = 0; c < n; ++c)
    for (int i = 0; i < P.mt; ++i) out[(size_t)c * P.mt + i] = P.scale * h[(size_t)i * nc + first_cell + c];
  GG_API_END
}

int gg_quadrature_points(gg_mesh* m, int32_t degree, double* out_chart, double* out_xyz) {
  GG_API_BEGIN
  GG_REQUIRE(m, GG_ERR_INVALID, "null mesh");
  gg_ctx* ctx = m->ctx;
  GG_CUDA(cudaSetDevice(ctx->device));
  int nq = num_points_1d(degree);
  GG_REQUIRE(degree >= 0 && nq <= (m->dim == 3 ? MAX_NQ : MAX_NQ_NORMS), GG_ERR_UNSUPPORTED, "quadrature degree out of range");
  if (m->dim == 3) {
    // [n_cells][nq^3][3]: chart (gamma, alpha, beta) / ambient, gamma index fastest
    hex_points(m, -1, nq, out_chart, out_xyz);
    return GG_OK;
  }
  auto T = get_tab(ctx, GG_SPACE_DG, 0, nq);
  int64_t n = m->n_cells * nq * nq;
  if (n == 0) return GG_OK;
  DevBuf<double> dc(out_chart ? 2 * n : 0), dx(out_xyz ? 3 * n : 0);
  k_qpoints<<<nblk(n, 256), 256, 0, ctx->stream>>>(geo_of(m), nq, T->xi.p, dc.p, dx.p);
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
  if (out_chart) dc.download(out_chart, ctx->stream);
  if (out_xyz) dx.download(out_xyz, ctx->stream);
  GG_API_END
}

int gg_integrate(gg_mesh* m, int32_t degree, gg_vec* qdata, double* out) {
  GG_API_BEGIN
  GG_REQUIRE(m && out, GG_ERR_INVALID, "null argument");
  gg_ctx* ctx = m->ctx;
  GG_CUDA(cudaSetDevice(ctx->device));
  int nq = num_points_1d(degree);
  auto T = get_tab(ctx, GG_SPACE_DG, 0, nq);
  int64_t nc = m->n_cells;
  if (qdata && m->dim == 2) GG_REQUIRE(gg_vec_size(qdata) == nc * nq * nq, GG_ERR_INVALID, "qdata has the wrong length");
  *out = 0.0;
  if (nc == 0) return GG_OK;
  DevBuf<double> sums(nc), total(1);
  if (m->dim == 3) {
    if (qdata) GG_REQUIRE(gg_vec_size(qdata) == nc * nq * nq * nq, GG_ERR_INVALID, "qdata has the wrong length");
    k_integrate_shell<<<nblk(nc, 128), 128, 0, ctx->stream>>>(nc, m->n, m->n_layers, m->radius, m->thickness, nq, T->xi.p, T->w.p,
                                                             m->d_col.p, m->d_layer.p, qdata ? qdata->d.p : nullptr, sums.p);
  } else
  k_integrate_cells<<<nblk(nc, 128), 128, 0, ctx->stream>>>(geo_of(m), nq, T->xi.p, T->w.p, qdata ? qdata->d.p : nullptr,
                                                            m->d_cell_owned.n ? m->d_cell_owned.p : nullptr, sums.p);
  size_t tb = 0;
  GG_CUDA(cub::DeviceReduce::Sum(nullptr, tb, sums.p, total.p, (int)nc, ctx->stream));
  DevBuf<uint8_t> tmp(tb);
  GG_CUDA(cub::DeviceReduce::Sum(tmp.p, tb, sums.p, total.p, (int)nc, ctx->stream));
  ctx->launches += 3;
  total.download(out, ctx->stream);
  GG_API_END
}

int gg_eval_geometry(gg_ctx* ctx, int panel, double radius, int64_t n, const double* pts, double* xyz, double* jac,
                     double* metric, double* inv_metric, double* sqrtg) {
  GG_API_BEGIN
  GG_REQUIRE(ctx && pts, GG_ERR_INVALID, "null argument");
  GG_REQUIRE(panel >= 1 && panel <= 6, GG_ERR_INVALID, "panel must be in 1..6");
  GG_REQUIRE(n >= 0, GG_ERR_INVALID, "negative n");
  GG_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) return GG_OK;
  cudaStream_t st = ctx->stream;
  DevBuf<double> dp; dp.upload(pts, 2 * n, st);
  DevBuf<double> dx(xyz ? 3 * n : 0), dj(jac ? 6 * n : 0), dg(metric ? 3 * n : 0), di(inv_metric ? 3 * n : 0), ds(sqrtg ? n : 0);
  k_eval_geometry<<<nblk(n, 128), 128, 0, st>>>(panel - 1, radius, n, dp.p, dx.p, dj.p, dg.p, di.p, ds.p);
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
  if (xyz) dx.download(xyz, st);
  if (jac) dj.download(jac, st);
  if (metric) dg.download(metric, st);
  if (inv_metric) di.download(inv_metric, st);
  if (sqrtg) ds.download(sqrtg, st);
  GG_CUDA(cudaStreamSynchronize(st));
  GG_API_END
}

}  // extern "C"
