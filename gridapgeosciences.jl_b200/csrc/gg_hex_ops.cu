// 3-D shell: the pieces of the Laplace-Beltrami driver around the assembled operator -- quadrature / interpolation points,
// source vector, the boundary term of the integration by parts on the bottom and top faces, error norms, zero-mean shift
// (test/Laplacian/LaplaceBeltrami.jl:19-98 with Dc = 3; geometry src/Fields/CubedSphereWithThickness{Map,Metric,InvMetric}.jl).
// Point data cannot use the Kronecker factorisation of the operator (gg_hex.cuh), so these are plain tensor-rule loops: one
// thread per (cell, test function) or per cell.  Geometry comes from the column / layer of the cell: chart box of the column on
// its panel, gamma in [layer, layer + 1] / L, x = (1 + t gamma / r) phi_panel(alpha, beta), sqrtg = t s^2 sqrtg_surf.
#include <cub/cub.cuh>
#include "gg_common.cuh"
#include "gg_refel.cuh"
#include "gg_geom.cuh"
#include "gg_hex_ops.cuh"

namespace gg {

static inline unsigned nblk(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

struct ShellGeo {
  int64_t nc;
  int n, L;
  double radius, t;
  const int32_t *col, *layer;
};
static ShellGeo shell_geo(gg_mesh* m) { return {m->n_cells, m->n, m->n_layers, m->radius, m->thickness, m->d_col.p, m->d_layer.p}; }

struct HexRef {
  DevBuf<double> pts, w, N, D;   // 1-D points / weights [np], 1-D basis values / derivatives [np][nb]
  DevBuf<int32_t> lex;           // local dof -> i0 + nb (i1 + nb i2)
  int np = 0, nb = 0;
};

// 1-D tabulation of the Lagrange basis of `order` at the Gauss points of the rule (nq > 0) or at the nodes (nq == 0)
static void hex_ref(gg_ctx* ctx, int order, int nq, HexRef& R) {
  const int nb = order + 1, np = nq > 0 ? nq : nb;
  std::vector<double> x(np), w(np, 0.0), N((size_t)np * nb), D((size_t)np * nb);
  if (nq > 0) gauss_legendre_01(nq, x.data(), w.data());
  else for (int i = 0; i < nb; ++i) x[i] = order > 0 ? (double)i / order : 0.5;
  for (int q = 0; q < np; ++q) lagrange_1d(order, x[q], &N[(size_t)q * nb], &D[(size_t)q * nb]);
  cudaStream_t st = ctx->stream;
  R.pts.upload(x, st); R.w.upload(w, st); R.N.upload(N, st); R.D.upload(D, st);
  R.lex.upload(hex_local_to_lex(order), st);
  R.np = np; R.nb = nb;
}

__device__ __forceinline__ void shell_cell_box(const ShellGeo& g, int64_t c, int* panel, double* a0, double* b0, double* ha, double* g0,
                                               double* hg) {
  const double H = 0.78539816339744830962;
  const int cc = g.col[c] % (g.n * g.n);
  *panel = g.col[c] / (g.n * g.n);
  *ha = 2.0 * H / g.n;
  *a0 = -H + (cc % g.n) * *ha;
  *b0 = -H + (cc / g.n) * *ha;
  *hg = 1.0 / g.L;
  *g0 = g.layer[c] * *hg;
}

// points of a tensor rule (or the nodes) of every cell: chart (gamma, alpha, beta) and ambient coordinates, first index fastest
__global__ void k_hex_points(ShellGeo g, int np, const double* __restrict__ x1, double* __restrict__ chart, double* __restrict__ xyz) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int P3 = np * np * np;
  if (t >= g.nc * P3) return;
  const int64_t c = t / P3;
  const int q = (int)(t - c * P3), q0 = q % np, q1 = (q / np) % np, q2 = q / (np * np);
  int panel;
  double a0, b0, ha, g0, hg;
  shell_cell_box(g, c, &panel, &a0, &b0, &ha, &g0, &hg);
  const double gam = g0 + hg * x1[q0], al = a0 + ha * x1[q1], be = b0 + ha * x1[q2];
  if (chart) { chart[3 * t] = gam; chart[3 * t + 1] = al; chart[3 * t + 2] = be; }
  if (xyz) {
    double X[3];
    csphere_point(panel, tan(al), tan(be), g.radius, X);
    const double s = 1.0 + g.t * gam / g.radius;
    xyz[3 * t] = s * X[0]; xyz[3 * t + 1] = s * X[1]; xyz[3 * t + 2] = s * X[2];
  }
}

// element vector of int f v sqrtg, one thread per (cell, test function); entry-major out[a][cell]
__global__ void k_hex_source(ShellGeo g, int nq, int nb, const double* __restrict__ x1, const double* __restrict__ w1,
                             const double* __restrict__ N, const int32_t* __restrict__ lex, const double* __restrict__ fq,
                             double* __restrict__ vbuf) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int m = nb * nb * nb;
  if (t >= g.nc * m) return;
  const int64_t c = t % g.nc;
  const int a = (int)(t / g.nc), l = lex[a], i0 = l % nb, i1 = (l / nb) % nb, i2 = l / (nb * nb);
  int panel;
  double a0, b0, ha, g0, hg;
  shell_cell_box(g, c, &panel, &a0, &b0, &ha, &g0, &hg);
  const int Q = nq * nq * nq;
  double acc = 0.0;
  for (int q2 = 0; q2 < nq; ++q2) {
    const double tb = tan(b0 + ha * x1[q2]);
    for (int q1 = 0; q1 < nq; ++q1) {
      const Metric2 mt = csphere_metric_terms(tan(a0 + ha * x1[q1]), tb, g.radius);
      const double w12 = w1[q1] * w1[q2] * ha * ha * mt.sqrtg * N[q1 * nb + i1] * N[q2 * nb + i2];
      for (int q0 = 0; q0 < nq; ++q0) {
        const double s = 1.0 + g.t * (g0 + hg * x1[q0]) / g.radius;
        acc = fma(w12 * w1[q0] * hg * g.t * s * s * N[q0 * nb + i0], fq[c * Q + q0 + nq * (q1 + nq * q2)], acc);
      }
    }
  }
  vbuf[(int64_t)a * g.nc + c] = acc;
}

// boundary term of the integration by parts: int_G ((ginv grad(u_h)).n_G) v sqrtg dG on gamma = 0 (n_G = -e_gamma) and gamma = 1
// (n_G = +e_gamma); ginv_gg = 1/t^2, sqrtg = t s^2 sqrtg_surf, dG = chart measure of the face.  One thread per (cell, test fn).
__global__ void k_hex_boundary_flux(ShellGeo g, int nq, int nb, const double* __restrict__ x1, const double* __restrict__ w1,
                                    const double* __restrict__ N, const double* __restrict__ Nn, const double* __restrict__ Dn,
                                    const int32_t* __restrict__ lex, const int32_t* __restrict__ dofs, const double* __restrict__ u,
                                    double dirichlet, double* __restrict__ vbuf) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int m = nb * nb * nb;
  if (t >= g.nc * m) return;
  const int64_t c = t % g.nc;
  const int a = (int)(t / g.nc), l = lex[a], i0 = l % nb, i1 = (l / nb) % nb, i2 = l / (nb * nb);
  double acc = 0.0;
  const int lay = g.layer[c];
  int panel;
  double a0, b0, ha, g0, hg;
  shell_cell_box(g, c, &panel, &a0, &b0, &ha, &g0, &hg);
  for (int side = 0; side < 2; ++side) {
    if ((side == 0 && lay != 0) || (side == 1 && lay != g.L - 1)) continue;
    const int node = side == 0 ? 0 : nb - 1;                      // 1-D node on the face: N_i0 = delta(i0, node)
    if (i0 != node) continue;
    const double sgn = side == 0 ? -1.0 : 1.0;
    const double s = 1.0 + g.t * (side == 0 ? 0.0 : 1.0) / g.radius;
    for (int q2 = 0; q2 < nq; ++q2) {
      const double tb = tan(b0 + ha * x1[q2]);
      for (int q1 = 0; q1 < nq; ++q1) {
        const Metric2 mt = csphere_metric_terms(tan(a0 + ha * x1[q1]), tb, g.radius);
        // d u_h / d gamma at the face point
        double du = 0.0;
        for (int b = 0; b < m; ++b) {
          const int lb = lex[b], j0 = lb % nb, j1 = (lb / nb) % nb, j2 = lb / (nb * nb);
          const int32_t d = dofs[c * m + b];
          du = fma(Dn[node * nb + j0] * N[q1 * nb + j1] * N[q2 * nb + j2], d >= 0 ? u[d] : dirichlet, du);
        }
        du /= hg;
        const double flux = sgn * du / (g.t * g.t);
        acc = fma(w1[q1] * w1[q2] * ha * ha * g.t * s * s * mt.sqrtg * flux, N[q1 * nb + i1] * N[q2 * nb + i2], acc);
      }
    }
  }
  (void)Nn;
  vbuf[(int64_t)a * g.nc + c] = acc;
}

// per-cell functionals of u_h.  MODE 0: int (u_h - f)^2 sqrtg; MODE 1: int u_h in the chart measure; MODE 2: chart volume
template <int MODE>
__global__ void k_hex_cell_functional(ShellGeo g, int nq, int nb, const double* __restrict__ x1, const double* __restrict__ w1,
                                      const double* __restrict__ N, const int32_t* __restrict__ lex, const int32_t* __restrict__ dofs,
                                      const double* __restrict__ u, double dirichlet, const double* __restrict__ fq,
                                      double* __restrict__ sums) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= g.nc) return;
  const int m = nb * nb * nb;
  int panel;
  double a0, b0, ha, g0, hg;
  shell_cell_box(g, c, &panel, &a0, &b0, &ha, &g0, &hg);
  const int Q = nq * nq * nq;
  double acc = 0.0;
  for (int q2 = 0; q2 < nq; ++q2) {
    const double tb = tan(b0 + ha * x1[q2]);
    for (int q1 = 0; q1 < nq; ++q1) {
      const Metric2 mt = csphere_metric_terms(tan(a0 + ha * x1[q1]), tb, g.radius);
      for (int q0 = 0; q0 < nq; ++q0) {
        const double wv = w1[q0] * w1[q1] * w1[q2] * hg * ha * ha;
        if (MODE == 2) { acc += wv; continue; }
        double uh = 0.0;
        for (int b = 0; b < m; ++b) {
          const int lb = lex[b];
          const int32_t d = dofs[c * m + b];
          uh = fma(N[q0 * nb + lb % nb] * N[q1 * nb + (lb / nb) % nb] * N[q2 * nb + lb / (nb * nb)], d >= 0 ? u[d] : dirichlet, uh);
        }
        if (MODE == 1) { acc = fma(wv, uh, acc); continue; }
        const double s = 1.0 + g.t * (g0 + hg * x1[q0]) / g.radius;
        const double e = uh - (fq ? fq[c * Q + q0 + nq * (q1 + nq * q2)] : 0.0);
        acc = fma(wv * g.t * s * s * mt.sqrtg, e * e, acc);
      }
    }
  }
  sums[c] = acc;
}

static double reduce_cells(gg_ctx* ctx, const double* sums, int64_t nc) {
  cudaStream_t st = ctx->stream;
  DevBuf<double> total(1);
  size_t tb = 0;
  GG_CUDA(cub::DeviceReduce::Sum(nullptr, tb, sums, total.p, (int)nc, st));
  DevBuf<uint8_t> tmp(tb);
  GG_CUDA(cub::DeviceReduce::Sum(tmp.p, tb, sums, total.p, (int)nc, st));
  ctx->launches += 2;
  double out = 0.0;
  total.download(&out, st);
  return out;
}

static void require_shell(gg_mesh* m) {
  GG_REQUIRE(m->dim == 3 && m->n > 0 && m->d_col.n == (size_t)m->n_cells, GG_ERR_UNSUPPORTED,
             "this operation needs a cubed-sphere shell built by gg_mesh_cubed_sphere (columns / layers known)");
}

void hex_points(gg_mesh* m, int order_or_neg, int nq, double* chart_host, double* xyz_host) {
  require_shell(m);
  gg_ctx* ctx = m->ctx;
  HexRef R;
  hex_ref(ctx, order_or_neg < 0 ? 1 : order_or_neg, order_or_neg < 0 ? nq : 0, R);
  const int np = R.np;
  const int64_t tot = m->n_cells * np * np * np;
  if (tot == 0) return;
  DevBuf<double> dc(chart_host ? 3 * tot : 0), dx(xyz_host ? 3 * tot : 0);
  k_hex_points<<<nblk(tot, 256), 256, 0, ctx->stream>>>(shell_geo(m), np, R.pts.p, chart_host ? dc.p : nullptr, xyz_host ? dx.p : nullptr);
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
  if (chart_host) dc.download(chart_host, ctx->stream);
  if (xyz_host) dx.download(xyz_host, ctx->stream);
  GG_CUDA(cudaStreamSynchronize(ctx->stream));
}

void hex_source_vectors(gg_space* s, int nq, const double* fq, double* vbuf) {
  gg_mesh* m = s->mesh;
  require_shell(m);
  gg_ctx* ctx = m->ctx;
  HexRef R;
  hex_ref(ctx, s->order, nq, R);
  k_hex_source<<<nblk(m->n_cells * s->ndofs_loc, 128), 128, 0, ctx->stream>>>(shell_geo(m), nq, R.nb, R.pts.p, R.w.p, R.N.p, R.lex.p, fq, vbuf);
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
  GG_CUDA(cudaStreamSynchronize(ctx->stream));   // R goes out of scope
}

void hex_boundary_flux_vectors(gg_space* s, int nq, const double* u, double dirichlet, double* vbuf) {
  gg_mesh* m = s->mesh;
  require_shell(m);
  gg_ctx* ctx = m->ctx;
  HexRef R, Rn;
  hex_ref(ctx, s->order, nq, R);
  hex_ref(ctx, s->order, 0, Rn);   // basis derivatives at the 1-D nodes (faces)
  k_hex_boundary_flux<<<nblk(m->n_cells * s->ndofs_loc, 128), 128, 0, ctx->stream>>>(shell_geo(m), nq, R.nb, R.pts.p, R.w.p, R.N.p, Rn.N.p,
                                                                                   Rn.D.p, R.lex.p, s->d_cell_dofs.p, u, dirichlet, vbuf);
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
  GG_CUDA(cudaStreamSynchronize(ctx->stream));
}

double hex_cell_functional(gg_space* s, int mode, int nq, const double* u, double dirichlet, const double* fq) {
  gg_mesh* m = s->mesh;
  require_shell(m);
  gg_ctx* ctx = m->ctx;
  const int64_t nc = m->n_cells;
  if (nc == 0) return 0.0;
  HexRef R;
  hex_ref(ctx, s->order, nq, R);
  DevBuf<double> sums(nc);
  ShellGeo g = shell_geo(m);
  const unsigned nbk = nblk(nc, 128);
#define GG_ARGS g, nq, R.nb, R.pts.p, R.w.p, R.N.p, R.lex.p, s->d_cell_dofs.p, u, dirichlet, fq, sums.p
  if (mode == 0) k_hex_cell_functional<0><<<nbk, 128, 0, ctx->stream>>>(GG_ARGS);
  else if (mode == 1) k_hex_cell_functional<1><<<nbk, 128, 0, ctx->stream>>>(GG_ARGS);
  else k_hex_cell_functional<2><<<nbk, 128, 0, ctx->stream>>>(GG_ARGS);
#undef GG_ARGS
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
  return reduce_cells(ctx, sums.p, nc);
}

}  // namespace gg
