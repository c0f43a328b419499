// DG upwind advection of a scalar on the cubed sphere: volume + skeleton terms and explicit RK stepping, on the device.
//
// Replaces, for Q = DG-Q_p, the weak form of test/Tutorials/AdvectionUpwinding.jl:107-137 evaluated by Gridap on
// Triangulation(model) and SkeletonTriangulation(model) (facet geometry: src/Geometry/AtlasDiscreteModels.jl:346-399, normals
// and facet measures: src/CellData/NormalsAreaForms.jl, src/CellData/CellFields.jl:40-78):
//     res(u,v) = int -(u (grad(v).beta)) meas dOmega + int my_mean((beta u).n_L) jump(v) meas_skel.plus dL
//                + int 0.5 |(beta.n_L).plus| jump(u) jump(v) meas_skel.plus dL,         mass = int dtu v meas dOmega
// with beta = pinvJ(J).beta_x the contravariant wind, `plus` the lower-numbered cell of a facet, n_L.plus / n_L.minus the outward
// normals of the two cells in their OWN charts and dL the chart length of the facet.  BASELINE config 3 marches it with an
// explicit RK scheme (the tutorial itself uses SDIRK): M k = -res(u), M block diagonal.
//
// B200 design: owner-computes, one thread per cell, no atomics and no facet loop: a cell integrates its volume term and its
// four facets itself, reading the neighbour's DOFs across each facet (both cells evaluate the same numerical flux from the
// same operands, so the scheme stays conservative to round-off), applies the inverse of its own mass block and writes the RK
// slope directly.  Everything that does not depend on u is tabulated once at set-up: contravariant wind at the cell points,
// beta.n and sqrtg at the facet points (per cell side), the basis at cell and facet points.
#include <cstring>
#include "gg_rk.cuh"
#include "gg_refel.cuh"
#include "gg_geom.cuh"
#include "gg_dist_host.cuh"

namespace gg {

static inline unsigned nblk(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

static const int QEDGE_V[4][2] = {{0, 1}, {2, 3}, {0, 2}, {1, 3}};
static const double QVERT[4][2] = {{0, 0}, {1, 0}, {0, 1}, {1, 1}};
static const double QNORM[4][2] = {{0, -1}, {0, 1}, {-1, 0}, {1, 0}};

// ambient points of the facet quadrature of every cell side: out[((c*4+e)*nqf+q)*3 + k]
__global__ void k_skeleton_points(int64_t nc, int nqf, const double* __restrict__ s, const double* __restrict__ x0,
                                  const double* __restrict__ y0, const double* __restrict__ hx, const double* __restrict__ hy,
                                  const int32_t* __restrict__ chart, double radius, double* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * 4 * nqf) return;
  const int q = (int)(t % nqf), e = (int)((t / nqf) % 4);
  const int64_t c = t / (4 * nqf);
  const double va[4][2] = {{0, 0}, {0, 1}, {0, 0}, {1, 0}}, vb[4][2] = {{1, 0}, {1, 1}, {0, 1}, {1, 1}};
  const double xi = va[e][0] + s[q] * (vb[e][0] - va[e][0]), eta = va[e][1] + s[q] * (vb[e][1] - va[e][1]);
  csphere_point(chart[c], tan(fma(hx[c], xi, x0[c])), tan(fma(hy[c], eta, y0[c])), radius, &out[3 * t]);
}

// contravariant wind pinvJ(J).beta = ginv (J beta) at the cell points: bq[(q*2+i)*nc + c], times w hx hy sqrtg
__global__ void k_adv_cell_wind(int64_t nc, int nq, const double* __restrict__ xi, const double* __restrict__ w,
                                const double* __restrict__ x0, const double* __restrict__ y0, const double* __restrict__ hx,
                                const double* __restrict__ hy, const int32_t* __restrict__ chart, double radius,
                                const double* __restrict__ beta, double* __restrict__ bq) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int Q = nq * nq;
  if (t >= nc * Q) return;
  const int64_t c = t / Q;
  const int q = (int)(t - c * Q), q1 = q % nq, q2 = q / nq;
  const double a = tan(fma(hx[c], xi[q1], x0[c])), b = tan(fma(hy[c], xi[q2], y0[c]));
  double J[2][3];
  csphere_jacobian(chart[c], a, b, radius, J);
  const Metric2 m = csphere_metric_terms(a, b, radius);
  const double* bv = beta + 3 * t;
  const double j0 = J[0][0] * bv[0] + J[0][1] * bv[1] + J[0][2] * bv[2], j1 = J[1][0] * bv[0] + J[1][1] * bv[1] + J[1][2] * bv[2];
  const double wq = w[q1] * w[q2] * hx[c] * hy[c] * m.sqrtg;
  // - u (grad(v).beta) meas dOmega: the reference gradient components are divided by hx, hy here
  bq[(int64_t)(q * 2) * nc + c] = wq * (m.i11 * j0 + m.i12 * j1) / hx[c];
  bq[(int64_t)(q * 2 + 1) * nc + c] = wq * (m.i12 * j0 + m.i22 * j1) / hy[c];
}

// beta.n (outward normal of this cell, its own chart) and sqrtg * w * chart length at the facet points of every cell side
__global__ void k_adv_facet_wind(int64_t nc, int nqf, const double* __restrict__ s, const double* __restrict__ wf,
                                 const double* __restrict__ x0, const double* __restrict__ y0, const double* __restrict__ hx,
                                 const double* __restrict__ hy, const int32_t* __restrict__ chart, double radius,
                                 const double* __restrict__ beta, double* __restrict__ bn, double* __restrict__ wsg) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * 4 * nqf) return;
  const int q = (int)(t % nqf), e = (int)((t / nqf) % 4);
  const int64_t c = t / (4 * nqf);
  const double va[4][2] = {{0, 0}, {0, 1}, {0, 0}, {1, 0}}, vb[4][2] = {{1, 0}, {1, 1}, {0, 1}, {1, 1}};
  const double nn[4][2] = {{0, -1}, {0, 1}, {-1, 0}, {1, 0}};
  const double xi = va[e][0] + s[q] * (vb[e][0] - va[e][0]), eta = va[e][1] + s[q] * (vb[e][1] - va[e][1]);
  const double a = tan(fma(hx[c], xi, x0[c])), b = tan(fma(hy[c], eta, y0[c]));
  double J[2][3];
  csphere_jacobian(chart[c], a, b, radius, J);
  const Metric2 m = csphere_metric_terms(a, b, radius);
  const double* bv = beta + 3 * t;
  const double j0 = J[0][0] * bv[0] + J[0][1] * bv[1] + J[0][2] * bv[2], j1 = J[1][0] * bv[0] + J[1][1] * bv[1] + J[1][2] * bv[2];
  const double bx = m.i11 * j0 + m.i12 * j1, by = m.i12 * j0 + m.i22 * j1;
  const double len = e < 2 ? hx[c] : hy[c];
  const int64_t o = (int64_t)(e * nqf + q) * nc + c;   // [edge][point][cell]
  bn[o] = bx * nn[e][0] + by * nn[e][1];
  wsg[o] = wf[q] * len * m.sqrtg;
}

struct AdvTabs {
  const double* val;   // [Q][M] basis at the cell points
  const double* der;   // [Q][M][2] reference gradients
  const double* fval;  // [4][nqf][M] basis at the facet points of the four local edges
};

// slope k = -M^-1 res(u) (kout) and / or the residual itself (rout), one thread per cell
template <int P>
__global__ void __launch_bounds__(128)
k_adv_slope(int64_t nc, int nq, int nqf, AdvTabs T, const double* __restrict__ bq, const double* __restrict__ bn,
            const double* __restrict__ wsg, const int32_t* __restrict__ nbr_cell, const int8_t* __restrict__ nbr_edge,
            const int8_t* __restrict__ is_plus, const double* __restrict__ MInv, const uint8_t* __restrict__ cell_owned,
            const double* __restrict__ u, double* __restrict__ kout, double* __restrict__ rout) {
  constexpr int M = (P + 1) * (P + 1);
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  double ue[M], r[M];
  if (cell_owned && !cell_owned[c]) {
    // ghost cell of a partitioned model: its slope belongs to another rank (the stage states are made consistent instead)
#pragma unroll
    for (int j = 0; j < M; ++j) {
      if (rout) rout[c * M + j] = 0.0;
      if (kout) kout[c * M + j] = 0.0;
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < M; ++j) { ue[j] = u[c * M + j]; r[j] = 0.0; }
  const int Q = nq * nq;
  // volume: - sum_q w dV sqrtg u_h (grad(phi_I).beta)
#pragma unroll 1
  for (int q = 0; q < Q; ++q) {
    const double* __restrict__ v = T.val + q * M;
    const double* __restrict__ d = T.der + q * M * 2;
    double uh = 0.0;
#pragma unroll
    for (int j = 0; j < M; ++j) uh = fma(v[j], ue[j], uh);
    const double fx = -uh * bq[(int64_t)(q * 2) * nc + c], fy = -uh * bq[(int64_t)(q * 2 + 1) * nc + c];
#pragma unroll
    for (int j = 0; j < M; ++j) r[j] = fma(fx, d[2 * j], fma(fy, d[2 * j + 1], r[j]));
  }
  // skeleton: this cell's share of its four facets
#pragma unroll 1
  for (int e = 0; e < 4; ++e) {
    const int64_t cn = nbr_cell[c * 4 + e];
    const int en = nbr_edge[c * 4 + e];
    const bool plus = is_plus[c * 4 + e];
    double un[M];
#pragma unroll
    for (int j = 0; j < M; ++j) un[j] = u[cn * M + j];
#pragma unroll 1
    for (int q = 0; q < nqf; ++q) {
      const double* __restrict__ vh = T.fval + (e * nqf + q) * M;
      const double* __restrict__ vn = T.fval + (en * nqf + q) * M;
      double uh = 0.0, ut = 0.0;
#pragma unroll
      for (int j = 0; j < M; ++j) { uh = fma(vh[j], ue[j], uh); ut = fma(vn[j], un[j], ut); }
      const int64_t oh = (int64_t)(e * nqf + q) * nc + c, on = (int64_t)(en * nqf + q) * nc + cn;
      const double bh = bn[oh], bt = bn[on];
      // plus / minus operands of the facet; the test function of this cell enters jump(v) with sign +1 (plus) or -1 (minus)
      const double up = plus ? uh : ut, um = plus ? ut : uh, bp = plus ? bh : bt, bm = plus ? bt : bh;
      const double flux = 0.5 * (bp * up - bm * um) + 0.5 * fabs(bp) * (up - um);
      const double f = (plus ? 1.0 : -1.0) * (plus ? wsg[oh] : wsg[on]) * flux;
#pragma unroll
      for (int j = 0; j < M; ++j) r[j] = fma(f, vh[j], r[j]);
    }
  }
  if (rout) {
#pragma unroll
    for (int j = 0; j < M; ++j) rout[c * M + j] = r[j];
  }
  if (kout) {
#pragma unroll
    for (int i = 0; i < M; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < M; ++j) acc = fma(MInv[(int64_t)(i * M + j) * nc + c], r[j], acc);
      kout[c * M + i] = -acc;
    }
  }
}

void adv_slope_async(gg_rk* rk, const double* u, double* kout, double* rout) {
  gg_ctx* ctx = rk->ctx;
  const int64_t nc = rk->mesh->n_cells;
  if (nc == 0) return;
  auto tab = get_tab(ctx, GG_SPACE_DG, rk->order, rk->nq1);
  AdvTabs T{tab->val.p, tab->der.p, rk->adv_fval.p};
  unsigned nb = nblk(nc, 128);
  ProfScope ps(ctx, "k_adv_slope");
#define GG_ARGS nc, rk->nq1, rk->nq1, T, rk->adv_bq.p, rk->adv_bn.p, rk->adv_wsg.p, rk->adv_nbr_cell.p, rk->adv_nbr_edge.p, \
                rk->adv_is_plus.p, rk->MpInv.p, rk->mesh->d_cell_owned.n ? rk->mesh->d_cell_owned.p : nullptr, u, kout, rout
  switch (rk->order) {
    case 0: k_adv_slope<0><<<nb, 128, 0, ctx->stream>>>(GG_ARGS); break;
    case 1: k_adv_slope<1><<<nb, 128, 0, ctx->stream>>>(GG_ARGS); break;
    case 2: k_adv_slope<2><<<nb, 128, 0, ctx->stream>>>(GG_ARGS); break;
    default: throw Error(GG_ERR_UNSUPPORTED, "DG advection order must be <= 2");
  }
#undef GG_ARGS
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
}

void adv_step_async(gg_rk* rk) {
  const int s = rk->n_stages;
  for (int i = 0; i < s; ++i) {
    const double* xs = rk->x.p;
    if (i > 0) {
      double co[4] = {0, 0, 0, 0};
      for (int j = 0; j < i && j < 4; ++j) co[j] = rk->dt * rk->A[i * s + j];
      rk_combine(rk, std::min(i, 4), co, rk->xi.p);
      vec_consistent_async(rk->Q, rk->xi.p);   // partitioned models: owner -> ghost cells, through peer memory
      xs = rk->xi.p;
    }
    adv_slope_async(rk, xs, rk->k[i].p, nullptr);
  }
  double co[4] = {0, 0, 0, 0};
  for (int j = 0; j < s && j < 4; ++j) co[j] = rk->dt * rk->b[j];
  rk_combine(rk, std::min(s, 4), co, rk->x.p);
  vec_consistent_async(rk->Q, rk->x.p);
}

// Facet connectivity and DG traces shared by the skeleton terms of the DG advection and thermal shallow-water steppers:
// every edge of the closed sphere has two cells, plus = the lower-numbered one; basis at the facet points of the 4 local edges.
void facet_tables(gg_rk* rk, int nq) {
  gg_ctx* ctx = rk->ctx;
  gg_mesh* m = rk->mesh;
  cudaStream_t st = ctx->stream;
  const int64_t nc = m->n_cells;
  const int M = rk->mq;
  gg_mesh_info(m, nullptr, nullptr, nullptr);
  int64_t ne = 0;
  gg_mesh_info(m, nullptr, nullptr, &ne);   // builds the edge topology
  std::vector<int32_t> nbr_cell(nc * 4);
  std::vector<int8_t> nbr_edge(nc * 4), is_plus(nc * 4);
  for (int64_t c = 0; c < nc; ++c)
    for (int e = 0; e < 4; ++e) {
      const int32_t ed = m->cell_edges[c * 4 + e];
      GG_REQUIRE(!m->cell_edge_flip[c * 4 + e], GG_ERR_UNSUPPORTED, "skeleton terms need shared edges with the same vertex order in both cells");
      const int32_t c0 = m->edge_cells[ed * 2], c1 = m->edge_cells[ed * 2 + 1];
      const bool owned = m->cell_owned.empty() || m->cell_owned[c];
      if (c1 < 0 && !owned) {
        // outer edge of the ghost layer of a partitioned model: never integrated (ghost cells are skipped)
        nbr_cell[c * 4 + e] = (int32_t)c; nbr_edge[c * 4 + e] = (int8_t)e; is_plus[c * 4 + e] = 1;
        continue;
      }
      GG_REQUIRE(c1 >= 0, GG_ERR_UNSUPPORTED, "boundary facets are not supported (the cubed sphere is closed)");
      const int32_t other = c0 == (int32_t)c ? c1 : c0;
      int eo = -1;
      for (int k = 0; k < 4; ++k) if (m->cell_edges[other * 4 + k] == ed) eo = k;
      nbr_cell[c * 4 + e] = other; nbr_edge[c * 4 + e] = (int8_t)eo; is_plus[c * 4 + e] = c0 == (int32_t)c;
    }
  rk->adv_nbr_cell.upload(nbr_cell, st); rk->adv_nbr_edge.upload(nbr_edge, st); rk->adv_is_plus.upload(is_plus, st);
  // basis at the facet points of the four local edges
  std::vector<double> s(nq), wf(nq);
  gauss_legendre_01(nq, s.data(), wf.data());
  const int nb1 = rk->order + 1;
  std::vector<int32_t> l2x = quad_local_to_lex(rk->order);
  std::vector<double> fval((size_t)4 * nq * M), N1(nb1), D1(nb1), N2(nb1), D2(nb1);
  for (int e = 0; e < 4; ++e)
    for (int q = 0; q < nq; ++q) {
      const double xi = QVERT[QEDGE_V[e][0]][0] + s[q] * (QVERT[QEDGE_V[e][1]][0] - QVERT[QEDGE_V[e][0]][0]);
      const double eta = QVERT[QEDGE_V[e][0]][1] + s[q] * (QVERT[QEDGE_V[e][1]][1] - QVERT[QEDGE_V[e][0]][1]);
      lagrange_1d(rk->order, xi, N1.data(), D1.data());
      lagrange_1d(rk->order, eta, N2.data(), D2.data());
      for (int j = 0; j < M; ++j) fval[((size_t)e * nq + q) * M + j] = N1[l2x[j] % nb1] * N2[l2x[j] / nb1];
    }
  (void)QNORM;
  rk->adv_fval.upload(fval, st);
}

// wind tabulations: contravariant wind times the cell weights, beta.n and the facet weights per cell side
static void adv_wind_tables(gg_mesh* m, int order, int nq, const double* velocity_cells, const double* velocity_facets,
                            DevBuf<double>& bq, DevBuf<double>& bn, DevBuf<double>& wsg) {
  gg_ctx* ctx = m->ctx;
  cudaStream_t st = ctx->stream;
  const int64_t nc = m->n_cells;
  const int Q = nq * nq;
  std::vector<double> s(nq), wf(nq);
  gauss_legendre_01(nq, s.data(), wf.data());
  auto tab = get_tab(ctx, GG_SPACE_DG, order, nq);
  DevBuf<double> dbc, dbf, ds, dw;
  dbc.upload(velocity_cells, (size_t)nc * Q * 3, st);
  dbf.upload(velocity_facets, (size_t)nc * 4 * nq * 3, st);
  ds.upload(s, st); dw.upload(wf, st);
  bq.alloc((size_t)2 * Q * nc); bn.alloc((size_t)4 * nq * nc); wsg.alloc((size_t)4 * nq * nc);
  if (nc) {
    k_adv_cell_wind<<<nblk(nc * Q, 128), 128, 0, st>>>(nc, nq, tab->xi.p, tab->w.p, m->d_x0.p, m->d_y0.p, m->d_hx.p, m->d_hy.p,
                                                      m->d_chart.p, m->radius, dbc.p, bq.p);
    k_adv_facet_wind<<<nblk(nc * 4 * nq, 128), 128, 0, st>>>(nc, nq, ds.p, dw.p, m->d_x0.p, m->d_y0.p, m->d_hx.p, m->d_hy.p,
                                                            m->d_chart.p, m->radius, dbf.p, bn.p, wsg.p);
    ctx->launches += 2;
    GG_CUDA(cudaGetLastError());
  }
  GG_CUDA(cudaStreamSynchronize(st));
}

void adv_create(gg_rk* rk, const gg_rk_args* a) {
  gg_ctx* ctx = rk->ctx;
  gg_mesh* m = rk->mesh;
  cudaStream_t st = ctx->stream;
  GG_REQUIRE(a->velocity_cells && a->velocity_facets, GG_ERR_INVALID,
             "DG advection needs the ambient wind at the cell and facet quadrature points");
  const int64_t nc = m->n_cells;
  const int nq = num_points_1d(a->degree), Q = nq * nq, M = rk->mq;
  rk->nq1 = nq;
  GG_REQUIRE(nq <= MAX_NQ, GG_ERR_UNSUPPORTED, "quadrature degree too high");
  facet_tables(rk, nq);
  std::vector<double> s(nq), wf(nq);
  gauss_legendre_01(nq, s.data(), wf.data());
  adv_wind_tables(m, rk->order, nq, a->velocity_cells, a->velocity_facets, rk->adv_bq, rk->adv_bn, rk->adv_wsg);
  (void)Q; (void)s; (void)wf; (void)st; (void)nc;
}

// ---- assembled operator (the tutorial marches implicitly: jac = a + b + s, test/Tutorials/AdvectionUpwinding.jl:124-131) ----
// DG pattern with facet coupling: row (c, i) holds the 5 blocks of c and its four facet neighbours, ordered by cell id;
// every row has exactly 5 M entries, so rowptr is a multiplication and a slot is (rank of the block) * M + j.
__device__ __forceinline__ void adv_block_ranks(int64_t c, const int32_t* __restrict__ nb, int rank_own_and_nbr[5]) {
  const int64_t ids[5] = {c, nb[0], nb[1], nb[2], nb[3]};
#pragma unroll
  for (int a = 0; a < 5; ++a) {
    int r = 0;
#pragma unroll
    for (int b = 0; b < 5; ++b) r += ids[b] < ids[a];
    rank_own_and_nbr[a] = r;
  }
}

__global__ void k_adv_pattern(int64_t nc, int M, const int32_t* __restrict__ nbr_cell, int32_t* __restrict__ rowptr,
                              int32_t* __restrict__ colind) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > nc * M) return;
  rowptr[t] = (int32_t)(t * 5 * M);
  if (t == nc * M) return;
  const int64_t c = t / M;
  int rk[5];
  adv_block_ranks(c, nbr_cell + c * 4, rk);
  const int64_t ids[5] = {c, nbr_cell[c * 4], nbr_cell[c * 4 + 1], nbr_cell[c * 4 + 2], nbr_cell[c * 4 + 3]};
  for (int a = 0; a < 5; ++a)
    for (int j = 0; j < M; ++j) colind[t * 5 * M + rk[a] * M + j] = (int32_t)(ids[a] * M + j);
}

// one thread per (cell, test function): row of the volume block and of the four facet couplings
template <int P>
__global__ void __launch_bounds__(128)
k_adv_matrix(int64_t nc, int nq, AdvTabs T, const double* __restrict__ bq, const double* __restrict__ bn,
             const double* __restrict__ wsg, const int32_t* __restrict__ nbr_cell, const int8_t* __restrict__ nbr_edge,
             const int8_t* __restrict__ is_plus, double* __restrict__ vals) {
  constexpr int M = (P + 1) * (P + 1);
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * M) return;
  const int64_t c = t / M;
  const int i = (int)(t - c * M);
  double own[M], nbr[4][M];
#pragma unroll
  for (int j = 0; j < M; ++j) { own[j] = 0.0; nbr[0][j] = 0.0; nbr[1][j] = 0.0; nbr[2][j] = 0.0; nbr[3][j] = 0.0; }
  const int Q = nq * nq;
#pragma unroll 1
  for (int q = 0; q < Q; ++q) {
    const double* __restrict__ v = T.val + q * M;
    const double* __restrict__ d = T.der + q * M * 2;
    const double g = -(bq[(int64_t)(q * 2) * nc + c] * d[2 * i] + bq[(int64_t)(q * 2 + 1) * nc + c] * d[2 * i + 1]);
#pragma unroll
    for (int j = 0; j < M; ++j) own[j] = fma(g, v[j], own[j]);
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int64_t cn = nbr_cell[c * 4 + e];
    const int en = nbr_edge[c * 4 + e];
    const bool plus = is_plus[c * 4 + e];
#pragma unroll 1
    for (int q = 0; q < nq; ++q) {
      const double* __restrict__ vh = T.fval + (e * nq + q) * M;
      const double* __restrict__ vn = T.fval + (en * nq + q) * M;
      const int64_t oh = (int64_t)(e * nq + q) * nc + c, on = (int64_t)(en * nq + q) * nc + cn;
      const double bh = bn[oh], bt = bn[on];
      const double bp = plus ? bh : bt, bm = plus ? bt : bh;
      const double cp = 0.5 * bp + 0.5 * fabs(bp), cm = -0.5 * bm - 0.5 * fabs(bp);      // d flux / d u_plus, d flux / d u_minus
      const double f = (plus ? wsg[oh] : -wsg[on]) * vh[i];
      const double co = f * (plus ? cp : cm), cnb = f * (plus ? cm : cp);
#pragma unroll
      for (int j = 0; j < M; ++j) { own[j] = fma(co, vh[j], own[j]); nbr[e][j] = fma(cnb, vn[j], nbr[e][j]); }
    }
  }
  int rk[5];
  adv_block_ranks(c, nbr_cell + c * 4, rk);
  double* row = vals + t * 5 * M;
#pragma unroll
  for (int j = 0; j < M; ++j) {
    row[rk[0] * M + j] = own[j];
#pragma unroll
    for (int e = 0; e < 4; ++e) row[rk[1 + e] * M + j] = nbr[e][j];
  }
}

}  // namespace gg

using namespace gg;

extern "C" {

int gg_pattern_skeleton(gg_space* Q, gg_csr** out) {
  GG_API_BEGIN
  GG_REQUIRE(Q && out, GG_ERR_INVALID, "null argument");
  gg_mesh* m = Q->mesh;
  GG_REQUIRE(Q->kind == GG_SPACE_DG && m->dim == 2 && Q->constraint == GG_CONSTRAINT_NONE && !m->plan && Q->order <= 2, GG_ERR_UNSUPPORTED,
             "skeleton patterns are implemented for unconstrained DG-Q0..Q2 spaces on the (unpartitioned) 2-D surface");
  gg_ctx* ctx = m->ctx;
  GG_CUDA(cudaSetDevice(ctx->device));
  const int64_t nc = m->n_cells;
  const int M = Q->ndofs_loc;
  GG_REQUIRE(nc * M * 5 * M < (int64_t)INT32_MAX, GG_ERR_UNSUPPORTED, "skeleton pattern too large for Int32 offsets");
  gg_rk tmp;   // only to reuse facet_tables()
  tmp.ctx = ctx; tmp.mesh = m; tmp.order = Q->order; tmp.mq = M;
  facet_tables(&tmp, 1);
  std::vector<int32_t> nb(nc * 4);
  tmp.adv_nbr_cell.download(nb.data(), ctx->stream);
  for (int64_t c = 0; c < nc; ++c)
    for (int a = 0; a < 4; ++a) {
      GG_REQUIRE(nb[c * 4 + a] != c, GG_ERR_UNSUPPORTED, "a cell is its own facet neighbour");
      for (int b = a + 1; b < 4; ++b) GG_REQUIRE(nb[c * 4 + a] != nb[c * 4 + b], GG_ERR_UNSUPPORTED, "two facets of a cell lead to the same neighbour");
    }
  std::unique_ptr<gg_csr> A(new gg_csr);
  A->ctx = ctx; A->test = Q; A->trial = Q; A->skeleton = true;
  A->n_rows = A->n_cols = Q->n_free; A->nnz = nc * M * 5 * M;
  A->d_rowptr.alloc(A->n_rows + 1); A->d_colind.alloc(A->nnz); A->d_vals.alloc(A->nnz);
  if (nc) {
    k_adv_pattern<<<nblk(nc * M + 1, 128), 128, 0, ctx->stream>>>(nc, M, tmp.adv_nbr_cell.p, A->d_rowptr.p, A->d_colind.p);
    ctx->launches++;
    GG_CUDA(cudaGetLastError());
    GG_CUDA(cudaMemsetAsync(A->d_vals.p, 0, A->nnz * sizeof(double), ctx->stream));
  } else {
    GG_CUDA(cudaMemsetAsync(A->d_rowptr.p, 0, sizeof(int32_t), ctx->stream));
  }
  GG_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = A.release();
  GG_API_END
}

int gg_assemble_advection(gg_space* Q, int32_t degree, const double* velocity_cells, const double* velocity_facets, gg_csr* A) {
  GG_API_BEGIN
  GG_REQUIRE(Q && velocity_cells && velocity_facets && A, GG_ERR_INVALID, "null argument");
  GG_REQUIRE(A->skeleton && A->test == Q, GG_ERR_INVALID, "the matrix must come from gg_pattern_skeleton on the same space");
  gg_mesh* m = Q->mesh;
  gg_ctx* ctx = m->ctx;
  GG_CUDA(cudaSetDevice(ctx->device));
  const int nq = num_points_1d(degree);
  GG_REQUIRE(degree >= 0 && nq <= MAX_NQ, GG_ERR_UNSUPPORTED, "quadrature degree out of range");
  const int64_t nc = m->n_cells;
  if (nc == 0) return GG_OK;
  gg_rk tmp;
  tmp.ctx = ctx; tmp.mesh = m; tmp.order = Q->order; tmp.mq = Q->ndofs_loc;
  facet_tables(&tmp, nq);
  adv_wind_tables(m, Q->order, nq, velocity_cells, velocity_facets, tmp.adv_bq, tmp.adv_bn, tmp.adv_wsg);
  auto tab = get_tab(ctx, GG_SPACE_DG, Q->order, nq);
  AdvTabs T{tab->val.p, tab->der.p, tmp.adv_fval.p};
  const unsigned nb = nblk(nc * Q->ndofs_loc, 128);
  ProfScope ps(ctx, "k_adv_matrix");
#define GG_ARGS nc, nq, T, tmp.adv_bq.p, tmp.adv_bn.p, tmp.adv_wsg.p, tmp.adv_nbr_cell.p, tmp.adv_nbr_edge.p, tmp.adv_is_plus.p, A->d_vals.p
  switch (Q->order) {
    case 0: k_adv_matrix<0><<<nb, 128, 0, ctx->stream>>>(GG_ARGS); break;
    case 1: k_adv_matrix<1><<<nb, 128, 0, ctx->stream>>>(GG_ARGS); break;
    case 2: k_adv_matrix<2><<<nb, 128, 0, ctx->stream>>>(GG_ARGS); break;
    default: throw Error(GG_ERR_UNSUPPORTED, "DG advection order must be <= 2");
  }
#undef GG_ARGS
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
  GG_CUDA(cudaStreamSynchronize(ctx->stream));
  GG_API_END
}

int gg_skeleton_points(gg_mesh* m, int32_t degree, int32_t* n_pts, double* xyz) {
  GG_API_BEGIN
  GG_REQUIRE(m, GG_ERR_INVALID, "null mesh");
  GG_REQUIRE(m->dim == 2, GG_ERR_UNSUPPORTED, "skeleton points are implemented on the 2-D surface");
  gg_ctx* ctx = m->ctx;
  GG_CUDA(cudaSetDevice(ctx->device));
  const int nq = num_points_1d(degree);
  GG_REQUIRE(degree >= 0 && nq <= MAX_NQ, GG_ERR_UNSUPPORTED, "quadrature degree out of range");
  if (n_pts) *n_pts = nq;
  const int64_t tot = m->n_cells * 4 * nq;
  if (!xyz || tot == 0) return GG_OK;
  std::vector<double> s(nq), w(nq);
  gauss_legendre_01(nq, s.data(), w.data());
  DevBuf<double> ds, out(3 * tot);
  ds.upload(s, ctx->stream);
  k_skeleton_points<<<nblk(tot, 256), 256, 0, ctx->stream>>>(m->n_cells, nq, ds.p, m->d_x0.p, m->d_y0.p, m->d_hx.p, m->d_hy.p,
                                                             m->d_chart.p, m->radius, out.p);
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
  out.download(xyz, ctx->stream);
  GG_API_END
}

}  // extern "C"
