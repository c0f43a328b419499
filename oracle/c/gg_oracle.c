/* CPU ORACLE (C restatement) -- TEST / BASELINE INFRASTRUCTURE ONLY, never linked into the product.
 *
 * Restates, cell by cell and quadrature point by quadrature point, what the reference does for the
 * Laplace-Beltrami path (test/Laplacian/LaplaceBeltrami.jl:47):
 *     Ke[i][j] += w_q |det grad(cell_map)| grad(phi_i) . (ginv(x_q) . grad(phi_j)) sqrtg(x_q)
 * with the analytic fields of src/Fields/CubedSphereInvMetric.jl:1-7 (ginv), src/Fields/CubedSphereMetric.jl:16-22
 * (g) and src/CellData/CellFields.jl:80-82 (sqrtg = sqrt(det(g))), each of which re-evaluates tan() as the
 * reference's separate Fields do ("recomputed by every field that needs it"); then
 * Gridap's IntegrationMap loop order (q outermost) and a COO -> CSR scatter-add in cell order.
 * PARITY: checked against the numpy oracle (tests/test_oracle_c.py), which is pinned to the reference's
 * known-answer tests.  It is deliberately the plain per-cell algorithm, not the GPU's factorised one.
 *
 * Used as `cpu_baseline` (kind "port") and by `bench.py --impl reference`: OpenMP over cells stands in
 * for the reference's MPI ranks (linear cell partition, src/Distributed/AtlasDistributedDiscreteModels.jl:31).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static void inv_metric(double r, double x, double y, double* h) { /* CubedSphereInvMetric.jl:1-7 */
  double a = tan(x), b = tan(y);
  double sa = 1 + a * a, sb = 1 + b * b, rho2 = 1 + a * a + b * b;
  double c = rho2 / (r * r * sa * sb);
  h[0] = c * sb; h[1] = c * a * b; h[2] = c * sa;
}

static void metric(double r, double x, double y, double* g) { /* CubedSphereMetric.jl:16-22 */
  double a = tan(x), b = tan(y);
  double sa = 1 + a * a, sb = 1 + b * b, rho4 = (1 + a * a + b * b) * (1 + a * a + b * b);
  double c = r * r * sa * sb / rho4;
  g[0] = c * sa; g[1] = -c * a * b; g[2] = c * sb;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Element matrices of cells [0,nc): chart_coords [nc][4][2] (BL,BR,TL,TR), pts [Q][2], w [Q],
 * gref [Q][m][2] reference gradients (Gridap local order); out Ke [nc][m][m].  nthreads <= 0: all. */
void oracle_lb_element_matrices(int64_t nc, const double* chart_coords, double radius, int Q, const double* pts,
                                const double* w, int m, const double* gref, double* Ke, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    double* grad = (double*)malloc(sizeof(double) * m * 2);
    double* flux = (double*)malloc(sizeof(double) * m * 2);
#pragma omp for schedule(static)
    for (int64_t c = 0; c < nc; ++c) {
      const double* cc = chart_coords + c * 8;
      double* K = Ke + (size_t)c * m * m;
      memset(K, 0, sizeof(double) * m * m);
      for (int q = 0; q < Q; ++q) {
        double xi = pts[2 * q], eta = pts[2 * q + 1];
        /* bilinear cell map and its gradient Jt[i][d] = d x_d / d xi_i */
        double N[4] = {(1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta};
        double dN[4][2] = {{-(1 - eta), -(1 - xi)}, {(1 - eta), -xi}, {-eta, (1 - xi)}, {eta, xi}};
        double x[2] = {0, 0}, Jt[2][2] = {{0, 0}, {0, 0}};
        for (int k = 0; k < 4; ++k)
          for (int d = 0; d < 2; ++d) {
            x[d] += N[k] * cc[2 * k + d];
            Jt[0][d] += dN[k][0] * cc[2 * k + d];
            Jt[1][d] += dN[k][1] * cc[2 * k + d];
          }
        double det = Jt[0][0] * Jt[1][1] - Jt[0][1] * Jt[1][0];
        double iJ[2][2] = {{Jt[1][1] / det, -Jt[0][1] / det}, {-Jt[1][0] / det, Jt[0][0] / det}};
        double h[3], g[3];
        inv_metric(radius, x[0], x[1], h);
        metric(radius, x[0], x[1], g);
        double meas = sqrt(g[0] * g[2] - g[1] * g[1]);
        double dV = w[q] * fabs(det);
        for (int i = 0; i < m; ++i) {
          double g0 = gref[(q * m + i) * 2], g1 = gref[(q * m + i) * 2 + 1];
          double gx = iJ[0][0] * g0 + iJ[0][1] * g1, gy = iJ[1][0] * g0 + iJ[1][1] * g1;
          grad[2 * i] = gx; grad[2 * i + 1] = gy;
          flux[2 * i] = h[0] * gx + h[1] * gy;
          flux[2 * i + 1] = h[1] * gx + h[2] * gy;
        }
        for (int i = 0; i < m; ++i)
          for (int j = 0; j < m; ++j)
            K[i * m + j] += (grad[2 * i] * flux[2 * j] + grad[2 * i + 1] * flux[2 * j + 1]) * meas * dV;
      }
    }
    free(grad);
    free(flux);
  }
}

/* Element residual vectors fe[c][i] = int grad(phi_i).(ginv.grad(u_h)) sqrtg, u_h from ue [nc][m]. */
void oracle_lb_element_vectors(int64_t nc, const double* chart_coords, double radius, int Q, const double* pts,
                               const double* w, int m, const double* gref, const double* ue, double* fe, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < nc; ++c) {
    const double* cc = chart_coords + c * 8;
    double* f = fe + (size_t)c * m;
    memset(f, 0, sizeof(double) * m);
    for (int q = 0; q < Q; ++q) {
      double xi = pts[2 * q], eta = pts[2 * q + 1];
      double N[4] = {(1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta};
      double dN[4][2] = {{-(1 - eta), -(1 - xi)}, {(1 - eta), -xi}, {-eta, (1 - xi)}, {eta, xi}};
      double x[2] = {0, 0}, Jt[2][2] = {{0, 0}, {0, 0}};
      for (int k = 0; k < 4; ++k)
        for (int d = 0; d < 2; ++d) {
          x[d] += N[k] * cc[2 * k + d];
          Jt[0][d] += dN[k][0] * cc[2 * k + d];
          Jt[1][d] += dN[k][1] * cc[2 * k + d];
        }
      double det = Jt[0][0] * Jt[1][1] - Jt[0][1] * Jt[1][0];
      double iJ[2][2] = {{Jt[1][1] / det, -Jt[0][1] / det}, {-Jt[1][0] / det, Jt[0][0] / det}};
      double h[3], g[3];
      inv_metric(radius, x[0], x[1], h);
      metric(radius, x[0], x[1], g);
      double meas = sqrt(g[0] * g[2] - g[1] * g[1]);
      double dV = w[q] * fabs(det);
      double ux = 0, uy = 0;
      for (int j = 0; j < m; ++j) {
        double g0 = gref[(q * m + j) * 2], g1 = gref[(q * m + j) * 2 + 1];
        ux += (iJ[0][0] * g0 + iJ[0][1] * g1) * ue[c * m + j];
        uy += (iJ[1][0] * g0 + iJ[1][1] * g1) * ue[c * m + j];
      }
      double fx = h[0] * ux + h[1] * uy, fy = h[1] * ux + h[2] * uy;
      for (int i = 0; i < m; ++i) {
        double g0 = gref[(q * m + i) * 2], g1 = gref[(q * m + i) * 2 + 1];
        double gx = iJ[0][0] * g0 + iJ[0][1] * g1, gy = iJ[1][0] * g0 + iJ[1][1] * g1;
        f[i] += (gx * fx + gy * fy) * meas * dV;
      }
    }
  }
}

/* COO -> CSR numeric phase: vals[slot(i,j)] += Ke[c][i][j] for all cells, rows/cols < 0 dropped.
 * The pattern (rowptr, colind sorted in each row) is given.  nthreads == 1: serial in cell order = sparse(I,J,V);
 * otherwise the cells are shared out over the threads and the adds are atomic (the summation order of a slot then
 * depends on the schedule: results agree to round-off, not bit for bit, with the serial order). */
void oracle_scatter_csr(int64_t nc, int m, const int64_t* cell_dofs, const double* Ke, const int64_t* rowptr,
                        const int64_t* colind, double* vals, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static) if (nthreads != 1)
  for (int64_t c = 0; c < nc; ++c)
    for (int i = 0; i < m; ++i) {
      int64_t r = cell_dofs[c * m + i];
      if (r < 0) continue;
      for (int j = 0; j < m; ++j) {
        int64_t col = cell_dofs[c * m + j];
        if (col < 0) continue;
        int64_t lo = rowptr[r], hi = rowptr[r + 1];
        while (lo < hi) {
          int64_t mid = (lo + hi) >> 1;
          if (colind[mid] < col) lo = mid + 1; else hi = mid;
        }
        const double v = Ke[((size_t)c * m + i) * m + j];
#pragma omp atomic
        vals[lo] += v;
      }
    }
}

void oracle_scatter_vector(int64_t nc, int m, const int64_t* cell_dofs, const double* fe, double* out, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static) if (nthreads != 1)
  for (int64_t c = 0; c < nc; ++c)
    for (int i = 0; i < m; ++i) {
      int64_t r = cell_dofs[c * m + i];
      if (r >= 0) {
        const double v = fe[c * m + i];
#pragma omp atomic
        out[r] += v;
      }
    }
}
