/* CPU ORACLE (C restatement) of the explicit-RK half of the hot path -- TEST / BASELINE INFRASTRUCTURE ONLY, never linked
 * into the product.
 *
 * Restates, cell by cell and quadrature point by quadrature point, what the reference assembles for
 *   the linear wave driver     mass / res            test/Transient/TransientWaveEquation.jl:63-66
 *   the shallow-water driver   res_y, jac_y, mass,   test/Transient/TransientShallowWater.jl:113-139
 *                              res_x
 * with the analytic metric fields of src/Fields/CubedSphereMetric.jl:16-22, CubedSphereInvMetric.jl:1-7 and
 * src/CellData/CellFields.jl:80-82 (each re-evaluating tan(), as the reference's separate Fields do), the contravariant
 * Piola map of the Raviart-Thomas basis (u = (1/det J) uhat . Jt, div u = divhat / det J, slave-facet sign flips) and
 * R v = (-v2, v1) (src/Helpers/Operators.jl:62-64).  Mass solves: Jacobi-preconditioned CG, the reference's own iterative
 * option (test/Tutorials/ShallowWater.jl:175-176, rtol 1e-12); the DG mass is block diagonal and solved cell by cell.
 * PARITY: every function is checked against the numpy oracle (oracle/rk.py, oracle/swe.py; tests/test_oracle_c.py), which is
 * pinned to the reference's regression numbers.
 *
 * OpenMP over cells stands in for the reference's MPI ranks (linear cell partition).  Used as `cpu_baseline` for the RK
 * steps/s half of the metric (bench.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static void set_threads(int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
}

typedef struct {
  double x[2], Jt[2][2], det, dV;  /* chart point, grad(cell_map)[i][d] = d x_d / d xi_i, its determinant, w |det| */
  double g[3], h[3], sqrtg;        /* metric (11,12,22), inverse metric, sqrt(det g) */
} QPoint;

static void metric_at(double r, double x, double y, double* g) { /* CubedSphereMetric.jl:16-22 */
  double a = tan(x), b = tan(y);
  double sa = 1 + a * a, sb = 1 + b * b, rho4 = (1 + a * a + b * b) * (1 + a * a + b * b);
  double c = r * r * sa * sb / rho4;
  g[0] = c * sa; g[1] = -c * a * b; g[2] = c * sb;
}

static void inv_metric_at(double r, double x, double y, double* h) { /* CubedSphereInvMetric.jl:1-7 */
  double a = tan(x), b = tan(y);
  double sa = 1 + a * a, sb = 1 + b * b, rho2 = 1 + a * a + b * b;
  double c = rho2 / (r * r * sa * sb);
  h[0] = c * sb; h[1] = c * a * b; h[2] = c * sa;
}

static void qpoint(const double* cc, double radius, double xi, double eta, double w, QPoint* P) {
  double N[4] = {(1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta};
  double dN[4][2] = {{-(1 - eta), -(1 - xi)}, {(1 - eta), -xi}, {-eta, (1 - xi)}, {eta, xi}};
  P->x[0] = P->x[1] = 0;
  P->Jt[0][0] = P->Jt[0][1] = P->Jt[1][0] = P->Jt[1][1] = 0;
  for (int k = 0; k < 4; ++k)
    for (int d = 0; d < 2; ++d) {
      P->x[d] += N[k] * cc[2 * k + d];
      P->Jt[0][d] += dN[k][0] * cc[2 * k + d];
      P->Jt[1][d] += dN[k][1] * cc[2 * k + d];
    }
  P->det = P->Jt[0][0] * P->Jt[1][1] - P->Jt[0][1] * P->Jt[1][0];
  P->dV = w * fabs(P->det);
  metric_at(radius, P->x[0], P->x[1], P->g);
  inv_metric_at(radius, P->x[0], P->x[1], P->h);
  P->sqrtg = sqrt(P->g[0] * P->g[2] - P->g[1] * P->g[1]);
}

/* physical RT basis at a point: u[m][d] = s_m (uhat[m] . Jt)[d] / det, div[m] = s_m divhat[m] / det */
static void rt_push(const QPoint* P, int mu, const double* rt_val_q /*[mu][2]*/, const double* rt_div_q /*[mu]*/,
                    const int8_t* sg, double* u /*[mu][2]*/, double* dv /*[mu]*/) {
  for (int m = 0; m < mu; ++m) {
    double s = sg ? (double)sg[m] : 1.0;
    double a0 = rt_val_q[2 * m], a1 = rt_val_q[2 * m + 1];
    u[2 * m] = s * (a0 * P->Jt[0][0] + a1 * P->Jt[1][0]) / P->det;
    u[2 * m + 1] = s * (a0 * P->Jt[0][1] + a1 * P->Jt[1][1]) / P->det;
    if (dv) dv[m] = s * rt_div_q[m] / P->det;
  }
}

/* ---- element matrices of the constant operators ------------------------------------------------------------------------ */

/* int (v . (g u)) / sqrtg  -- RT block of `mass` (TransientWaveEquation.jl:63); Ke [nc][mu][mu] */
void oracle_rt_mass_matrices(int64_t nc, const double* chart_coords, double radius, int Q, const double* pts, const double* w,
                             int mu, const double* rt_val, const int8_t* signs, double* Ke, int nthreads) {
  set_threads(nthreads);
#pragma omp parallel
  {
    double* u = (double*)malloc(sizeof(double) * mu * 2);
    double* gu = (double*)malloc(sizeof(double) * mu * 2);
#pragma omp for schedule(static)
    for (int64_t c = 0; c < nc; ++c) {
      double* K = Ke + (size_t)c * mu * mu;
      memset(K, 0, sizeof(double) * mu * mu);
      for (int q = 0; q < Q; ++q) {
        QPoint P;
        qpoint(chart_coords + c * 8, radius, pts[2 * q], pts[2 * q + 1], w[q], &P);
        rt_push(&P, mu, rt_val + (size_t)q * mu * 2, NULL, signs ? signs + c * mu : NULL, u, NULL);
        for (int m = 0; m < mu; ++m) {
          gu[2 * m] = P.g[0] * u[2 * m] + P.g[1] * u[2 * m + 1];
          gu[2 * m + 1] = P.g[1] * u[2 * m] + P.g[2] * u[2 * m + 1];
        }
        double f = P.dV / P.sqrtg;
        for (int a = 0; a < mu; ++a)
          for (int b = 0; b < mu; ++b) K[a * mu + b] += f * (u[2 * a] * gu[2 * b] + u[2 * a + 1] * gu[2 * b + 1]);
      }
    }
    free(u); free(gu);
  }
}

/* int q p sqrtg [coef]  -- scalar block of `mass`, and jac_y's potential-vorticity block with coef = depth at the points
 * (TransientShallowWater.jl:118); val [Q][m]; coef [nc][Q] or NULL; Ke [nc][m][m] */
void oracle_scalar_mass_matrices(int64_t nc, const double* chart_coords, double radius, int Q, const double* pts,
                                 const double* w, int m, const double* val, const double* coef, double* Ke, int nthreads) {
  set_threads(nthreads);
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < nc; ++c) {
    double* K = Ke + (size_t)c * m * m;
    memset(K, 0, sizeof(double) * m * m);
    for (int q = 0; q < Q; ++q) {
      QPoint P;
      qpoint(chart_coords + c * 8, radius, pts[2 * q], pts[2 * q + 1], w[q], &P);
      double f = P.dV * P.sqrtg * (coef ? coef[c * Q + q] : 1.0);
      for (int a = 0; a < m; ++a)
        for (int b = 0; b < m; ++b) K[a * m + b] += f * val[q * m + a] * val[q * m + b];
    }
  }
}

/* ---- linear wave: res(t,(u,p),(v,q)) = int q div(u) - int p div(v)  (TransientWaveEquation.jl:64) -------------------------
 * ue [nc][mu], pe [nc][mp] cell DOF values; ru [nc][mu], rp [nc][mp] element vectors */
void oracle_wave_residual_vectors(int64_t nc, const double* chart_coords, int Q, const double* pts, const double* w, int mu,
                                  const double* rt_div, const int8_t* signs, int mp, const double* dg_val, const double* ue,
                                  const double* pe, double* ru, double* rp, int nthreads) {
  set_threads(nthreads);
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < nc; ++c) {
    double* r_u = ru + c * mu;
    double* r_p = rp + c * mp;
    memset(r_u, 0, sizeof(double) * mu);
    memset(r_p, 0, sizeof(double) * mp);
    for (int q = 0; q < Q; ++q) {
      QPoint P;
      qpoint(chart_coords + c * 8, 1.0, pts[2 * q], pts[2 * q + 1], w[q], &P);
      double divu = 0, ph = 0;
      for (int m = 0; m < mu; ++m) divu += signs[c * mu + m] * rt_div[q * mu + m] / P.det * ue[c * mu + m];
      for (int a = 0; a < mp; ++a) ph += dg_val[q * mp + a] * pe[c * mp + a];
      for (int m = 0; m < mu; ++m) r_u[m] -= P.dV * ph * signs[c * mu + m] * rt_div[q * mu + m] / P.det;
      for (int a = 0; a < mp; ++a) r_p[a] += P.dV * dg_val[q * mp + a] * divu;
    }
  }
}

/* ---- shallow water: right-hand sides of the diagnostic system res_y (TransientShallowWater.jl:113-126) --------------------
 *   q:    int q p w sqrtg = int f w sqrtg + int ((R u) . ginv) . grad(w) sqrtg
 *   F:    int F.(g v)/sqrtg = int p u.(g v)/sqrtg
 *   Phi:  int Phi psi sqrtg = int g_r (p + b) psi sqrtg + int 0.5 (u.(g u)) psi / sqrtg
 * fq, bq [nc][Q] Coriolis parameter and topography at the points; outputs element vectors and the depth at the points pq */
void oracle_swe_diag_vectors(int64_t nc, const double* chart_coords, double radius, int Q, const double* pts, const double* w,
                             int mu, const double* rt_val, const int8_t* signs, int mp, const double* dg_val, int mr,
                             const double* cg_val, const double* cg_grad, const double* fq, const double* bq, double gravity,
                             const double* ue, const double* pe, double* rhs_q, double* rhs_F, double* rhs_P, double* pq,
                             int nthreads) {
  set_threads(nthreads);
#pragma omp parallel
  {
    double* u = (double*)malloc(sizeof(double) * mu * 2);
#pragma omp for schedule(static)
    for (int64_t c = 0; c < nc; ++c) {
      double* rq = rhs_q + c * mr;
      double* rF = rhs_F + c * mu;
      double* rP = rhs_P + c * mp;
      memset(rq, 0, sizeof(double) * mr);
      memset(rF, 0, sizeof(double) * mu);
      memset(rP, 0, sizeof(double) * mp);
      for (int q = 0; q < Q; ++q) {
        QPoint P;
        qpoint(chart_coords + c * 8, radius, pts[2 * q], pts[2 * q + 1], w[q], &P);
        rt_push(&P, mu, rt_val + (size_t)q * mu * 2, NULL, signs + c * mu, u, NULL);
        double uq[2] = {0, 0}, ph = 0;
        for (int m = 0; m < mu; ++m) { uq[0] += u[2 * m] * ue[c * mu + m]; uq[1] += u[2 * m + 1] * ue[c * mu + m]; }
        for (int a = 0; a < mp; ++a) ph += dg_val[q * mp + a] * pe[c * mp + a];
        pq[c * Q + q] = ph;
        double Ru[2] = {-uq[1], uq[0]};
        double flux[2] = {Ru[0] * P.h[0] + Ru[1] * P.h[1], Ru[0] * P.h[1] + Ru[1] * P.h[2]};
        double iJ[2][2] = {{P.Jt[1][1] / P.det, -P.Jt[0][1] / P.det}, {-P.Jt[1][0] / P.det, P.Jt[0][0] / P.det}};
        double wq = P.dV * P.sqrtg;
        for (int a = 0; a < mr; ++a) {
          double g0 = cg_grad[(q * mr + a) * 2], g1 = cg_grad[(q * mr + a) * 2 + 1];
          double gx = iJ[0][0] * g0 + iJ[0][1] * g1, gy = iJ[1][0] * g0 + iJ[1][1] * g1;
          rq[a] += wq * fq[c * Q + q] * cg_val[q * mr + a] + wq * (flux[0] * gx + flux[1] * gy);
        }
        double gu[2] = {P.g[0] * uq[0] + P.g[1] * uq[1], P.g[1] * uq[0] + P.g[2] * uq[1]};
        double fF = P.dV / P.sqrtg * ph;
        for (int m = 0; m < mu; ++m) rF[m] += fF * (gu[0] * u[2 * m] + gu[1] * u[2 * m + 1]);
        double ke = 0.5 * (uq[0] * gu[0] + uq[1] * gu[1]);
        double fP = P.dV * (gravity * (ph + (bq ? bq[c * Q + q] : 0.0)) * P.sqrtg + ke / P.sqrtg);
        for (int a = 0; a < mp; ++a) rP[a] += fP * dg_val[q * mp + a];
      }
    }
    free(u);
  }
}

/* ---- shallow water: res_x + mass(0) (TransientShallowWater.jl:128-136) as element vectors ------------------------------------
 *   r_u = int (q - tau (q - q0)/dt - tau u.grad(q)/sqrtg) (R F).v dV - int Phi div(v) dV ;  r_p = int psi div(F) dV */
void oracle_swe_residual_vectors(int64_t nc, const double* chart_coords, double radius, int Q, const double* pts, const double* w,
                                 int mu, const double* rt_val, const double* rt_div, const int8_t* signs, int mp,
                                 const double* dg_val, int mr, const double* cg_val, const double* cg_grad, double tau, double dt,
                                 const double* ue, const double* Fe, const double* Pe, const double* qe, const double* q0e,
                                 double* ru, double* rp, int nthreads) {
  set_threads(nthreads);
#pragma omp parallel
  {
    double* u = (double*)malloc(sizeof(double) * mu * 2);
    double* dv = (double*)malloc(sizeof(double) * mu);
#pragma omp for schedule(static)
    for (int64_t c = 0; c < nc; ++c) {
      double* r_u = ru + c * mu;
      double* r_p = rp + c * mp;
      memset(r_u, 0, sizeof(double) * mu);
      memset(r_p, 0, sizeof(double) * mp);
      for (int q = 0; q < Q; ++q) {
        QPoint P;
        qpoint(chart_coords + c * 8, radius, pts[2 * q], pts[2 * q + 1], w[q], &P);
        rt_push(&P, mu, rt_val + (size_t)q * mu * 2, rt_div + (size_t)q * mu, signs + c * mu, u, dv);
        double uq[2] = {0, 0}, Fq[2] = {0, 0}, divF = 0, Ph = 0, qq = 0, q0q = 0, gq[2] = {0, 0};
        for (int m = 0; m < mu; ++m) {
          uq[0] += u[2 * m] * ue[c * mu + m]; uq[1] += u[2 * m + 1] * ue[c * mu + m];
          Fq[0] += u[2 * m] * Fe[c * mu + m]; Fq[1] += u[2 * m + 1] * Fe[c * mu + m];
          divF += dv[m] * Fe[c * mu + m];
        }
        for (int a = 0; a < mp; ++a) Ph += dg_val[q * mp + a] * Pe[c * mp + a];
        double iJ[2][2] = {{P.Jt[1][1] / P.det, -P.Jt[0][1] / P.det}, {-P.Jt[1][0] / P.det, P.Jt[0][0] / P.det}};
        for (int a = 0; a < mr; ++a) {
          double g0 = cg_grad[(q * mr + a) * 2], g1 = cg_grad[(q * mr + a) * 2 + 1];
          qq += cg_val[q * mr + a] * qe[c * mr + a];
          q0q += cg_val[q * mr + a] * q0e[c * mr + a];
          gq[0] += (iJ[0][0] * g0 + iJ[0][1] * g1) * qe[c * mr + a];
          gq[1] += (iJ[1][0] * g0 + iJ[1][1] * g1) * qe[c * mr + a];
        }
        double coef = qq - tau * (qq - q0q) / dt - tau * (uq[0] * gq[0] + uq[1] * gq[1]) / P.sqrtg;
        double RF[2] = {-Fq[1], Fq[0]};
        for (int m = 0; m < mu; ++m) r_u[m] += P.dV * coef * (RF[0] * u[2 * m] + RF[1] * u[2 * m + 1]) - P.dV * Ph * dv[m];
        for (int a = 0; a < mp; ++a) r_p[a] += P.dV * dg_val[q * mp + a] * divF;
      }
    }
    free(u); free(dv);
  }
}

/* ---- solvers -------------------------------------------------------------------------------------------------------------- */

/* Jacobi-preconditioned CG on a CSR matrix, x = initial guess in / solution out; returns the iteration count (negative: not
 * converged).  Stopping test |r| <= rtol |b|. */
int oracle_csr_pcg(int64_t n, const int64_t* rowptr, const int64_t* colind, const double* vals, const double* b, double* x,
                   double rtol, int max_iter, int nthreads) {
  set_threads(nthreads);
  double* r = (double*)malloc(sizeof(double) * n);
  double* z = (double*)malloc(sizeof(double) * n);
  double* p = (double*)malloc(sizeof(double) * n);
  double* Ap = (double*)malloc(sizeof(double) * n);
  double* dinv = (double*)malloc(sizeof(double) * n);
  double bb = 0, rz = 0, rr = 0;
#pragma omp parallel for schedule(static) reduction(+ : bb, rz, rr)
  for (int64_t i = 0; i < n; ++i) {
    double s = 0, d = 1.0;
    for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
      s += vals[k] * x[colind[k]];
      if (colind[k] == i) d = vals[k];
    }
    dinv[i] = 1.0 / d;
    r[i] = b[i] - s;
    z[i] = dinv[i] * r[i];
    p[i] = z[i];
    bb += b[i] * b[i]; rz += r[i] * z[i]; rr += r[i] * r[i];
  }
  int it = 0;
  double tol2 = rtol * rtol * bb;
  while (rr > tol2 && it < max_iter) {
    double pAp = 0;
#pragma omp parallel for schedule(static) reduction(+ : pAp)
    for (int64_t i = 0; i < n; ++i) {
      double s = 0;
      for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) s += vals[k] * p[colind[k]];
      Ap[i] = s;
      pAp += p[i] * s;
    }
    double alpha = rz / pAp, rz_new = 0;
    rr = 0;
#pragma omp parallel for schedule(static) reduction(+ : rz_new, rr)
    for (int64_t i = 0; i < n; ++i) {
      x[i] += alpha * p[i];
      r[i] -= alpha * Ap[i];
      z[i] = dinv[i] * r[i];
      rz_new += r[i] * z[i];
      rr += r[i] * r[i];
    }
    double beta = rz_new / rz;
    rz = rz_new;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
    ++it;
  }
  free(r); free(z); free(p); free(Ap); free(dinv);
  return rr > tol2 ? -it : it;
}

/* block-diagonal solve (DG mass): Me [nc][m][m] (destroyed), be [nc][m] -> xe [nc][m]; Gaussian elimination with partial pivoting */
void oracle_block_solve(int64_t nc, int m, double* Me, const double* be, double* xe, int nthreads) {
  set_threads(nthreads);
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < nc; ++c) {
    double* A = Me + (size_t)c * m * m;
    double* x = xe + c * m;
    for (int i = 0; i < m; ++i) x[i] = be[c * m + i];
    for (int k = 0; k < m; ++k) {
      int piv = k;
      for (int i = k + 1; i < m; ++i)
        if (fabs(A[i * m + k]) > fabs(A[piv * m + k])) piv = i;
      if (piv != k) {
        for (int j = 0; j < m; ++j) { double t = A[k * m + j]; A[k * m + j] = A[piv * m + j]; A[piv * m + j] = t; }
        double t = x[k]; x[k] = x[piv]; x[piv] = t;
      }
      for (int i = k + 1; i < m; ++i) {
        double f = A[i * m + k] / A[k * m + k];
        for (int j = k; j < m; ++j) A[i * m + j] -= f * A[k * m + j];
        x[i] -= f * x[k];
      }
    }
    for (int i = m - 1; i >= 0; --i) {
      double s = x[i];
      for (int j = i + 1; j < m; ++j) s -= A[i * m + j] * x[j];
      x[i] = s / A[i * m + i];
    }
  }
}
