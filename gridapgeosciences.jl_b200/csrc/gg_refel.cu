// Host-side construction of quadrature rules and reference bases (tiny, one-off).
#include <cmath>
#include <cstring>
#include "gg_common.cuh"
#include "gg_refel.cuh"

namespace gg {

int num_points_1d(int degree) { return (degree + 2) / 2; }  // ceil((degree+1)/2)

void gauss_legendre_01(int nq, double* x, double* w) {
  // Newton iteration on P_n with the Chebyshev initial guess; converges to machine precision
  for (int i = 0; i < nq; ++i) {
    double z = std::cos(M_PI * (i + 0.75) / (nq + 0.5));
    double pp = 0.0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 0; j < nq; ++j) {
        double p3 = p2;
        p2 = p1;
        p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0);
      }
      pp = nq * (z * p1 - p2) / (z * z - 1.0);
      double dz = p1 / pp;
      z -= dz;
      if (std::fabs(dz) < 1e-16) break;
    }
    // roots come out in decreasing order; store ascending
    int k = nq - 1 - i;
    x[k] = 0.5 * (z + 1.0);
    w[k] = 1.0 / ((1.0 - z * z) * pp * pp);  // = (2/((1-z^2)pp^2))/2
  }
  // symmetrise exactly
  for (int i = 0; i < nq / 2; ++i) {
    double xs = 0.5 * (x[i] + (1.0 - x[nq - 1 - i]));
    x[i] = xs;
    x[nq - 1 - i] = 1.0 - xs;
    double ws = 0.5 * (w[i] + w[nq - 1 - i]);
    w[i] = w[nq - 1 - i] = ws;
  }
  if (nq % 2) x[nq / 2] = 0.5;
}

void lagrange_1d(int p, double x, double* val, double* der) {
  if (p == 0) {
    val[0] = 1.0;
    der[0] = 0.0;
    return;
  }
  int n = p + 1;
  for (int i = 0; i < n; ++i) {
    double xi = (double)i / p;
    double v = 1.0;
    for (int j = 0; j < n; ++j)
      if (j != i) v *= (x - (double)j / p) / (xi - (double)j / p);
    val[i] = v;
    double d = 0.0;
    for (int j = 0; j < n; ++j) {
      if (j == i) continue;
      double t = 1.0 / (xi - (double)j / p);
      for (int l = 0; l < n; ++l)
        if (l != i && l != j) t *= (x - (double)l / p) / (xi - (double)l / p);
      d += t;
    }
    der[i] = d;
  }
}

std::vector<int32_t> quad_local_to_lex(int p) {
  // vertices (x fastest), edges [1,2],[3,4],[1,3],[2,4] interiors, cell interior (x fastest)
  std::vector<int32_t> out;
  int nb = p + 1;
  if (p == 0) return {0};
  auto lex = [&](int ix, int iy) { return ix + nb * iy; };
  out.push_back(lex(0, 0));
  out.push_back(lex(p, 0));
  out.push_back(lex(0, p));
  out.push_back(lex(p, p));
  for (int t = 1; t < p; ++t) out.push_back(lex(t, 0));
  for (int t = 1; t < p; ++t) out.push_back(lex(t, p));
  for (int t = 1; t < p; ++t) out.push_back(lex(0, t));
  for (int t = 1; t < p; ++t) out.push_back(lex(p, t));
  for (int iy = 1; iy < p; ++iy)
    for (int ix = 1; ix < p; ++ix) out.push_back(lex(ix, iy));
  return out;
}

void build_scalar_tables(int order, int nq, ScalarTables* t) {
  GG_REQUIRE(order >= 0 && order + 1 <= MAX_NB, GG_ERR_UNSUPPORTED, "Lagrangian order must be in 0..4");
  GG_REQUIRE(nq >= 1 && nq <= MAX_NQ, GG_ERR_UNSUPPORTED, "quadrature degree too high (max 31)");
  std::memset(t, 0, sizeof(*t));
  t->nb = order + 1;
  t->nq = nq;
  gauss_legendre_01(nq, t->xi, t->w);
  for (int q = 0; q < nq; ++q) lagrange_1d(order, t->xi[q], t->N[q], t->D[q]);
  int nb = t->nb;
  for (int i = 0; i < nb; ++i)
    for (int j = 0; j < nb; ++j)
      for (int q = 0; q < nq; ++q) {
        t->NN[i * nb + j][q] = t->w[q] * t->N[q][i] * t->N[q][j];
        t->ND[i * nb + j][q] = t->w[q] * t->N[q][i] * t->D[q][j];
        t->DD[i * nb + j][q] = t->w[q] * t->D[q][i] * t->D[q][j];
      }
}

// ---- Raviart-Thomas ---------------------------------------------------------------------------------

static void invert_dense(int n, std::vector<double>& A) {
  // Gauss-Jordan with partial pivoting, in place (row-major)
  std::vector<double> B(n * n, 0.0);
  for (int i = 0; i < n; ++i) B[i * n + i] = 1.0;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(A[r * n + c]) > std::fabs(A[piv * n + c])) piv = r;
    GG_REQUIRE(std::fabs(A[piv * n + c]) > 1e-14, GG_ERR_INVALID, "singular moment matrix");
    if (piv != c)
      for (int k = 0; k < n; ++k) {
        std::swap(A[c * n + k], A[piv * n + k]);
        std::swap(B[c * n + k], B[piv * n + k]);
      }
    double d = 1.0 / A[c * n + c];
    for (int k = 0; k < n; ++k) {
      A[c * n + k] *= d;
      B[c * n + k] *= d;
    }
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      double f = A[r * n + c];
      if (f == 0.0) continue;
      for (int k = 0; k < n; ++k) {
        A[r * n + k] -= f * A[c * n + k];
        B[r * n + k] -= f * B[c * n + k];
      }
    }
  }
  A = B;
}

static inline double ipow(double x, int e) {
  double r = 1.0;
  for (int i = 0; i < e; ++i) r *= x;
  return r;
}

RTRef build_rt_ref(int k) {
  GG_REQUIRE(k >= 0 && k <= 3, GG_ERR_UNSUPPORTED, "Raviart-Thomas order must be in 0..3");
  RTRef R;
  R.k = k;
  for (int j = 0; j <= k; ++j)
    for (int i = 0; i <= k + 1; ++i) { R.comp.push_back(0); R.ex.push_back(i); R.ey.push_back(j); }
  for (int j = 0; j <= k + 1; ++j)
    for (int i = 0; i <= k; ++i) { R.comp.push_back(1); R.ex.push_back(i); R.ey.push_back(j); }
  int n = (int)R.comp.size();
  R.ndofs = n;
  std::vector<double> M(n * n, 0.0);
  // facet moments: facets [1,2] (y=0, n=(0,-1)), [3,4] (y=1, n=(0,1)), [1,3] (x=0, n=(-1,0)), [2,4] (x=1, n=(1,0)),
  // parametrised from the first to the second vertex, against s^m, m = 0..k
  const double va[4][2] = {{0, 0}, {0, 1}, {0, 0}, {1, 0}};
  const double vb[4][2] = {{1, 0}, {1, 1}, {0, 1}, {1, 1}};
  const double nn[4][2] = {{0, -1}, {0, 1}, {-1, 0}, {1, 0}};
  int nqf = k + 2;
  double xq[MAX_NQ], wq[MAX_NQ];
  gauss_legendre_01(nqf, xq, wq);
  int row = 0;
  for (int f = 0; f < 4; ++f)
    for (int m = 0; m <= k; ++m) {
      for (int q = 0; q < nqf; ++q) {
        double x = va[f][0] + xq[q] * (vb[f][0] - va[f][0]);
        double y = va[f][1] + xq[q] * (vb[f][1] - va[f][1]);
        double wt = wq[q] * ipow(xq[q], m);
        for (int c = 0; c < n; ++c) M[row * n + c] += wt * nn[f][R.comp[c]] * ipow(x, R.ex[c]) * ipow(y, R.ey[c]);
      }
      ++row;
    }
  if (k > 0) {
    int nqc = num_points_1d(2 * (k + 1));
    double xc[MAX_NQ], wc[MAX_NQ];
    gauss_legendre_01(nqc, xc, wc);
    for (int comp = 0; comp < 2; ++comp) {
      int imax = comp == 0 ? k - 1 : k, jmax = comp == 0 ? k : k - 1;
      for (int j = 0; j <= jmax; ++j)
        for (int i = 0; i <= imax; ++i) {
          for (int qy = 0; qy < nqc; ++qy)
            for (int qx = 0; qx < nqc; ++qx) {
              double x = xc[qx], y = xc[qy];
              double wt = wc[qx] * wc[qy] * ipow(x, i) * ipow(y, j);
              for (int c = 0; c < n; ++c)
                if (R.comp[c] == comp) M[row * n + c] += wt * ipow(x, R.ex[c]) * ipow(y, R.ey[c]);
            }
          ++row;
        }
    }
  }
  GG_REQUIRE(row == n, GG_ERR_INVALID, "RT moment count mismatch");
  // C = M^-1: phi_j = sum_i C[i][j] psi_i.  The monomial moment matrix is mildly ill-conditioned (1e4 for k = 2):
  // polish the inverse with two Newton-Schulz steps C <- C + C (I - M C) in extended precision.
  std::vector<double> Minv = M;
  invert_dense(n, Minv);
  std::vector<long double> Cl(Minv.begin(), Minv.end()), Rl(n * n), Tl(n * n);
  for (int it = 0; it < 2; ++it) {
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        long double s = (i == j) ? 1.0L : 0.0L;
        for (int k2 = 0; k2 < n; ++k2) s -= (long double)M[i * n + k2] * Cl[k2 * n + j];
        Rl[i * n + j] = s;
      }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        long double s = 0.0L;
        for (int k2 = 0; k2 < n; ++k2) s += Cl[i * n + k2] * Rl[k2 * n + j];
        Tl[i * n + j] = s;
      }
    for (int i = 0; i < n * n; ++i) Cl[i] += Tl[i];
  }
  R.C.resize(n * n);
  for (int i = 0; i < n * n; ++i) R.C[i] = (double)Cl[i];
  return R;
}

// Points and weights of the RT_k moment functionals (same facet / interior rules as build_rt_ref):
// dof_i(f) = sum_n W[(i*N + n)*2 + c] f_c(pts[n]) for a reference-space vector field f.
void rt_moment_functionals(int k, std::vector<double>& pts, std::vector<double>& W) {
  GG_REQUIRE(k >= 0 && k <= 3, GG_ERR_UNSUPPORTED, "Raviart-Thomas order must be in 0..3");
  const double va[4][2] = {{0, 0}, {0, 1}, {0, 0}, {1, 0}};
  const double vb[4][2] = {{1, 0}, {1, 1}, {0, 1}, {1, 1}};
  const double nn[4][2] = {{0, -1}, {0, 1}, {-1, 0}, {1, 0}};
  int nqf = k + 2, nqc = k > 0 ? num_points_1d(2 * (k + 1)) : 0;
  int N = 4 * nqf + nqc * nqc;
  int ndofs = 2 * (k + 1) * (k + 2);
  double xq[MAX_NQ], wq[MAX_NQ], xc[MAX_NQ], wc[MAX_NQ];
  gauss_legendre_01(nqf, xq, wq);
  if (nqc) gauss_legendre_01(nqc, xc, wc);
  pts.assign((size_t)N * 2, 0.0);
  W.assign((size_t)ndofs * N * 2, 0.0);
  for (int f = 0; f < 4; ++f)
    for (int q = 0; q < nqf; ++q) {
      pts[2 * (f * nqf + q)] = va[f][0] + xq[q] * (vb[f][0] - va[f][0]);
      pts[2 * (f * nqf + q) + 1] = va[f][1] + xq[q] * (vb[f][1] - va[f][1]);
    }
  for (int qy = 0; qy < nqc; ++qy)
    for (int qx = 0; qx < nqc; ++qx) {
      pts[2 * (4 * nqf + qx + nqc * qy)] = xc[qx];
      pts[2 * (4 * nqf + qx + nqc * qy) + 1] = xc[qy];
    }
  int row = 0;
  for (int f = 0; f < 4; ++f)
    for (int m = 0; m <= k; ++m) {
      for (int q = 0; q < nqf; ++q)
        for (int c = 0; c < 2; ++c) W[((size_t)row * N + f * nqf + q) * 2 + c] = wq[q] * ipow(xq[q], m) * nn[f][c];
      ++row;
    }
  if (k > 0)
    for (int comp = 0; comp < 2; ++comp) {
      int imax = comp == 0 ? k - 1 : k, jmax = comp == 0 ? k : k - 1;
      for (int j = 0; j <= jmax; ++j)
        for (int i = 0; i <= imax; ++i) {
          for (int qy = 0; qy < nqc; ++qy)
            for (int qx = 0; qx < nqc; ++qx)
              W[((size_t)row * N + 4 * nqf + qx + nqc * qy) * 2 + comp] = wc[qx] * wc[qy] * ipow(xc[qx], i) * ipow(xc[qy], j);
          ++row;
        }
    }
  GG_REQUIRE(row == ndofs, GG_ERR_INVALID, "RT moment count mismatch");
}

void RTRef::tabulate(int npts, const double* pts, std::vector<double>& val, std::vector<double>& div) const {
  int n = ndofs;
  val.assign((size_t)npts * n * 2, 0.0);
  div.assign((size_t)npts * n, 0.0);
  std::vector<double> pv(n), pd(n);
  for (int q = 0; q < npts; ++q) {
    double x = pts[2 * q], y = pts[2 * q + 1];
    for (int i = 0; i < n; ++i) {
      pv[i] = ipow(x, ex[i]) * ipow(y, ey[i]);
      if (comp[i] == 0)
        pd[i] = ex[i] > 0 ? ex[i] * ipow(x, ex[i] - 1) * ipow(y, ey[i]) : 0.0;
      else
        pd[i] = ey[i] > 0 ? ey[i] * ipow(x, ex[i]) * ipow(y, ey[i] - 1) : 0.0;
    }
    for (int j = 0; j < n; ++j) {
      double v0 = 0, v1 = 0, d = 0;
      for (int i = 0; i < n; ++i) {
        double c = C[i * n + j];
        if (comp[i] == 0) v0 += c * pv[i]; else v1 += c * pv[i];
        d += c * pd[i];
      }
      val[((size_t)q * n + j) * 2 + 0] = v0;
      val[((size_t)q * n + j) * 2 + 1] = v1;
      div[(size_t)q * n + j] = d;
    }
  }
}

// ---- Raviart-Thomas on the cube ------------------------------------------------------------------------------------------

static const int HEXF_AXIS[6] = {2, 2, 1, 1, 0, 0};
static const int HEXF_VAL[6] = {0, 1, 0, 1, 0, 1};

void rt_hex_moment_functionals(int k, std::vector<double>& pts, std::vector<double>& W) {
  GG_REQUIRE(k >= 0 && k <= 1, GG_ERR_UNSUPPORTED, "hexahedral Raviart-Thomas order must be 0 or 1");
  const int nqf = k + 2, nqc = k > 0 ? num_points_1d(2 * (k + 1)) : 0;
  const int nfp = nqf * nqf, N = 6 * nfp + nqc * nqc * nqc;
  const int ndofs = 3 * (k + 2) * (k + 1) * (k + 1);
  double xq[MAX_NQ], wq[MAX_NQ], xc[MAX_NQ], wc[MAX_NQ];
  gauss_legendre_01(nqf, xq, wq);
  if (nqc) gauss_legendre_01(nqc, xc, wc);
  pts.assign((size_t)N * 3, 0.0);
  W.assign((size_t)ndofs * N * 3, 0.0);
  for (int f = 0; f < 6; ++f) {
    const int ax = HEXF_AXIS[f];
    int sp[2], t = 0;
    for (int a = 0; a < 3; ++a) if (a != ax) sp[t++] = a;
    for (int q2 = 0; q2 < nqf; ++q2)
      for (int q1 = 0; q1 < nqf; ++q1) {
        double* p = &pts[(size_t)(f * nfp + q1 + nqf * q2) * 3];
        p[ax] = HEXF_VAL[f]; p[sp[0]] = xq[q1]; p[sp[1]] = xq[q2];
      }
  }
  for (int q3 = 0; q3 < nqc; ++q3)
    for (int q2 = 0; q2 < nqc; ++q2)
      for (int q1 = 0; q1 < nqc; ++q1) {
        double* p = &pts[(size_t)(6 * nfp + q1 + nqc * (q2 + nqc * q3)) * 3];
        p[0] = xc[q1]; p[1] = xc[q2]; p[2] = xc[q3];
      }
  int row = 0;
  for (int f = 0; f < 6; ++f)
    for (int j = 0; j <= k; ++j)
      for (int i = 0; i <= k; ++i) {
        for (int q2 = 0; q2 < nqf; ++q2)
          for (int q1 = 0; q1 < nqf; ++q1)
            W[((size_t)row * N + f * nfp + q1 + nqf * q2) * 3 + HEXF_AXIS[f]] =
                wq[q1] * wq[q2] * ipow(xq[q1], i) * ipow(xq[q2], j) * (HEXF_VAL[f] ? 1.0 : -1.0);
        ++row;
      }
  if (k > 0)
    for (int c = 0; c < 3; ++c) {
      int deg[3] = {k, k, k};
      deg[c] = k - 1;
      for (int l = 0; l <= deg[2]; ++l)
        for (int j = 0; j <= deg[1]; ++j)
          for (int i = 0; i <= deg[0]; ++i) {
            for (int q3 = 0; q3 < nqc; ++q3)
              for (int q2 = 0; q2 < nqc; ++q2)
                for (int q1 = 0; q1 < nqc; ++q1)
                  W[((size_t)row * N + 6 * nfp + q1 + nqc * (q2 + nqc * q3)) * 3 + c] =
                      wc[q1] * wc[q2] * wc[q3] * ipow(xc[q1], i) * ipow(xc[q2], j) * ipow(xc[q3], l);
            ++row;
          }
    }
  GG_REQUIRE(row == ndofs, GG_ERR_INVALID, "hexahedral RT moment count mismatch");
}

RTRefHex build_rt_hex_ref(int k) {
  GG_REQUIRE(k >= 0 && k <= 1, GG_ERR_UNSUPPORTED, "hexahedral Raviart-Thomas order must be 0 or 1");
  RTRefHex R;
  R.k = k;
  for (int c = 0; c < 3; ++c) {
    int deg[3] = {k, k, k};
    deg[c] = k + 1;
    for (int l = 0; l <= deg[2]; ++l)
      for (int j = 0; j <= deg[1]; ++j)
        for (int i = 0; i <= deg[0]; ++i) { R.comp.push_back(c); R.ex.push_back(i); R.ey.push_back(j); R.ez.push_back(l); }
  }
  const int n = (int)R.comp.size();
  R.ndofs = n;
  std::vector<double> pts, W;
  rt_hex_moment_functionals(k, pts, W);
  const int N = (int)(pts.size() / 3);
  std::vector<double> M((size_t)n * n, 0.0);
  for (int m = 0; m < n; ++m)
    for (int q = 0; q < N; ++q) {
      const double* p = &pts[(size_t)q * 3];
      for (int c = 0; c < n; ++c) {
        const double w = W[((size_t)m * N + q) * 3 + R.comp[c]];
        if (w != 0.0) M[(size_t)m * n + c] += w * ipow(p[0], R.ex[c]) * ipow(p[1], R.ey[c]) * ipow(p[2], R.ez[c]);
      }
    }
  std::vector<double> Minv = M;
  invert_dense(n, Minv);
  std::vector<long double> Cl(Minv.begin(), Minv.end()), Rl((size_t)n * n), Tl((size_t)n * n);
  for (int it = 0; it < 2; ++it) {
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        long double s = (i == j) ? 1.0L : 0.0L;
        for (int k2 = 0; k2 < n; ++k2) s -= (long double)M[(size_t)i * n + k2] * Cl[(size_t)k2 * n + j];
        Rl[(size_t)i * n + j] = s;
      }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        long double s = 0.0L;
        for (int k2 = 0; k2 < n; ++k2) s += Cl[(size_t)i * n + k2] * Rl[(size_t)k2 * n + j];
        Tl[(size_t)i * n + j] = s;
      }
    for (size_t i = 0; i < (size_t)n * n; ++i) Cl[i] += Tl[i];
  }
  R.C.resize((size_t)n * n);
  for (size_t i = 0; i < (size_t)n * n; ++i) R.C[i] = (double)Cl[i];
  return R;
}

void RTRefHex::tabulate(int npts, const double* pts, std::vector<double>& val, std::vector<double>& div) const {
  const int n = ndofs;
  val.assign((size_t)npts * n * 3, 0.0);
  div.assign((size_t)npts * n, 0.0);
  std::vector<double> pv(n), pd(n);
  for (int q = 0; q < npts; ++q) {
    const double x = pts[3 * q], y = pts[3 * q + 1], z = pts[3 * q + 2];
    for (int i = 0; i < n; ++i) {
      pv[i] = ipow(x, ex[i]) * ipow(y, ey[i]) * ipow(z, ez[i]);
      int e[3] = {ex[i], ey[i], ez[i]};
      if (e[comp[i]] > 0) {
        const double f = e[comp[i]];
        e[comp[i]] -= 1;
        pd[i] = f * ipow(x, e[0]) * ipow(y, e[1]) * ipow(z, e[2]);
      } else {
        pd[i] = 0.0;
      }
    }
    for (int j = 0; j < n; ++j) {
      double v[3] = {0, 0, 0}, d = 0;
      for (int i = 0; i < n; ++i) {
        const double c = C[(size_t)i * n + j];
        v[comp[i]] += c * pv[i];
        d += c * pd[i];
      }
      for (int a = 0; a < 3; ++a) val[((size_t)q * n + j) * 3 + a] = v[a];
      div[(size_t)q * n + j] = d;
    }
  }
}

}  // namespace gg
