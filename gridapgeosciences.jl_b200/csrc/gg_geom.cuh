// Device (and host) evaluation of the analytic cubed-sphere geometry in FP64 registers.
// Formula-for-formula the reference's
//   src/Fields/CubedSphereMap.jl:19-49        gnomonic panel map and Jacobian, panel permutation table
//   src/Fields/CubedSphereMetric.jl:16-22     g = r^2 sa sb / rho^4 [[sa, -ab], [-ab, sb]]
//   src/Fields/CubedSphereInvMetric.jl:1-7    g^-1 = rho^2 / (r^2 sa sb) [[sb, ab], [ab, sa]]
//   src/CellData/CellFields.jl:80-82          sqrt(det g)  (= r^2 sa sb / rho^3)
// with a = tan(alpha), b = tan(beta), sa = 1 + a^2, sb = 1 + b^2, rho^2 = 1 + a^2 + b^2.
#pragma once
#include <cuda_runtime.h>

namespace gg {

// (j1,s1, j2,s2, j3,s3), 0-based j: output component k = s_k * h[j_k]
static __constant__ int c_perm_j[6][3] = {{0, 1, 2}, {2, 1, 0}, {1, 0, 2}, {0, 2, 1}, {1, 2, 0}, {2, 0, 1}};
static __constant__ double c_perm_s[6][3] = {{1, 1, 1}, {-1, 1, 1}, {-1, 1, 1}, {-1, 1, 1}, {-1, 1, -1}, {-1, -1, 1}};

struct Metric2 {
  double g11, g12, g22;  // metric
  double i11, i12, i22;  // inverse metric
  double sqrtg;          // sqrt(det g)
};

__device__ __forceinline__ Metric2 csphere_metric_terms(double a, double b, double r) {
  Metric2 m;
  double sa = 1.0 + a * a, sb = 1.0 + b * b;
  double rho2 = 1.0 + a * a + b * b;
  double r2 = r * r;
  double c = r2 * sa * sb / (rho2 * rho2);
  m.g11 = c * sa; m.g12 = -c * a * b; m.g22 = c * sb;
  double ci = rho2 / (r2 * sa * sb);
  m.i11 = ci * sb; m.i12 = ci * a * b; m.i22 = ci * sa;
  m.sqrtg = sqrt(m.g11 * m.g22 - m.g12 * m.g12);
  return m;
}

// xyz = r * perm_p(1, a, b) / rho
__device__ __forceinline__ void csphere_point(int panel, double a, double b, double r, double* xyz) {
  double rho = sqrt(1.0 + a * a + b * b);
  double h[3] = {1.0 / rho, a / rho, b / rho};
#pragma unroll
  for (int k = 0; k < 3; ++k) xyz[k] = r * c_perm_s[panel][k] * h[c_perm_j[panel][k]];
}

// J[i][k] = d phi_k / d x_i, i in {alpha, beta}
__device__ __forceinline__ void csphere_jacobian(int panel, double a, double b, double r, double J[2][3]) {
  double rho2 = 1.0 + a * a + b * b;
  double rho3 = rho2 * sqrt(rho2);
  double sa = 1.0 + a * a, sb = 1.0 + b * b;
  double c[3][2] = {{-a * sa, -b * sb}, {sa * sb, -a * b * sb}, {-a * b * sa, sa * sb}};
  double f = r / rho3;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int j = c_perm_j[panel][k];
    double s = c_perm_s[panel][k];
    J[0][k] = f * s * c[j][0];
    J[1][k] = f * s * c[j][1];
  }
}

// Extrinsic ("ambient model") metric terms from the 2x3 Jacobian of cell_map = ambient o chart:
// G = J J^T (pseudo-determinant meas = sqrt(det G)), Ginv = (J J^T)^-1 (what the pseudo-inverse gradient contracts to)
__device__ __forceinline__ Metric2 extrinsic_metric_terms(int panel, double a, double b, double r) {
  double J[2][3];
  csphere_jacobian(panel, a, b, r, J);
  Metric2 m;
  m.g11 = J[0][0] * J[0][0] + J[0][1] * J[0][1] + J[0][2] * J[0][2];
  m.g12 = J[0][0] * J[1][0] + J[0][1] * J[1][1] + J[0][2] * J[1][2];
  m.g22 = J[1][0] * J[1][0] + J[1][1] * J[1][1] + J[1][2] * J[1][2];
  double det = m.g11 * m.g22 - m.g12 * m.g12;
  double id = 1.0 / det;
  m.i11 = m.g22 * id; m.i12 = -m.g12 * id; m.i22 = m.g11 * id;
  m.sqrtg = sqrt(det);
  return m;
}


// Second derivatives of the panel map: H[l][i][k] = d^2 phi_k / (d x_i d x_l)  (src/Fields/CubedSphereMap.jl:97-123)
__device__ __forceinline__ void csphere_hessian(int panel, double a, double b, double r, double H[2][2][3]) {
  const double sa = 1.0 + a * a, sb = 1.0 + b * b, rho2 = 1.0 + a * a + b * b;
  const double rho5 = rho2 * rho2 * sqrt(rho2);
  const double D = rho2 + 3.0 * a * a * b * b, E = 3.0 * a * b * sa * sb;
  const double F = a * sa * sb * (2.0 * b * b - sa), G = b * sa * sb * (2.0 * a * a - sb);
  const double n[3][4] = {{-sa * D, E, E, -sb * D}, {F, G, G, -a * sb * D}, {-b * sa * D, F, F, G}};
  const double fac = r / rho5;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int j = c_perm_j[panel][k];
    const double s = fac * c_perm_s[panel][k];
    H[0][0][k] = s * n[j][0]; H[1][0][k] = s * n[j][1]; H[0][1][k] = s * n[j][2]; H[1][1][k] = s * n[j][3];
  }
}

// dg[k] = (d g11, d g12, d g22) / d x_k  (src/Fields/CubedSphereMetric.jl:52-71) and
// dgi[k] = (d ginv11, d ginv12, d ginv22) / d x_k  (src/Fields/CubedSphereInvMetric.jl:37-56)
__device__ __forceinline__ void csphere_metric_gradients(double a, double b, double r, double dg[2][3], double dgi[2][3]) {
  const double sa = 1.0 + a * a, sb = 1.0 + b * b, rho2 = 1.0 + a * a + b * b;
  const double c6 = r * r / (rho2 * rho2 * rho2);
  dg[0][0] = 4.0 * c6 * a * sa * sa * sb * b * b;
  dg[1][0] = 2.0 * c6 * b * sa * sa * sb * (a * a - sb);
  dg[0][1] = -c6 * b * sa * sb * (rho2 + a * a * (3.0 * b * b - sa));
  dg[1][1] = -c6 * a * sa * sb * (rho2 + b * b * (3.0 * a * a - sb));
  dg[0][2] = 2.0 * c6 * a * sa * sb * sb * (b * b - sa);
  dg[1][2] = 4.0 * c6 * b * sa * sb * sb * a * a;
  const double ci = 1.0 / (r * r * sa * sb);
  dgi[0][0] = -2.0 * ci * a * b * b * sb;
  dgi[1][0] = 2.0 * ci * b * sb * sb;
  dgi[0][1] = ci * b * (sa * rho2 - 2.0 * a * a * b * b);
  dgi[1][1] = ci * a * (sb * rho2 - 2.0 * a * a * b * b);
  dgi[0][2] = 2.0 * ci * a * sa * sa;
  dgi[1][2] = -2.0 * ci * b * a * a * sa;
}

}  // namespace gg
