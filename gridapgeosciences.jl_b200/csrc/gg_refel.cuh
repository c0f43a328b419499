// Reference-element tables: tensor Gauss-Legendre quadrature and 1-D Lagrange bases on [0,1].
// The reference takes these from Gridap (Measure / LagrangianRefFE; call sites
// test/Laplacian/LaplaceBeltrami.jl:29-35).  Conventions: SURVEY.md App. B.1-B.3.
#pragma once
#include <vector>
#include <stdint.h>

namespace gg {

constexpr int MAX_NB = 5;   // 1-D basis size (order <= 4)
constexpr int MAX_NQ = 16;  // 1-D quadrature points of the form kernels (degree <= 31: fixed-size tables)
// error measures, integrals and point sets work from dynamically sized tabulations: the Stokes drivers measure errors with
// Measure(., 2 degree) = 36 and Measure(., 3 degree) = 54 (test/SurfaceStokes/VectorLaplacian2D.jl:30-33, SurfaceStokes_TaylorHood.jl:45-48)
constexpr int MAX_NQ_NORMS = 32;  // degree <= 63

// nq = ceil((degree+1)/2)
int num_points_1d(int degree);
// Gauss-Legendre points/weights mapped to [0,1]
void gauss_legendre_01(int nq, double* x, double* w);
// Lagrange basis on the equispaced nodes i/p (p=0: the constant 1), values and derivatives at x
void lagrange_1d(int p, double x, double* val, double* der);
// Gridap local node order of Q_p on the square -> lexicographic tensor index ix + (p+1)*iy
std::vector<int32_t> quad_local_to_lex(int p);

// Tables of one (order, nq) pair as laid out in __constant__ memory for the scalar Q_p kernels.
// Pair tables carry the quadrature weight: NN[i*nb+j][q] = w_q N_i(x_q) N_j(x_q), ND = w N_i N'_j, DD = w N'_i N'_j.
struct ScalarTables {
  int32_t nb, nq;
  double xi[MAX_NQ], w[MAX_NQ];
  double N[MAX_NQ][MAX_NB], D[MAX_NQ][MAX_NB];
  double NN[MAX_NB * MAX_NB][MAX_NQ];
  double ND[MAX_NB * MAX_NB][MAX_NQ];
  double DD[MAX_NB * MAX_NB][MAX_NQ];
};
void build_scalar_tables(int order, int nq, ScalarTables* t);

// Raviart-Thomas RT_k on the square (moment-based, outward reference normals):
// coefficient matrix of the nodal basis in the monomial prebasis Q_{k+1,k} x Q_{k,k+1}.
struct RTRef {
  int k = 0, ndofs = 0;
  std::vector<int> comp, ex, ey;  // prebasis monomials: component, x exponent, y exponent
  std::vector<double> C;          // [ndofs(prebasis)][ndofs(basis)]: phi_j = sum_i C[i][j] psi_i
  // tabulate at reference points: val [npts][ndofs][2], div [npts][ndofs]
  void tabulate(int npts, const double* pts, std::vector<double>& val, std::vector<double>& div) const;
};
RTRef build_rt_ref(int k);
// points [N][2] and weights [ndofs][N][2] of the moment functionals of RT_k (interpolation of a reference-space field)
void rt_moment_functionals(int k, std::vector<double>& pts, std::vector<double>& W);

// Raviart-Thomas RT_k on the cube (same conventions one dimension up): prebasis Q_{k+1,k,k} x Q_{k,k+1,k} x Q_{k,k,k+1};
// faces in Gridap order (z=0, z=1, y=0, y=1, x=0, x=1), each parametrised by its two spanning axes in increasing order, with
// (k+1)^2 moments of the outward normal flux against s^i t^j (i fastest); then the interior moments of component a against
// x^i y^j z^l with degree k-1 in x_a.  3 (k+2)(k+1)^2 DOFs.
struct RTRefHex {
  int k = 0, ndofs = 0;
  std::vector<int> comp, ex, ey, ez;
  std::vector<double> C;          // [ndofs(prebasis)][ndofs(basis)]
  // tabulate at reference points [npts][3]: val [npts][ndofs][3], div [npts][ndofs]
  void tabulate(int npts, const double* pts, std::vector<double>& val, std::vector<double>& div) const;
};
RTRefHex build_rt_hex_ref(int k);
// points [N][3] and weights [ndofs][N][3] of the moment functionals of the hexahedral RT_k
void rt_hex_moment_functionals(int k, std::vector<double>& pts, std::vector<double>& W);

}  // namespace gg
