// Hot kernels: Laplace-Beltrami element matrices / action for scalar Q_p (p = 1, 2) on axis-aligned chart cells.
//
// What the reference computes per cell and quadrature point (test/Laplacian/LaplaceBeltrami.jl:47, evaluated by
// Gridap's lazy operation tree; geometry from src/Fields/CubedSphereInvMetric.jl:1-7 and
// src/CellData/CellFields.jl:80-82):
//     Ke[I,J] += w_q |det grad(cell_map)| grad(phi_I) . (ginv(x_q) grad(phi_J)) sqrtg(x_q)
// B200 design: ONE THREAD PER CELL, everything in FP64 registers, no shared memory, no synchronisation.
//   * geometry: ginv*sqrtg = (1/rho) [[sb, ab],[ab, sa]] (the radius cancels in 2-D).  tan() is not evaluated in the
//     kernel at all: an axis-aligned cell only needs the tangents of its 2*NQ one-dimensional abscissae, and cells
//     of one mesh column/row share them, so they are tabulated once per (mesh, rule) for the distinct chart
//     intervals (k_tan_table) and read back through L1 (a few KB for a structured panel);
//   * 1/rho: MUFU.RSQ64H seed + one cubically convergent FP64 step, no range branches (rho^2 is in [1,3]);
//   * sum factorisation: first contract the y-direction into T11/T12/T22 (NQ*(2*NS+NB^2) FMAs per x-abscissa), then
//     the x-direction into the packed symmetric element matrix (4 FMAs per unique entry and abscissa);
//   * table operands sit in a compact __constant__ block laid out in use order ([abscissa][NN.. DD.. ND..], 1.7 KB for
//     Q2/degree 18, so it stays in the constant cache and adjacent operands can be fetched by one 128-bit uniform load);
//   * the symmetric half (M(M+1)/2 entries) is written entry-major [pair][cell]: consecutive threads write
//     consecutive addresses (fully coalesced FP64 stores); the atomic-free CSR fill is k_gather_chunks (gg_pattern.cu).
// The fused residual r_e = Ke u_e is formed from the registers that already hold Ke (LB is linear).
#pragma once
#include "gg_common.cuh"
#include "gg_refel.cuh"

namespace gg {

// generic 1-D tables of the resident (order, nq) pair (used by the action kernel); rewritten when the pair changes
__constant__ ScalarTables c_tab;
// compact pair tables of the resident (order, nq) pair: row q = [NN sym (NS) | DD sym (NS) | ND (NB*NB)], weights folded
__constant__ double c_lb[MAX_NQ * (MAX_NB * (MAX_NB + 1) + MAX_NB * MAX_NB)];
// packed-pair -> plane index of the symmetric entry-major buffer, and lexicographic -> Gridap local dof
__constant__ int32_t c_plane[(MAX_NB * MAX_NB) * (MAX_NB * MAX_NB + 1) / 2];
__constant__ int32_t c_lex2loc[MAX_NB * MAX_NB];

template <int NB>
__host__ __device__ constexpr int sym_index(int i, int j) {  // i <= j
  return i * NB - (i * (i - 1)) / 2 + (j - i);
}

// 1/sqrt(x) for x in a benign range: hardware seed (rel. error ~2^-22) + one cubic Newton step -> < 1 ulp
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(x, -(y * y), 1.0);
  const double p = fma(e, 0.375, 0.5);
  return fma(p, y * e, y);
}

// tangents of the 1-D abscissae of every distinct chart interval: out[i*nq + q] = tan(lo[i] + h[i] * xi[q])
__global__ void k_tan_table(int64_t n_int, int nq, const double* __restrict__ lo, const double* __restrict__ h,
                            const double* __restrict__ xi, double* __restrict__ out) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_int * nq) return;
  int64_t i = t / nq;
  int q = (int)(t - i * nq);
  out[t] = tan(fma(h[i], xi[q], lo[i]));
}

// EXTRINSIC: geometry from the 2x3 Jacobian of ambient o chart (pseudo-determinant path) instead of the analytic metric
// CLS: geometry-class mode -- thread t computes the element matrix of the representative cell cell_list[t] of class t and
// stores it in slot t of a buffer with `nc` = number of classes planes stride (no residual in this mode)
template <int NB, int NQ, bool EXTRINSIC, bool CLS = false>
__global__ void __launch_bounds__(128, 3)
k_lb_fast(int64_t nc, int64_t c0, int64_t c1, CellGeo g, const int32_t* __restrict__ cell_dofs,
          const double* __restrict__ u, double dirichlet, double* __restrict__ ebuf, double* __restrict__ vbuf,
          const int32_t* __restrict__ cell_list = nullptr) {
  constexpr int M = NB * NB, NS = NB * (NB + 1) / 2, NP = M * (M + 1) / 2;
  constexpr int ROW = 2 * NS + M;  // doubles per abscissa in c_lb
  constexpr int ONN = 0, ODD = NS, OND = 2 * NS;
  // cells [c0, c1) of the nc cells of the mesh (the numeric phase is pipelined over cell batches)
  const int64_t slot = c0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= c1) return;
  const int64_t c = CLS ? (int64_t)cell_list[slot] : slot;
  const double HX = g.hx[c], HY = g.hy[c];
  const double rx = HY / HX, ry = HX / HY;
  const double* __restrict__ ta = g.tx + (int64_t)g.ix[c] * NQ;
  const double* __restrict__ tb = g.ty + (int64_t)g.iy[c] * NQ;
  int panel = 0;
  if (EXTRINSIC) panel = g.chart[c];

  double b[NQ], sb[NQ];
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    b[j] = tb[j];
    sb[j] = fma(b[j], b[j], 1.0);
  }
  double acc[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) acc[p] = 0.0;

#pragma unroll 1
  for (int q1 = 0; q1 < NQ; ++q1) {
    const double a = ta[q1];
    const double sa = fma(a, a, 1.0);
    double T11[NS], T22[NS], T12[M];
#pragma unroll
    for (int k = 0; k < NS; ++k) { T11[k] = 0.0; T22[k] = 0.0; }
#pragma unroll
    for (int k = 0; k < M; ++k) T12[k] = 0.0;
#pragma unroll
    for (int q2 = 0; q2 < NQ; ++q2) {
      double G11, G12, G22;
      if (!EXTRINSIC) {
        const double rinv = fast_rsqrt(fma(b[q2], b[q2], sa));
        const double ar = a * rinv;
        G11 = sb[q2] * rinv;
        G12 = ar * b[q2];
        G22 = sa * rinv;
      } else {
        // chart-space factors of the reference-to-ambient Jacobian are applied through rx/ry below
        Metric2 m = extrinsic_metric_terms(panel, a, b[q2], g.radius);
        G11 = m.i11 * m.sqrtg; G12 = m.i12 * m.sqrtg; G22 = m.i22 * m.sqrtg;
      }
#pragma unroll
      for (int k = 0; k < NS; ++k) T11[k] = fma(G11, c_lb[q2 * ROW + ONN + k], T11[k]);
#pragma unroll
      for (int k = 0; k < NS; ++k) T22[k] = fma(G22, c_lb[q2 * ROW + ODD + k], T22[k]);
#pragma unroll
      for (int k = 0; k < M; ++k) T12[k] = fma(G12, c_lb[q2 * ROW + OND + k], T12[k]);
    }
#pragma unroll
    for (int k = 0; k < NS; ++k) { T11[k] *= rx; T22[k] *= ry; }
    const double* __restrict__ row = c_lb + q1 * ROW;  // warp-uniform
    // second contraction into the packed upper triangle, I = i1 + NB*i2 <= J = j1 + NB*j2
    int p = 0;
#pragma unroll
    for (int i2 = 0; i2 < NB; ++i2)
#pragma unroll
      for (int i1 = 0; i1 < NB; ++i1)
#pragma unroll
        for (int j2 = i2; j2 < NB; ++j2)
#pragma unroll
          for (int j1 = 0; j1 < NB; ++j1) {
            if (j2 == i2 && j1 < i1) continue;
            const int s1 = i1 <= j1 ? sym_index<NB>(i1, j1) : sym_index<NB>(j1, i1);
            const int s2 = sym_index<NB>(i2, j2);
            double v = acc[p];
            v = fma(row[ODD + s1], T11[s2], v);
            v = fma(row[OND + j1 * NB + i1], T12[i2 * NB + j2], v);
            v = fma(row[OND + i1 * NB + j1], T12[j2 * NB + i2], v);
            v = fma(row[ONN + s1], T22[s2], v);
            acc[p] = v;
            ++p;
          }
  }
  // coalesced entry-major store of the packed element matrix
  if (ebuf) {
#pragma unroll
    for (int p = 0; p < NP; ++p) ebuf[(int64_t)c_plane[p] * nc + slot] = acc[p];
  }
  if (!CLS && vbuf) {
    // fused residual r_e = Ke u_e (u gathered through the cell DOF ids; lexicographic <-> Gridap local order)
    double ue[M], re[M];
#pragma unroll
    for (int I = 0; I < M; ++I) {
      int32_t d = cell_dofs[c * M + c_lex2loc[I]];
      ue[I] = d >= 0 ? u[d] : dirichlet;
      re[I] = 0.0;
    }
    int p = 0;
#pragma unroll
    for (int I = 0; I < M; ++I)
#pragma unroll
      for (int J = I; J < M; ++J) {
        re[I] = fma(acc[p], ue[J], re[I]);
        if (J != I) re[J] = fma(acc[p], ue[I], re[J]);
        ++p;
      }
#pragma unroll
    for (int I = 0; I < M; ++I) vbuf[(int64_t)c_lex2loc[I] * nc + c] = re[I];
  }
}

// Residual r_e = Ke u_e from the element matrices stored once per geometry class (symmetric packed, entry-major [plane][class])
template <int NB>
__global__ void __launch_bounds__(128)
k_lb_residual_classes(int64_t nc, int64_t ncls, const int32_t* __restrict__ cls, const double* __restrict__ ebuf,
                      const int32_t* __restrict__ cell_dofs, const double* __restrict__ u, double dirichlet,
                      double* __restrict__ vbuf) {
  constexpr int M = NB * NB;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const int64_t k = cls[c];
  double ue[M], re[M];
#pragma unroll
  for (int I = 0; I < M; ++I) {
    const int32_t d = cell_dofs[c * M + c_lex2loc[I]];
    ue[I] = d >= 0 ? u[d] : dirichlet;
    re[I] = 0.0;
  }
  int p = 0;
#pragma unroll
  for (int I = 0; I < M; ++I)
#pragma unroll
    for (int J = I; J < M; ++J) {
      const double v = ebuf[(int64_t)c_plane[p] * ncls + k];
      re[I] = fma(v, ue[J], re[I]);
      if (J != I) re[J] = fma(v, ue[I], re[J]);
      ++p;
    }
#pragma unroll
  for (int I = 0; I < M; ++I) vbuf[(int64_t)c_lex2loc[I] * nc + c] = re[I];
}

// Matrix-free action r_e = a(u_h, phi_I) by sum factorisation, never forming Ke (3 contractions per abscissa)
template <int NB, int NQ>
__global__ void __launch_bounds__(128)
k_lb_action_fast(int64_t nc, CellGeo g, const int32_t* __restrict__ cell_dofs, const double* __restrict__ u,
                 double dirichlet, double* __restrict__ vbuf) {
  constexpr int M = NB * NB;
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const double HX = g.hx[c], HY = g.hy[c];
  const double rx = HY / HX, ry = HX / HY;
  const double* __restrict__ ta = g.tx + (int64_t)g.ix[c] * NQ;
  const double* __restrict__ tb = g.ty + (int64_t)g.iy[c] * NQ;
  double ue[M], re[M];
#pragma unroll
  for (int I = 0; I < M; ++I) {
    int32_t d = cell_dofs[c * M + c_lex2loc[I]];
    ue[I] = d >= 0 ? u[d] : dirichlet;
    re[I] = 0.0;
  }
  double b[NQ], sb[NQ];
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    b[j] = tb[j];
    sb[j] = fma(b[j], b[j], 1.0);
  }
#pragma unroll 1
  for (int q1 = 0; q1 < NQ; ++q1) {
    const double a = ta[q1];
    const double sa = fma(a, a, 1.0);
    const double w1 = c_tab.w[q1];
    double n1[NB], d1[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) { n1[i] = c_tab.N[q1][i]; d1[i] = c_tab.D[q1][i]; }
    double U[NB], Ux[NB], V1[NB], V2[NB];
#pragma unroll
    for (int i2 = 0; i2 < NB; ++i2) {
      U[i2] = 0.0; Ux[i2] = 0.0; V1[i2] = 0.0; V2[i2] = 0.0;
#pragma unroll
      for (int i1 = 0; i1 < NB; ++i1) {
        U[i2] = fma(n1[i1], ue[i1 + NB * i2], U[i2]);
        Ux[i2] = fma(d1[i1], ue[i1 + NB * i2], Ux[i2]);
      }
    }
#pragma unroll
    for (int q2 = 0; q2 < NQ; ++q2) {
      const double rinv = fast_rsqrt(fma(b[q2], b[q2], sa));
      const double ww = w1 * c_tab.w[q2] * rinv;
      const double G11 = sb[q2] * ww * rx, G12 = a * b[q2] * ww, G22 = sa * ww * ry;
      double ux = 0.0, uy = 0.0;
#pragma unroll
      for (int i2 = 0; i2 < NB; ++i2) {
        ux = fma(c_tab.N[q2][i2], Ux[i2], ux);
        uy = fma(c_tab.D[q2][i2], U[i2], uy);
      }
      const double F1 = fma(G11, ux, G12 * uy), F2 = fma(G12, ux, G22 * uy);
#pragma unroll
      for (int i2 = 0; i2 < NB; ++i2) {
        V1[i2] = fma(F1, c_tab.N[q2][i2], V1[i2]);
        V2[i2] = fma(F2, c_tab.D[q2][i2], V2[i2]);
      }
    }
#pragma unroll
    for (int i2 = 0; i2 < NB; ++i2)
#pragma unroll
      for (int i1 = 0; i1 < NB; ++i1)
        re[i1 + NB * i2] = fma(d1[i1], V1[i2], fma(n1[i1], V2[i2], re[i1 + NB * i2]));
  }
#pragma unroll
  for (int I = 0; I < M; ++I) vbuf[(int64_t)c_lex2loc[I] * nc + c] = re[I];
}

}  // namespace gg
