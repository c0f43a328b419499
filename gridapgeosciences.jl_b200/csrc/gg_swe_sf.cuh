// Sum-factorised variants of the shallow-water cell kernels (included by gg_swe.cu).
//
// The direct kernels (k_swe_diag, k_swe_stage) evaluate every field at every quadrature point from the tabulated basis:
// 2 x MU FMAs per RT field and point, 3 x MR per CG field, and the same again on the test side -- 9.4 kFMA per cell for
// RT_2 / CG_3 at 7 x 7 points.  All three element families are tensor products in the reference coordinates:
//   RT_k  component x in Q_{k+1,k}, component y in Q_{k,k+1};   CG_{k+1} in Q_{k+1,k+1};   DG_k in Q_{k,k}.
// So a field is first expressed by its values at the tensor Lagrange nodes of its polynomial space (for RT: 2 x MU/2 x MU FMAs
// with the node values of the Gridap basis functions, a well-conditioned change of basis, unlike monomial coefficients), then
// interpolated row by row: contract the y-direction once per row of points, the x-direction per point (K + 2 FMAs per field
// component instead of MU).  The test side runs the same contractions transposed.  4.1 kFMA per cell instead of 9.4.
#pragma once

namespace gg {

constexpr int SF_MAXNQ = 16;
__constant__ double c_sf_phix[12 * 24];   // [node b * (K+2) + a][j]: x-component of reference RT basis function j at node (x_a, y_b)
__constant__ double c_sf_phiy[12 * 24];   // [node b' * (K+1) + a'][j]: y-component at node (x_a', y_b'), b' = 0..K+1
__constant__ double c_sf_la[SF_MAXNQ * 4];   // [q][a] 1-D Lagrange basis of order K+1 at the Gauss points
__constant__ double c_sf_da[SF_MAXNQ * 4];   // ... its derivative
__constant__ double c_sf_lb[SF_MAXNQ * 3];   // [q][b] 1-D Lagrange basis of order K (K = 0: the constant 1)

// Gridap local node order of Q_P on the square -> lexicographic ix + (P+1) iy (same as quad_local_to_lex)
__host__ __device__ constexpr int sf_loc2lex(int P, int a) {
  if (P == 0) return 0;
  const int nb = P + 1, ne = P - 1;
  if (a == 0) return 0;
  if (a == 1) return P;
  if (a == 2) return nb * P;
  if (a == 3) return P + nb * P;
  int e = a - 4;
  if (e < ne) return e + 1;
  e -= ne;
  if (e < ne) return e + 1 + nb * P;
  e -= ne;
  if (e < ne) return nb * (e + 1);
  e -= ne;
  if (e < ne) return P + nb * (e + 1);
  e -= ne;
  return (e % ne + 1) + nb * (e / ne + 1);
}

// node values of an RT field from its (signed) element DOFs
template <int K>
__device__ __forceinline__ void sf_rt_nodes(const double (&e)[2 * (K + 1) * (K + 2)], double (&nx)[(K + 1) * (K + 2)],
                                            double (&ny)[(K + 1) * (K + 2)]) {
  constexpr int MU = 2 * (K + 1) * (K + 2), NX = (K + 1) * (K + 2);
#pragma unroll
  for (int i = 0; i < NX; ++i) { nx[i] = 0.0; ny[i] = 0.0; }
#pragma unroll
  for (int j = 0; j < MU; ++j)
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      nx[i] = fma(c_sf_phix[i * MU + j], e[j], nx[i]);
      ny[i] = fma(c_sf_phiy[i * MU + j], e[j], ny[i]);
    }
}

// Stage residual (same contract as k_swe_stage): eru [MU][nc], rp, kp
// NQ > 0: points per direction known at compile time -- the x-loops are unrolled, so the 1-D tables become immediate constant
// operands (no LDC on the critical path of every point) and the points of a row are in flight together; NQ = 0: any rule.
template <int K, bool ALIAS, int NQ>
__global__ void __launch_bounds__(64, 4)
k_swe_stage_sf(int64_t nc, int nq_rt, CellGeo g, const double* __restrict__ w1, const int32_t* __restrict__ vdofs,
               const int8_t* __restrict__ vsigns, const int32_t* __restrict__ rdofs, const double* __restrict__ u,
               const double* __restrict__ qv, const double* __restrict__ q0v, const double* __restrict__ Fv,
               const double* __restrict__ Phi, const double* __restrict__ Bhat, const double* __restrict__ MpInv, double tau,
               double tau_over_dt, double* __restrict__ eru, double* __restrict__ rp, double* __restrict__ kp,
               double* __restrict__ gscratch) {
  constexpr int MU = 2 * (K + 1) * (K + 2), MP = (K + 1) * (K + 1), MR = (K + 2) * (K + 2);
  constexpr int NA = K + 2, NB = K + 1, NX = NA * NB;
  constexpr int UQ = NQ > 0 ? NQ : 1;
  const int nq = NQ > 0 ? NQ : nq_rt;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const double HX = g.hx[c], HY = g.hy[c], iHX = 1.0 / HX, iHY = 1.0 / HY;
  const double* __restrict__ ta = g.tx + (int64_t)g.ix[c] * nq;
  const double* __restrict__ tb = g.ty + (int64_t)g.iy[c] * nq;
  Stage2 stg = stage2_of(gscratch, nc, c);     // shared memory ([point][component][thread]) or the global scratch
  double Fnx[NX], Fny[NX];
  {
    double Fe[MU];
#pragma unroll
    for (int j = 0; j < MU; ++j) Fe[j] = (double)vsigns[c * MU + j] * Fv[vdofs[c * MU + j]];
    // linear, constant-coefficient parts through the reference divergence matrix: -int Phi div(v), int r div(F)
    {
      double pe[MP], r[MP];
#pragma unroll
      for (int i = 0; i < MP; ++i) { pe[i] = Phi[c * MP + i]; r[i] = 0.0; }
#pragma unroll
      for (int j = 0; j < MU; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < MP; ++i) {
          const double bij = Bhat[i * MU + j];
          acc = fma(bij, pe[i], acc);
          r[i] = fma(bij, Fe[j], r[i]);
        }
        eru[(int64_t)j * nc + c] = -(double)vsigns[c * MU + j] * acc;
      }
#pragma unroll
      for (int i = 0; i < MP; ++i) rp[c * MP + i] = r[i];
      if (kp) {
#pragma unroll
        for (int i = 0; i < MP; ++i) {
          double acc = 0.0;
#pragma unroll
          for (int j = 0; j < MP; ++j) acc = fma(MpInv[(int64_t)(i * MP + j) * nc + c], r[j], acc);
          kp[c * MP + i] = -acc;
        }
      }
    }
    sf_rt_nodes<K>(Fe, Fnx, Fny);
  }
  // ---- pass 1: coefficient of (R F).v at every point --------------------------------------------------------------------------
  {
    double Unx[NX], Uny[NX];
    {
      double ue[MU];
#pragma unroll
      for (int j = 0; j < MU; ++j) ue[j] = (double)vsigns[c * MU + j] * u[vdofs[c * MU + j]];
      sf_rt_nodes<K>(ue, Unx, Uny);
    }
    double qn[MR], qc[ALIAS ? 1 : MR];   // nodal values of q (and of q - tau (q - q0)/dt), lexicographic
#pragma unroll
    for (int i = 0; i < MR; ++i) {
      const int32_t d = rdofs[c * MR + i];
      const double v = qv[d];
      qn[sf_loc2lex(K + 1, i)] = v;
      if (!ALIAS) qc[sf_loc2lex(K + 1, i)] = fma(tau_over_dt, q0v[d] - v, v);
    }
    const double ir2 = 1.0 / (g.radius * g.radius);
#pragma unroll 1
    for (int q2 = 0; q2 < nq; ++q2) {
      const double b = tb[q2], isb = 1.0 / fma(b, b, 1.0);
      const double* __restrict__ la2 = c_sf_la + q2 * NA;
      const double* __restrict__ da2 = c_sf_da + q2 * NA;
      const double* __restrict__ lb2 = c_sf_lb + q2 * NB;
      double TUx[NA], TFx[NA], TUy[NB], TFy[NB], Tq[NA], Tqy[NA], Tqc[ALIAS ? 1 : NA];
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        double su = 0.0, sf = 0.0, sq = 0.0, sy = 0.0, sc = 0.0;
#pragma unroll
        for (int bb = 0; bb < NB; ++bb) { su = fma(Unx[bb * NA + a], lb2[bb], su); sf = fma(Fnx[bb * NA + a], lb2[bb], sf); }
#pragma unroll
        for (int i2 = 0; i2 < NA; ++i2) {
          sq = fma(qn[a + NA * i2], la2[i2], sq);
          sy = fma(qn[a + NA * i2], da2[i2], sy);
          if (!ALIAS) sc = fma(qc[a + NA * i2], la2[i2], sc);
        }
        TUx[a] = su; TFx[a] = sf; Tq[a] = sq; Tqy[a] = sy;
        if (!ALIAS) Tqc[a] = sc;
      }
#pragma unroll
      for (int a = 0; a < NB; ++a) {
        double su = 0.0, sf = 0.0;
#pragma unroll
        for (int bb = 0; bb < NA; ++bb) { su = fma(Uny[bb * NB + a], la2[bb], su); sf = fma(Fny[bb * NB + a], la2[bb], sf); }
        TUy[a] = su; TFy[a] = sf;
      }
#pragma unroll UQ
      for (int q1 = 0; q1 < nq; ++q1) {
        const double* __restrict__ la1 = c_sf_la + q1 * NA;
        const double* __restrict__ da1 = c_sf_da + q1 * NA;
        const double* __restrict__ lb1 = c_sf_lb + q1 * NB;
        double Ux = 0.0, Fx = 0.0, qq = 0.0, qx = 0.0, qy = 0.0, Uy = 0.0, Fy = 0.0;
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          Ux = fma(TUx[a], la1[a], Ux); Fx = fma(TFx[a], la1[a], Fx);
          qq = fma(ALIAS ? Tq[a] : Tqc[a], la1[a], qq);
          qx = fma(Tq[a], da1[a], qx);
          qy = fma(Tqy[a], la1[a], qy);
        }
#pragma unroll
        for (int a = 0; a < NB; ++a) { Uy = fma(TUy[a], lb1[a], Uy); Fy = fma(TFy[a], lb1[a], Fy); }
        // 1 / sqrtg = rho^3 / (r^2 sa sb)
        const double t = ta[q1], sa = fma(t, t, 1.0), rho2 = sa + b * b;
        const double isqrtg = ir2 * isb * rho2 * rho2 * swe_rsqrt(rho2) / sa;
        // u.grad(q) in chart coordinates: (Ux/HY)(qx/HX) + (Uy/HX)(qy/HY)
        const double ugq = (Ux * qx + Uy * qy) * iHX * iHY;
        const double coef = w1[q1] * w1[q2] * HX * HY * (qq - tau * ugq * isqrtg);
        const int q = q1 + nq * q2;
        stg.at(q, 0) = -coef * Fy * iHX * iHY;
        stg.at(q, 1) = coef * Fx * iHX * iHY;
      }
    }
  }
  // ---- pass 2: moments of the staged coefficients against the nodal bases, then back to the RT basis ---------------------------
  double mX[NX], mY[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) { mX[i] = 0.0; mY[i] = 0.0; }
#pragma unroll 1
  for (int q2 = 0; q2 < nq; ++q2) {
    const double* __restrict__ la2 = c_sf_la + q2 * NA;
    const double* __restrict__ lb2 = c_sf_lb + q2 * NB;
    double Sx[NA], Sy[NB];
#pragma unroll
    for (int a = 0; a < NA; ++a) Sx[a] = 0.0;
#pragma unroll
    for (int a = 0; a < NB; ++a) Sy[a] = 0.0;
#pragma unroll UQ
    for (int q1 = 0; q1 < nq; ++q1) {
      const int q = q1 + nq * q2;
      const double fx = stg.at(q, 0), fy = stg.at(q, 1);
      const double* __restrict__ la1 = c_sf_la + q1 * NA;
      const double* __restrict__ lb1 = c_sf_lb + q1 * NB;
#pragma unroll
      for (int a = 0; a < NA; ++a) Sx[a] = fma(fx, la1[a], Sx[a]);
#pragma unroll
      for (int a = 0; a < NB; ++a) Sy[a] = fma(fy, lb1[a], Sy[a]);
    }
#pragma unroll
    for (int bb = 0; bb < NB; ++bb)
#pragma unroll
      for (int a = 0; a < NA; ++a) mX[bb * NA + a] = fma(Sx[a], lb2[bb], mX[bb * NA + a]);
#pragma unroll
    for (int bb = 0; bb < NA; ++bb)
#pragma unroll
      for (int a = 0; a < NB; ++a) mY[bb * NB + a] = fma(Sy[a], la2[bb], mY[bb * NB + a]);
  }
#pragma unroll
  for (int j = 0; j < MU; ++j) {
    double r = 0.0;
#pragma unroll
    for (int i = 0; i < NX; ++i) r = fma(c_sf_phix[i * MU + j], mX[i], fma(c_sf_phiy[i * MU + j], mY[i], r));
    eru[(int64_t)j * nc + c] += (double)vsigns[c * MU + j] * r;
  }
}

}  // namespace gg
