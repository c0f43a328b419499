// Shallow-water cell kernels with several lanes per cell (included by gg_swe.cu after gg_swe_sf.cuh).
//
// One thread per cell (k_swe_diag, k_swe_stage_sf) needs 170-255 registers for RT_2 / CG_3 and runs at 12-18 % occupancy: the FP64
// pipe waits on its own dependent chains.  Here a cell is owned by a group of G = 4 or 8 lanes of one warp and the sum-factorised
// evaluation is split along the rows of the tensor rule.
//
// The Raviart-Thomas basis needs no change of basis for that: every moment functional of RT_k on the square (build_rt_ref,
// gg_refel.cu) is a tensor product of 1-D functionals -- for the x-component Lambda_alpha (x) mu_m with
//   Lambda_0 f = -f(0), Lambda_1 f = f(1), Lambda_{2+i} f = int x^i f (i < k)   on P_{k+1},      mu_m f = int s^m f (m <= k) on P_k,
// and mu_m (x) Lambda_beta for the y-component -- so the dual basis is phi^x_{alpha,m} = A_alpha(x) B_m(y), phi^y_{m,beta} =
// B_m(x) A_beta(y) with the 1-D dual bases A, B (swe_sf_tabs checks this against the tabulated basis).  Hence
//   P0  the lanes of a group load the (signed) element DOFs, the Coriolis parameter and the tangents of the abscissae of their
//       cell into the warp's slice of shared memory (32-byte segments per cell and instruction: whole sectors);
//   P2  lane l owns the row q2 = l of the rule: it contracts the y-direction once (K + 1 or K + 2 FMAs per 1-D coefficient), then
//       walks the NQ points of its row with the 1-D tables as immediate constant operands, evaluates the metric and the nonlinear
//       coefficients and contracts the test side along x again: one short record per row, left in shared memory;
//   P3  the y-direction of the test side: a lane takes one column of the row records and produces the K + 1 or K + 2 entries that
//       share it (tables again immediate operands: one shared-memory read per K + 1 .. 2 (K + 2) FMAs);
//   P5  the lanes write the entry-major element vectors ([entry][cell]: the 4 or 8 cells of the warp are consecutive, one sector per
//       entry) and apply the inverse DG mass blocks, one row per lane.
// A warp never waits for another one: no __syncthreads, the 4 warps of a CTA only share the 1-D tables.
// (Measured and dropped: prefetch.global.L2 of the warp's next batch at the top of the loop, 0.375 -> 0.39 ms; and a register
// software pipeline -- indices of the next batch loaded before the row phase, values right after it, committed to shared memory
// at the top of the next iteration -- 0.355 -> 0.364 ms with 150 bytes more spills: the long-scoreboard samples on the P0 loads
// are warps that wait while others run, not lost issue slots.)
// The constant-coefficient terms -int Phi div(v) and int r div(F) ride on the same rows (div phi^x = A'_alpha B_m is a tensor
// product too), the Jacobian determinant cancels under the contravariant Piola map: plain reference weights, exact for NQ >= K + 1.
// Nothing is staged per quadrature point; 96 registers, 5 CTAs of 128 threads per SM.
// The forms are those of k_swe_diag / k_swe_stage (test/Transient/TransientShallowWater.jl:113-139).
#pragma once

namespace gg {

__constant__ double c_sf_w[SF_MAXNQ];        // 1-D Gauss weights on [0, 1]
__constant__ double c_sf_ra[SF_MAXNQ * 4];   // [q][alpha] A_alpha at the Gauss points
__constant__ double c_sf_rda[SF_MAXNQ * 4];  // ... A'_alpha
__constant__ double c_sf_rb[SF_MAXNQ * 3];   // [q][m] B_m
// Register budget: 5 CTAs of 128 threads per SM (96 registers).  At 80 registers (6 CTAs) the row of the diagnostics kernel spills
// 0.4-0.8 KB per thread, and with 200 KB of the SM's L1 carved out for shared memory those spills are served by L2: measured
// 0.70 (6 CTAs, point loop unrolled) / 0.43 (6 CTAs, rolled) / 0.37 ms (5 CTAs, rolled) per launch at 256^2 cells, RT_2, 7 x 7 points.
#ifndef SWE_LANES_MINB
#define SWE_LANES_MINB 5
#endif
// the point loop of the diagnostics row stays rolled (18 accumulators + 10 row coefficients live across it; unrolled, the
// scheduler hoists the 7 points' operands and spills them)
#define GG_Q1_LOOP _Pragma("unroll 1")

template <int K, int NQ_>
struct SweLanes {
  static constexpr int NQ = NQ_;
  static constexpr int G = NQ <= 4 ? 4 : 8;   // lanes per cell (NQ <= G)
  static constexpr int THREADS = 128, CPB = THREADS / G;
  static constexpr int MU = 2 * (K + 1) * (K + 2), MP = (K + 1) * (K + 1), MR = (K + 2) * (K + 2);
  static constexpr int NA = K + 2, NB = K + 1, Q = NQ * NQ;
  // 1-D tables a lane indexes by its own row, copied to shared memory once per CTA (same content as the __constant__ copies)
  static constexpr int T_LA = 0, T_DA = T_LA + NQ * NA, T_LB = T_DA + NQ * NA, T_RA = T_LB + NQ * NB, T_RDA = T_RA + NQ * NA,
                       T_RB = T_RDA + NQ * NA, T_W = T_RB + NQ * NB, T_SIZE = (T_W + NQ + 1) & ~1;
  static constexpr int cmax(int a, int b) { return a > b ? a : b; }
  // local RT DOF of the tensor coefficient (alpha, m) of the x-component / (m, beta) of the y-component (build_rt_ref order:
  // facets y = 0, y = 1, x = 0, x = 1 with K + 1 moments each, then the interior moments of the x-, then of the y-component)
  __host__ __device__ static constexpr int jx(int al, int m) {
    return al == 0 ? 2 * NB + m : al == 1 ? 3 * NB + m : 4 * NB + m * K + (al - 2);
  }
  __host__ __device__ static constexpr int jy(int m, int be) {
    return be == 0 ? m : be == 1 ? NB + m : 4 * NB + K * NB + (be - 2) * NB + m;
  }
  // ---- diagnostics: per-cell shared layout (doubles) ----
  static constexpr int D_RS = (3 * NA + 2 * NB) | 1;   // row record: A1[NA] A2[NA] Sx[NA] Sy[NB] SP[NB]
  static constexpr int D_R = cmax(NQ * D_RS, MU + MP); // rows; before the row phase: Ue [MU], Pe [MP]
  static constexpr int D_UE = 0, D_PE = MU;
  static constexpr int D_EF = D_R, D_AP = D_EF + MU, D_EQ = D_AP + MP, D_FQ = D_EQ + MR;              // Coriolis parameter at the points, overwritten by the depth (pq_out)
  static constexpr int D_GEO = D_FQ + Q;               // hx, hy, tangents of the abscissae in x [NQ] and y [NQ]
  // cell stride = 8 mod 16 doubles: every shared-memory access has the lanes of a group on consecutive (or the same) doubles, so
  // the two cells of a half-warp must sit half a bank row apart
  static constexpr int pad8(int raw) { return (raw + 7) / 16 * 16 + 8; }
  static constexpr int D_S = pad8(D_GEO + 2 + 2 * NQ);
  static constexpr size_t diag_smem() { return (size_t)(T_SIZE + CPB * D_S) * sizeof(double); }
  // ---- stage residual ----
  static constexpr int S_RS = (NA + 3 * NB) | 1;       // row record: Sx[NA] Sy[NB] SyD[NB] SPr[NB]
  static constexpr int S_FE = 0, S_UE = MU, S_QN = 2 * MU, S_QC = S_QN + MR, S_PE = S_QC + MR;
  static constexpr int S_R = cmax(NQ * S_RS, S_PE + MP);
  static constexpr int S_EF = S_R, S_RP = S_EF + MU, S_GEO = S_RP + MP, S_S = pad8(S_GEO + 2 + 2 * NQ);
  static constexpr size_t stage_smem() { return (size_t)(T_SIZE + CPB * S_S) * sizeof(double); }
};

// Right-hand sides of the three diagnostic equations (same contract as k_swe_diag)
template <int K, int NQ>
__global__ void __launch_bounds__(128, SWE_LANES_MINB)
k_swe_diag_lanes(int64_t nc, CellGeo g, const double* __restrict__ tabs, const int32_t* __restrict__ vdofs,
                 const int8_t* __restrict__ vsigns, const double* __restrict__ u, const double* __restrict__ p,
                 const double* __restrict__ fq, const double* __restrict__ bq, double gravity, const double* __restrict__ MpInv,
                 double* __restrict__ pq_out, double* __restrict__ eq, double* __restrict__ eF, double* __restrict__ Phi) {
  using C = SweLanes<K, NQ>;
  constexpr int G = C::G, CPB = C::CPB, TH = C::THREADS, MU = C::MU, MP = C::MP, MR = C::MR, NA = C::NA, NB = C::NB, Q = C::Q,
                S = C::D_S, RS = C::D_RS;
  extern __shared__ double lanes_sm[];
  double* s_tab = lanes_sm;
  double* s_cell = lanes_sm + C::T_SIZE;
  const int tid = threadIdx.x;
  for (int i = tid; i < C::T_SIZE; i += TH) s_tab[i] = tabs[i];
  __shared__ int s_lexp[MP], s_lexr[MR];   // Gridap local node order -> lexicographic, DG_K and CG_{K+1}
  if (tid < MP) s_lexp[tid] = sf_loc2lex(K, tid);
  if (tid < MR) s_lexr[tid] = sf_loc2lex(K + 1, tid);
  __syncthreads();
  const int cell = tid / G, l = tid % G;
  double* sc = s_cell + cell * S;
  const double r2 = g.radius * g.radius;
  constexpr int CW = 32 / G;                         // cells per warp
  const int64_t nwb = (nc + CW - 1) / CW;
  const int64_t wstride = (int64_t)gridDim.x * (TH / 32);
  for (int64_t wb = (int64_t)blockIdx.x * (TH / 32) + tid / 32; wb < nwb; wb += wstride) {
    const int64_t c = wb * CW + (tid % 32) / G;
    const bool valid = c < nc;
    // ---- P0: inputs of the group's cell ----
    int8_t sg[(MU + G - 1) / G];
#pragma unroll
    for (int m = 0; m < (MU + G - 1) / G; ++m) {
      const int j = l + G * m;
      sg[m] = 0;
      if (j < MU) {
        double v = 0.0;
        if (valid) { sg[m] = vsigns[c * MU + j]; v = (double)sg[m] * u[vdofs[c * MU + j]]; }
        sc[C::D_UE + j] = v;
      }
    }
#pragma unroll
    for (int m = 0; m < (MP + G - 1) / G; ++m) {
      const int i = l + G * m;
      if (i < MP) sc[C::D_PE + s_lexp[i]] = valid ? p[c * MP + i] : 0.0;
    }
#pragma unroll
    for (int m = 0; m < (Q + G - 1) / G; ++m) {
      const int q = l + G * m;
      if (q < Q) sc[C::D_FQ + q] = valid ? fq[c * Q + q] : 0.0;
    }
    if (valid) {
      if (l < NQ) {
        sc[C::D_GEO + 2 + l] = g.tx[(int64_t)g.ix[c] * NQ + l];
        sc[C::D_GEO + 2 + NQ + l] = g.ty[(int64_t)g.iy[c] * NQ + l];
      }
      if (l == 0) { sc[C::D_GEO] = g.hx[c]; sc[C::D_GEO + 1] = g.hy[c]; }
    }
    __syncwarp();
    // ---- P2: lane l walks the row q2 = l (shared memory and registers only) ----
    const bool act = l < NQ && valid;
    double A1[NA], A2[NA], Sx[NA], Sy[NB], SP[NB];
#pragma unroll
    for (int a = 0; a < NA; ++a) { A1[a] = 0.0; A2[a] = 0.0; Sx[a] = 0.0; }
#pragma unroll
    for (int a = 0; a < NB; ++a) { Sy[a] = 0.0; SP[a] = 0.0; }
    if (act) {
      double TUx[NA], TUy[NB], Tp[NB];
      {
        const double* __restrict__ ra2 = s_tab + C::T_RA + l * NA;
        const double* __restrict__ rb2 = s_tab + C::T_RB + l * NB;
        const double* __restrict__ lb2 = s_tab + C::T_LB + l * NB;
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          double su = 0.0;
#pragma unroll
          for (int m = 0; m < NB; ++m) su = fma(sc[C::D_UE + C::jx(a, m)], rb2[m], su);
          TUx[a] = su;
        }
#pragma unroll
        for (int m = 0; m < NB; ++m) {
          double su = 0.0, sp = 0.0;
#pragma unroll
          for (int be = 0; be < NA; ++be) su = fma(sc[C::D_UE + C::jy(m, be)], ra2[be], su);
#pragma unroll
          for (int bb = 0; bb < NB; ++bb) sp = fma(sc[C::D_PE + m + NB * bb], lb2[bb], sp);
          TUy[m] = su; Tp[m] = sp;
        }
      }
      const double HX = sc[C::D_GEO], HY = sc[C::D_GEO + 1], iHX = 1.0 / HX, iHY = 1.0 / HY;
      const double* __restrict__ ta = sc + C::D_GEO + 2;
      const double b = sc[C::D_GEO + 2 + NQ + l], sb = fma(b, b, 1.0), bb2 = b * b;
      const double w2 = s_tab[C::T_W + l] * HX * HY;
      double* __restrict__ fqc = sc + C::D_FQ + l * NQ;   // f in, depth out
      const double* __restrict__ bqc = bq ? bq + c * Q + l * NQ : nullptr;
      GG_Q1_LOOP
      for (int q1 = 0; q1 < NQ; ++q1) {
        const double* __restrict__ la1 = c_sf_la + q1 * NA;
        const double* __restrict__ da1 = c_sf_da + q1 * NA;
        const double* __restrict__ lb1 = c_sf_lb + q1 * NB;
        const double* __restrict__ ra1 = c_sf_ra + q1 * NA;
        const double* __restrict__ rb1 = c_sf_rb + q1 * NB;
        double Ux = 0.0, Uy = 0.0, ph = 0.0;
#pragma unroll
        for (int a = 0; a < NA; ++a) Ux = fma(TUx[a], ra1[a], Ux);
#pragma unroll
        for (int a = 0; a < NB; ++a) { Uy = fma(TUy[a], rb1[a], Uy); ph = fma(Tp[a], lb1[a], ph); }
        // metric from one rsqrt: sqrtg = r^2 sa sb / rho^3, g / sqrtg = [[sa, -ab], [-ab, sb]] / rho, ginv sqrtg = [[sb, ab], [ab, sa]] / rho
        const double t = ta[q1], sa = fma(t, t, 1.0), ab = t * b;
        const double ri = swe_rsqrt(sa + bb2);
        const double sqrtg = r2 * sa * sb * ri * ri * ri;
        const double ux = Ux * iHY, uy = Uy * iHX;   // contravariant Piola, axis-aligned cell
        const double wq = c_sf_w[q1] * w2, wr = wq * ri;
        // vorticity equation: f w sqrtg + ((R u).ginv).grad(w) sqrtg, R u = (-u_y, u_x)
        const double f0 = wq * sqrtg * fqc[q1];
        fqc[q1] = ph;
        const double fx = wr * (ux * ab - uy * sb) * iHX, fy = wr * (ux * sa - uy * ab) * iHY;
#pragma unroll
        for (int a = 0; a < NA; ++a) { A1[a] = fma(f0, la1[a], fma(fx, da1[a], A1[a])); A2[a] = fma(fy, la1[a], A2[a]); }
        // mass flux p u.(g v)/sqrtg, (g u)/sqrtg = (sa ux - ab uy, sb uy - ab ux) / rho
        const double gux = sa * ux - ab * uy, guy = sb * uy - ab * ux;
        const double cx = wr * ph * gux * iHY, cy = wr * ph * guy * iHX;
#pragma unroll
        for (int a = 0; a < NA; ++a) Sx[a] = fma(cx, ra1[a], Sx[a]);
        // Bernoulli potential: g_r (p + b) sqrtg + 1/2 u.(g u)/sqrtg
        const double bt = bqc ? bqc[q1] : 0.0;
        const double cP = wq * (gravity * (ph + bt) * sqrtg + 0.5 * ri * (ux * gux + uy * guy));
#pragma unroll
        for (int a = 0; a < NB; ++a) { Sy[a] = fma(cy, rb1[a], Sy[a]); SP[a] = fma(cP, lb1[a], SP[a]); }
      }
    }
    __syncwarp();   // the row records overlay Ue / Pe
    if (l < NQ) {
      double* __restrict__ rec = sc + l * RS;
#pragma unroll
      for (int a = 0; a < NA; ++a) { rec[a] = A1[a]; rec[NA + a] = A2[a]; rec[2 * NA + a] = Sx[a]; }
#pragma unroll
      for (int a = 0; a < NB; ++a) { rec[3 * NA + a] = Sy[a]; rec[3 * NA + NB + a] = SP[a]; }
    }
    __syncwarp();
    // ---- P3: y-direction of the test side, one column of the row records per task ----
#pragma unroll
    for (int t = l; t < 2 * NA + 2 * NB; t += G) {
      if (t < NA) {                       // CG_{K+1} element vector, lexicographic ix + NA iy
        double acc[NA];
#pragma unroll
        for (int iy = 0; iy < NA; ++iy) acc[iy] = 0.0;
#pragma unroll
        for (int q2 = 0; q2 < NQ; ++q2) {
          const double r1 = sc[q2 * RS + t], r2v = sc[q2 * RS + NA + t];
#pragma unroll
          for (int iy = 0; iy < NA; ++iy) acc[iy] = fma(r1, c_sf_la[q2 * NA + iy], fma(r2v, c_sf_da[q2 * NA + iy], acc[iy]));
        }
#pragma unroll
        for (int iy = 0; iy < NA; ++iy) sc[C::D_EQ + t + NA * iy] = acc[iy];
      } else if (t < 2 * NA) {            // x-component entries (alpha, m), m = 0..K
        const int al = t - NA;
        double acc[NB];
#pragma unroll
        for (int m = 0; m < NB; ++m) acc[m] = 0.0;
#pragma unroll
        for (int q2 = 0; q2 < NQ; ++q2) {
          const double r1 = sc[q2 * RS + 2 * NA + al];
#pragma unroll
          for (int m = 0; m < NB; ++m) acc[m] = fma(r1, c_sf_rb[q2 * NB + m], acc[m]);
        }
#pragma unroll
        for (int m = 0; m < NB; ++m) sc[C::D_EF + C::jx(al, m)] = acc[m];
      } else if (t < 2 * NA + NB) {       // y-component entries (m, beta), beta = 0..K+1
        const int m = t - 2 * NA;
        double acc[NA];
#pragma unroll
        for (int be = 0; be < NA; ++be) acc[be] = 0.0;
#pragma unroll
        for (int q2 = 0; q2 < NQ; ++q2) {
          const double r1 = sc[q2 * RS + 3 * NA + m];
#pragma unroll
          for (int be = 0; be < NA; ++be) acc[be] = fma(r1, c_sf_ra[q2 * NA + be], acc[be]);
        }
#pragma unroll
        for (int be = 0; be < NA; ++be) sc[C::D_EF + C::jy(m, be)] = acc[be];
      } else {                            // DG_K right-hand side of the Bernoulli potential, lexicographic a + NB b
        const int a = t - 2 * NA - NB;
        double acc[NB];
#pragma unroll
        for (int bb = 0; bb < NB; ++bb) acc[bb] = 0.0;
#pragma unroll
        for (int q2 = 0; q2 < NQ; ++q2) {
          const double r1 = sc[q2 * RS + 3 * NA + NB + a];
#pragma unroll
          for (int bb = 0; bb < NB; ++bb) acc[bb] = fma(r1, c_sf_lb[q2 * NB + bb], acc[bb]);
        }
#pragma unroll
        for (int bb = 0; bb < NB; ++bb) sc[C::D_AP + a + NB * bb] = acc[bb];
      }
    }
    __syncwarp();
    // ---- P5: entry-major element vectors, depth at the points, Phi = Mp^-1 aP ----
    if (valid) {
#pragma unroll
      for (int m = 0; m < (MU + G - 1) / G; ++m) {
        const int j = l + G * m;
        if (j < MU) eF[(int64_t)j * nc + c] = (double)sg[m] * sc[C::D_EF + j];
      }
#pragma unroll
      for (int m = 0; m < (MR + G - 1) / G; ++m) {
        const int i = l + G * m;
        if (i < MR) eq[(int64_t)i * nc + c] = sc[C::D_EQ + s_lexr[i]];
      }
#pragma unroll
      for (int m = 0; m < (Q + G - 1) / G; ++m) {
        const int q = l + G * m;
        if (q < Q) pq_out[c * Q + q] = sc[C::D_FQ + q];
      }
#pragma unroll
      for (int m = 0; m < (MP + G - 1) / G; ++m) {
        const int i = l + G * m;
        if (i < MP) {
          double acc = 0.0;
#pragma unroll
          for (int j = 0; j < MP; ++j) acc = fma(MpInv[(int64_t)(i * MP + j) * nc + c], sc[C::D_AP + sf_loc2lex(K, j)], acc);
          Phi[c * MP + i] = acc;
        }
      }
    }
    __syncwarp();
  }
}

// Stage residual (same contract as k_swe_stage_sf, but eru is written, not accumulated)
template <int K, bool ALIAS, int NQ>
__global__ void __launch_bounds__(128, SWE_LANES_MINB)
k_swe_stage_lanes(int64_t nc, CellGeo g, const double* __restrict__ tabs, const int32_t* __restrict__ vdofs,
                  const int8_t* __restrict__ vsigns, const int32_t* __restrict__ rdofs, const double* __restrict__ u,
                  const double* __restrict__ qv, const double* __restrict__ q0v, const double* __restrict__ Fv,
                  const double* __restrict__ Phi, const double* __restrict__ MpInv, double tau, double tau_over_dt,
                  double* __restrict__ eru, double* __restrict__ rp, double* __restrict__ kp) {
  using C = SweLanes<K, NQ>;
  constexpr int G = C::G, CPB = C::CPB, TH = C::THREADS, MU = C::MU, MP = C::MP, MR = C::MR, NA = C::NA, NB = C::NB, S = C::S_S,
                RS = C::S_RS;
  extern __shared__ double lanes_sm[];
  double* s_tab = lanes_sm;
  double* s_cell = lanes_sm + C::T_SIZE;
  const int tid = threadIdx.x;
  for (int i = tid; i < C::T_SIZE; i += TH) s_tab[i] = tabs[i];
  __shared__ int s_lexp[MP], s_lexr[MR];   // Gridap local node order -> lexicographic, DG_K and CG_{K+1}
  if (tid < MP) s_lexp[tid] = sf_loc2lex(K, tid);
  if (tid < MR) s_lexr[tid] = sf_loc2lex(K + 1, tid);
  __syncthreads();
  const int cell = tid / G, l = tid % G;
  double* sc = s_cell + cell * S;
  const double ir2 = 1.0 / (g.radius * g.radius);
  constexpr int CW = 32 / G;                         // cells per warp
  const int64_t nwb = (nc + CW - 1) / CW;
  const int64_t wstride = (int64_t)gridDim.x * (TH / 32);
  for (int64_t wb = (int64_t)blockIdx.x * (TH / 32) + tid / 32; wb < nwb; wb += wstride) {
    const int64_t c = wb * CW + (tid % 32) / G;
    const bool valid = c < nc;
    // ---- P0 ----
    int8_t sg[(MU + G - 1) / G];
#pragma unroll
    for (int m = 0; m < (MU + G - 1) / G; ++m) {
      const int j = l + G * m;
      sg[m] = 0;
      if (j < MU) {
        double vf = 0.0, vu = 0.0;
        if (valid) {
          sg[m] = vsigns[c * MU + j];
          const int32_t d = vdofs[c * MU + j];
          vf = (double)sg[m] * Fv[d]; vu = (double)sg[m] * u[d];
        }
        sc[C::S_FE + j] = vf;
        sc[C::S_UE + j] = vu;
      }
    }
#pragma unroll
    for (int m = 0; m < (MR + G - 1) / G; ++m) {
      const int i = l + G * m;
      if (i < MR) {
        double v = 0.0, vc = 0.0;
        if (valid) {
          const int32_t d = rdofs[c * MR + i];
          v = qv[d];
          // q - tau (q - q0)/dt = (1 - tau/dt) q + (tau/dt) q0
          if (!ALIAS) vc = fma(tau_over_dt, q0v[d] - v, v);
        }
        sc[C::S_QN + s_lexr[i]] = v;
        if (!ALIAS) sc[C::S_QC + s_lexr[i]] = vc;
      }
    }
#pragma unroll
    for (int m = 0; m < (MP + G - 1) / G; ++m) {
      const int i = l + G * m;
      if (i < MP) sc[C::S_PE + s_lexp[i]] = valid ? Phi[c * MP + i] : 0.0;
    }
    if (valid) {
      if (l < NQ) {
        sc[C::S_GEO + 2 + l] = g.tx[(int64_t)g.ix[c] * NQ + l];
        sc[C::S_GEO + 2 + NQ + l] = g.ty[(int64_t)g.iy[c] * NQ + l];
      }
      if (l == 0) { sc[C::S_GEO] = g.hx[c]; sc[C::S_GEO + 1] = g.hy[c]; }
    }
    __syncwarp();
    // ---- P2 ----
    // ---- P2 ----
    const bool act = l < NQ && valid;
    double Sx[NA], Sy[NB], SyD[NB], SPr[NB];
#pragma unroll
    for (int a = 0; a < NA; ++a) Sx[a] = 0.0;
#pragma unroll
    for (int a = 0; a < NB; ++a) { Sy[a] = 0.0; SyD[a] = 0.0; SPr[a] = 0.0; }
    if (act) {
      const double* __restrict__ la2 = s_tab + C::T_LA + l * NA;
      const double* __restrict__ da2 = s_tab + C::T_DA + l * NA;
      const double* __restrict__ lb2 = s_tab + C::T_LB + l * NB;
      const double* __restrict__ ra2 = s_tab + C::T_RA + l * NA;
      const double* __restrict__ rda2 = s_tab + C::T_RDA + l * NA;
      const double* __restrict__ rb2 = s_tab + C::T_RB + l * NB;
      const double HX = sc[C::S_GEO], HY = sc[C::S_GEO + 1], iHX = 1.0 / HX, iHY = 1.0 / HY, ih = iHX * iHY;
      const double w2r = s_tab[C::T_W + l];
      // pass A: the scalar coefficient w (q - tau (q - q0)/dt - tau u.grad(q)/sqrtg) at the points of the row (the fields of u and q
      // are dead before the ones of F come alive: the row fits in 80 registers)
      double coefq[NQ];
      {
        double TUx[NA], TUy[NB], Tq[NA], Tqy[NA], Tqc[ALIAS ? 1 : NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          double su = 0.0, sq = 0.0, sy = 0.0, sq0 = 0.0;
#pragma unroll
          for (int m = 0; m < NB; ++m) su = fma(sc[C::S_UE + C::jx(a, m)], rb2[m], su);
#pragma unroll
          for (int i2 = 0; i2 < NA; ++i2) {
            const double qn = sc[C::S_QN + a + NA * i2];
            sq = fma(qn, la2[i2], sq);
            sy = fma(qn, da2[i2], sy);
            if (!ALIAS) sq0 = fma(sc[C::S_QC + a + NA * i2], la2[i2], sq0);
          }
          TUx[a] = su; Tq[a] = sq; Tqy[a] = sy;
          if (!ALIAS) Tqc[a] = sq0;
        }
#pragma unroll
        for (int m = 0; m < NB; ++m) {
          double su = 0.0;
#pragma unroll
          for (int be = 0; be < NA; ++be) su = fma(sc[C::S_UE + C::jy(m, be)], ra2[be], su);
          TUy[m] = su;
        }
        const double* __restrict__ ta = sc + C::S_GEO + 2;
        const double b = sc[C::S_GEO + 2 + NQ + l], bb2 = b * b, isb = ir2 / fma(b, b, 1.0);
        const double w2 = w2r * HX * HY * ih;
#pragma unroll
        for (int q1 = 0; q1 < NQ; ++q1) {
          const double* __restrict__ la1 = c_sf_la + q1 * NA;
          const double* __restrict__ da1 = c_sf_da + q1 * NA;
          const double* __restrict__ ra1 = c_sf_ra + q1 * NA;
          const double* __restrict__ rb1 = c_sf_rb + q1 * NB;
          double Ux = 0.0, qq = 0.0, qx = 0.0, qy = 0.0, Uy = 0.0;
#pragma unroll
          for (int a = 0; a < NA; ++a) {
            Ux = fma(TUx[a], ra1[a], Ux);
            qq = fma(ALIAS ? Tq[a] : Tqc[a], la1[a], qq);
            qx = fma(Tq[a], da1[a], qx);
            qy = fma(Tqy[a], la1[a], qy);
          }
#pragma unroll
          for (int a = 0; a < NB; ++a) Uy = fma(TUy[a], rb1[a], Uy);
          // 1 / sqrtg = rho^3 / (r^2 sa sb)
          const double t = ta[q1], sa = fma(t, t, 1.0), rho2 = sa + bb2;
          const double isqrtg = isb * rho2 * rho2 * swe_rsqrt(rho2) / sa;
          // u.grad(q) in chart coordinates: (Ux/HY)(qx/HX) + (Uy/HX)(qy/HY)
          const double ugq = (Ux * qx + Uy * qy) * ih;
          coefq[q1] = c_sf_w[q1] * w2 * (qq - tau * ugq * isqrtg);
        }
      }
      // pass B: (R F).v_I with F = (Fx/HY, Fy/HX), v_I = (vx/HY, vy/HX): (-Fy/HX)(vx/HY) + (Fx/HY)(vy/HX); and the reference
      // divergence of F against the DG basis (int r div(F))
      {
        double TFx[NA], TFy[NB], TFd[NB];
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          double sf = 0.0;
#pragma unroll
          for (int m = 0; m < NB; ++m) sf = fma(sc[C::S_FE + C::jx(a, m)], rb2[m], sf);
          TFx[a] = sf;
        }
#pragma unroll
        for (int m = 0; m < NB; ++m) {
          double sf = 0.0, sd = 0.0;
#pragma unroll
          for (int be = 0; be < NA; ++be) {
            const double fe = sc[C::S_FE + C::jy(m, be)];
            sf = fma(fe, ra2[be], sf);
            sd = fma(fe, rda2[be], sd);
          }
          TFy[m] = sf; TFd[m] = sd;
        }
#pragma unroll
        for (int q1 = 0; q1 < NQ; ++q1) {
          const double* __restrict__ lb1 = c_sf_lb + q1 * NB;
          const double* __restrict__ ra1 = c_sf_ra + q1 * NA;
          const double* __restrict__ rda1 = c_sf_rda + q1 * NA;
          const double* __restrict__ rb1 = c_sf_rb + q1 * NB;
          double Fx = 0.0, Fy = 0.0, dv = 0.0;
#pragma unroll
          for (int a = 0; a < NA; ++a) { Fx = fma(TFx[a], ra1[a], Fx); dv = fma(TFx[a], rda1[a], dv); }
#pragma unroll
          for (int a = 0; a < NB; ++a) { Fy = fma(TFy[a], rb1[a], Fy); dv = fma(TFd[a], rb1[a], dv); }
          const double fx = -coefq[q1] * Fy, fy = coefq[q1] * Fx, wd = c_sf_w[q1] * w2r * dv;
#pragma unroll
          for (int a = 0; a < NA; ++a) Sx[a] = fma(fx, ra1[a], Sx[a]);
#pragma unroll
          for (int a = 0; a < NB; ++a) { Sy[a] = fma(fy, rb1[a], Sy[a]); SPr[a] = fma(wd, lb1[a], SPr[a]); }
        }
      }
      // pass C: -int Phi div(v)
      {
        double TP[NB];
#pragma unroll
        for (int a = 0; a < NB; ++a) {
          double sp = 0.0;
#pragma unroll
          for (int bb = 0; bb < NB; ++bb) sp = fma(sc[C::S_PE + a + NB * bb], lb2[bb], sp);
          TP[a] = sp;
        }
#pragma unroll
        for (int q1 = 0; q1 < NQ; ++q1) {
          const double* __restrict__ lb1 = c_sf_lb + q1 * NB;
          const double* __restrict__ rda1 = c_sf_rda + q1 * NA;
          const double* __restrict__ rb1 = c_sf_rb + q1 * NB;
          double ph = 0.0;
#pragma unroll
          for (int a = 0; a < NB; ++a) ph = fma(TP[a], lb1[a], ph);
          const double wp = -c_sf_w[q1] * w2r * ph;
#pragma unroll
          for (int a = 0; a < NA; ++a) Sx[a] = fma(wp, rda1[a], Sx[a]);
#pragma unroll
          for (int a = 0; a < NB; ++a) SyD[a] = fma(wp, rb1[a], SyD[a]);
        }
      }
    }
    __syncwarp();   // the row records overlay the element DOFs
    if (l < NQ) {
      double* __restrict__ rec = sc + l * RS;
#pragma unroll
      for (int a = 0; a < NA; ++a) rec[a] = Sx[a];
#pragma unroll
      for (int a = 0; a < NB; ++a) { rec[NA + a] = Sy[a]; rec[NA + NB + a] = SyD[a]; rec[NA + 2 * NB + a] = SPr[a]; }
    }
    __syncwarp();
    // ---- P3 ----
#pragma unroll
    for (int t = l; t < NA + 2 * NB; t += G) {
      if (t < NA) {                       // x-component entries (alpha, m)
        double acc[NB];
#pragma unroll
        for (int m = 0; m < NB; ++m) acc[m] = 0.0;
#pragma unroll
        for (int q2 = 0; q2 < NQ; ++q2) {
          const double r1 = sc[q2 * RS + t];
#pragma unroll
          for (int m = 0; m < NB; ++m) acc[m] = fma(r1, c_sf_rb[q2 * NB + m], acc[m]);
        }
#pragma unroll
        for (int m = 0; m < NB; ++m) sc[C::S_EF + C::jx(t, m)] = acc[m];
      } else if (t < NA + NB) {           // y-component entries (m, beta)
        const int m = t - NA;
        double acc[NA];
#pragma unroll
        for (int be = 0; be < NA; ++be) acc[be] = 0.0;
#pragma unroll
        for (int q2 = 0; q2 < NQ; ++q2) {
          const double r1 = sc[q2 * RS + NA + m], r2v = sc[q2 * RS + NA + NB + m];
#pragma unroll
          for (int be = 0; be < NA; ++be) acc[be] = fma(r1, c_sf_ra[q2 * NA + be], fma(r2v, c_sf_rda[q2 * NA + be], acc[be]));
        }
#pragma unroll
        for (int be = 0; be < NA; ++be) sc[C::S_EF + C::jy(m, be)] = acc[be];
      } else {                            // r_p, lexicographic a + NB b
        const int a = t - NA - NB;
        double acc[NB];
#pragma unroll
        for (int bb = 0; bb < NB; ++bb) acc[bb] = 0.0;
#pragma unroll
        for (int q2 = 0; q2 < NQ; ++q2) {
          const double r1 = sc[q2 * RS + NA + 2 * NB + a];
#pragma unroll
          for (int bb = 0; bb < NB; ++bb) acc[bb] = fma(r1, c_sf_lb[q2 * NB + bb], acc[bb]);
        }
#pragma unroll
        for (int bb = 0; bb < NB; ++bb) sc[C::S_RP + a + NB * bb] = acc[bb];
      }
    }
    __syncwarp();
    // ---- P5 ----
    if (valid) {
#pragma unroll
      for (int m = 0; m < (MU + G - 1) / G; ++m) {
        const int j = l + G * m;
        if (j < MU) eru[(int64_t)j * nc + c] = (double)sg[m] * sc[C::S_EF + j];
      }
#pragma unroll
      for (int m = 0; m < (MP + G - 1) / G; ++m) {
        const int i = l + G * m;
        if (i < MP) {
          rp[c * MP + i] = sc[C::S_RP + s_lexp[i]];
          if (kp) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < MP; ++j) acc = fma(MpInv[(int64_t)(i * MP + j) * nc + c], sc[C::S_RP + sf_loc2lex(K, j)], acc);
            kp[c * MP + i] = -acc;
          }
        }
      }
    }
    __syncwarp();
  }
}

}  // namespace gg
