// 3-D shell: Laplace-Beltrami element matrices of hexahedral Q_p cells by an exact Kronecker factorisation.
//
// What the reference computes per hexahedron (test/Laplacian/LaplaceBeltrami.jl:47 on AtlasDiscreteModel(
// CubedSphereWithThicknessMesh(r, t), l), geometry src/Fields/CubedSphereWithThickness{Metric,InvMetric}.jl):
//     Ke[I,J] = sum over nq^3 points of  w |det grad(cell_map)| grad(phi_I).(ginv grad(phi_J)) sqrtg      (1000 points at degree 18)
// with the block-diagonal shell metric  ginv = diag(1/t^2, gsurf^-1 / s^2),  sqrtg = t s^2 sqrtg_surf,  s = 1 + t gamma / r.
// So  ginv sqrtg = diag( (s^2/t) sqrtg_surf ,  t gsurf^-1 sqrtg_surf ):  the (gamma, gamma) entry is a product f(gamma) h(alpha, beta)
// and the surface block does not depend on gamma at all (s^2 cancels).  With phi_I = N_i0(gamma) phi2_A(alpha, beta) the tensor
// Gauss rule factorises EXACTLY:
//     Ke[(i0,A),(j0,B)] = (1/t) Dg[i0,j0] M2[A,B] + t Ng[i0,j0] K2[A,B]
//     Dg = int N'_i0 N'_j0 s^2 / h_gamma   Ng = int N_i0 N_j0 h_gamma   (per layer, 1-D rule)
//     M2 = int phi2_A phi2_B sqrtg_surf    K2 = int grad(phi2_A).(gsurf^-1 grad(phi2_B)) sqrtg_surf   (per column, 2-D rule)
// i.e. the 4.9 MFLOP per cell of SURVEY.md §8(d) collapse to 2 FMAs per entry; the kernel is bound by writing the 378 packed
// entries (3 KB per cell).  K2 comes from the 2-D hot kernel (gg_lb_fast.cuh), M2 from the generic scalar kernel; both are
// re-integrated on every call (6 n^2 columns against 6 n^2 L cells).
#pragma once
#include "gg_common.cuh"

namespace gg {

// tables of the resident order: hex local index of (i0, A) and the packed plane of ((i0,A),(j0,B)) for A <= B (Gridap orders)
__constant__ int32_t c_hex_t3[3 * 9];
__constant__ int32_t c_hex_plane[45 * 9];

template <int NB>
__global__ void __launch_bounds__(128)
k_hex_kron(int64_t nc, int64_t ncol, const int32_t* __restrict__ col, const int32_t* __restrict__ layer,
           const double* __restrict__ K2, const double* __restrict__ M2, const double* __restrict__ Dg,
           const double* __restrict__ Ng, double t, double* __restrict__ ebuf3) {
  constexpr int M2N = NB * NB, G = NB * NB;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const int64_t cl = col[c];
  const double* __restrict__ dg = Dg + (int64_t)layer[c] * G;
  const double* __restrict__ ng = Ng + (int64_t)layer[c] * G;
  double dl[G], nl[G];
  const double it = 1.0 / t;
#pragma unroll
  for (int k = 0; k < G; ++k) { dl[k] = it * dg[k]; nl[k] = t * ng[k]; }
  int p2 = 0;
#pragma unroll
  for (int a = 0; a < M2N; ++a)
#pragma unroll
    for (int b = a; b < M2N; ++b) {
      const double k2 = K2[(int64_t)p2 * ncol + cl];          // packed symmetric, entry-major
      const double m2 = M2[(int64_t)(a * M2N + b) * ncol + cl];  // full, entry-major
#pragma unroll
      for (int k0 = 0; k0 < G; ++k0) {
        const int pl = c_hex_plane[p2 * G + k0];
        if (pl >= 0) ebuf3[(int64_t)pl * nc + c] = fma(dl[k0], m2, nl[k0] * k2);
      }
      ++p2;
    }
}

// r_e = Ke u_e through the same factorisation (never forming Ke)
template <int NB>
__global__ void __launch_bounds__(128)
k_hex_residual(int64_t nc, int64_t ncol, const int32_t* __restrict__ col, const int32_t* __restrict__ layer,
               const double* __restrict__ K2, const double* __restrict__ M2, const double* __restrict__ Dg,
               const double* __restrict__ Ng, double t, const int32_t* __restrict__ cell_dofs, const double* __restrict__ u,
               double dirichlet, double* __restrict__ vbuf3) {
  constexpr int M2N = NB * NB, M3N = NB * NB * NB, G = NB * NB;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const int64_t cl = col[c];
  const double* __restrict__ dg = Dg + (int64_t)layer[c] * G;
  const double* __restrict__ ng = Ng + (int64_t)layer[c] * G;
  double U[NB][M2N], WK[NB][M2N], WM[NB][M2N];
#pragma unroll
  for (int i0 = 0; i0 < NB; ++i0)
#pragma unroll
    for (int A = 0; A < M2N; ++A) {
      const int32_t d = cell_dofs[c * M3N + c_hex_t3[i0 * M2N + A]];
      U[i0][A] = d >= 0 ? u[d] : dirichlet;
      WK[i0][A] = 0.0; WM[i0][A] = 0.0;
    }
  int p2 = 0;
#pragma unroll
  for (int a = 0; a < M2N; ++a)
#pragma unroll
    for (int b = a; b < M2N; ++b) {
      const double k2 = K2[(int64_t)p2 * ncol + cl];
      const double m2 = M2[(int64_t)(a * M2N + b) * ncol + cl];
#pragma unroll
      for (int i0 = 0; i0 < NB; ++i0) {
        WK[i0][a] = fma(k2, U[i0][b], WK[i0][a]);
        WM[i0][a] = fma(m2, U[i0][b], WM[i0][a]);
        if (a != b) {
          WK[i0][b] = fma(k2, U[i0][a], WK[i0][b]);
          WM[i0][b] = fma(m2, U[i0][a], WM[i0][b]);
        }
      }
      ++p2;
    }
  const double it = 1.0 / t;
#pragma unroll
  for (int i0 = 0; i0 < NB; ++i0)
#pragma unroll
    for (int A = 0; A < M2N; ++A) {
      double r = 0.0;
#pragma unroll
      for (int j0 = 0; j0 < NB; ++j0) r = fma(it * dg[i0 * NB + j0], WM[j0][A], fma(t * ng[i0 * NB + j0], WK[j0][A], r));
      vbuf3[(int64_t)c_hex_t3[i0 * M2N + A] * nc + c] = r;
    }
}

// ---- tensor-structured numeric phase of the assembled shell matrix ----------------------------------------------------------
// Q_p on the shell is the tensor product (radial Q_p on the layers) x (surface Q_p), so the ASSEMBLED matrix factorises too:
//     A[(rho_i, sigma_i), (rho_j, sigma_j)] = (1/t) DG[rho_i, rho_j] M2G[sigma_i, sigma_j] + t NG[rho_i, rho_j] K2G[sigma_i, sigma_j]
// with DG, NG the assembled radial matrices and M2G, K2G the assembled SURFACE mass / stiffness matrices (cell-wise quadrature
// + atomic-free CSR fill on the 6 n^2 surface cells, every call).  One thread per CSR slot: 14 B of traffic per non-zero
// (slot of the surface pattern, radial pair, value) instead of writing and re-gathering 3 KB of element matrix per cell.

// one thread per row: for every slot of the row, the slot of (sigma_i, sigma_j) in the surface pattern and the radial pair
__global__ void k_hex_tensor_maps(int64_t n_rows, int NR, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                                  const int32_t* __restrict__ rho, const int32_t* __restrict__ sig,
                                  const int32_t* __restrict__ rowptr2, const int32_t* __restrict__ colind2,
                                  int32_t* __restrict__ slot2d, uint16_t* __restrict__ rr, int* __restrict__ bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  const int32_t ri = rho[i], si = sig[i];
  const int32_t b2 = rowptr2[si], e2 = rowptr2[si + 1];
  for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
    const int32_t j = colind[k], sj = sig[j];
    int32_t lo = b2, hi = e2;
    while (lo < hi) {
      const int32_t mid = (lo + hi) >> 1;
      if (colind2[mid] < sj) lo = mid + 1; else hi = mid;
    }
    const bool ok = lo < e2 && colind2[lo] == sj;
    if (!ok) atomicAdd(bad, 1);
    slot2d[k] = ok ? lo : -1;
    rr[k] = (uint16_t)(ri * NR + rho[j]);
  }
}

// interleave the two surface value arrays so that a slot needs one 16-byte gather instead of two 8-byte ones
__global__ void k_hex_interleave(int64_t n, const double* __restrict__ a, const double* __restrict__ b, double2* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = make_double2(a[k], b[k]);
}

// One thread per CSR slot: 6 B of maps read + 8 B written per non-zero, two 16-byte gathers from tables whose entries for one
// row are a few cache lines.  DN[r] = (DG[r] / t, t NG[r]); KM[s2] = (M2G[s2], K2G[s2]).
// (Tried and slower on B200: one warp per row with a 1-byte code per slot, 1.50 ms against 1.29 ms at 48^3 cells per panel --
// short rows leave lanes idle; a CTA-level row search in shared memory, 3.3 ms -- the serial search stalls the CTA.)
__global__ void __launch_bounds__(256)
k_hex_tensor_fill(int64_t nnz, const int32_t* __restrict__ slot2d, const uint16_t* __restrict__ rr, const double2* __restrict__ DN,
                  const double2* __restrict__ KM, double scale, int add, double* __restrict__ vals) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  // streaming operands bypass L1 allocation (evict-first) so that L1 keeps the gathered factor tables
  const int32_t s2 = __ldcs(slot2d + k);
  const int r = __ldcs(rr + k);
  double v = 0.0;
  if (s2 >= 0) {
    const double2 dn = __ldg(DN + r), km = __ldg(KM + s2);
    v = scale * fma(dn.x, km.x, dn.y * km.y);
  }
  if (add) v += vals[k];
  __stcs(vals + k, v);
}

// volume of the shell cells: sum_q w h_gamma h_alpha h_beta t s^2 sqrtg_surf (optionally times a coefficient at the points)
__global__ void k_integrate_shell(int64_t nc, int n, int L, double radius, double t, int nq, const double* __restrict__ xi,
                                  const double* __restrict__ w, const int32_t* __restrict__ col,
                                  const int32_t* __restrict__ layer, const double* __restrict__ qdata,
                                  double* __restrict__ cell_sums) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const double H = 0.78539816339744830962, ha = 2.0 * H / n, hg = 1.0 / L;
  const int cc = col[c] % (n * n), ia = cc % n, jb = cc / n, kg = layer[c];
  const int Q = nq * nq * nq;
  double acc = 0.0;
  for (int q2 = 0; q2 < nq; ++q2) {
    const double b = tan(-H + (jb + xi[q2]) * ha), sb = 1.0 + b * b;
    for (int q1 = 0; q1 < nq; ++q1) {
      const double a = tan(-H + (ia + xi[q1]) * ha), sa = 1.0 + a * a;
      const double rho2 = sa + b * b;
      const double sg = radius * radius * sa * sb / (rho2 * sqrt(rho2));
      for (int q0 = 0; q0 < nq; ++q0) {
        const double s = 1.0 + t * (kg + xi[q0]) * hg / radius;
        const double f = qdata ? qdata[c * Q + q0 + nq * (q1 + nq * q2)] : 1.0;
        acc = fma(w[q0] * w[q1] * w[q2] * hg * ha * ha * t * s * s * sg, f, acc);
      }
    }
  }
  cell_sums[c] = acc;
}

}  // namespace gg
