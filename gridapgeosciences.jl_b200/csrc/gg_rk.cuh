// Explicit Runge-Kutta stepper state shared by the linear-wave (gg_rk.cu) and shallow-water (gg_swe.cu) drivers.
#pragma once
#include "gg_common.cuh"
#include "gg_ebe.cuh"

// solutions of the same solve (same stage, same unknown) in the previous steps: extrapolated initial guesses
constexpr int GG_HIST_LEVELS = 6;
struct gg_rk_hist {
  gg::DevBuf<double> h[GG_HIST_LEVELS];
  int head = 0, count = 0;   // h[head] is the newest
};

struct gg_rk {
  gg_ctx* ctx = nullptr;
  gg_mesh* mesh = nullptr;
  int problem = 0;
  gg_space *V = nullptr, *Q = nullptr;  // RT_k, DG_k (prognostic u, p)
  gg_csr* Mu = nullptr;                 // RT mass matrix (constant)
  gg_solver* solver = nullptr;          // block-Jacobi PCG on Mu (assembled or matrix-free operator)
  gg::EbeSolver* ebe_u = nullptr;       // default: statically condensed element-by-element PCG on the RT mass
  gg::EbeSolver* ebe_q = nullptr;       // ... and on the potential-vorticity mass (shallow water)
  int order = 0, degree = 0, n_stages = 0, mq = 0, mu = 0;
  std::vector<double> A, b, c;
  double dt = 0.0;
  int64_t nu = 0, np = 0;
  int64_t nB = 0;                       // thermal shallow water: the second DG prognostic field B (= np), else 0
  bool thermal = false;                 // shallow water with the buoyancy fields of test/Tutorials/ThermalShallowWater.jl
  bool warm_start = true;               // mass solves start from the previous content of their output vector
  int extrap_order = 4;                 // ... or from a polynomial extrapolation of the last 1..6 solutions of the same solve
  std::map<int, gg_rk_hist> hist;       // (GG_EXTRAPOLATE=0..6: history levels; 0 = plain warm start)
  int diag_slot = -1;                   // history slot of the diagnostic solves in flight (-1: none)
  gg::DevBuf<double> x, xi, r, MpInv, Bhat;
  std::vector<gg::DevBuf<double>> k;
  gg::DevBuf<double> iter_accum;  // [0] total PCG iterations, [1] number of solves, [2] solves that stopped at max_iter
  double noconv_seen = 0.0;       // [2] as last reported by gg_rk_step
  // ---- shallow water (DAE) only -------------------------------------------------------------------
  gg_space* R = nullptr;          // CG_{k+1} (potential vorticity)
  gg_csr* Mq = nullptr;           // int q p w sqrtg, re-assembled for every diagnostic solve
  gg_solver* solver_q = nullptr;
  gg_vec* pq = nullptr;           // depth at the quadrature points [n_cells][Q] (coefficient of Mq)
  int mr = 0, nq1 = 0;
  int64_t nr = 0;
  double gravity = 0.0, tau = 0.0;
  int q0_mode = 0;
  gg::DevBuf<double> y;           // diagnostics [q (nr); F (nu); Phi (np)]
  gg::DevBuf<double> q0;          // potential vorticity at the start of the step (q0_mode 1)
  gg::DevBuf<double> fq, bq;      // Coriolis parameter / topography at the quadrature points
  gg::DevBuf<double> rhs_q, rhs_F;
  gg::DevBuf<double> evec_q, evec_u;  // entry-major element-vector scratch
  gg::DevBuf<double> stage2;          // [point][component][cell] coefficient staging of the cell kernels (large rules only)
  // ---- DG advection only (gg_adv.cu) ------------------------------------------------------------------
  gg::DevBuf<int32_t> adv_nbr_cell;            // [n_cells][4] cell across each local edge
  gg::DevBuf<int8_t> adv_nbr_edge, adv_is_plus;  // its local edge there; whether this cell is the facet's plus side
  gg::DevBuf<double> adv_bq;                   // [Q][2][n_cells] w dV sqrtg times the contravariant wind / (hx, hy)
  gg::DevBuf<double> adv_bn, adv_wsg;          // [4][nqf][n_cells] beta.n and w dL sqrtg at the facet points of each side
  gg::DevBuf<double> adv_fval;                 // [4][nqf][m] basis at the facet points
  // ---- thermal shallow water only (gg_tswe.cu) -----------------------------------------------------------
  gg::DevBuf<double> tsw_fvn, tsw_wf;          // [4][nqf][mu] reference RT normal component at the facet points; [nqf] weights
  // ---- linear Boussinesq on the shell only (gg_bq.cu) ----------------------------------------------------------
  gg::DevBuf<double> bq_xi, bq_w, bq_rt_val, bq_rt_div, bq_dg_val;   // 1-D rule; reference tabulations at the nq^3 points
  double bq_c2 = 1.0, bq_N2 = 1.0;
  int64_t diag_solves = 0;
  bool y_valid = false;           // y == diagnostics(x) (lets a step reuse the final diagnostics of the previous one)
};

inline int64_t rk_nx(const gg_rk* rk) { return rk->nu + rk->np + rk->nB; }   // prognostic state [u; p; (B)]

namespace gg {
// thermal shallow water (gg_tswe.cu)
void tswe_create(gg_rk* rk);
void tswe_diagnostics_async(gg_rk* rk, const double* x, double* y);
void tswe_residual_async(gg_rk* rk, const double* x, const double* y, double* rB, double* kB);
// reference divergence matrix Bhat[I][J] = sum_q w_q psi_I(xi_q) divhat(phihat_J)(xi_q)  (mq x mu, row-major)
std::vector<double> build_bhat(int order, int degree);
// inverses of the DG mass blocks int psi_i psi_j sqrtg, entry-major [(i*mq+j)][cell] -> rk->MpInv
void build_dg_mass_inverse(gg_rk* rk);
void rk_accumulate_stats(gg_rk* rk, const double* solver_stats, double rtol);
// k = A^-1 (scale b) with the RT mass (which = 0) or the potential-vorticity mass (which = 1), whichever solver is active
// slot >= 0 names the solve across steps (stage / unknown) for the extrapolated initial guess
// be != nullptr (element-by-element solver only): right-hand side as entry-major element vectors, gathered inside the solve
void rk_mass_solve(gg_rk* rk, int which, const double* b, double scale, double* x, int slot = -1, const double* be = nullptr);
void rk_combine(gg_rk* rk, int ns, const double* coef, double* out);  // out = x + sum_j coef[j] k_j
// shallow water (gg_swe.cu)
void swe_create(gg_rk* rk, const gg_rk_args* a);
void swe_step_async(gg_rk* rk);
void swe_diagnostics_async(gg_rk* rk, const double* x, double* y);
// gather_u = false leaves r_u in element form (rk->evec_u) for a solve that gathers it itself
void swe_residual_async(gg_rk* rk, const double* x, const double* y, const double* q0, double* r, double* kp, bool gather_u = true);
void swe_destroy(gg_rk* rk);
// linear Boussinesq on the 3-D shell (gg_bq.cu)
void bq_create(gg_rk* rk, const gg_rk_args* a);
void bq_step_async(gg_rk* rk);
void bq_residual_async(gg_rk* rk, const double* x, double* r, double* k_pb);
int bq_interp_points(gg_space* s, double* chart, double* xyz);
void bq_interpolate(gg_space* s, const double* vals, double* out);
// DG advection (gg_adv.cu)
void adv_create(gg_rk* rk, const gg_rk_args* a);
void facet_tables(gg_rk* rk, int nqf);  // adv_nbr_cell / adv_nbr_edge / adv_is_plus / adv_fval
void adv_step_async(gg_rk* rk);
void adv_slope_async(gg_rk* rk, const double* u, double* kout, double* rout);
}  // namespace gg
