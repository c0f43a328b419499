"""Oracle (test infrastructure): rotating shallow water as a DAE with explicit Runge-Kutta stepping.

Restates the driver /root/reference/test/Transient/TransientShallowWater.jl:49-186 (spaces :65-78, initial state :88-95,
forms :113-139, time step :147-150) on top of the DAE-aware stepper of the reference:
  /root/reference/src/ODEs/RungeKuttaEX.jl:48-134   diagnostic solve, then per stage: stage state, diagnostic solve, slope solve;
                                                    final update and a last diagnostic solve
  /root/reference/src/ODEs/DAE.jl:241-274            diagnostics: x = 0, b = res_y(x), A = jac_y, solve A x = -b (one linear solve)
  /root/reference/src/ODEs/StageOperators.jl:55-88   slope: x = 0, b = res_x + mass(0), A = mass, solve A x = -b
Williamson test case 2 fields: /root/reference/test/Geophysical/Williamson_functions.jl:1-37 and the
non-dimensionalisation of TransientShallowWater.jl:27-46.

q0 handling: the reference passes the SAME array as `diagnostics` and `diagnostics0`
(RungeKuttaEX.jl:75 -> ODEOpsFromTFEOps.jl:214-215) and later overwrites it in place, so q0 aliases q and the term
-tau ((q - q0)/dt) (R F).v vanishes at every stage.  `q0_mode="alias"` (default) reproduces that;
`q0_mode="start_of_step"` keeps the diagnostics of the beginning of the step (the apparent intent).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import forms
from . import geometry as geo
from .rk import TABLEAUS, interpolate_rt, interpolate_scalar
from .spaces import CGSpace, DGSpace, RTSpace

# ---- TransientShallowWater.jl:27-46 -------------------------------------------------------------------------------
a_e = 6.37e6
g_dim = 9.8
omega_dim = 7.29e-5
T_dim = 12 * 24 * 3600
H_0 = 2.94e4 / g_dim
u_0 = 2 * np.pi * a_e / T_dim
TF = 1 * (3600 * 24)
L = a_e
_tau_scale = 1 / omega_dim
_a = a_e / L
_g = g_dim * _tau_scale**2 / L
_omega = omega_dim * _tau_scale
_H_0 = H_0 / L
_T = T_dim / _tau_scale
_u0 = u_0 / L * _tau_scale
_tF = TF / _tau_scale


# ---- Williamson_functions.jl (zeta = 0 in the test) -------------------------------------------------------------------
def coriolis(zeta=0.0):
    def f(X):
        th, ph, _ = geo.xyz2thetaphir(X)
        return 2.0 * _omega * (-np.cos(th) * np.cos(ph) * np.sin(zeta) + np.sin(ph) * np.cos(zeta))
    return f


def velocity(zeta=0.0):
    def f(X):
        th, ph, _ = geo.xyz2thetaphir(X)
        u = _u0 * (np.cos(ph) * np.cos(zeta) + np.cos(th) * np.sin(ph) * np.sin(zeta))
        v = -_u0 * np.sin(th) * np.sin(zeta)
        # _spherical_to_cartesian_matrix(thetaphir) . (u, v, 0): TensorValue is column-major, so the rows of the
        # 3x3 written in the source are its columns
        e_th = np.stack([-np.sin(th), np.cos(th), 0 * th], -1)
        e_ph = np.stack([-np.sin(ph) * np.cos(th), -np.sin(ph) * np.sin(th), np.cos(ph)], -1)
        return u[..., None] * e_th + v[..., None] * e_ph
    return f


def depth(zeta=0.0):
    def f(X):
        th, ph, _ = geo.xyz2thetaphir(X)
        h = -np.cos(th) * np.cos(ph) * np.sin(zeta) + np.sin(ph) * np.cos(zeta)
        return _H_0 - (_omega * _u0 + 0.5 * _u0 * _u0) * h * h / _g
    return f


def tangent_component(vfun):
    """src/Helpers/SphereSurfaceFunctions.jl:2-7."""
    def f(X):
        n = X / np.linalg.norm(X, axis=-1, keepdims=True)
        v = vfun(X)
        return v - np.einsum("...k,...k->...", v, n)[..., None] * n
    return f


class ShallowWaterProblem:
    def __init__(self, mesh, p_fe, gravity=_g, f=None, b=None, degree=None, q0_mode="alias"):
        self.mesh, self.p = mesh, p_fe
        self.R = CGSpace(mesh, p_fe + 1)
        self.Q = DGSpace(mesh, p_fe)
        self.V = RTSpace(mesh, p_fe)
        self.degree = 4 * (p_fe + 1) if degree is None else degree
        self.quad = q = forms.CellQuadrature(mesh, self.degree)
        self.gravity = gravity
        self.q0_mode = q0_mode
        X = q.ambient()
        self.f_q = (f or coriolis())(X)
        self.b_q = b(X) if b is not None else np.zeros(X.shape[:2])
        self.nu, self.np_, self.nq = self.V.num_free, self.Q.num_free, self.R.num_free
        # constant operators
        self.Mu = forms.assemble_matrix(self.V, self.V, forms.rt_mass_element_matrices(self.V, q)).tocsc()
        self.Mp = forms.assemble_matrix(self.Q, self.Q, forms.scalar_mass_element_matrices(self.Q, self.Q, q)).tocsc()
        self.D = forms.assemble_matrix(self.Q, self.V, forms.div_element_matrices(self.Q, self.V, q))
        self.lu_u, self.lu_p = spla.splu(self.Mu), spla.splu(self.Mp)
        # tabulations
        self.rt_u, self.rt_div = forms.rt_basis(self.V, q)               # (Nc,Q,m,2), (Nc,Q,m)
        self.dg_val = self.Q.ref.tabulate(q.pts)[0]                      # (Q,m)
        self.cg_val, self.cg_grad = forms.scalar_basis(self.R, q)        # (Q,m), (Nc,Q,m,2)
        self.tau = None
        self.dt = None

    # -- evaluation helpers
    def _u(self, u):
        return np.einsum("cqmi,cm->cqi", self.rt_u, u[self.V.cell_dofs])

    def _dg(self, p):
        return np.einsum("qm,cm->cq", self.dg_val, p[self.Q.cell_dofs])

    def _cg(self, qv):
        xe = qv[self.R.cell_dofs]
        return np.einsum("qm,cm->cq", self.cg_val, xe), np.einsum("cqmi,cm->cqi", self.cg_grad, xe)

    def diagnostics(self, x):
        """y = (q, F, Phi) from one linear solve of the diagnostic system at state x = (u, p)."""
        q = self.quad
        u, p = x[: self.nu], x[self.nu:]
        uq, pq = self._u(u), self._dg(p)
        Ru = geo.perp(uq)
        # vorticity: int q p w sqrtg = int f w sqrtg + int ((R u).ginv).grad(w) sqrtg
        Mq = forms.assemble_matrix(self.R, self.R, forms.scalar_mass_element_matrices(self.R, self.R, q, coef=pq))
        flux = np.einsum("cqi,cqij->cqj", Ru, q.ginv)
        wq = q.dV * q.sqrtg
        rhs_q = np.einsum("cq,qm->cm", wq * self.f_q, self.cg_val) + np.einsum("cq,cqj,cqmj->cm", wq, flux, self.cg_grad)
        qv = spla.spsolve(Mq.tocsc(), forms.assemble_vector(self.R, rhs_q))
        # mass flux: int F.(g v)/sqrtg = int p u.(g v)/sqrtg
        gu = np.einsum("cqij,cqj->cqi", q.g, uq)
        rhs_F = np.einsum("cq,cqi,cqmi->cm", q.dV / q.sqrtg * pq, gu, self.rt_u)
        Fv = self.lu_u.solve(forms.assemble_vector(self.V, rhs_F))
        # Bernoulli potential: int Phi psi sqrtg = int g_r (p + b) psi sqrtg + int 0.5 (u.(g u)) psi / sqrtg
        ke = 0.5 * np.einsum("cqi,cqi->cq", uq, gu)
        rhs_P = np.einsum("cq,qm->cm", q.dV * (self.gravity * (pq + self.b_q) * q.sqrtg + ke / q.sqrtg), self.dg_val)
        Pv = self.lu_p.solve(forms.assemble_vector(self.Q, rhs_P))
        return np.concatenate([qv, Fv, Pv])

    def split_y(self, y):
        return y[: self.nq], y[self.nq: self.nq + self.nu], y[self.nq + self.nu:]

    def residual_x(self, x, y, y0):
        """res_x + mass(0) as the vector [r_u; r_p] (TransientShallowWater.jl:128-136)."""
        q = self.quad
        u = x[: self.nu]
        qv, Fv, Pv = self.split_y(y)
        q0v = self.split_y(y0)[0]
        uq = self._u(u)
        Fq = self._u(Fv)
        RF = geo.perp(Fq)
        qq, gq = self._cg(qv)
        q0q, _ = self._cg(q0v)
        coef = qq - self.tau * (qq - q0q) / self.dt - self.tau * np.einsum("cqi,cqi->cq", uq, gq) / q.sqrtg
        r_u = forms.assemble_vector(self.V, np.einsum("cq,cqi,cqmi->cm", q.dV * coef, RF, self.rt_u)) - self.D.T @ Pv
        r_p = self.D @ Fv
        return np.concatenate([r_u, r_p])

    def slope(self, x, y, y0):
        r = self.residual_x(x, y, y0)
        return np.concatenate([self.lu_u.solve(-r[: self.nu]), self.lu_p.solve(-r[self.nu:])])

    def step(self, x0, dt, tableau="EXRK_SSP_3_3"):
        """One ode_march! (RungeKuttaEX.jl:48-134).  Returns (x_F, y_F)."""
        A, b, c = TABLEAUS[tableau]
        self.dt, self.tau = dt, dt / 2                      # TransientShallowWater.jl:150
        y_start = self.diagnostics(x0)                      # :70-75
        ks = []
        for i in range(len(c)):
            xi = x0.copy()
            for j in range(i):
                xi += dt * A[i, j] * ks[j]
            y = self.diagnostics(xi)                        # :91-93
            y0 = y if self.q0_mode == "alias" else y_start
            ks.append(self.slope(xi, y, y0))                # :100-117
        xF = x0.copy()
        for i in range(len(c)):
            xF += dt * b[i] * ks[i]
        return xF, self.diagnostics(xF)                     # :121-127

    def initial_state(self, vfun=None, hfun=None, bfun=None):
        u = interpolate_rt(self.V, self.mesh, tangent_component(vfun or velocity()))
        h = interpolate_scalar(self.Q, self.mesh, hfun or depth())
        if bfun is not None:
            h = h - interpolate_scalar(self.Q, self.mesh, bfun)
        return np.concatenate([u, h])

    def errors(self, x0, xF):
        qe = forms.CellQuadrature(self.mesh, 2 * self.degree)
        e = x0 - xF
        eu, _ = forms.eval_rt(self.V, qe, e[: self.nu])
        ep = forms.eval_scalar(self.Q, qe, e[self.nu:])
        geu = np.einsum("cqij,cqj->cqi", qe.g, eu)
        e_u = np.sqrt((np.einsum("cqi,cqi->cq", eu, geu) / qe.sqrtg * qe.dV).sum())
        e_p = np.sqrt((ep * ep * qe.sqrtg * qe.dV).sum())
        return e_u, e_p


def reference_dt(mesh, p_fe, CFL=0.1, gravity=_g, tF=_tF):
    """TransientShallowWater.jl:147-149."""
    dx = np.sqrt(4 * np.pi * mesh.radius**2 / mesh.num_cells)
    _dt = dx * CFL / (p_fe * np.sqrt(gravity * _H_0))
    return tF / np.floor(tF / _dt)
