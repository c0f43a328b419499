"""Oracle (test infrastructure): explicit Runge-Kutta drivers (linear wave, shallow-water DAE).

Restates
  Gridap.ODEs stock EXRungeKutta for a semilinear operator with constant mass (wave driver,
    /root/reference/test/Transient/TransientWaveEquation.jl:63-77): per stage solve M k_i = -res(t_i, u_i),
    u_i = u_0 + dt sum_j A_ij k_j, u_F = u_0 + dt sum_i b_i k_i;
  /root/reference/src/ODEs/RungeKuttaEX.jl:48-134 (DAE-aware march: diagnostic solve before every stage
    and after the update), /root/reference/src/ODEs/DAE.jl:241-274 (one linear solve from x = 0),
    /root/reference/src/ODEs/StageOperators.jl:55-88 (slope_i = -M^-1 r).
Linear solves use scipy's sparse LU, mirroring the reference's `LUSolver()`.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from . import forms
from . import geometry as geo

# SURVEY.md App. A.6 (Gridap tableau names)
TABLEAUS = {
    "EXRK_SSP_3_3": (np.array([[0, 0, 0], [1, 0, 0], [0.25, 0.25, 0]]), np.array([1 / 6, 1 / 6, 2 / 3]), np.array([0, 1, 0.5])),
    "EXRK_SSP_2_2": (np.array([[0, 0], [1, 0]]), np.array([0.5, 0.5]), np.array([0, 1.0])),
    "EXRK_Euler_1_1": (np.array([[0.0]]), np.array([1.0]), np.array([0.0])),
}


def dx(mesh):
    """src/ConvergenceTools/Tools.jl:228-232."""
    return np.sqrt(4 * np.pi * mesh.radius**2 / mesh.num_cells)


def snapped_dt(dt0, tF):
    """dt = tF / floor(tF / dt0) (TransientWaveEquation.jl:71-73)."""
    return tF / np.floor(tF / dt0)


def interpolate_scalar(space, mesh, fun):
    """Lagrangian interpolation of an ambient function fun(xyz) (DG or CG)."""
    nodes = space.ref.nodes                                       # (m, 2) reference
    xi, eta = nodes[:, 0], nodes[:, 1]
    N = np.stack([(1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta], axis=1)
    x = np.einsum("qk,ckd->cqd", N, mesh.cell_chart_coords)
    vals = np.empty(x.shape[:2])
    for p in range(6):
        sel = mesh.cell_to_chart == p
        vals[sel] = fun(geo.csphere_eval(p + 1, mesh.radius, x[sel]))
    out = np.zeros(space.num_free)
    ids = space.cell_dofs
    out[ids[ids >= 0]] = vals[ids >= 0]
    return out


def interpolate_rt(space, mesh, vfun):
    """RT moment interpolation of u = sqrtg * pinvJ(J) . v(xyz)  (TransientShallowWater.jl:89-90).

    The chart-space field is pulled back to the reference cell with the inverse contravariant Piola map
    before the reference moments are applied (Gridap's TransformRTDofBasis)."""
    pts, W = space.ref.moment_points()                            # (N,2), (m,N,2)
    xi, eta = pts[:, 0], pts[:, 1]
    N = np.stack([(1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta], axis=1)
    dN = np.stack([
        np.stack([-(1 - eta), (1 - eta), -eta, eta], axis=1),
        np.stack([-(1 - xi), -xi, (1 - xi), xi], axis=1)], axis=2)
    c = mesh.cell_chart_coords
    x = np.einsum("qk,ckd->cqd", N, c)
    Jt = np.einsum("qki,ckd->cqid", dN, c)
    detJ = np.linalg.det(Jt)
    u = np.empty(x.shape)
    for p in range(6):
        sel = mesh.cell_to_chart == p
        X = geo.csphere_eval(p + 1, mesh.radius, x[sel])
        Jg = geo.csphere_jac(p + 1, mesh.radius, x[sel])
        P = geo.pinvJ(Jg)                                          # (.., 2, 3)
        u[sel] = geo.csphere_sqrtg(mesh.radius, x[sel])[..., None] * np.einsum("...ik,...k->...i", P, vfun(X))
    # uhat = det J * inv(J) u, with J[d,i] = Jt[i,d]  => uhat_i = det * sum_d u_d invJt[d,i]
    uhat = detJ[..., None] * np.einsum("cqd,cqdi->cqi", u, np.linalg.inv(Jt))
    dofs = np.einsum("mqi,cqi->cm", W, uhat) * space.signs
    out = np.zeros(space.num_free)
    out[space.cell_dofs] = dofs   # shared facet DOFs: both cells give the same value
    return out


class WaveProblem:
    """Linear wave equation with RT_k x DG_k (TransientWaveEquation.jl:28-77)."""

    def __init__(self, mesh, p_fe, degree=None):
        from .spaces import RTSpace, DGSpace
        self.mesh = mesh
        self.V = RTSpace(mesh, p_fe)
        self.Q = DGSpace(mesh, p_fe)
        self.degree = 5 * (p_fe + 1) if degree is None else degree
        self.quad = forms.CellQuadrature(mesh, self.degree)
        q = self.quad
        self.Mu = forms.assemble_matrix(self.V, self.V, forms.rt_mass_element_matrices(self.V, q))
        self.Mp = forms.assemble_matrix(self.Q, self.Q, forms.scalar_mass_element_matrices(self.Q, self.Q, q))
        self.D = forms.assemble_matrix(self.Q, self.V, forms.div_element_matrices(self.Q, self.V, q))
        self.nu, self.np_ = self.V.num_free, self.Q.num_free
        self.M = sp.block_diag([self.Mu, self.Mp]).tocsc()
        self.lu = spla.splu(self.M)

    def residual(self, x):
        """res(t,(u,p),(v,q)) = int q div u - int p div v, as a vector [r_v; r_q]."""
        u, p = x[: self.nu], x[self.nu:]
        return np.concatenate([-(self.D.T @ p), self.D @ u])

    def step(self, x0, dt, tableau="EXRK_SSP_3_3"):
        A, b, c = TABLEAUS[tableau]
        ks = []
        for i in range(len(c)):
            xi = x0.copy()
            for j in range(i):
                xi += dt * A[i, j] * ks[j]
            ks.append(self.lu.solve(-self.residual(xi)))
        xF = x0.copy()
        for i in range(len(c)):
            xF += dt * b[i] * ks[i]
        return xF

    def errors(self, x0, xF):
        from .forms import CellQuadrature, eval_rt, eval_scalar
        qe = CellQuadrature(self.mesh, 2 * self.degree)
        e = x0 - xF
        eu, _ = eval_rt(self.V, qe, e[: self.nu])
        ep = eval_scalar(self.Q, qe, e[self.nu:])
        geu = np.einsum("cqij,cqj->cqi", qe.g, eu)
        e_u = np.sqrt((np.einsum("cqi,cqi->cq", eu, geu) / qe.sqrtg * qe.dV).sum())
        e_p = np.sqrt((ep * ep * qe.sqrtg * qe.dV).sum())
        return e_u, e_p


def gridap_num_steps(t0, tF, dt):
    """Number of steps Gridap's ODE-solution iterator takes: it stops when t >= tF - 100 eps, with t
    accumulated as t += dt.  For the reference wave test (nref 3, CFL 0.1) this is 348, not 347: the
    accumulated time after 347 steps is 2.8e-14 short of 2 pi; the reference's regression numbers
    (TransientWaveEquation.jl:128-129) are only reproduced with the extra step."""
    eps = 100 * np.finfo(np.float64).eps
    t, n = t0, 0
    while not (t >= tF - eps):
        t = t + dt
        n += 1
    return n
