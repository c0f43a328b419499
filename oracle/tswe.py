"""Oracle (test infrastructure): thermal shallow water as a DAE with explicit Runge-Kutta stepping.

Restates the weak forms of /root/reference/test/Tutorials/ThermalShallowWater.jl on top of oracle/swe.py (same stepper,
src/ODEs/RungeKuttaEX.jl:48-134; same q0 aliasing quirk):
  prognostic  x = (u in RT_k, p in DG_k, B in DG_k)                                  :74-75
  diagnostic  y = (q in CG_{k+1}, F in RT_k, Phi in DG_k, b in DG_k)                 :77-78
  res_y       resq, resF as in the shallow-water driver;
              resPhi = int Phi psi sqrtg - int 0.5 B psi sqrtg - int 0.5 (u.(g u)) psi / sqrtg     :173
              resb   = int b p r sqrtg - int B r sqrtg                                             :174
  res_x       res_u = (shallow-water terms with Phi) + int 0.5 b (grad(th).v) - int 0.5 b th (div v) - int 0.5 th (grad(b).v),
              th = 0.5 p                                                                            :216-224
              u_s1  = int_L -0.5 my_mean((v b).n) jump(th) + 0.5 my_mean((v th).n) jump(b)         :226-229
              u_s2  = int_L -0.5 upwinding_sign((F.n).plus) (v.n).plus jump(b) jump(th)            :231
              res_p = int r div(F)                                                                  :196
              res_B = int -0.5 b (grad(w).F) + 0.5 b w div(F) + 0.5 w (grad(b).F)                  :234-238
              B_s1  = int_L 0.5 my_mean((F b).n) jump(w) - 0.5 my_mean((F w).n) jump(b)            :240-243
              B_s2  = int_L 0.5 upwinding_sign((F.n).plus) (F.n).plus jump(b) jump(w)              :245
              my_mean(x) = 0.5 (x.plus - x.minus) :198-202, upwinding_sign(Fn) = -0.5 / 0 / +0.5 with threshold 1e-4 :205-213
  mass        RT mass for u, DG mass for p and B                                                    :189-193
  rules       dOmega = Measure(Omega, 4 order), dL = Measure(Lambda, 4 order)                       :166-167
  initial state: thermogeostrophic balance, u0 / h0 of Williamson 2, b0 = g (1 + 0.05 H0^2 / h^2), B0 = b0 h0   :100-139
RT functions carry sqrtg (contravariant Piola of the reference + u_cf = meas * pinvJ . u), so no metric appears in the skeleton
terms; n_L is the outward unit normal of each cell in its own chart and dL the chart length of the edge (both sides agree on
the uniform panels).  `plus` = the lower-numbered cell of a facet.  No reference test pins numbers for this tutorial:
PARITY UNPINNED beyond the properties checked in tests (the balanced state is steady, int p sqrtg and int B sqrtg conserved).
"""
import numpy as np
import scipy.sparse.linalg as spla

from . import forms
from . import geometry as geo
from . import swe
from .refel import QUAD_EDGES, QUAD_OUT_NORMALS, QUAD_VERTS, gauss_legendre_01, num_points_1d
from .rk import TABLEAUS, interpolate_scalar

C_BUOY = 0.05


def buoyancy(hfun=None):
    def f(X):
        h = (hfun or swe.depth())(X)
        return swe._g * (1.0 + C_BUOY * swe._H_0**2 / h**2)
    return f


def weighted_buoyancy(hfun=None):
    def f(X):
        return buoyancy(hfun)(X) * (hfun or swe.depth())(X)
    return f


def upwinding_sign(Fn):
    return np.where(Fn < -1e-4, -0.5, np.where(Fn > 1e-4, 0.5, 0.0))


class ThermalShallowWaterProblem(swe.ShallowWaterProblem):
    def __init__(self, mesh, p_fe, f=None, degree=None, q0_mode="alias"):
        degree = (4 * p_fe if p_fe > 0 else 2) if degree is None else degree
        super().__init__(mesh, p_fe, gravity=0.0, f=f, degree=degree, q0_mode=q0_mode)
        q = self.quad
        self.dg_grad = forms.scalar_basis(self.Q, q)[1]                   # (Nc,Q,m,2) chart gradients
        # ---- skeleton tabulations ------------------------------------------------------------------------------------
        Nc, ne = mesh.num_cells, mesh.num_edges
        fc = mesh.facet_cells()
        self.cp, self.cm = fc[:, 0], fc[:, 1]
        loc = {}
        for c in range(Nc):
            for e in range(4):
                loc[(c, mesh.cell_edges[c, e])] = e
        self.ep = np.array([loc[(self.cp[k], k)] for k in range(ne)])
        self.em = np.array([loc[(self.cm[k], k)] for k in range(ne)])
        nqf = num_points_1d(self.degree)
        s, wf = gauss_legendre_01(nqf)
        cc = mesh.cell_chart_coords
        h = cc[:, 3, :] - cc[:, 0, :]                                      # (Nc,2) chart sizes of the axis-aligned cells
        self.dg_f, self.rtn_f = [], []
        for e in range(4):
            a, b = QUAD_VERTS[QUAD_EDGES[e][0]], QUAD_VERTS[QUAD_EDGES[e][1]]
            ref = a[None, :] + s[:, None] * (b - a)[None, :]
            self.dg_f.append(self.Q.ref.tabulate(ref)[0])                  # (nqf,m)
            self.rtn_f.append(self.V.ref.tabulate(ref)[0])                 # (nqf,mu,2) reference values
        self.dg_f, self.rtn_f = np.stack(self.dg_f), np.stack(self.rtn_f)

        def side(cells, ledge):
            va = QUAD_VERTS[[QUAD_EDGES[e][0] for e in ledge]]
            vb = QUAD_VERTS[[QUAD_EDGES[e][1] for e in ledge]]
            hc = h[cells]
            length = np.abs(np.einsum("ki,ki->k", vb - va, hc))
            nrm = QUAD_OUT_NORMALS[ledge]                                  # (ne,2)
            # contravariant Piola on an axis-aligned cell: v = (vhat_x hx, vhat_y hy) / (hx hy); normal component, signs applied
            piola = hc / np.prod(hc, axis=1)[:, None]
            vn = np.einsum("kqmi,ki,ki->kqm", self.rtn_f[ledge], piola, nrm) * self.V.signs[cells][:, None, :]
            return length, vn, self.dg_f[ledge]

        self.len_f, self.vn_p, self.w_p = side(self.cp, self.ep)
        _, self.vn_m, self.w_m = side(self.cm, self.em)
        self.wL = wf[None, :] * self.len_f[:, None]                        # (ne,nqf)
        self.nB = self.np_

    # ---- diagnostics ----------------------------------------------------------------------------------------------------
    def split_x(self, x):
        return x[: self.nu], x[self.nu: self.nu + self.np_], x[self.nu + self.np_:]

    def split_y(self, y):
        a, b, c = self.nq, self.nq + self.nu, self.nq + self.nu + self.np_
        return y[:a], y[a:b], y[b:c], y[c:]

    def diagnostics(self, x):
        q = self.quad
        u, p, B = self.split_x(x)
        uq, pq, Bq = self._u(u), self._dg(p), self._dg(B)
        wq = q.dV * q.sqrtg
        Mq = forms.assemble_matrix(self.R, self.R, forms.scalar_mass_element_matrices(self.R, self.R, q, coef=pq))
        flux = np.einsum("cqi,cqij->cqj", geo.perp(uq), q.ginv)
        rhs_q = np.einsum("cq,qm->cm", wq * self.f_q, self.cg_val) + np.einsum("cq,cqj,cqmj->cm", wq, flux, self.cg_grad)
        qv = spla.spsolve(Mq.tocsc(), forms.assemble_vector(self.R, rhs_q))
        gu = np.einsum("cqij,cqj->cqi", q.g, uq)
        rhs_F = np.einsum("cq,cqi,cqmi->cm", q.dV / q.sqrtg * pq, gu, self.rt_u)
        Fv = self.lu_u.solve(forms.assemble_vector(self.V, rhs_F))
        ke = 0.5 * np.einsum("cqi,cqi->cq", uq, gu)
        rhs_P = np.einsum("cq,qm->cm", q.dV * (0.5 * Bq * q.sqrtg + ke / q.sqrtg), self.dg_val)
        Pv = self.lu_p.solve(forms.assemble_vector(self.Q, rhs_P))
        Mb = forms.assemble_matrix(self.Q, self.Q, forms.scalar_mass_element_matrices(self.Q, self.Q, q, coef=pq))
        bv = spla.spsolve(Mb.tocsc(), forms.assemble_vector(self.Q, np.einsum("cq,qm->cm", wq * Bq, self.dg_val)))
        return np.concatenate([qv, Fv, Pv, bv])

    # ---- stage residual -------------------------------------------------------------------------------------------------
    def _dg_grad(self, v):
        return np.einsum("cqmi,cm->cqi", self.dg_grad, v[self.Q.cell_dofs])

    def residual_x(self, x, y, y0):
        q = self.quad
        u, p, B = self.split_x(x)
        qv, Fv, Pv, bv = self.split_y(y)
        q0v = self.split_y(y0)[0]
        uq, Fq = self._u(u), self._u(Fv)
        divF = np.einsum("cqm,cm->cq", self.rt_div, Fv[self.V.cell_dofs])
        RF = geo.perp(Fq)
        qq, gq = self._cg(qv)
        q0q, _ = self._cg(q0v)
        th, gth = 0.5 * self._dg(p), 0.5 * self._dg_grad(p)
        bq, gb = self._dg(bv), self._dg_grad(bv)
        coef = qq - self.tau * (qq - q0q) / self.dt - self.tau * np.einsum("cqi,cqi->cq", uq, gq) / q.sqrtg
        vec = coef[..., None] * RF + 0.5 * bq[..., None] * gth - 0.5 * th[..., None] * gb          # multiplies v
        eu = np.einsum("cq,cqi,cqmi->cm", q.dV, vec, self.rt_u) - np.einsum("cq,cqm->cm", q.dV * 0.5 * bq * th, self.rt_div)
        r_u = forms.assemble_vector(self.V, eu) - self.D.T @ Pv
        r_p = self.D @ Fv
        eB = (np.einsum("cq,cqi,cqmi->cm", -0.5 * q.dV * bq, Fq, self.dg_grad)
              + np.einsum("cq,qm->cm", q.dV * (0.5 * bq * divF + 0.5 * np.einsum("cqi,cqi->cq", gb, Fq)), self.dg_val))
        # ---- skeleton ----------------------------------------------------------------------------------------------------
        cp, cm = self.cp, self.cm
        trace = lambda v, cells, w: np.einsum("kqm,km->kq", w, v[self.Q.cell_dofs[cells]])          # noqa: E731
        bP, bM = trace(bv, cp, self.w_p), trace(bv, cm, self.w_m)
        tP, tM = 0.5 * trace(p, cp, self.w_p), 0.5 * trace(p, cm, self.w_m)
        FnP = np.einsum("kqm,km->kq", self.vn_p, Fv[self.V.cell_dofs[cp]])       # (F.n).plus; the DOF signs are in vn
        FnM = np.einsum("kqm,km->kq", self.vn_m, Fv[self.V.cell_dofs[cm]])
        jb, jt = bP - bM, tP - tM
        sg = upwinding_sign(FnP)
        wL = self.wL
        # u equation: element vectors of the plus and the minus cell of every facet
        fu_p = np.einsum("kq,kqm->km", wL * (-0.25 * bP * jt + 0.25 * tP * jb - 0.5 * sg * jb * jt), self.vn_p)
        fu_m = np.einsum("kq,kqm->km", wL * (0.25 * bM * jt - 0.25 * tM * jb), self.vn_m)
        np.add.at(r_u, self.V.cell_dofs[cp], fu_p)
        np.add.at(r_u, self.V.cell_dofs[cm], fu_m)
        # B equation
        mFb = 0.5 * (FnP * bP - FnM * bM)
        up = 0.5 * sg * FnP * jb
        fB_p = np.einsum("kq,kqm->km", wL * (0.5 * mFb - 0.25 * FnP * jb + up), self.w_p)
        fB_m = np.einsum("kq,kqm->km", wL * (-0.5 * mFb + 0.25 * FnM * jb - up), self.w_m)
        r_B = forms.assemble_vector(self.Q, eB)
        np.add.at(r_B, self.Q.cell_dofs[cp], fB_p)
        np.add.at(r_B, self.Q.cell_dofs[cm], fB_m)
        return np.concatenate([r_u, r_p, r_B])

    def slope(self, x, y, y0):
        r = self.residual_x(x, y, y0)
        a, b = self.nu, self.nu + self.np_
        return np.concatenate([self.lu_u.solve(-r[:a]), self.lu_p.solve(-r[a:b]), self.lu_p.solve(-r[b:])])

    def initial_state(self, vfun=None, hfun=None, Bfun=None):
        x = super().initial_state(vfun, hfun)
        return np.concatenate([x, interpolate_scalar(self.Q, self.mesh, Bfun or weighted_buoyancy(hfun))])

    def total(self, v):
        """int v_h sqrtg of a DG field."""
        return float(np.ones(self.np_) @ (self.Mp @ v))
