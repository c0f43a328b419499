"""Oracle (test infrastructure): DG upwind advection of a scalar on the cubed sphere, volume + skeleton terms.

Restates the weak form of /root/reference/test/Tutorials/AdvectionUpwinding.jl:107-137
    beta   = (pinvJ o transpose o grad(ambient_map)) . beta_x(ambient_map)         contravariant components        (:107-109)
    a(u,v) = int -(u (grad(v).beta)) meas dOmega                                                                  (:124)
    b(u,v) = int my_mean((beta u).n_L) jump(v) meas_skel.plus dL,   my_mean(p) = 0.5 (p.plus - p.minus)          (:113-117,125)
    s(u,v) = int 0.5 |(beta.n_L).plus| jump(u) jump(v) meas_skel.plus dL                                         (:120,126)
    mass   = int dtu v meas dOmega                                                                                (:128)
on SkeletonTriangulation(model): every edge of the closed cubed sphere is an interior facet; `plus` is the first (lower id)
cell around it, `minus` the second; n_L.plus / n_L.minus are the outward unit normals of the two cells in their OWN charts
(they differ across panel edges), quadrature on the facet is Gauss-Legendre of the given degree, mapped into both cells by
the facet parameter (all shared edges have the same vertex order in both cells on this mesh,
test/AtlasDiscreteModels/seq/DarcyCubedSphereTests.jl:25-32), and dL is the chart length of the edge on the plus side.
The tutorial marches with an implicit SDIRK scheme; BASELINE config 3 asks for explicit RK, which is what `step` does
(M k = -res, Gridap.ODEs EXRungeKutta with constant mass).  No reference test pins numbers for this path: PARITY UNPINNED
beyond the structural properties checked in tests (conservation, constants are steady for a solenoidal wind).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import forms
from . import geometry as geo
from .refel import QUAD_EDGES, QUAD_OUT_NORMALS, QUAD_VERTS, gauss_legendre_01, num_points_1d
from .rk import TABLEAUS
from .spaces import DGSpace


def contravariant(mesh, cells, x, beta_amb):
    """pinvJ(J) . beta for chart points x (n, Q, 2) of the given cells: g^-1 (J beta), J[i,k] = d phi_k / d x_i."""
    out = np.empty(x.shape)
    for p in range(6):
        sel = mesh.cell_to_chart[cells] == p
        if not sel.any():
            continue
        X = geo.csphere_eval(p + 1, mesh.radius, x[sel])
        J = geo.csphere_jac(p + 1, mesh.radius, x[sel])
        gi = geo.csphere_inv_metric(mesh.radius, x[sel])
        out[sel] = np.einsum("...ij,...j->...i", gi, np.einsum("...ik,...k->...i", J, beta_amb(X)))
    return out


class AdvectionProblem:
    def __init__(self, mesh, order, beta_amb, degree=None):
        self.mesh, self.order = mesh, order
        self.Q = DGSpace(mesh, order)
        self.degree = 4 * order if degree is None else degree
        q = self.quad = forms.CellQuadrature(mesh, self.degree)
        Nc, m = mesh.num_cells, self.Q.ndofs_loc
        val, grad = forms.scalar_basis(self.Q, q)                       # (Q,m), (Nc,Q,m,2)
        allc = np.arange(Nc)
        beta = contravariant(mesh, allc, q.x, beta_amb)                 # (Nc,Q,2)
        # volume term: A[I,J] = -sum_q dV sqrtg phi_J (grad phi_I . beta)
        Ke = -np.einsum("cq,qj,cqi->cij", q.dV * q.sqrtg, val, np.einsum("cqmi,cqi->cqm", grad, beta))
        rows = np.repeat(self.Q.cell_dofs, m, axis=1).reshape(-1)
        cols = np.tile(self.Q.cell_dofs, (1, m)).reshape(-1)
        A = sp.coo_matrix((Ke.reshape(-1), (rows, cols)), shape=(self.Q.num_free,) * 2).tocsr()
        # skeleton terms
        fc = mesh.facet_cells()                                         # edge -> (plus, minus)
        nqf = num_points_1d(self.degree)
        s, wf = gauss_legendre_01(nqf)
        loc_edge = {}
        for c in range(Nc):
            for e in range(4):
                loc_edge[(c, mesh.cell_edges[c, e])] = e
        ne = mesh.num_edges
        cp, cm = fc[:, 0], fc[:, 1]
        ep = np.array([loc_edge[(cp[k], k)] for k in range(ne)])
        em = np.array([loc_edge[(cm[k], k)] for k in range(ne)])

        def side(cells, ledge):
            va = QUAD_VERTS[[QUAD_EDGES[e][0] for e in ledge]]          # (ne,2) reference end points
            vb = QUAD_VERTS[[QUAD_EDGES[e][1] for e in ledge]]
            ref = va[:, None, :] + s[None, :, None] * (vb - va)[:, None, :]          # (ne,nqf,2)
            cc = mesh.cell_chart_coords[cells]
            lo, hi = cc[:, 0, :], cc[:, 3, :]
            x = lo[:, None, :] + ref * (hi - lo)[:, None, :]
            nrm = QUAD_OUT_NORMALS[ledge]                                # outward parametric normal (unit in the chart)
            bn = np.einsum("kqi,ki->kq", contravariant(mesh, cells, x, beta_amb), nrm)
            sg = geo.csphere_sqrtg(mesh.radius, x)
            length = np.abs(np.einsum("ki,ki->k", (vb - va), (hi - lo)))             # chart length of the edge
            phi = np.stack([self.Q.ref.tabulate(ref[k])[0] for k in range(ne)])      # (ne,nqf,m)
            return bn, sg, length, phi

        bnp, sgp, lenp, phip = side(cp, ep)
        bnm, _, _, phim = side(cm, em)
        w = wf[None, :] * lenp[:, None] * sgp                           # dL * meas_skel.plus
        up = 0.5 * np.abs(bnp)
        # u-coefficients of  mean + upwind*jump(u):  plus side 0.5 bn+ + up,  minus side -0.5 bn- - up
        cu_p, cu_m = 0.5 * bnp + up, -0.5 * bnm - up
        dp, dm = self.Q.cell_dofs[cp], self.Q.cell_dofs[cm]             # (ne,m)
        blocks = []
        for (tv, td, ts) in ((phip, dp, 1.0), (phim, dm, -1.0)):        # jump(v) = v+ - v-
            for (uv, ud, cu) in ((phip, dp, cu_p), (phim, dm, cu_m)):
                K = ts * np.einsum("kq,kqi,kqj->kij", w * cu, tv, uv)
                blocks.append(sp.coo_matrix((K.reshape(-1), (np.repeat(td, m, axis=1).reshape(-1), np.tile(ud, (1, m)).reshape(-1))),
                                            shape=A.shape))
        self.A = (A + sum(blocks)).tocsr()
        self.M = forms.assemble_matrix(self.Q, self.Q, forms.scalar_mass_element_matrices(self.Q, self.Q, q)).tocsc()
        self.lu = spla.splu(self.M)

    def residual(self, u):
        return self.A @ u

    def slope(self, u):
        return self.lu.solve(-(self.A @ u))

    def step(self, u0, dt, tableau="EXRK_SSP_3_3"):
        A, b, c = TABLEAUS[tableau]
        ks = []
        for i in range(len(c)):
            ui = u0.copy()
            for j in range(i):
                ui += dt * A[i, j] * ks[j]
            ks.append(self.slope(ui))
        return u0 + dt * sum(b[i] * ks[i] for i in range(len(c)))

    def total_mass(self, u):
        return float(np.ones(self.Q.num_free) @ (self.M @ u))
