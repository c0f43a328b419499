"""Oracle (test infrastructure): linearised Boussinesq equations on the 3-D cubed-sphere shell, RT_k x DG_k x DG_k.

Restates /root/reference/test/Tutorials/LinearBoussinesq.jl:
  spaces  :50-61   RT_k (H(div), no flux through the bottom and top boundaries), DG_k pressure, DG_k buoyancy
  fields  :100-113 metric g, meas = sqrt(det g), unit radial normal n = (1,0,0) / sqrt(ginv_11), initial state :63-91,115
  forms   :112-134 mass = int v.(g u)/meas + int q p meas + int r b meas
                   resu = int omega (n x (g u / meas)).(g v)/meas - int p div(v) - int b (n.v)
                   resp = int c^2 q div(u),   resb = int N^2 r (n.u)
with the shell geometry of src/Fields/CubedSphereWithThickness{Map,Metric,InvMetric}.jl (oracle/geometry.py csphere3d_*) and
the contravariant Piola map of the Raviart-Thomas basis on the axis-aligned chart boxes (u = (1/det J) J uhat, div u =
divhat / det J; sign flip on the slave cell of every face).  The tutorial marches with ThetaMethod(Newton(LU), dt, 0.5); both
that march and the explicit Runge-Kutta march of the hot path (mass constant, slopes k_i = -M^-1 res) are restated here.

PARITY UNPINNED: the tutorial is a smoke test (no numbers to reproduce) and Gridap's hexahedral Raviart-Thomas element
(moment basis, DOF order, face orientation rule) is not in the reference tree; the element here follows the conventions of the
2-D element (oracle/refel.py RaviartThomasQuad).  The oracle is checked through the properties the
scheme must have: unisolvence and H(div) conformity of the element, the divergence theorem cell by cell, the exact energy
balance  u.r_u + p.r_p/c^2 + b.r_b/N^2 = 0  of the semi-discrete system (the Coriolis term is skew, the pressure and buoyancy
couplings cancel), second-order energy conservation of the theta = 1/2 march, and convergence of the RT interpolant.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import geometry as geo
from .refel import LagrangeQ, gauss_legendre_01, tensor_quadrature
from .rk import TABLEAUS
from .shell import entity_ids

# faces of the reference hexahedron in Gridap order (vertex sets (1,2,3,4), (5,6,7,8), (1,2,5,6), (3,4,7,8), (1,3,5,7),
# (2,4,6,8)): (fixed axis, value); the two spanning axes in increasing order parametrise the face
HEX_FACES = ((2, 0), (2, 1), (1, 0), (1, 1), (0, 0), (0, 1))


class RaviartThomasHex:
    """RT_k on [0,1]^3: 3 (k+2)(k+1)^2 DOFs -- (k+1)^2 moments of the outward normal flux against the monomials s^i t^j of every
    face (i fastest), face by face, then the interior moments of component a against x^i y^j z^l with degree k-1 in x_a."""

    def __init__(self, k):
        self.k = k
        self.pre = []
        for c in range(3):
            deg = [k, k, k]
            deg[c] = k + 1
            for l in range(deg[2] + 1):
                for j in range(deg[1] + 1):
                    for i in range(deg[0] + 1):
                        self.pre.append((c, i, j, l))
        self.ndofs = len(self.pre)
        self.nface_dofs = (k + 1) ** 2
        pts, W = self.moment_points()
        P = self._prebasis(pts)                                    # (N, n, 3)
        M = np.einsum("mnc,nic->mi", W, P)
        Ml = M.astype(np.longdouble)
        Cl = np.linalg.inv(M).astype(np.longdouble)
        for _ in range(2):
            Cl = Cl + Cl @ (np.eye(self.ndofs, dtype=np.longdouble) - Ml @ Cl)
        self.C = Cl.astype(np.float64)

    def _prebasis(self, pts):
        out = np.zeros((pts.shape[0], self.ndofs, 3))
        for n, (c, i, j, l) in enumerate(self.pre):
            out[:, n, c] = pts[:, 0] ** i * pts[:, 1] ** j * pts[:, 2] ** l
        return out

    def _prebasis_div(self, pts):
        out = np.zeros((pts.shape[0], self.ndofs))
        for n, (c, i, j, l) in enumerate(self.pre):
            e = [i, j, l]
            if e[c] > 0:
                f = float(e[c])
                e[c] -= 1
                out[:, n] = f * pts[:, 0] ** e[0] * pts[:, 1] ** e[1] * pts[:, 2] ** e[2]
        return out

    def tabulate(self, pts):
        """values (Q, n, 3), divergence (Q, n) in reference coordinates."""
        pts = np.asarray(pts, dtype=np.float64)
        return np.einsum("qic,ij->qjc", self._prebasis(pts), self.C), self._prebasis_div(pts) @ self.C

    def moment_points(self):
        """pts (N, 3) and W (ndofs, N, 3) with dof_m(f) = sum_{n,c} W[m,n,c] f(pts[n])[c]: faces first, then the interior."""
        k = self.k
        xq, wq = gauss_legendre_01(k + 2)
        S, T = np.meshgrid(xq, xq, indexing="ij")
        s, t = S.flatten(order="F"), T.flatten(order="F")             # s fastest
        ws = np.outer(wq, wq).flatten(order="F")
        pts_list = []
        for ax, val in HEX_FACES:
            sp_ax = [a for a in range(3) if a != ax]
            p = np.zeros((len(s), 3))
            p[:, ax] = val
            p[:, sp_ax[0]], p[:, sp_ax[1]] = s, t
            pts_list.append(p)
        nfp = len(s)
        if k > 0:
            pi, wi = tensor_quadrature(3, 2 * (k + 1))
            pts_list.append(pi)
        pts = np.concatenate(pts_list, axis=0)
        W = np.zeros((self.ndofs, pts.shape[0], 3))
        row = 0
        for f, (ax, val) in enumerate(HEX_FACES):
            sl = slice(f * nfp, (f + 1) * nfp)
            for j in range(k + 1):
                for i in range(k + 1):
                    W[row, sl, ax] = ws * s**i * t**j * (1.0 if val == 1 else -1.0)     # outward normal
                    row += 1
        if k > 0:
            for c in range(3):
                deg = [k, k, k]
                deg[c] = k - 1
                for l in range(deg[2] + 1):
                    for j in range(deg[1] + 1):
                        for i in range(deg[0] + 1):
                            W[row, 6 * nfp:, c] = wi * pi[:, 0] ** i * pi[:, 1] ** j * pi[:, 2] ** l
                            row += 1
        assert row == self.ndofs
        return pts, W


class HexRTSpace:
    """H(div)-conforming RT_k on the shell; face DOFs negated on the slave (second) cell of a face; the DOFs of the faces on the
    bottom (gamma = 0) and top (gamma = 1) boundaries are Dirichlet (-1) when `no_flux` (dirichlet_tags of the tutorial)."""
    kind = "RT"

    def __init__(self, mesh, k, no_flux=True):
        assert mesh.D == 3
        self.mesh, self.k = mesh, k
        self.ref = RaviartThomasHex(k)
        Nc = mesh.num_cells
        cf, nfaces = entity_ids(mesh.cell_nodes, 3, 2)                # (Nc, 6) in Gridap local order
        nf, ni = (k + 1) ** 2, 3 * k * (k + 1) ** 2
        first = np.full(nfaces, -1, dtype=np.int64)
        count = np.zeros(nfaces, dtype=np.int64)
        for c in range(Nc):
            for f in range(6):
                if first[cf[c, f]] < 0:
                    first[cf[c, f]] = c
                count[cf[c, f]] += 1
        boundary = count == 1
        # free faces are numbered in the order of the global face ids, skipping the Dirichlet ones
        fid = np.full(nfaces, -1, dtype=np.int64)
        free = ~boundary if no_flux else np.ones(nfaces, dtype=bool)
        fid[free] = np.arange(free.sum())
        nfree_faces = int(free.sum())
        ids = np.empty((Nc, 6 * nf + ni), dtype=np.int64)
        signs = np.ones((Nc, 6 * nf + ni))
        col = 0
        for f in range(6):
            slave = first[cf[:, f]] != np.arange(Nc)
            for t in range(nf):
                ids[:, col] = np.where(fid[cf[:, f]] >= 0, fid[cf[:, f]] * nf + t, -1)
                signs[slave, col] = -1.0
                col += 1
        for t in range(ni):
            ids[:, col] = nfree_faces * nf + np.arange(Nc) * ni + t
            col += 1
        self.cell_faces, self.num_faces, self.boundary_faces = cf, nfaces, boundary
        self.cell_dofs, self.signs = ids, signs
        self.num_free = nfree_faces * nf + Nc * ni
        self.ndofs_loc = ids.shape[1]


class HexDGSpace:
    kind = "DG"

    def __init__(self, mesh, p):
        self.mesh, self.p = mesh, p
        self.ref = LagrangeQ(3, p)
        m = self.ref.ndofs
        self.cell_dofs = np.arange(mesh.num_cells * m).reshape(mesh.num_cells, m)
        self.num_free, self.ndofs_loc, self.signs = mesh.num_cells * m, m, None


class ShellRTQuadrature:
    """Measure(Omega, degree) on the shell with what the RT forms need at the points: g, meas, the radial unit normal."""

    def __init__(self, mesh, degree):
        self.mesh, self.degree = mesh, degree
        self.pts, self.w = tensor_quadrature(3, degree)
        c = mesh.cell_chart_coords
        lo, hi = c[:, 0, :], c[:, 7, :]
        self.h = hi - lo
        self.x = lo[:, None, :] + self.pts[None, :, :] * self.h[:, None, :]
        self.detJ = np.prod(self.h, axis=1)
        self.dV = self.w[None, :] * self.detJ[:, None]
        r, t = mesh.radius, mesh.thickness
        self.g = geo.csphere3d_metric(r, t, self.x)
        self.ginv = geo.csphere3d_inv_metric(r, t, self.x)
        self.meas = geo.csphere3d_sqrtg(r, t, self.x)
        n3 = np.zeros(self.x.shape)
        n3[..., 0] = 1.0 / np.sqrt(self.ginv[..., 0, 0])             # normal_3D_cf = n3D / sqrt(n3D.(ginv.n3D))
        self.normal = n3

    def ambient(self):
        out = np.empty(self.x.shape)
        for p in range(6):
            sel = self.mesh.cell_to_chart == p
            out[sel] = geo.csphere3d_eval(p + 1, self.mesh.radius, self.mesh.thickness, self.x[sel])
        return out


def rt_basis(space, quad):
    """u (Nc,Q,m,3) = s (1/det J) J uhat with J = diag(h); div u (Nc,Q,m) = s divhat / det J."""
    val, div = space.ref.tabulate(quad.pts)
    u = val[None] * quad.h[:, None, None, :] / quad.detJ[:, None, None, None]
    d = div[None] / quad.detJ[:, None, None]
    s = space.signs
    return u * s[:, None, :, None], d * s[:, None, :]


def _assemble(test, trial, Ke):
    rows = np.repeat(test.cell_dofs[:, :, None], trial.ndofs_loc, axis=2)
    cols = np.repeat(trial.cell_dofs[:, None, :], test.ndofs_loc, axis=1)
    keep = (rows >= 0) & (cols >= 0)
    A = sp.coo_matrix((Ke[keep], (rows[keep], cols[keep])), shape=(test.num_free, trial.num_free)).tocsr()
    A.sort_indices()
    return A


def _assemble_vec(space, fe):
    out = np.zeros(space.num_free)
    keep = space.cell_dofs >= 0
    np.add.at(out, space.cell_dofs[keep], fe[keep])
    return out


def interpolate_rt(space, vfun):
    """Moment interpolation of u = meas * (pinvJ(covariant basis) . v(xyz))  (LinearBoussinesq.jl:107-108): the chart-space
    contravariant components, pulled back to the reference cell (uhat = det J J^-1 u) before the moments are taken."""
    mesh = space.mesh
    pts, W = space.ref.moment_points()
    c = mesh.cell_chart_coords
    lo, hi = c[:, 0, :], c[:, 7, :]
    h = hi - lo
    x = lo[:, None, :] + pts[None, :, :] * h[:, None, :]
    u = np.empty(x.shape)
    for p in range(6):
        sel = mesh.cell_to_chart == p
        X = geo.csphere3d_eval(p + 1, mesh.radius, mesh.thickness, x[sel])
        J = geo.csphere3d_jac(p + 1, mesh.radius, mesh.thickness, x[sel])         # J[i,k] = d X_k / d x_i
        # pinvJ of the square Jacobian is its inverse: contravariant components c with sum_i c_i J[i,:] = v
        comp = np.linalg.solve(np.swapaxes(J, -1, -2), vfun(X)[..., None])[..., 0]
        u[sel] = geo.csphere3d_sqrtg(mesh.radius, mesh.thickness, x[sel])[..., None] * comp
    uhat = np.prod(h, axis=1)[:, None, None] * u / h[:, None, :]
    dofs = np.einsum("mnc,bnc->bm", W, uhat) * space.signs
    out = np.zeros(space.num_free)
    keep = space.cell_dofs >= 0
    out[space.cell_dofs[keep]] = dofs[keep]
    return out


def interpolate_scalar(space, fun):
    mesh = space.mesh
    nodes = space.ref.nodes
    c = mesh.cell_chart_coords
    lo, hi = c[:, 0, :], c[:, 7, :]
    x = lo[:, None, :] + nodes[None, :, :] * (hi - lo)[:, None, :]
    vals = np.empty(x.shape[:2])
    for p in range(6):
        sel = mesh.cell_to_chart == p
        vals[sel] = fun(geo.csphere3d_eval(p + 1, mesh.radius, mesh.thickness, x[sel]))
    out = np.zeros(space.num_free)
    out[space.cell_dofs] = vals
    return out


# ---- initial state and parameters of the tutorial (LinearBoussinesq.jl:63-91, 126-130) ------------------------------------------
def b0(radius=1.0):
    def f(X):
        th, ph, _ = geo.xyz2thetaphir(X)
        thc, phc, d, zeta = 2 * np.pi / 3, 0.0, 0.095, 0.38
        k = np.sqrt((X ** 2).sum(-1)) - radius
        q = np.arccos(np.clip(np.sin(phc) * np.sin(ph) + np.cos(phc) * np.cos(ph) * np.cos(th - thc), -1.0, 1.0))
        return d * d / (d * d + q * q) * np.sin(2 * np.pi * k / zeta)
    return f


def u0(X):
    return 0.058 * np.stack([-X[..., 1], X[..., 0], 0 * X[..., 0]], -1)


def omega(X):
    _, ph, _ = geo.xyz2thetaphir(X)
    return 2 * 0.01 * np.sin(ph)


class BoussinesqProblem:
    def __init__(self, mesh, k, degree=None, c=1.0, N=1.48, omega_fun=omega):
        self.mesh, self.k, self.c2, self.N2 = mesh, k, c * c, N * N
        self.V, self.Q = HexRTSpace(mesh, k), HexDGSpace(mesh, k)
        self.degree = 4 * (k + 1) if degree is None else degree
        self.quad = q = ShellRTQuadrature(mesh, self.degree)
        self.om = omega_fun(q.ambient())
        self.nu, self.np_ = self.V.num_free, self.Q.num_free
        self.u, self.div = rt_basis(self.V, q)                        # (Nc,Q,m,3), (Nc,Q,m)
        self.dg = self.Q.ref.tabulate(q.pts)[0]                       # (Q,mq)
        self.gu = np.einsum("cqij,cqmj->cqmi", q.g, self.u)
        self.Mu_e = np.einsum("cq,cqai,cqbi->cab", q.dV / q.meas, self.u, self.gu)
        self.Mp_e = np.einsum("cq,qa,qb->cab", q.dV * q.meas, self.dg, self.dg)
        self.Mu = _assemble(self.V, self.V, self.Mu_e).tocsc()
        self.Mp = _assemble(self.Q, self.Q, self.Mp_e).tocsc()
        self.lu_u, self.lu_p = spla.splu(self.Mu), spla.splu(self.Mp)
        self._J = None

    def split(self, x):
        return x[: self.nu], x[self.nu: self.nu + self.np_], x[self.nu + self.np_:]

    def _gather_u(self, u):
        ids = self.V.cell_dofs
        return np.where(ids >= 0, u[np.maximum(ids, 0)], 0.0)

    def element_residuals(self, x):
        """(r_u (Nc,mu), r_p (Nc,mq), r_b (Nc,mq)) of res = resu + resp + resb."""
        q = self.quad
        u, p, b = self.split(x)
        ue, pe, be = self._gather_u(u), p[self.Q.cell_dofs], b[self.Q.cell_dofs]
        guq = np.einsum("cqmi,cm->cqi", self.gu, ue)                  # g . u
        uq = np.einsum("cqmi,cm->cqi", self.u, ue)
        divu = np.einsum("cqm,cm->cq", self.div, ue)
        pq, bq = np.einsum("qa,ca->cq", self.dg, pe), np.einsum("qa,ca->cq", self.dg, be)
        cor = np.cross(q.normal, guq / q.meas[..., None])             # n x (g u / meas)
        r_u = (np.einsum("cq,cqi,cqmi->cm", q.dV * self.om / q.meas, cor, self.gu)
               - np.einsum("cq,cqm->cm", q.dV * pq, self.div)
               - np.einsum("cq,cqi,cqmi->cm", q.dV * bq, q.normal, self.u))
        r_p = self.c2 * np.einsum("cq,qa->ca", q.dV * divu, self.dg)
        r_b = self.N2 * np.einsum("cq,qa->ca", q.dV * np.einsum("cqi,cqi->cq", q.normal, uq), self.dg)
        return r_u, r_p, r_b

    def residual(self, x):
        r_u, r_p, r_b = self.element_residuals(x)
        return np.concatenate([_assemble_vec(self.V, r_u), _assemble_vec(self.Q, r_p), _assemble_vec(self.Q, r_b)])

    def jacobian(self):
        """The (constant) matrix of the linear residual, column by column of the element operators."""
        if self._J is None:
            q = self.quad
            n = q.normal
            cor = np.cross(n[:, :, None, :], self.gu / q.meas[..., None, None])                   # (Nc,Q,mb,3)
            Kuu = np.einsum("cq,cqbi,cqai->cab", q.dV * self.om / q.meas, cor, self.gu)
            Kup = -np.einsum("cq,cqa,qb->cab", q.dV, self.div, self.dg)
            Kub = -np.einsum("cq,cqi,cqai,qb->cab", q.dV, n, self.u, self.dg)
            Kpu = self.c2 * np.einsum("cq,qa,cqb->cab", q.dV, self.dg, self.div)
            Kbu = self.N2 * np.einsum("cq,qa,cqi,cqbi->cab", q.dV, self.dg, n, self.u)
            V, Q = self.V, self.Q
            Z = sp.csr_matrix((self.np_, self.np_))
            self._J = sp.bmat([[_assemble(V, V, Kuu), _assemble(V, Q, Kup), _assemble(V, Q, Kub)],
                               [_assemble(Q, V, Kpu), Z, Z], [_assemble(Q, V, Kbu), Z, Z]]).tocsr()
        return self._J

    def mass_matrix(self):
        return sp.block_diag([self.Mu, self.Mp, self.Mp]).tocsc()

    def slope(self, x):
        r = self.residual(x)
        ru, rp, rb = self.split(r)
        return np.concatenate([self.lu_u.solve(-ru), self.lu_p.solve(-rp), self.lu_p.solve(-rb)])

    def step(self, x0, dt, tableau="EXRK_SSP_3_3"):
        A, b, c = TABLEAUS[tableau]
        ks = []
        for i in range(len(c)):
            xi = x0.copy()
            for j in range(i):
                xi += dt * A[i, j] * ks[j]
            ks.append(self.slope(xi))
        xF = x0.copy()
        for i in range(len(c)):
            xF += dt * b[i] * ks[i]
        return xF

    def theta_step(self, x0, dt, theta=0.5):
        """ThetaMethod(Newton(LU), dt, theta) on the linear system (LinearBoussinesq.jl:150-152): M (x_th - x_0)/(th dt) +
        J x_th = 0, x_1 = x_0 + (x_th - x_0)/th."""
        M, J = self.mass_matrix(), self.jacobian()
        if getattr(self, "_theta_key", None) != (dt, theta):
            self._theta_lu = spla.splu((M / (theta * dt) + J).tocsc())
            self._theta_key = (dt, theta)
        xth = self._theta_lu.solve(M @ x0 / (theta * dt))
        return x0 + (xth - x0) / theta

    def energy(self, x):
        u, p, b = self.split(x)
        return 0.5 * (u @ (self.Mu @ u) + p @ (self.Mp @ p) / self.c2 + b @ (self.Mp @ b) / self.N2)

    def initial_state(self):
        return np.concatenate([interpolate_rt(self.V, u0), np.zeros(self.np_), interpolate_scalar(self.Q, b0(self.mesh.radius))])
