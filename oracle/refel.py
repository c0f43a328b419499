"""Oracle (test infrastructure): quadrature and reference finite elements on the n-cube.

The reference delegates all of this to Gridap 0.20.9 (un-vendored: /root/reference/Project.toml:19-31);
the conventions restated here are SURVEY.md App. B.1-B.4/B.8 and are PARITY UNPINNED at the level of
local DOF order / RT moment bases (physics-level results do not depend on them).

  * quadrature: tensor Gauss-Legendre on [0,1]^D, n1 = ceil((degree+1)/2), first coordinate fastest
    (call sites: test/Laplacian/LaplaceBeltrami.jl:29-32 `Measure(Omega, degree)`).
  * Lagrangian Q_p: equispaced nodes; local order = vertices, edge interiors (edge by edge), [faces], interior.
  * Raviart-Thomas RT_k on the square: prebasis Q_{k+1,k} x Q_{k,k+1}; facet moments against monomials
    of degree <= k with the OUTWARD reference normal, interior moments against Q_{k-1,k} x Q_{k,k-1}.
"""
import itertools
import math
import numpy as np


# ---- quadrature ----------------------------------------------------------------------------------

def gauss_legendre_01(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return (x + 1.0) / 2.0, w / 2.0


def num_points_1d(degree):
    return int(math.ceil((degree + 1) / 2.0))


def tensor_quadrature(D, degree):
    """points (Q, D), weights (Q,), first coordinate fastest."""
    n = num_points_1d(degree)
    x, w = gauss_legendre_01(n)
    grids = np.meshgrid(*([x] * D), indexing="ij")  # axis d <-> coordinate d
    wg = np.meshgrid(*([w] * D), indexing="ij")
    # flatten with first coordinate fastest -> Fortran order
    pts = np.stack([g.flatten(order="F") for g in grids], axis=-1)
    wts = np.ones(n**D)
    for g in wg:
        wts = wts * g.flatten(order="F")
    return pts, wts


# ---- 1-D Lagrange --------------------------------------------------------------------------------

def lagrange_1d(p, x):
    """values (len(x), p+1) and derivatives of the Lagrange basis on nodes i/p."""
    x = np.asarray(x, dtype=np.float64)
    nodes = np.arange(p + 1) / p if p > 0 else np.array([0.5])
    n = len(nodes)
    V = np.ones((len(x), n))
    dV = np.zeros((len(x), n))
    for i in range(n):
        for j in range(n):
            if j != i:
                V[:, i] *= (x - nodes[j]) / (nodes[i] - nodes[j])
        for j in range(n):
            if j == i:
                continue
            term = np.full(len(x), 1.0 / (nodes[i] - nodes[j]))
            for l in range(n):
                if l != i and l != j:
                    term *= (x - nodes[l]) / (nodes[i] - nodes[l])
            dV[:, i] += term
    return V, dV


# ---- n-cube topology (Gridap ordering) -------------------------------------------

def ncube_faces(D):
    """List over d of lists of (spanning_axes, fixed dict axis->0/1) in Gridap order."""
    out = []
    for d in range(D + 1):
        faces = []
        for S in itertools.combinations(range(D), d):
            fixed_axes = [a for a in range(D) if a not in S]
            for bits in itertools.product((0, 1), repeat=len(fixed_axes)):
                bits = bits[::-1]  # first fixed axis fastest
                faces.append((S, dict(zip(fixed_axes, bits))))
        out.append(faces)
    return out


def lagrange_node_multi_indices(D, p):
    """Local node -> tensor multi-index (ix, iy[, iz]) in {0..p}^D, Gridap local order; plus node -> (d, face)."""
    nodes, owner = [], []
    for d, faces in enumerate(ncube_faces(D)):
        for f, (S, fixed) in enumerate(faces):
            for inner in itertools.product(range(1, p), repeat=len(S)):
                inner = inner[::-1]  # first spanning axis fastest
                mi = [0] * D
                for a, v in fixed.items():
                    mi[a] = v * p
                for a, v in zip(S, inner):
                    mi[a] = v
                nodes.append(tuple(mi))
                owner.append((d, f))
    if p == 0:
        return [tuple([0] * D)], [(D, 0)]
    return nodes, owner


class LagrangeQ:
    """Scalar Lagrangian Q_p on [0,1]^D (CG when glued, DG when conformity=:L2)."""

    def __init__(self, D, p):
        self.D, self.p = D, p
        self.mi, self.owner = lagrange_node_multi_indices(D, p)
        self.ndofs = len(self.mi)
        if p > 0:
            self.nodes = np.array(self.mi, dtype=np.float64) / p
        else:
            self.nodes = np.full((1, D), 0.5)

    def tabulate(self, pts):
        """values (Q, m), gradients (Q, m, D) in reference coordinates."""
        pts = np.asarray(pts, dtype=np.float64)
        Q = pts.shape[0]
        V1 = [lagrange_1d(self.p, pts[:, a]) for a in range(self.D)]
        val = np.ones((Q, self.ndofs))
        grad = np.ones((Q, self.ndofs, self.D))
        for k, mi in enumerate(self.mi):
            for a in range(self.D):
                val[:, k] *= V1[a][0][:, mi[a]]
                for b in range(self.D):
                    grad[:, k, b] *= V1[a][1][:, mi[a]] if a == b else V1[a][0][:, mi[a]]
        return val, grad


# ---- Raviart-Thomas on the square ---------------------------------------------------------------

QUAD_EDGES = ((0, 1), (2, 3), (0, 2), (1, 3))          # local vertex pairs (0-based), Gridap order
QUAD_VERTS = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]])
QUAD_OUT_NORMALS = np.array([[0.0, -1.0], [0.0, 1.0], [-1.0, 0.0], [1.0, 0.0]])


class RaviartThomasQuad:
    """RT_k on [0,1]^2 (Gridap `ReferenceFE(raviart_thomas, Float64, k)`), 2(k+1)(k+2) DOFs.

    DOF order: (k+1) moments per facet, facet by facet, then 2k(k+1) interior moments."""

    def __init__(self, k):
        self.k = k
        # prebasis exponents: component 0 in Q_{k+1,k}, component 1 in Q_{k,k+1}
        self.pre = []
        for j in range(k + 1):
            for i in range(k + 2):
                self.pre.append((0, i, j))
        for j in range(k + 2):
            for i in range(k + 1):
                self.pre.append((1, i, j))
        self.ndofs = len(self.pre)
        self.nfacet_dofs = k + 1
        M = np.zeros((self.ndofs, self.ndofs))
        xq, wq = gauss_legendre_01(k + 2)
        row = 0
        for f, (va, vb) in enumerate(QUAD_EDGES):
            a, b = QUAD_VERTS[va], QUAD_VERTS[vb]
            pts = a[None, :] + xq[:, None] * (b - a)[None, :]
            P = self._prebasis(pts)  # (Q, n, 2)
            flux = P @ QUAD_OUT_NORMALS[f]  # (Q, n)
            for m in range(k + 1):
                M[row, :] = (wq * xq**m) @ flux
                row += 1
        if k > 0:
            pts2, w2 = tensor_quadrature(2, 2 * (k + 1))
            P = self._prebasis(pts2)
            for comp, (imax, jmax) in ((0, (k - 1, k)), (1, (k, k - 1))):
                for j in range(jmax + 1):
                    for i in range(imax + 1):
                        q = pts2[:, 0] ** i * pts2[:, 1] ** j
                        M[row, :] = (w2 * q) @ P[:, :, comp]
                        row += 1
        assert row == self.ndofs
        # phi_j = sum_i C[i, j] psi_i; the monomial moment matrix has cond ~1e4 for k = 2, so polish the
        # inverse with two Newton-Schulz steps in extended precision
        Ml = M.astype(np.longdouble)
        Cl = np.linalg.inv(M).astype(np.longdouble)
        for _ in range(2):
            Cl = Cl + Cl @ (np.eye(self.ndofs, dtype=np.longdouble) - Ml @ Cl)
        self.C = Cl.astype(np.float64)

    def _prebasis(self, pts):
        out = np.zeros((pts.shape[0], self.ndofs, 2))
        for n, (c, i, j) in enumerate(self.pre):
            out[:, n, c] = pts[:, 0] ** i * pts[:, 1] ** j
        return out

    def _prebasis_div(self, pts):
        out = np.zeros((pts.shape[0], self.ndofs))
        for n, (c, i, j) in enumerate(self.pre):
            if c == 0 and i > 0:
                out[:, n] = i * pts[:, 0] ** (i - 1) * pts[:, 1] ** j
            if c == 1 and j > 0:
                out[:, n] = j * pts[:, 0] ** i * pts[:, 1] ** (j - 1)
        return out

    def tabulate(self, pts):
        """values (Q, n, 2), divergence (Q, n) in reference coordinates."""
        pts = np.asarray(pts, dtype=np.float64)
        val = np.einsum("qic,ij->qjc", self._prebasis(pts), self.C)
        div = self._prebasis_div(pts) @ self.C
        return val, div

    def dofs_of(self, fun):
        """Moment DOFs of a reference-space vector field fun(pts)->(Q,2) (used for interpolation)."""
        k = self.k
        out = np.zeros(self.ndofs)
        xq, wq = gauss_legendre_01(k + 2)
        row = 0
        for f, (va, vb) in enumerate(QUAD_EDGES):
            a, b = QUAD_VERTS[va], QUAD_VERTS[vb]
            pts = a[None, :] + xq[:, None] * (b - a)[None, :]
            flux = fun(pts) @ QUAD_OUT_NORMALS[f]
            for m in range(k + 1):
                out[row] = (wq * xq**m) @ flux
                row += 1
        if k > 0:
            pts2, w2 = tensor_quadrature(2, 2 * (k + 1))
            F = fun(pts2)
            for comp, (imax, jmax) in ((0, (k - 1, k)), (1, (k, k - 1))):
                for j in range(jmax + 1):
                    for i in range(imax + 1):
                        out[row] = (w2 * pts2[:, 0] ** i * pts2[:, 1] ** j) @ F[:, comp]
                        row += 1
        return out

    def moment_points(self):
        """All points at which `dofs_of` samples (facets then interior) and the linear functional weights.

        Returns pts (N,2) and W (ndofs, N, 2) such that dof_i(f) = sum_{n,c} W[i,n,c] f(pts[n])[c]."""
        k = self.k
        xq, wq = gauss_legendre_01(k + 2)
        pts_list, blocks = [], []
        for f, (va, vb) in enumerate(QUAD_EDGES):
            a, b = QUAD_VERTS[va], QUAD_VERTS[vb]
            pts_list.append(a[None, :] + xq[:, None] * (b - a)[None, :])
        nfp = len(xq) * 4
        if k > 0:
            pts2, w2 = tensor_quadrature(2, 2 * (k + 1))
            pts_list.append(pts2)
        pts = np.concatenate(pts_list, axis=0)
        W = np.zeros((self.ndofs, pts.shape[0], 2))
        row = 0
        for f in range(4):
            sl = slice(f * len(xq), (f + 1) * len(xq))
            for m in range(k + 1):
                W[row, sl, :] = (wq * xq**m)[:, None] * QUAD_OUT_NORMALS[f][None, :]
                row += 1
        if k > 0:
            for comp, (imax, jmax) in ((0, (k - 1, k)), (1, (k, k - 1))):
                for j in range(jmax + 1):
                    for i in range(imax + 1):
                        W[row, nfp:, comp] = w2 * pts2[:, 0] ** i * pts2[:, 1] ** j
                        row += 1
        return pts, W
