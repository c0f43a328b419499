"""Oracle (test infrastructure): the 3-D cubed-sphere shell -- mesh, hexahedral Q_p numbering, Laplace-Beltrami forms.

Restates
  /root/reference/src/Geometry/CubedSphereWithThicknessMesh.jl:33-124  16-node / 6-hex coarse shell, chart box
                                                                       [0,1] x [-pi/4,pi/4]^2, coordinate order (gamma, alpha, beta)
  /root/reference/src/Geometry/AtlasGrids.jl:227-304                   one-shot refinement by n in every direction, parent-major
                                                                       cell order, reference-grid (x = gamma fastest) order inside
  /root/reference/src/Fields/CubedSphereWithThickness{Map,Metric,InvMetric}.jl  (geometry: oracle/geometry.py csphere3d_*)
  /root/reference/test/Laplacian/LaplaceBeltrami.jl:47                 a(u,v) = int grad(v).(ginv.grad(u)) sqrtg, here in 3-D
PARITY UNPINNED (Gridap internals): local order of hex nodes = vertices (x fastest), edges (4 x-dir,
4 y-dir, 4 z-dir), faces (z=0, z=1, y=0, y=1, x=0, x=1), interior; edge / face ids by first encounter over cells and local
entities.  Fine-mesh NODE ids are numbered by first encounter too (cells in order, local vertices in order), with nodes
identified exactly through their integer position on the cube surface lattice; the 2-D mesh uses the Q_n numbering of the
coarse cube instead -- both are restatement choices the reference tests do not pin.
Written as brute-force tensor quadrature over all nq^3 points (the GPU path factorises the integrals instead).
"""
import numpy as np

from . import geometry as geo
from .refel import LagrangeQ, lagrange_node_multi_indices, ncube_faces, tensor_quadrature

H = geo.CUBE_HALF_EDGE


def _cube_lattice_point(panel, i, j, n):
    """Integer position on the cube surface of the panel lattice point (i, j): perm_p(n, 2i - n, 2j - n) (App. A.1)."""
    h = (n, 2 * i - n, 2 * j - n)
    jj, ss = geo._perm(panel + 1)
    return tuple(int(ss[k]) * h[jj[k]] for k in range(3))


class CubedSphereShellMesh3D:
    """AtlasDiscreteModel(CubedSphereWithThicknessMesh(radius, thickness), l): n x n cells per panel, L cells through the shell
    (the reference refines isotropically: L = n)."""

    def __init__(self, n, radius=1.0, thickness=0.1, n_layers=None, ordering=0):
        """ordering 2: the cell order of a uniformly refined p8est forest
        (/root/reference/src/Distributed/AtlasOctreeDistributedDiscreteModels.jl:164-208): Morton order inside each tree, octant
        coordinates (x, y, z) = (gamma, alpha, beta) -- the coarse hexahedron's vertices are handed to p8est in Gridap's
        lexicographic order -- with the x bit lowest."""
        L = n if n_layers is None else n_layers
        assert ordering in (0, 2) and (ordering == 0 or (L == n and n & (n - 1) == 0))
        self.n, self.L, self.radius, self.thickness, self.D = n, L, radius, thickness, 3
        self.num_cells = 6 * n * n * L
        nodes = {}
        cn = np.empty((self.num_cells, 8), dtype=np.int64)
        cc = np.empty((self.num_cells, 8, 3))
        c2c = np.empty(self.num_cells, dtype=np.int64)
        col = np.empty(self.num_cells, dtype=np.int64)
        lay = np.empty(self.num_cells, dtype=np.int64)
        c = 0
        # cell order inside a panel: gamma fastest, then alpha, then beta
        def cells_of_a_panel():
            if ordering == 0:
                for jb in range(n):
                    for ia in range(n):
                        for kg in range(L):
                            yield kg, ia, jb
            else:
                for t in range(n * n * L):
                    x = [0, 0, 0]
                    for b in range(n.bit_length() - 1):
                        for a in range(3):
                            x[a] |= ((t >> (3 * b + a)) & 1) << b
                    yield x[0], x[1], x[2]

        for p in range(6):
            for kg, ia, jb in cells_of_a_panel():
                if True:
                    if True:
                        for v in range(8):
                            dg, da, db = v & 1, (v >> 1) & 1, (v >> 2) & 1
                            key = _cube_lattice_point(p, ia + da, jb + db, n) + (kg + dg,)
                            if key not in nodes:
                                nodes[key] = len(nodes)
                            cn[c, v] = nodes[key]
                            cc[c, v] = ((kg + dg) / L, -H + (ia + da) * 2 * H / n, -H + (jb + db) * 2 * H / n)
                        c2c[c] = p
                        col[c] = p * n * n + jb * n + ia
                        lay[c] = kg
                        c += 1
        self.cell_nodes, self.cell_chart_coords, self.cell_to_chart = cn, cc, c2c
        self.cell_column, self.cell_layer = col, lay
        self.num_nodes = len(nodes)


def entity_ids(cell_verts, D, d):
    """ids of the d-dimensional entities (edges d=1, faces d=2) by first encounter; returns (Nc, n_local) ids and the count."""
    faces = ncube_faces(D)[d]
    keys = {}
    out = np.empty((cell_verts.shape[0], len(faces)), dtype=np.int64)
    for c in range(cell_verts.shape[0]):
        for f, (S, fixed) in enumerate(faces):
            vs = []
            for v in range(1 << D):
                if all(((v >> a) & 1) == b for a, b in fixed.items()):
                    vs.append(cell_verts[c, v])
            k = tuple(sorted(vs))
            if k not in keys:
                keys[k] = len(keys)
            out[c, f] = keys[k]
    return out, len(keys)


class HexCGSpace:
    """H1-conforming Q_p (p = 1, 2) on the hexahedral shell mesh; `zeromean=True` fixes the last DOF."""
    kind = "CG"

    def __init__(self, mesh, p, zeromean=False):
        assert mesh.D == 3 and p in (1, 2)
        self.mesh, self.p = mesh, p
        self.ref = LagrangeQ(3, p)
        Nc = mesh.num_cells
        ids = [mesh.cell_nodes]
        n = mesh.num_nodes
        if p == 2:
            e, ne = entity_ids(mesh.cell_nodes, 3, 1)
            f, nf = entity_ids(mesh.cell_nodes, 3, 2)
            ids += [n + e, n + ne + f, (n + ne + nf + np.arange(Nc))[:, None]]
            n = n + ne + nf + Nc
            self.num_edges, self.num_faces = ne, nf
        ids = np.concatenate(ids, axis=1)
        self.num_dofs_total = n
        if zeromean:
            ids = np.where(ids == n - 1, -1, ids)
            n -= 1
        self.cell_dofs, self.num_free, self.ndofs_loc, self.signs = ids, n, ids.shape[1], None


class ShellQuadrature:
    """Measure(Omega, degree) on the shell: tensor Gauss-Legendre with nq^3 points per hexahedron, geometry at the points."""

    def __init__(self, mesh, degree):
        self.mesh, self.degree = mesh, degree
        self.pts, self.w = tensor_quadrature(3, degree)           # reference points, first coordinate (gamma) fastest
        c = mesh.cell_chart_coords
        lo, hi = c[:, 0, :], c[:, 7, :]
        self.h = hi - lo                                           # (Nc, 3)
        self.x = lo[:, None, :] + self.pts[None, :, :] * self.h[:, None, :]
        self.dV = self.w[None, :] * np.prod(self.h, axis=1)[:, None]
        self.ginv = geo.csphere3d_inv_metric(mesh.radius, mesh.thickness, self.x)
        self.sqrtg = geo.csphere3d_sqrtg(mesh.radius, mesh.thickness, self.x)

    def ambient(self):
        out = np.empty(self.x.shape)
        for p in range(6):
            sel = self.mesh.cell_to_chart == p
            out[sel] = geo.csphere3d_eval(p + 1, self.mesh.radius, self.mesh.thickness, self.x[sel])
        return out


def lb_element_matrices(space, quad):
    """Ke[c, I, J] = sum_q dV grad(phi_I).(ginv grad(phi_J)) sqrtg (test/Laplacian/LaplaceBeltrami.jl:47 with the shell metric)."""
    _, gref = space.ref.tabulate(quad.pts)                        # (Q, m, 3)
    out = np.empty((space.mesh.num_cells, space.ndofs_loc, space.ndofs_loc))
    for c in range(space.mesh.num_cells):                         # cell by cell: (Q, m, m) temporaries stay small
        g = gref / quad.h[c][None, None, :]
        flux = np.einsum("qij,qmj->qmi", quad.ginv[c], g)
        out[c] = np.einsum("q,qai,qbi->ab", quad.dV[c] * quad.sqrtg[c], g, flux)
    return out


def mass_element_matrices(space, quad):
    """Me[c, I, J] = sum_q dV phi_I phi_J sqrtg (the u*v*meas_cf term of test/Laplacian/Helmholtz.jl:43 with the shell measure)."""
    val, _ = space.ref.tabulate(quad.pts)                         # (Q, m)
    return np.einsum("cq,qa,qb->cab", quad.dV * quad.sqrtg, val, val)


def helmholtz_element_matrices(space, quad):
    """int u v sqrtg - int grad(v).(ginv grad(u)) sqrtg -- test/Laplacian/Helmholtz.jl:43."""
    return mass_element_matrices(space, quad) - lb_element_matrices(space, quad)


def source_element_vectors(space, quad, fq):
    """int f v sqrtg with f at the nq^3 points (Nc, Q) -- test/Laplacian/LaplaceBeltrami.jl:53."""
    val, _ = space.ref.tabulate(quad.pts)
    return np.einsum("cq,qa->ca", quad.dV * quad.sqrtg * fq, val)


def interpolate(space, fun):
    """Nodal interpolation of an ambient function (interpolate(f o ambient_map_cf, U), LaplaceBeltrami.jl:62)."""
    mesh = space.mesh
    cc = mesh.cell_chart_coords
    lo, hi = cc[:, 0, :], cc[:, 7, :]
    x = lo[:, None, :] + space.ref.nodes[None, :, :] * (hi - lo)[:, None, :]               # (Nc, m, 3) chart nodes
    U = np.zeros(space.num_dofs_total)
    for p in range(6):
        sel = mesh.cell_to_chart == p
        vals = fun(geo.csphere3d_eval(p + 1, mesh.radius, mesh.thickness, x[sel]))
        ids = space.cell_dofs[sel]
        U[np.where(ids >= 0, ids, space.num_dofs_total - 1)] = vals          # the fixed DOF of a zeromean space is the last one
    return U


def boundary_flux_element_vectors(space, degree, U):
    """int_G ((ginv grad(u_h)).n_G) v sqrtg dG over the bottom (gamma = 0) and top (gamma = 1) faces of the shell, u_h given by its
    values U on ALL dofs -- the boundary term of the 3-D driver, test/Laplacian/LaplaceBeltrami.jl:58-64.  n_G is the outward unit
    normal in chart space (-/+ e_gamma), dG the chart measure of the face; ginv_gg = 1/t^2, sqrtg = t s^2 sqrtg_surf."""
    mesh = space.mesh
    pts2, w2 = tensor_quadrature(2, degree)
    out = np.zeros((mesh.num_cells, space.ndofs_loc))
    cc = mesh.cell_chart_coords
    lo, hi = cc[:, 0, :], cc[:, 7, :]
    h = hi - lo
    for side, gref, sign in ((0, 0.0, -1.0), (mesh.L - 1, 1.0, 1.0)):
        cells = np.where(mesh.cell_layer == side)[0]
        ref = np.concatenate([np.full((len(pts2), 1), gref), pts2], axis=1)              # (Q2, 3) reference points on the face
        val, grad = space.ref.tabulate(ref)                                             # (Q2, m), (Q2, m, 3)
        x = lo[cells, None, :] + ref[None, :, :] * h[cells, None, :]
        sq = geo.csphere3d_sqrtg(mesh.radius, mesh.thickness, x)
        gi = geo.csphere3d_inv_metric(mesh.radius, mesh.thickness, x)
        Ue = U[np.where(space.cell_dofs[cells] >= 0, space.cell_dofs[cells], len(U) - 1)]
        gu = np.einsum("qmi,cm->cqi", grad, Ue) / h[cells, None, :]                      # chart gradient of u_h
        flux = sign * np.einsum("cqj,cqj->cq", gi[..., 0, :], gu)                        # (ginv grad u) . (+-e_gamma)
        dG = w2[None, :] * (h[cells, 1] * h[cells, 2])[:, None]
        out[cells] += np.einsum("cq,qa->ca", flux * sq * dG, val)
    return out


def eval_scalar(space, quad, U):
    val, _ = space.ref.tabulate(quad.pts)
    ids = space.cell_dofs
    return np.einsum("qa,ca->cq", val, U[np.where(ids >= 0, ids, len(U) - 1)])


def solve_lb(mesh, p, f, lap_f, degree=None):
    """laplace_beltrami_solver on the shell (test/Laplacian/LaplaceBeltrami.jl:19-98 with Dc = 3): returns the L2 error."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    degree = 6 * (p + 1) if degree is None else degree
    V = HexCGSpace(mesh, p, zeromean=True)
    q = ShellQuadrature(mesh, degree)
    m = V.ndofs_loc
    ids = V.cell_dofs
    Ke = lb_element_matrices(V, q)
    keep = (ids[:, :, None] >= 0) & (ids[:, None, :] >= 0)
    rows = np.broadcast_to(ids[:, :, None], Ke.shape)[keep]
    cols = np.broadcast_to(ids[:, None, :], Ke.shape)[keep]
    A = sp.coo_matrix((Ke[keep], (rows, cols)), shape=(V.num_free,) * 2).tocsc()
    X = q.ambient()
    fe = source_element_vectors(V, q, -lap_f(X))
    Uint = interpolate(V, f)
    fe = fe + boundary_flux_element_vectors(V, degree, Uint)
    b = np.zeros(V.num_free)
    np.add.at(b, ids[ids >= 0], fe[ids >= 0])
    x = spla.spsolve(A, b)
    xf = np.append(x, 0.0)
    # zero-mean shift in the CHART measure (Gridap ZeroMeanFESpace)
    val, _ = V.ref.tabulate(q.pts)
    vol_i = np.zeros(V.num_dofs_total)
    np.add.at(vol_i, np.where(ids >= 0, ids, V.num_dofs_total - 1), np.einsum("cq,qa->ca", q.dV, val))
    xf = xf - (vol_i @ xf) / vol_i.sum()
    qe = ShellQuadrature(mesh, min(2 * degree, 24))
    e = f(qe.ambient()) - eval_scalar(V, qe, xf)
    return float(np.sqrt((e * e * qe.dV * qe.sqrtg).sum()))


def shell_volume(mesh, degree):
    q = ShellQuadrature(mesh, degree)
    return float((q.dV * q.sqrtg).sum())
