"""Oracle (test infrastructure): vector-valued H1 Lagrangian space on the intrinsic cubed sphere, with the panel-interface
change of basis, and the L2 projection driver that exercises it.

Restates
  /root/reference/src/FESpaces/GradConformingFESpaces.jl:1-24    master cell of a vertex / edge = the surrounding cell with the
                                                                 LARGEST cell id
  /root/reference/src/FESpaces/GradConformingFESpaces.jl:60-113  per-cell change-of-basis matrix: identity, except the 2x2 block
                                                                 of every node owned by an entity whose master cell lies on
                                                                 another panel: coeffs = pinvJ(J_slave)(x_node) . J_master^T(x_node)
                                                                 (the DOFs hold the contravariant components in the MASTER's chart)
  /root/reference/src/FESpaces/GradConformingFESpaces.jl:146-156 shape functions are combined with C, the DOF basis with C^-T
  /root/reference/test/Projection/L2_projection_Lagrangian_vector.jl:17-58
        a(u,v) = int (u.(g v)) sqrtg,  l(v) = int (f.(g v)) sqrtg,  f = pinvJ . vecX (contravariant components),
        error = sqrt(int e.(g e) sqrtg) with Measure(Omega, 2 degree), degree = 4 (p + 1)
  /root/reference/test/Projection/L2ProjectionTests.jl:13-15,36   vecX = (-y, x, 0), conformity H1, p in (1, 2), slope >= p
PARITY UNPINNED (Gridap internals): local DOF order of ReferenceFE(lagrangian, VectorValue{2}, p) = component-major
(dof = node + n_nodes * comp); global numbering = entity by entity (vertices, edges, cell interiors in the scalar Q_p order),
component-major inside an entity.
"""
import numpy as np
import scipy.sparse.linalg as spla

from . import forms
from . import geometry as geo
from .spaces import CGSpace, Space


class VectorCGSpace(Space):
    kind = "CGV"

    def __init__(self, mesh, p):
        self.mesh, self.p = mesh, p
        S = self.scalar = CGSpace(mesh, p)
        self.ref = S.ref
        nn = self.nn = S.ndofs_loc
        Nc, Nv, Ne = mesh.num_cells, mesh.num_nodes, mesh.num_edges
        ne, ni = p - 1, (p - 1) ** 2
        ids = np.empty((Nc, 2 * nn), dtype=np.int64)
        for a in range(nn):
            d, f = S.ref.owner[a]
            sid = S.cell_dofs[:, a]
            for comp in range(2):
                if d == 0:
                    g = 2 * sid + comp
                elif d == 1:
                    e, k = np.divmod(sid - Nv, ne)
                    g = 2 * Nv + e * 2 * ne + comp * ne + k
                else:
                    c, k = np.divmod(sid - Nv - Ne * ne, ni)
                    g = 2 * Nv + 2 * Ne * ne + c * 2 * ni + comp * ni + k
                ids[:, a + nn * comp] = g
        self.cell_dofs, self.ndofs_loc, self.num_free, self.signs = ids, 2 * nn, 2 * S.num_dofs_total, None
        self.cob = self._change_of_basis()

    def _node_chart_coords(self):
        cc = self.mesh.cell_chart_coords
        lo, hi = cc[:, 0, :], cc[:, 3, :]
        return lo[:, None, :] + self.ref.nodes[None, :, :] * (hi - lo)[:, None, :]          # (Nc, nn, 2)

    def _change_of_basis(self):
        """cob[c, a] = 2x2 block of node a in cell c (identity where the owning entity's master cell is on the same panel)."""
        mesh, nn = self.mesh, self.nn
        Nc = mesh.num_cells
        cob = np.tile(np.eye(2), (Nc, nn, 1, 1))
        x = self._node_chart_coords()
        panel = mesh.cell_to_chart
        # master cells: largest cell id around each vertex / edge
        vmaster = np.full(mesh.num_nodes, -1, dtype=np.int64)
        emaster = np.full(mesh.num_edges, -1, dtype=np.int64)
        for c in range(Nc):
            vmaster[mesh.cell_nodes[c]] = np.maximum(vmaster[mesh.cell_nodes[c]], c)
            emaster[mesh.cell_edges[c]] = np.maximum(emaster[mesh.cell_edges[c]], c)
        own_nodes = {}
        for a in range(nn):
            own_nodes.setdefault(self.ref.owner[a], []).append(a)
        for c in range(Nc):
            for (d, f), nodes in own_nodes.items():
                if d == 2:
                    continue
                ent = mesh.cell_nodes[c, f] if d == 0 else mesh.cell_edges[c, f]
                m = vmaster[ent] if d == 0 else emaster[ent]
                if panel[m] == panel[c]:
                    continue
                ents_m = mesh.cell_nodes[m] if d == 0 else mesh.cell_edges[m]
                fm = int(np.where(ents_m == ent)[0][0])
                for a_s, a_m in zip(nodes, own_nodes[(d, fm)]):
                    JM = geo.csphere_jac(panel[m] + 1, mesh.radius, x[m, a_m])                # (2,3): J[i,k]
                    JS = geo.csphere_jac(panel[c] + 1, mesh.radius, x[c, a_s])
                    cob[c, a_s] = geo.pinvJ(JS) @ JM.T                                       # (2,3) . (3,2)
        return cob

    def cell_matrix(self):
        """C[c] (m x m): shape function j of the space = sum_i C[i, j] (reference shape function i)."""
        Nc, nn = self.mesh.num_cells, self.nn
        C = np.zeros((Nc, 2 * nn, 2 * nn))
        for a in range(nn):
            for i in range(2):
                for j in range(2):
                    C[:, a + nn * i, a + nn * j] = self.cob[:, a, i, j]
        return C

    def eval(self, quad, U):
        """chart-space contravariant components (Nc, Q, 2) of the FE function with DOF vector U."""
        val = self.ref.tabulate(quad.pts)[0]                                                 # (Q, nn)
        Ue = U[self.cell_dofs].reshape(-1, 2, self.nn).transpose(0, 2, 1)                    # (Nc, nn, 2) master-frame components
        loc = np.einsum("caij,caj->cai", self.cob, Ue)                                       # this cell's chart
        return np.einsum("qa,cai->cqi", val, loc)


def contravariant(mesh, quad, vecX):
    out = np.empty(quad.x.shape)
    X, J = quad.ambient(), quad.ambient_jac()
    out[:] = np.einsum("...ik,...k->...i", geo.pinvJ(J), vecX(X))
    return out


def mass_element_matrices(V, quad):
    """int (u.(g v)) sqrtg in the space's basis: C^T M_ref C with M_ref[(a,i),(b,j)] = int N_a N_b g_ij sqrtg."""
    val = V.ref.tabulate(quad.pts)[0]
    nn = V.nn
    blocks = np.einsum("cq,qa,qb,cqij->caibj", quad.dV * quad.sqrtg, val, val, quad.g)       # (Nc, nn, 2, nn, 2)
    M = blocks.transpose(0, 2, 1, 4, 3).reshape(-1, 2 * nn, 2 * nn)                          # component-major local order
    C = V.cell_matrix()
    return np.einsum("cki,ckl,clj->cij", C, M, C)


def source_element_vectors(V, quad, fq):
    """int (f.(g v)) sqrtg with the contravariant field f (Nc, Q, 2) at the quadrature points."""
    val = V.ref.tabulate(quad.pts)[0]
    gf = np.einsum("cqij,cqj->cqi", quad.g, fq)
    b = np.einsum("cq,qa,cqi->cia", quad.dV * quad.sqrtg, val, gf).reshape(-1, 2 * V.nn)
    return np.einsum("cki,ck->ci", V.cell_matrix(), b)


def interpolate(V, vecX):
    """DOF values = C^-1 applied to the nodal contravariant components in the cell's own chart."""
    mesh = V.mesh
    x = V._node_chart_coords()
    U = np.zeros(V.num_free)
    for p in range(6):
        sel = np.where(mesh.cell_to_chart == p)[0]
        X = geo.csphere_eval(p + 1, mesh.radius, x[sel])
        J = geo.csphere_jac(p + 1, mesh.radius, x[sel])
        loc = np.einsum("...ik,...k->...i", geo.pinvJ(J), vecX(X))                           # (n, nn, 2)
        mas = np.einsum("caij,caj->cai", np.linalg.inv(V.cob[sel]), loc)
        U[V.cell_dofs[sel]] = mas.transpose(0, 2, 1).reshape(len(sel), -1)                   # all cells around a node agree
    return U


def l2_error(V, quad, U, fq):
    e = V.eval(quad, U) - fq
    return np.sqrt((np.einsum("cqi,cqij,cqj->cq", e, quad.g, e) * quad.dV * quad.sqrtg).sum())


def l2_projection(mesh, p, vecX):
    """L2_projection_Lagrangian_vector(..., :H1): returns (projection error, interpolation error, space, DOF vector)."""
    degree = 4 * (p + 1)
    V = VectorCGSpace(mesh, p)
    q = forms.CellQuadrature(mesh, degree)
    fq = contravariant(mesh, q, vecX)
    A = forms.assemble_matrix(V, V, mass_element_matrices(V, q))
    b = forms.assemble_vector(V, source_element_vectors(V, q, fq))
    U = spla.spsolve(A.tocsc(), b)
    qe = forms.CellQuadrature(mesh, 2 * degree)
    fe = contravariant(mesh, qe, vecX)
    return l2_error(V, qe, U, fe), l2_error(V, qe, interpolate(V, vecX), fe), V, U
