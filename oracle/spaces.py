"""Oracle (test infrastructure): FE spaces on the atlas mesh (DOF numbering, signs, interpolation).

The reference obtains these from Gridap (`TestFESpace(model, ReferenceFE(...); conformity, constraint)`,
call sites test/Laplacian/LaplaceBeltrami.jl:34-35, test/Transient/TransientWaveEquation.jl:40-47,
test/Transient/TransientShallowWater.jl:65-78).  PARITY UNPINNED: numbering follows SURVEY.md App. B.5/B.8.
"""
import numpy as np
from .refel import LagrangeQ, RaviartThomasQuad
from .mesh import cg_numbering


class Space:
    kind = None


class CGSpace(Space):
    """H1-conforming Lagrangian Q_p.  `zeromean=True` fixes the last free DOF (Gridap ZeroMeanFESpace)."""
    kind = "CG"

    def __init__(self, mesh, p, zeromean=False):
        self.mesh, self.p = mesh, p
        self.ref = LagrangeQ(mesh.D, p)
        ids, n = cg_numbering(mesh.cell_nodes, p, n_verts=mesh.num_nodes)
        self.num_dofs_total = n
        self.zeromean = zeromean
        if zeromean:
            self.fixed_dof = n - 1
            ids = np.where(ids == n - 1, -1, ids)  # -1 = the single Dirichlet DOF
            self.num_free = n - 1
        else:
            self.num_free = n
        self.cell_dofs = ids
        self.ndofs_loc = ids.shape[1]
        self.signs = None


class DGSpace(Space):
    kind = "DG"

    def __init__(self, mesh, p):
        self.mesh, self.p = mesh, p
        self.ref = LagrangeQ(mesh.D, p)
        m = self.ref.ndofs
        self.cell_dofs = np.arange(mesh.num_cells * m, dtype=np.int64).reshape(mesh.num_cells, m)
        self.num_free = mesh.num_cells * m
        self.ndofs_loc = m
        self.signs = None


class RTSpace(Space):
    """H(div)-conforming RT_k; facet DOFs negated on the slave (higher-id) cell of each facet."""
    kind = "RT"

    def __init__(self, mesh, k):
        assert mesh.D == 2
        self.mesh, self.k = mesh, k
        self.ref = RaviartThomasQuad(k)
        Nc, ne = mesh.num_cells, mesh.num_edges
        assert not mesh.cell_edge_flip.any(), "shared edges must have the same vertex order in both cells"
        nf, ni = k + 1, 2 * k * (k + 1)
        ids = np.empty((Nc, 4 * nf + ni), dtype=np.int64)
        signs = np.ones((Nc, 4 * nf + ni))
        fc = mesh.facet_cells()
        col = 0
        for e in range(4):
            slave = fc[mesh.cell_edges[:, e], 1] == np.arange(Nc)
            for t in range(nf):
                ids[:, col] = mesh.cell_edges[:, e] * nf + t
                signs[slave, col] = -1.0
                col += 1
        for t in range(ni):
            ids[:, col] = ne * nf + np.arange(Nc) * ni + t
            col += 1
        self.cell_dofs, self.signs = ids, signs
        self.num_free = ne * nf + Nc * ni
        self.ndofs_loc = ids.shape[1]
