"""Oracle (test infrastructure): cubed-sphere atlas mesh, topology and DOF numbering.

Restates
  /root/reference/src/Geometry/CubedSphereMesh.jl:83-116   (8-node / 6-cell coarse cube, panel corners)
  /root/reference/src/Geometry/AtlasGrids.jl:227-304       (one-shot refinement by n, parent-major cell order,
                                                           chart coordinates by tiling the reference grid)
  /root/reference/src/Distributed/AtlasDistributedDiscreteModels.jl:20-73 (linear partition + vertex-adjacent ghosts)
and, PARITY UNPINNED (Gridap 0.20.9 internals): edge ids in order of first
encounter scanning cells and local edges; conforming DOF ids = vertices, then edges, then cell
interiors; fine-mesh node ids = CG-Q_n numbering on the coarse cube.  Unlike the reference's public
constructors (n = 2^l only, AtlasGrids.jl:254-256) `n` is arbitrary here.
All ids are 0-based in this module; the C ABI and Gridap use 1-based ids.
"""
import numpy as np
from .refel import QUAD_EDGES, lagrange_node_multi_indices
from .geometry import CUBE_HALF_EDGE

# src/Geometry/CubedSphereMesh.jl:94 (1-based there)
COARSE_CELLS = np.array(
    [[1, 2, 3, 4], [3, 4, 5, 6], [2, 7, 4, 6], [8, 5, 7, 6], [1, 8, 2, 7], [1, 3, 8, 5]], dtype=np.int64
) - 1


def quad_edges(cell_verts):
    """Edge ids by first encounter.  Returns cell_edges (Nc,4), edge_verts (Ne,2) (direction of the
    first-encounter cell), cell_edge_flip (Nc,4) (True where this cell sees the edge reversed)."""
    Nc = cell_verts.shape[0]
    ga = cell_verts[:, [e[0] for e in QUAD_EDGES]].reshape(-1)
    gb = cell_verts[:, [e[1] for e in QUAD_EDGES]].reshape(-1)
    lo, hi = np.minimum(ga, gb), np.maximum(ga, gb)
    nv = int(cell_verts.max()) + 1
    key = lo * nv + hi
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")          # unique keys sorted by first occurrence
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    edge_id = rank[inv]
    first_sorted = first[order]
    edge_verts = np.stack([ga[first_sorted], gb[first_sorted]], axis=1)
    flip = ga != edge_verts[edge_id, 0]
    return edge_id.reshape(Nc, 4), edge_verts, flip.reshape(Nc, 4)


def cg_numbering(cell_verts, p, n_verts=None):
    """Conforming Q_p DOF ids (Nc, (p+1)^2) in Gridap local order; returns (ids, ndofs)."""
    Nc = cell_verts.shape[0]
    nv = int(cell_verts.max()) + 1 if n_verts is None else n_verts
    mi, owner = lagrange_node_multi_indices(2, p)
    m = len(mi)
    ids = np.empty((Nc, m), dtype=np.int64)
    if p == 1:
        return cell_verts.copy(), nv
    cell_edges, edge_verts, flip = quad_edges(cell_verts)
    ne = edge_verts.shape[0]
    k = 0
    for v in range(4):
        ids[:, k] = cell_verts[:, v]
        k += 1
    for e in range(4):
        for t in range(p - 1):
            tt = np.where(flip[:, e], p - 2 - t, t)
            ids[:, k] = nv + cell_edges[:, e] * (p - 1) + tt
            k += 1
    ni = (p - 1) ** 2
    for t in range(ni):
        ids[:, k] = nv + ne * (p - 1) + np.arange(Nc) * ni + t
        k += 1
    return ids, nv + ne * (p - 1) + Nc * ni


class CubedSphereMesh2D:
    """AtlasDiscreteModel(CubedSphereMesh(radius), l) with n = 2^l (or any n) cells per panel edge."""

    def __init__(self, n, radius=1.0):
        self.n, self.radius, self.D = n, radius, 2
        self.num_cells = 6 * n * n
        # fine nodes = CG-Q_n DOFs of the coarse cube
        coarse_ids, self.num_nodes = cg_numbering(COARSE_CELLS, n, n_verts=8)
        mi, _ = lagrange_node_multi_indices(2, n)
        lut = np.empty((n + 1, n + 1), dtype=np.int64)
        for k, (ix, iy) in enumerate(mi):
            lut[ix, iy] = k
        j, i = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")  # cell (i,j), i fastest
        i, j = i.reshape(-1), j.reshape(-1)
        cv = np.empty((6, n * n, 4), dtype=np.int64)
        for p in range(6):
            loc = np.stack([lut[i, j], lut[i + 1, j], lut[i, j + 1], lut[i + 1, j + 1]], axis=1)
            cv[p] = coarse_ids[p][loc]
        self.cell_nodes = cv.reshape(-1, 4)
        self.cell_to_chart = np.repeat(np.arange(6), n * n)  # 0-based panel
        self.cell_ij = np.stack([np.tile(i, 6), np.tile(j, 6)], axis=1)
        # chart corners = bilinear panel map (corners +-pi/4) applied to the reference grid k/n
        # (AtlasGrids.jl:277-280: Psi_k = linear_combination(corners, shapefuns) evaluated on ref_coords)
        H = CUBE_HALF_EDGE
        pc = np.array([[-H, -H], [H, -H], [-H, H], [H, H]])

        def psi(xi, eta):
            N = np.stack([(1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta], axis=1)
            return N @ pc

        ti, tj = np.tile(i, 6), np.tile(j, 6)
        c = np.empty((self.num_cells, 4, 2))
        c[:, 0] = psi(ti / n, tj / n)
        c[:, 1] = psi((ti + 1) / n, tj / n)
        c[:, 2] = psi(ti / n, (tj + 1) / n)
        c[:, 3] = psi((ti + 1) / n, (tj + 1) / n)
        self.cell_chart_coords = c
        self.cell_edges, self.edge_nodes, self.cell_edge_flip = quad_edges(self.cell_nodes)
        self.num_edges = self.edge_nodes.shape[0]

    def facet_cells(self):
        """edge -> (master cell, slave cell): cells around each edge in increasing cell id."""
        ne = self.num_edges
        fc = np.full((ne, 2), -1, dtype=np.int64)
        for c in range(self.num_cells):
            for e in self.cell_edges[c]:
                if fc[e, 0] < 0:
                    fc[e, 0] = c
                else:
                    fc[e, 1] = c
        return fc


def linear_partition(num_cells, nparts):
    """cell_to_part[i] = div((i-1)*nparts, ncells)+1 (1-based i) -- AtlasDistributedDiscreteModels.jl:31."""
    i = np.arange(num_cells, dtype=np.int64)
    return (i * nparts) // num_cells


def ghost_cells(cell_nodes, cell_to_part, part):
    """Cells not owned by `part` that share a vertex with an owned cell (:34-49), ascending ids."""
    owned = cell_to_part == part
    nn = int(cell_nodes.max()) + 1
    touched = np.zeros(nn, dtype=bool)
    touched[cell_nodes[owned].reshape(-1)] = True
    near = touched[cell_nodes].any(axis=1)
    return np.nonzero(near & ~owned)[0]
