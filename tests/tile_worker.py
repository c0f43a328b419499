"""Worker of tests/test_gpu_assembly.py::test_tile_path_on_small_meshes: run with GG_LB_TILES=1 so that the fused tile assembly
(gg_tile.cuh / gg_lb_tile.cuh) -- chosen by default only for slices of several waves of tiles -- is exercised on small meshes,
partial tiles, zeromean spaces, sub-meshes, the Morton cell order and the extrinsic form, against the numpy oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from oracle import forms  # noqa: E402
from oracle.mesh import CubedSphereMesh2D, ghost_cells, linear_partition  # noqa: E402
from oracle.spaces import CGSpace  # noqa: E402


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def main():
    assert os.environ.get("GG_LB_TILES") == "1"
    gg = load_package()
    rng = np.random.default_rng(3)
    for n, p, degree, zeromean in ((1, 1, 4, False), (2, 2, 8, True), (5, 2, 18, False), (8, 1, 12, True), (20, 2, 12, False), (33, 2, 8, False)):
        mesh = CubedSphereMesh2D(n, 1.0)
        model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=n)
        V = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="H1", constraint="zeromean" if zeromean else None)
        dΩ = gg.Measure(gg.Triangulation(model), degree)
        Vo = CGSpace(mesh, p, zeromean=zeromean)
        Ao = forms.assemble_matrix(Vo, Vo, forms.lb_element_matrices(Vo, forms.CellQuadrature(mesh, degree)))
        u = rng.standard_normal(V.num_free_dofs)
        assem = gg.SparseMatrixAssembler(V, V)
        gg.assemble_matrix_and_vector(gg.LaplaceBeltrami(dΩ, u=u), assem)       # first call: one-off tables and the tile plan
        l0 = model.ctx.launches
        A, b = gg.assemble_matrix_and_vector(gg.LaplaceBeltrami(dΩ, u=u), assem)
        assert model.ctx.launches - l0 <= 2, "the fused tile path (one or two launches) was not taken"
        A = A.to_scipy()
        assert (A.indptr == Ao.indptr).all() and (A.indices == Ao.indices).all()
        assert rel(A.data, Ao.data) < 1e-12 and rel(b.get(), Ao @ u) < 1e-12, (n, p)
        # matrix only, scaled and added: A <- A + 2 a(.,.)
        A2 = gg.assemble_matrix(gg.LaplaceBeltrami(dΩ, scale=2.0), assem, add=True).to_scipy()
        assert rel(A2.data, 3.0 * Ao.data) < 1e-12
        # the extrinsic form gives the same matrix (AmbientLaplaceBeltramiTests.jl:28-30)
        if p == 2 and n <= 8:
            Ae = gg.assemble_matrix(gg.LaplaceBeltrami(dΩ, extrinsic=True), V, V).to_scipy()
            assert rel(Ae.data, Ao.data) < 1e-11
    # a rank's slice (owned + ghost cells, global node ids) and the Morton cell order
    n, p, degree = 8, 2, 8
    mesh = CubedSphereMesh2D(n, 1.0)
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=n)
    part = linear_partition(mesh.num_cells, 3)
    owned = np.nonzero(part == 1)[0]
    cells = np.sort(np.concatenate([owned, ghost_cells(mesh.cell_nodes, part, 1)]))
    sub = model.restrict(cells)
    Vs = gg.TestFESpace(sub, gg.ReferenceFE(gg.lagrangian, float, p), conformity="H1")
    V = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="H1")
    dΩs = gg.Measure(gg.Triangulation(sub), degree)
    As = gg.assemble_matrix(gg.LaplaceBeltrami(dΩs), Vs, Vs).to_scipy()
    Ke = gg.element_matrices(gg.LaplaceBeltrami(gg.Measure(gg.Triangulation(model), degree)), V, V)[cells]
    import scipy.sparse as sp
    ids = Vs.get_cell_dof_ids() - 1
    m = ids.shape[1]
    Ao = sp.coo_matrix((Ke.reshape(-1), (np.repeat(ids, m, axis=1).reshape(-1), np.tile(ids, (1, m)).reshape(-1))), shape=As.shape).tocsr()
    Ao.sort_indices()
    assert (As.indptr == Ao.indptr).all() and (As.indices == Ao.indices).all() and rel(As.data, Ao.data) < 1e-12
    m1 = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=n, ordering=2)
    V1 = gg.TestFESpace(m1, gg.ReferenceFE(gg.lagrangian, float, 1), conformity="H1")
    V0 = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, 1), conformity="H1")
    A0 = gg.assemble_matrix(gg.LaplaceBeltrami(gg.Measure(gg.Triangulation(model), 6)), V0, V0).to_scipy()
    A1 = gg.assemble_matrix(gg.LaplaceBeltrami(gg.Measure(gg.Triangulation(m1), 6)), V1, V1).to_scipy()
    assert abs(A0 - A1).max() < 1e-13 * np.abs(A0.data).max()
    print("TILE_OK")


if __name__ == "__main__":
    main()
