"""GPU parity tests: the CUDA path (through the C ABI) against the numpy oracle on the same inputs.

Tolerances: bit-exact for the CSR pattern and DOF numbering; 1e-12 relative (to the largest entry of the
element matrix / vector, north_star) for floating point."""
import numpy as np
import pytest

from __graft_entry__ import load_package
from oracle import forms, geometry as geo
from oracle.mesh import CubedSphereMesh2D, ghost_cells, linear_partition
from oracle.spaces import CGSpace, DGSpace, RTSpace

pytestmark = pytest.mark.gpu
RTOL = 1e-12


@pytest.fixture(scope="module")
def gg():
    return load_package()


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def gpu_spaces(gg, model, kind, p, **kw):
    if kind == "CG":
        return gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="H1", **kw)
    if kind == "DG":
        return gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="L2")
    return gg.TestFESpace(model, gg.ReferenceFE(gg.raviart_thomas, float, p), conformity="HDiv")


def test_geometry_fields(gg):
    # MetricFieldsTests.jl:93-122 on the device
    H = np.pi / 4
    pts = np.array([[0.0, 0.0], [H / 2, H / 3], [-H / 3, H / 2], [-0.6 * H, -0.4 * H], [H, -H], [-H, H]])
    r = 1.2
    for panel in range(1, 7):
        xyz, jac, g, gi, sq = gg.eval_geometry(panel, r, pts)
        assert np.abs(xyz - geo.csphere_eval(panel, r, pts)).max() < 1e-14
        assert np.abs(jac - geo.csphere_jac(panel, r, pts)).max() < 1e-13
        G = geo.csphere_metric(r, pts)
        Gi = geo.csphere_inv_metric(r, pts)
        assert np.abs(g - np.stack([G[:, 0, 0], G[:, 0, 1], G[:, 1, 1]], 1)).max() < 1e-13
        assert np.abs(gi - np.stack([Gi[:, 0, 0], Gi[:, 0, 1], Gi[:, 1, 1]], 1)).max() < 1e-13
        JJt = jac @ np.swapaxes(jac, 1, 2)
        assert np.abs(JJt[:, 0, 0] - g[:, 0]).max() < 1e-12 and np.abs(JJt[:, 0, 1] - g[:, 1]).max() < 1e-12
        assert np.abs(sq - geo.csphere_sqrtg(r, pts)).max() < 1e-13


@pytest.mark.parametrize("n", [1, 2, 5, 8])
def test_mesh_and_numbering_bit_exact(gg, n):
    mesh = CubedSphereMesh2D(n, 1.0)
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=n)
    assert (model.num_cells, model.num_nodes, model.num_edges) == (mesh.num_cells, mesh.num_nodes, mesh.num_edges)
    assert (model.get_cell_node_ids() - 1 == mesh.cell_nodes).all()
    assert (model.get_cell_to_chart() - 1 == mesh.cell_to_chart).all()
    assert np.abs(model.get_cell_coordinates() - mesh.cell_chart_coords).max() < 1e-15
    for kind, p, O in (("CG", 1, CGSpace), ("CG", 2, CGSpace), ("CG", 3, CGSpace), ("DG", 1, DGSpace), ("DG", 2, DGSpace),
                       ("RT", 0, RTSpace), ("RT", 1, RTSpace), ("RT", 2, RTSpace)):
        S = gpu_spaces(gg, model, kind, p)
        So = O(mesh, p)
        assert S.num_free_dofs == So.num_free
        assert (S.get_cell_dof_ids() - 1 == So.cell_dofs).all(), (kind, p)
        if kind == "RT":
            assert (S.get_cell_dof_signs() == So.signs).all()
    Z = gpu_spaces(gg, model, "CG", 2, constraint="zeromean")
    Zo = CGSpace(mesh, 2, zeromean=True)
    ids = Z.get_cell_dof_ids()
    assert Z.num_free_dofs == Zo.num_free and ((ids > 0) == (Zo.cell_dofs >= 0)).all()
    assert (np.where(ids > 0, ids - 1, -1) == Zo.cell_dofs).all()


@pytest.mark.parametrize("p,degree", [(1, 6), (1, 12), (2, 8), (2, 12), (2, 18), (1, 4), (2, 5), (3, 8)])
@pytest.mark.parametrize("extrinsic", [False, True])
def test_lb_element_matrices(gg, p, degree, extrinsic):
    n, r = 4, 1.3
    mesh = CubedSphereMesh2D(n, r)
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(r), n=n)
    V = gpu_spaces(gg, model, "CG", p)
    dΩ = gg.Measure(gg.Triangulation(model), degree)
    Ke = gg.element_matrices(gg.LaplaceBeltrami(dΩ, extrinsic=extrinsic), V, V)
    Vo = CGSpace(mesh, p)
    q = forms.CellQuadrature(mesh, degree)
    Ko = forms.lb_extrinsic_element_matrices(Vo, q) if extrinsic else forms.lb_element_matrices(Vo, q)
    scale = np.abs(Ko).max(axis=(1, 2), keepdims=True)
    assert (np.abs(Ke - Ko) / scale).max() < RTOL


@pytest.mark.parametrize("form", ["helmholtz", "mass", "mass_coef", "mass_dg_cg"])
def test_scalar_forms_element_matrices(gg, form):
    n, r, degree = 3, 0.9, 12
    mesh = CubedSphereMesh2D(n, r)
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(r), n=n)
    dΩ = gg.Measure(gg.Triangulation(model), degree)
    q = forms.CellQuadrature(mesh, degree)
    V, Vo = gpu_spaces(gg, model, "CG", 2), CGSpace(mesh, 2)
    if form == "helmholtz":
        Ke = gg.element_matrices(gg.Helmholtz(dΩ), V, V)
        Ko = forms.scalar_mass_element_matrices(Vo, Vo, q) - forms.lb_element_matrices(Vo, q)
    elif form == "mass":
        Ke = gg.element_matrices(gg.ScalarMass(dΩ), V, V)
        Ko = forms.scalar_mass_element_matrices(Vo, Vo, q)
    elif form == "mass_coef":
        coef = np.random.default_rng(1).standard_normal((mesh.num_cells, q.pts.shape[0]))
        Ke = gg.element_matrices(gg.ScalarMass(dΩ, qdata=coef), V, V)
        Ko = forms.scalar_mass_element_matrices(Vo, Vo, q, coef=coef)
    else:
        W, Wo = gpu_spaces(gg, model, "DG", 1), DGSpace(mesh, 1)
        Ke = gg.element_matrices(gg.ScalarMass(dΩ), V, W)          # rows: test DG1, cols: trial CG2
        Ko = forms.scalar_mass_element_matrices(Wo, Vo, q)
    scale = np.abs(Ko).max(axis=(1, 2), keepdims=True)
    assert (np.abs(Ke - Ko) / scale).max() < RTOL


@pytest.mark.parametrize("k", [0, 1, 2])
def test_rt_forms_element_matrices(gg, k):
    n, r, degree = 3, 1.1, 5 * (k + 1)
    mesh = CubedSphereMesh2D(n, r)
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(r), n=n)
    dΩ = gg.Measure(gg.Triangulation(model), degree)
    q = forms.CellQuadrature(mesh, degree)
    V, Vo = gpu_spaces(gg, model, "RT", k), RTSpace(mesh, k)
    Q, Qo = gpu_spaces(gg, model, "DG", k), DGSpace(mesh, k)
    Me, Mo = gg.element_matrices(gg.RTMass(dΩ), V, V), forms.rt_mass_element_matrices(Vo, q)
    assert (np.abs(Me - Mo) / np.abs(Mo).max(axis=(1, 2), keepdims=True)).max() < RTOL
    De, Do = gg.element_matrices(gg.Div(dΩ), V, Q), forms.div_element_matrices(Qo, Vo, q)
    assert (np.abs(De - Do) / np.abs(Do).max(axis=(1, 2), keepdims=True)).max() < RTOL


@pytest.mark.parametrize("n,p,degree,zeromean", [(1, 1, 4), (2, 1, 12), (4, 2, 18), (8, 2, 12), (5, 2, 8)][:0] or
                         [(1, 1, 4, False), (2, 1, 12, True), (4, 2, 18, True), (8, 2, 12, False), (5, 2, 8, False)])
def test_lb_assembly_pattern_and_values(gg, n, p, degree, zeromean):
    mesh = CubedSphereMesh2D(n, 1.0)
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=n)
    V = gpu_spaces(gg, model, "CG", p, constraint="zeromean" if zeromean else None)
    dΩ = gg.Measure(gg.Triangulation(model), degree)
    A = gg.assemble_matrix(gg.LaplaceBeltrami(dΩ), V, V).to_scipy()
    Vo = CGSpace(mesh, p, zeromean=zeromean)
    q = forms.CellQuadrature(mesh, degree)
    Ao = forms.assemble_matrix(Vo, Vo, forms.lb_element_matrices(Vo, q))
    assert A.shape == Ao.shape
    assert (A.indptr == Ao.indptr).all() and (A.indices == Ao.indices).all()      # bit-exact pattern
    assert rel(A.data, Ao.data) < RTOL


def test_survey_sizes_cfg1(gg):
    # cfg1 (32x32, CG-Q1) matrix is 6146^2 with nnz 55 298
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), 5)
    V = gpu_spaces(gg, model, "CG", 1)
    A = gg.SparseMatrixAssembler(V, V).matrix
    assert A.shape == (6146, 6146) and A.nnz == 55298


@pytest.mark.parametrize("p,degree", [(1, 12), (2, 18), (2, 12), (3, 8)])
def test_residual_and_fused_assembly(gg, p, degree):
    n = 4
    mesh = CubedSphereMesh2D(n, 1.0)
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=n)
    V = gpu_spaces(gg, model, "CG", p, constraint="zeromean")
    dΩ = gg.Measure(gg.Triangulation(model), degree)
    Vo = CGSpace(mesh, p, zeromean=True)
    q = forms.CellQuadrature(mesh, degree)
    Ke = forms.lb_element_matrices(Vo, q)
    Ao = forms.assemble_matrix(Vo, Vo, Ke)
    u = np.random.default_rng(20261019).standard_normal(Vo.num_free)
    dirichlet = 0.37
    ue = forms.gather(Vo, u, dirichlet)
    ro = forms.assemble_vector(Vo, np.einsum("cab,cb->ca", Ke, ue))
    # matrix-free action
    r = gg.assemble_vector(gg.LaplaceBeltrami(dΩ, u=u, dirichlet=dirichlet), V).get()
    assert rel(r, ro) < RTOL
    if p <= 2:
        assem = gg.SparseMatrixAssembler(V, V)
        A, b = gg.assemble_matrix_and_vector(gg.LaplaceBeltrami(dΩ, u=u, dirichlet=dirichlet), assem)
        assert rel(A.to_scipy().data, Ao.data) < RTOL
        assert rel(b.get(), ro) < RTOL
        # linearity of the action and add semantics
        b2 = gg.assemble_vector(gg.LaplaceBeltrami(dΩ, u=2 * u, dirichlet=2 * dirichlet), V, out=b, add=True).get()
        assert rel(b2, 3 * ro) < 1e-11


def test_source_vector_and_area(gg):
    # l(v) = ∫ f v sqrtg with the manufactured rhs 12 xyz / r^2 (= -Δs(xyz)); area = 4 pi r^2 (SurfaceArea.jl:25-38)
    n, r, degree = 4, 2.0, 12
    mesh = CubedSphereMesh2D(n, r)
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(r), n=n)
    dΩ = gg.Measure(gg.Triangulation(model), degree)
    V = gpu_spaces(gg, model, "CG", 2, constraint="zeromean")
    chart, xyz = dΩ.quadrature_points()
    q = forms.CellQuadrature(mesh, degree)
    assert np.abs(chart - q.x).max() < 1e-15 and np.abs(xyz - q.ambient()).max() < 1e-14
    f = 12 * xyz[..., 0] * xyz[..., 1] * xyz[..., 2] / r**2
    b = gg.assemble_vector(gg.Source(dΩ, f), V).get()
    Vo = CGSpace(mesh, 2, zeromean=True)
    bo = forms.assemble_vector(Vo, forms.source_element_vectors(Vo, q, f))
    assert rel(b, bo) < RTOL
    for deg in (2, 4, 6, 8):
        area = gg.Measure(gg.Triangulation(model), deg).integrate()
        assert abs(area - 4 * np.pi * r**2) / (4 * np.pi * r**2) < 1e-2
    assert abs(dΩ.integrate() - (q.dV * q.sqrtg).sum()) < 1e-12 * 4 * np.pi * r**2
    assert abs(dΩ.integrate(f)) < 1e-12


def test_connectivity_as_input_and_submesh(gg):
    # the ABI takes the caller's numbering: permute the DOF ids, and restrict to a rank's owned+ghost cells
    n, p, degree = 4, 2, 8
    mesh = CubedSphereMesh2D(n, 1.0)
    Vo = CGSpace(mesh, p)
    perm = np.random.default_rng(3).permutation(Vo.num_free)
    ids = perm[Vo.cell_dofs] + 1
    model = gg.AtlasDiscreteModel.from_arrays(gg.CubedSphereMesh(1.0), mesh.cell_to_chart + 1, mesh.cell_chart_coords,
                                              mesh.cell_nodes + 1, mesh.num_nodes)
    V = gg.FESpace.from_arrays(model, 0, p, Vo.num_free, ids)
    dΩ = gg.Measure(gg.Triangulation(model), degree)
    A = gg.assemble_matrix(gg.LaplaceBeltrami(dΩ), V, V).to_scipy()
    q = forms.CellQuadrature(mesh, degree)
    Ao = forms.assemble_matrix(Vo, Vo, forms.lb_element_matrices(Vo, q)).tocsr()
    P = np.empty_like(perm)
    P[perm] = np.arange(len(perm))
    Aop = Ao[P][:, P].tocsr()
    Aop.sort_indices()
    assert (A.indptr == Aop.indptr).all() and (A.indices == Aop.indices).all()
    assert rel(A.data, Aop.data) < RTOL
    # sub-mesh of part 1 of 4: owned + ghost cells, element matrices equal the global ones
    full = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=n)
    part = linear_partition(mesh.num_cells, 4)
    a, b = gg.partition_range(mesh.num_cells, 4, 1)
    assert (np.nonzero(part == 1)[0] == np.arange(a, b)).all()
    gh = full.ghost_cells(a, b)
    assert (gh == ghost_cells(mesh.cell_nodes, part, 1)).all()
    cells = np.sort(np.concatenate([np.arange(a, b), gh]))
    sub = full.restrict(cells)
    Vs = gg.FESpace.from_arrays(sub, 0, p, Vo.num_free, Vo.cell_dofs[cells] + 1)
    Ks = gg.element_matrices(gg.LaplaceBeltrami(gg.Measure(gg.Triangulation(sub), degree)), Vs, Vs)
    Ko = forms.lb_element_matrices(Vo, q)[cells]
    assert (np.abs(Ks - Ko) / np.abs(Ko).max(axis=(1, 2), keepdims=True)).max() < RTOL


def test_error_behaviour(gg):
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=2)
    V = gpu_spaces(gg, model, "CG", 2)
    with pytest.raises(gg.GGError):
        gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, 7), conformity="H1")       # unsupported order
    with pytest.raises(gg.GGError):
        gg.assemble_vector(gg.LaplaceBeltrami(gg.Measure(gg.Triangulation(model), 8), u=np.zeros(3)), V)  # wrong length
    with pytest.raises(gg.GGError):
        gg.AtlasDiscreteModel(gg.CubedSphereMesh(-1.0), n=2)
    # non axis-aligned cell is rejected, not silently mis-assembled
    xy = model.get_cell_coordinates().copy()
    xy[0, 3, 0] += 0.01
    with pytest.raises(gg.GGError):
        gg.AtlasDiscreteModel.from_arrays(gg.CubedSphereMesh(1.0), model.get_cell_to_chart(), xy, model.get_cell_node_ids(), model.num_nodes)


@pytest.mark.parametrize("ordering", [1, 2])
def test_recursive_and_morton_cell_orderings(gg, ordering):
    """ordering 1 (successive `refine`) and 2 (p4est Morton): children of a parent stay together in Gridap order BL,BR,TL,TR
    (test/AtlasDiscreteModels/seq/RefinementOrderingTests.jl:42-48,93-100); the mesh is the one-shot mesh with its cells
    permuted, so every assembled quantity is the same up to a symmetric permutation."""
    n = 8
    m0 = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=n)
    m1 = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=n, ordering=ordering)
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    mort = np.zeros_like(i)
    for b in range(3):
        mort |= (((i >> b) & 1) << (2 * b)) | (((j >> b) & 1) << (2 * b + 1))
    perm = np.concatenate([p * n * n + mort.reshape(-1) for p in range(6)])       # lexicographic cell -> its Morton position
    assert (m1.get_cell_node_ids()[perm] == m0.get_cell_node_ids()).all()
    assert np.abs(m1.get_cell_coordinates()[perm] - m0.get_cell_coordinates()).max() == 0
    assert (m1.get_cell_to_chart()[perm] == m0.get_cell_to_chart()).all()
    # parent-major recursion: cells 4k..4k+3 are the BL, BR, TL, TR children of one parent
    xy = m1.get_cell_coordinates().reshape(-1, 4, 4, 2)
    h = xy[:, 0, 1, 0] - xy[:, 0, 0, 0]
    assert np.abs(xy[:, 1, 0] - (xy[:, 0, 0] + np.stack([h, 0 * h], 1))).max() < 1e-14
    assert np.abs(xy[:, 2, 0] - (xy[:, 0, 0] + np.stack([0 * h, h], 1))).max() < 1e-14
    assert np.abs(xy[:, 3, 0] - (xy[:, 0, 0] + np.stack([h, h], 1))).max() < 1e-14
    # same operator up to the permutation
    dΩ0, dΩ1 = gg.Measure(gg.Triangulation(m0), 6), gg.Measure(gg.Triangulation(m1), 6)
    V0 = gg.TestFESpace(m0, gg.ReferenceFE(gg.lagrangian, float, 1), conformity="H1")
    V1 = gg.TestFESpace(m1, gg.ReferenceFE(gg.lagrangian, float, 1), conformity="H1")
    A0 = gg.assemble_matrix(gg.LaplaceBeltrami(dΩ0), V0, V0).to_scipy()
    A1 = gg.assemble_matrix(gg.LaplaceBeltrami(dΩ1), V1, V1).to_scipy()
    assert abs(A0 - A1).max() < 1e-13 * np.abs(A0.data).max()                    # Q1 DOFs are the nodes: identical numbering
    with pytest.raises(gg.GGError):
        gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=6, ordering=ordering)


@pytest.mark.parametrize("n,p,degree", [(4, 2, 18), (6, 1, 12), (5, 2, 8)])
def test_geometry_class_assembly_is_bitwise_the_per_cell_assembly(gg, n, p, degree):
    """gg_set_geometry_classes: one element matrix per (column, row) class of the 6 congruent panels; values and residual are
    bit for bit those of the per-cell path, for repeated calls with different states."""
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.3), n=n)
    V = gpu_spaces(gg, model, "CG", p, constraint="zeromean")
    dΩ = gg.Measure(gg.Triangulation(model), degree)
    assem = gg.SparseMatrixAssembler(V, V)
    rng = np.random.default_rng(7)
    ctx = model.ctx
    try:
        for _ in range(2):
            u = rng.standard_normal(V.num_free_dofs)
            ctx.set_geometry_classes(False)
            A, b = gg.assemble_matrix_and_vector(gg.LaplaceBeltrami(dΩ, u=u, dirichlet=0.4), assem)
            v0, b0 = A.values().copy(), b.get().copy()
            ctx.set_geometry_classes(True)
            A, b = gg.assemble_matrix_and_vector(gg.LaplaceBeltrami(dΩ, u=u, dirichlet=0.4), assem)
            assert (A.values() == v0).all()
            assert np.abs(b.get() - b0).max() <= 1e-15 * np.abs(b0).max()
            A2 = gg.assemble_matrix(gg.LaplaceBeltrami(dΩ), assem)
            assert (A2.values() == v0).all()
    finally:
        ctx.set_geometry_classes(False)


def test_tile_path_on_small_meshes():
    """The fused tile assembly is chosen by size (several waves of 128-cell tiles); GG_LB_TILES=1 forces it, so that partial tiles,
    single-tile panels, constrained spaces, sub-meshes and the Morton order go through it too (tests/tile_worker.py)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tests", "tile_worker.py")], capture_output=True, text=True, timeout=600,
                         cwd=root, env=dict(os.environ, GG_LB_TILES="1"))
    assert out.returncode == 0 and "TILE_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
