"""GPU parity on the 3-D cubed-sphere shell (BASELINE config 5): hexahedral mesh and Q1/Q2 numbering bit-exact against the
oracle, Laplace-Beltrami element matrices / assembled matrix / residual (Kronecker-factorised on the GPU, brute-force
nq^3-point quadrature in the oracle) to 1e-12, and the shell volume of test/Distributed/DistributedSurfaceAreaTests.jl:120-134."""
import numpy as np
import pytest

from __graft_entry__ import load_package
from oracle import forms, shell

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gg():
    return load_package()


def _models(gg, n, L, r, t):
    return (gg.AtlasDiscreteModel(gg.CubedSphereWithThicknessMesh(r, t), n=n, n_layers=L),
            shell.CubedSphereShellMesh3D(n, r, t, n_layers=L))


@pytest.mark.parametrize("n,L", [(1, 1), (2, 2), (3, 2)])
def test_shell_mesh_and_numbering_bit_exact(gg, n, L):
    model, mesh = _models(gg, n, L, 1.0, 0.19)
    assert (model.num_cells, model.num_nodes) == (mesh.num_cells, mesh.num_nodes)
    assert (model.get_cell_node_ids() - 1 == mesh.cell_nodes).all()
    assert np.abs(model.get_cell_coordinates() - mesh.cell_chart_coords).max() < 1e-15
    for p in (1, 2):
        V = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="H1")
        Vo = shell.HexCGSpace(mesh, p)
        assert V.num_free_dofs == Vo.num_free and (V.get_cell_dof_ids() - 1 == Vo.cell_dofs).all()
    Vz = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, 2), conformity="H1", constraint="zeromean")
    Vzo = shell.HexCGSpace(mesh, 2, zeromean=True)
    ids = Vz.get_cell_dof_ids()
    assert Vz.num_free_dofs == Vzo.num_free and (np.where(ids < 0, -1, ids - 1) == Vzo.cell_dofs).all()


@pytest.mark.parametrize("n,L,p,degree", [(2, 2, 2, 18), (2, 3, 2, 8), (3, 2, 1, 12), (2, 1, 1, 6)])
def test_shell_laplace_beltrami_matches_oracle(gg, n, L, p, degree):
    r, t = 1.3, 0.19
    model, mesh = _models(gg, n, L, r, t)
    dΩ = gg.Measure(gg.Triangulation(model), degree)
    V = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="H1")
    Vo = shell.HexCGSpace(mesh, p)
    Keo = shell.lb_element_matrices(Vo, shell.ShellQuadrature(mesh, degree))
    Ke = gg.element_matrices(gg.LaplaceBeltrami(dΩ), V, V)
    assert np.abs(Ke - Keo).max() < 1e-12 * np.abs(Keo).max()
    assem = gg.SparseMatrixAssembler(V, V)
    u = np.random.default_rng(20261019).standard_normal(V.num_free_dofs)
    # on a complete shell the CSR values come from the tensor-structured numeric phase (assembled surface x radial factors);
    # the element matrices above and the sub-mesh test below cover the cell-wise path
    A, b = gg.assemble_matrix_and_vector(gg.LaplaceBeltrami(dΩ, u=u), assem)
    A, b = A.to_scipy(), b.get()
    Ao = forms.assemble_matrix(Vo, Vo, Keo)
    assert (A.indptr == Ao.indptr).all() and (A.indices == Ao.indices).all()
    assert np.abs(A.data - Ao.data).max() < 1e-12 * np.abs(Ao.data).max()
    assert np.abs(b - Ao @ u).max() < 1e-12 * np.abs(Ao @ u).max()
    # residual alone (matrix-free action) and the null space: constants
    b2 = gg.assemble_vector(gg.LaplaceBeltrami(dΩ, u=u), V).get()
    assert np.abs(b2 - b).max() < 1e-13 * np.abs(b).max()
    assert np.abs(A @ np.ones(V.num_free_dofs)).max() < 1e-11 * np.abs(A.data).max()


def test_shell_volume(gg):
    # int sqrtg = 4/3 pi ((r + t)^3 - r^3), rel 1e-2 in the reference (coarse meshes, low degrees)
    r, t = 2.0, 0.25
    exact = 4.0 / 3.0 * np.pi * ((r + t) ** 3 - r ** 3)
    for n, degree, tol in ((2, 2, 1e-2), (4, 4, 1e-4), (8, 8, 1e-9)):
        model = gg.AtlasDiscreteModel(gg.CubedSphereWithThicknessMesh(r, t), n=n)
        vol = gg.Measure(gg.Triangulation(model), degree).integrate()
        assert abs(vol - exact) < tol * exact, (n, degree, vol, exact)
    mesh = shell.CubedSphereShellMesh3D(2, r, t)
    model = gg.AtlasDiscreteModel(gg.CubedSphereWithThicknessMesh(r, t), n=2)
    assert abs(gg.Measure(gg.Triangulation(model), 4).integrate() - shell.shell_volume(mesh, 4)) < 1e-13 * exact


def test_unsupported_requests_on_the_shell_fail_loudly(gg):
    model = gg.AtlasDiscreteModel(gg.CubedSphereWithThicknessMesh(1.0, 0.1), n=2)
    V = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, 1), conformity="H1")
    dΩ = gg.Measure(gg.Triangulation(model), 4)
    with pytest.raises(gg.GGError):
        gg.assemble_matrix(gg.ScalarMass(dΩ, qdata=np.ones(model.num_cells * dΩ.num_points)), V, V)
    with pytest.raises(gg.GGError):
        gg.TestFESpace(model, gg.ReferenceFE(gg.raviart_thomas, float, 2), conformity="HDiv")    # hexahedral RT0 / RT1 only
    with pytest.raises(gg.GGError):
        gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, 3), conformity="H1")


def test_shell_slices_assemble_the_owned_rows(gg):
    """Sharding as in bench.py: a rank assembles its slice of the linear partition plus the vertex-adjacent ghost layer; the
    element matrices of the sub-mesh are those of the same cells of the full mesh (global node ids are kept)."""
    n, L, p, degree = 3, 2, 2, 8
    model, mesh = _models(gg, n, L, 1.0, 0.19)
    a, b = gg.partition_range(model.num_cells, 3, 1)
    ghosts = model.ghost_cells(a, b)
    assert len(ghosts) > 0 and not ((ghosts >= a) & (ghosts < b)).any()
    cells = np.sort(np.concatenate([np.arange(a, b), ghosts]))
    sub = model.restrict(cells)
    assert sub.num_cells == len(cells) and (sub.get_cell_node_ids() == model.get_cell_node_ids()[cells]).all()
    V, Vs = (gg.TestFESpace(m, gg.ReferenceFE(gg.lagrangian, float, p), conformity="H1") for m in (model, sub))
    Ke = gg.element_matrices(gg.LaplaceBeltrami(gg.Measure(gg.Triangulation(model), degree)), V, V)
    Ks = gg.element_matrices(gg.LaplaceBeltrami(gg.Measure(gg.Triangulation(sub), degree)), Vs, Vs)
    assert np.abs(Ks - Ke[cells]).max() < 1e-15 * np.abs(Ke).max()
    A = gg.assemble_matrix(gg.LaplaceBeltrami(gg.Measure(gg.Triangulation(sub), degree)), Vs, Vs).to_scipy()
    assert A.nnz > 0 and np.isfinite(A.data).all()
    # cell-wise CSR fill of the sub-mesh against scipy's COO assembly of the same element matrices
    import scipy.sparse as sp
    ids = Vs.get_cell_dof_ids() - 1
    m = ids.shape[1]
    Ao = sp.coo_matrix((Ks.reshape(-1), (np.repeat(ids, m, axis=1).reshape(-1), np.tile(ids, (1, m)).reshape(-1))),
                       shape=A.shape).tocsr()
    Ao.sort_indices()
    assert (A.indptr == Ao.indptr).all() and (A.indices == Ao.indices).all()
    assert np.abs(A.data - Ao.data).max() < 1e-13 * np.abs(Ao.data).max()


def test_p8est_ordering_and_sfc_partition(gg):
    """BASELINE configs[4]: the distributed shell is a uniformly refined p8est forest -- Morton order inside each of the 6 trees
    (src/Distributed/AtlasOctreeDistributedDiscreteModels.jl:164-208), partitioned into equal-count chunks of that order
    (p4est's partition, :286-294) with a vertex-adjacent ghost layer.  The Morton-ordered mesh is the lexicographic mesh with
    its cells permuted; a rank's slice assembles the same element matrices; the assembled operator is the same up to the
    permutation of the DOFs that the different first-encounter numbering induces (checked through its action on a nodal field)."""
    n, r, t, p, degree = 4, 1.0, 0.19, 2, 8
    m0 = gg.AtlasDiscreteModel(gg.CubedSphereWithThicknessMesh(r, t), n=n)
    m2 = gg.AtlasDiscreteModel(gg.CubedSphereWithThicknessMesh(r, t), n=n, ordering=2)
    o2 = shell.CubedSphereShellMesh3D(n, r, t, ordering=2)
    assert (m2.num_cells, m2.num_nodes) == (o2.num_cells, o2.num_nodes) == (m0.num_cells, m0.num_nodes)
    assert (m2.get_cell_node_ids() - 1 == o2.cell_nodes).all()
    assert np.abs(m2.get_cell_coordinates() - o2.cell_chart_coords).max() < 1e-15
    # Morton position of the lexicographic cell (kg, ia, jb) of a panel: bits interleaved, gamma lowest
    kg, ia, jb = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    lex = (kg + n * (ia + n * jb)).reshape(-1)
    mort = np.zeros_like(lex)
    for b in range(2):
        mort |= (((kg.reshape(-1) >> b) & 1) << (3 * b)) | (((ia.reshape(-1) >> b) & 1) << (3 * b + 1)) | (((jb.reshape(-1) >> b) & 1) << (3 * b + 2))
    perm = np.empty(6 * n ** 3, dtype=np.int64)                 # lexicographic cell -> its position in the p8est order
    for pn in range(6):
        perm[pn * n ** 3 + lex] = pn * n ** 3 + mort
    assert np.abs(m2.get_cell_coordinates()[perm] - m0.get_cell_coordinates()).max() == 0
    assert (m2.get_cell_to_chart()[perm] == m0.get_cell_to_chart()).all()
    # the 8 children of a parent octant are consecutive, x (gamma) fastest
    cc = m2.get_cell_coordinates().reshape(-1, 8, 8, 3)[:, :, 0, :]
    h = cc[:, 1, 0] - cc[:, 0, 0]
    assert (h > 0).all() and np.abs(cc[:, 2, 1] - cc[:, 0, 1] - (cc[:, 2, 1] - cc[:, 0, 1])[0]).max() < 1e-14
    assert np.abs(cc[:, 1, 1:] - cc[:, 0, 1:]).max() == 0 and np.abs(cc[:, 4, 2] - cc[:, 0, 2]).min() > 0
    with pytest.raises(gg.GGError):
        gg.AtlasDiscreteModel(gg.CubedSphereWithThicknessMesh(r, t), n=6, ordering=2)
    # SFC chunk of rank 2 of 8 + ghost layer: same element matrices as the same cells of the complete forest
    V2 = gg.TestFESpace(m2, gg.ReferenceFE(gg.lagrangian, float, p), conformity="H1")
    dO2 = gg.Measure(gg.Triangulation(m2), degree)
    Ke = gg.element_matrices(gg.LaplaceBeltrami(dO2), V2, V2)
    a, b = gg.partition_range(m2.num_cells, 8, 2)
    assert b - a == m2.num_cells // 8
    cells = np.sort(np.concatenate([np.arange(a, b), m2.ghost_cells(a, b)]))
    sub = m2.restrict(cells)
    Vs = gg.TestFESpace(sub, gg.ReferenceFE(gg.lagrangian, float, p), conformity="H1")
    Ks = gg.element_matrices(gg.LaplaceBeltrami(gg.Measure(gg.Triangulation(sub), degree)), Vs, Vs)
    assert np.abs(Ks - Ke[cells]).max() < 1e-15 * np.abs(Ke).max()
    # same operator: u^T A u for the interpolant of an ambient function does not depend on the cell order
    V0 = gg.TestFESpace(m0, gg.ReferenceFE(gg.lagrangian, float, p), conformity="H1")
    f = lambda X: X[..., 0] * X[..., 1] + X[..., 2] ** 2   # noqa: E731
    en = []
    for m_, V_ in ((m0, V0), (m2, V2)):
        dO = gg.Measure(gg.Triangulation(m_), degree)
        u = gg.interpolate(f, V_)
        A = gg.assemble_matrix(gg.LaplaceBeltrami(dO), V_, V_).to_scipy()
        uh = u.get()
        en.append(float(uh @ (A @ uh)))
    assert abs(en[0] - en[1]) < 1e-12 * abs(en[0])


@pytest.mark.parametrize("form", ["helmholtz", "mass"])
@pytest.mark.parametrize("n,L,p,degree", [(2, 2, 2, 8), (2, 3, 1, 6)])
def test_shell_helmholtz_and_mass_match_oracle(gg, form, n, L, p, degree):
    """test/Laplacian/Helmholtz.jl:43 on the shell: the mass term keeps the Kronecker structure (sqrtg = t s^2 sqrtg_surf)."""
    model, mesh = _models(gg, n, L, 1.3, 0.25)
    V = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="H1")
    dΩ = gg.Measure(gg.Triangulation(model), degree)
    Vo = shell.HexCGSpace(mesh, p)
    q = shell.ShellQuadrature(mesh, degree)
    Ko = shell.helmholtz_element_matrices(Vo, q) if form == "helmholtz" else shell.mass_element_matrices(Vo, q)
    F = gg.Helmholtz if form == "helmholtz" else gg.ScalarMass
    Ke = gg.element_matrices(F(dΩ), V, V)
    assert np.abs(Ke - Ko).max() <= 1e-12 * np.abs(Ko).max()
    # assembled (tensor-structured fill) and action
    import scipy.sparse as sp
    rows = np.repeat(Vo.cell_dofs, Vo.ndofs_loc, axis=1).reshape(-1)
    cols = np.tile(Vo.cell_dofs, (1, Vo.ndofs_loc)).reshape(-1)
    Ao = sp.coo_matrix((Ko.reshape(-1), (rows, cols)), shape=(Vo.num_free,) * 2).tocsr()
    A = gg.assemble_matrix(F(dΩ), V, V).to_scipy()
    assert np.abs((A - Ao).data).max() <= 1e-12 * np.abs(Ao.data).max()
    u = np.random.default_rng(3).standard_normal(V.num_free_dofs)
    r = gg.assemble_vector(F(dΩ, u=u), V).get()
    assert np.abs(r - Ao @ u).max() <= 1e-12 * np.abs(Ao @ u).max()


def _fX(X):
    return X[..., 0] * X[..., 1] * X[..., 2]


def _hfX(X):
    x, y, z = X[..., 0], X[..., 1], X[..., 2]
    H = np.zeros(X.shape + (3,))
    H[..., 0, 1] = H[..., 1, 0] = z
    H[..., 0, 2] = H[..., 2, 0] = y
    H[..., 1, 2] = H[..., 2, 1] = x
    return H


def test_shell_driver_pieces_match_oracle(gg):
    """points, source vector, boundary term, interpolation, error norm and zero-mean shift of the 3-D Laplace-Beltrami driver."""
    n, L, r, t, p, degree = 2, 3, 1.3, 0.25, 2, 6
    model, mesh = _models(gg, n, L, r, t)
    V = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="H1", constraint="zeromean")
    Vo = shell.HexCGSpace(mesh, p, zeromean=True)
    dΩ = gg.Measure(gg.Triangulation(model), degree)
    q = shell.ShellQuadrature(mesh, degree)
    chart, xyz = dΩ.quadrature_points()
    assert np.abs(chart - q.x).max() < 1e-14 and np.abs(xyz - q.ambient()).max() < 1e-13
    fq = np.cos(xyz[..., 0]) + xyz[..., 1] * xyz[..., 2]
    b = gg.assemble_vector(gg.Source(dΩ, fq), V).get()
    fe = shell.source_element_vectors(Vo, q, fq)
    bo = np.zeros(Vo.num_free)
    np.add.at(bo, Vo.cell_dofs[Vo.cell_dofs >= 0], fe[Vo.cell_dofs >= 0])
    assert np.abs(b - bo).max() <= 1e-12 * np.abs(bo).max()
    Uo = shell.interpolate(Vo, _fX)
    U = gg.interpolate(_fX, V)
    assert np.abs(U.get() - Uo[:-1]).max() <= 1e-13 * np.abs(Uo).max()
    # the fixed (last) DOF of the interpolant enters the boundary term with its own value
    g = gg.assemble_vector(gg.ShellBoundaryFlux(dΩ, u=U, dirichlet=float(Uo[-1])), V).get()
    fe = shell.boundary_flux_element_vectors(Vo, degree, Uo)
    go = np.zeros(Vo.num_free)
    np.add.at(go, Vo.cell_dofs[Vo.cell_dofs >= 0], fe[Vo.cell_dofs >= 0])
    assert np.abs(g - go).max() <= 1e-12 * np.abs(go).max()
    rng = np.random.default_rng(2)
    u = rng.standard_normal(V.num_free_dofs)
    uq = shell.eval_scalar(Vo, q, np.append(u, 0.3))
    e_ref = np.sqrt((((uq - fq) ** 2) * q.dV * q.sqrtg).sum())
    assert abs(gg.l2_error(u, V, dΩ, fq=fq, dirichlet=0.3) - e_ref) <= 1e-12 * e_ref
    mean_ref = (uq * q.dV).sum() / q.dV.sum()
    assert abs(gg.chart_mean(u, V, dirichlet=0.3) - mean_ref) <= 1e-12 * abs(mean_ref)


def test_shell_laplace_beltrami_driver_converges(gg):
    """3-D run of test/Laplacian/LaplaceBeltrami.jl (test/Laplacian/mpi/LaplaceBeltramiTests.jl:23-28): slope >= p - 0.1, and the
    errors of the device driver equal the oracle's."""
    errs = []
    for n in (2, 4):
        model = gg.AtlasDiscreteModel(gg.CubedSphereWithThicknessMesh(1.0, 0.19), n=n)
        e, uh, V, ns = gg.laplace_beltrami_solver(model, 2, _fX, None, _hfX, degree=8)
        assert ns.rel_res < 1e-12
        errs.append(e)
    e_ref = shell.solve_lb(shell.CubedSphereShellMesh3D(2, 1.0, 0.19), 2, _fX, lambda X: np.zeros(X.shape[:-1]), degree=8)
    assert abs(errs[0] - e_ref) <= 1e-5 * e_ref        # error rule of degree 16 here, 16 in the oracle; iterative solve
    assert np.log(errs[0] / errs[1]) / np.log(2.0) >= 2 - 0.1
