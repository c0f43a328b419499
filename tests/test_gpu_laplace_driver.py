"""The whole Laplace-Beltrami driver on the device (test/Laplacian/LaplaceBeltrami.jl:19-105): manufactured right-hand side
from the closed-form surface Laplacian, assembly, PCG solve, zero-mean shift, L2 error and the convergence-slope criterion."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def fX(X):
    return X[..., 0] * X[..., 1] * X[..., 2]


def grad_fX(X):
    x, y, z = X[..., 0], X[..., 1], X[..., 2]
    return np.stack([y * z, x * z, x * y], axis=-1)


def hess_fX(X):
    x, y, z = X[..., 0], X[..., 1], X[..., 2]
    H = np.zeros(X.shape + (3,))
    H[..., 0, 1] = H[..., 1, 0] = z
    H[..., 0, 2] = H[..., 2, 0] = y
    H[..., 1, 2] = H[..., 2, 1] = x
    return H


@pytest.fixture(scope="module")
def gg():
    from __graft_entry__ import load_package
    return load_package()


def test_surface_laplacian_matches_oracle_and_closed_form(gg):
    from oracle import forms
    from oracle.mesh import CubedSphereMesh2D
    r = 1.3
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(r), n=4)
    dΩ = gg.Measure(gg.Triangulation(model), 6)
    lap = gg.surface_laplacian(dΩ, grad_fX, hess_fX).get().reshape(model.num_cells, -1)
    mesh = CubedSphereMesh2D(4, r)
    q = forms.CellQuadrature(mesh, 6)
    ref = forms.surface_laplacian(fX, grad_fX, hess_fX, mesh, q.x, mesh.cell_to_chart)
    assert np.abs(lap - ref).max() <= 1e-12 * np.abs(ref).max()
    # xyz is a degree-3 spherical harmonic: Delta_s f = -12 f / r^2
    assert np.abs(lap + 12 * fX(q.ambient()) / r**2).max() < 1e-12


def test_l2_error_and_chart_mean_match_oracle(gg):
    from oracle import forms
    from oracle.mesh import CubedSphereMesh2D
    from oracle.spaces import CGSpace
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=4)
    V = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, 2), conformity="H1")
    dΩ = gg.Measure(gg.Triangulation(model), 10)
    rng = np.random.default_rng(5)
    u = rng.standard_normal(V.num_free_dofs)
    mesh = CubedSphereMesh2D(4)
    Vo = CGSpace(mesh, 2)
    q = forms.CellQuadrature(mesh, 10)
    f = fX(q.ambient())
    uh = forms.eval_scalar(Vo, q, u)
    e_ref = np.sqrt((((f - uh) ** 2) * q.dV * q.sqrtg).sum())
    assert abs(gg.l2_error(u, V, dΩ, fq=f) - e_ref) <= 1e-12 * e_ref
    vol_i = forms.assemble_vector(Vo, np.einsum("cq,qa->ca", q.dV, Vo.ref.tabulate(q.pts)[0]))
    assert abs(gg.chart_mean(u, V) - (vol_i @ u) / vol_i.sum()) < 1e-13


def test_laplace_beltrami_driver_convergence(gg):
    # test/Laplacian/LaplaceBeltrami.jl:101-105 + src/ConvergenceTools/Tools.jl:4-35: slope >= p over four refinements
    from test_oracle_forms import solve_lb
    from oracle.mesh import CubedSphereMesh2D
    errs, dxs = [], []
    for lvl in (1, 2, 3, 4):
        n = 2**lvl
        model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=n)
        e, uh, V, ns = gg.laplace_beltrami_solver(model, 2, fX, grad_fX, hess_fX)
        assert ns.rel_res < 1e-12
        errs.append(e)
        dxs.append(gg.dx(model))
        if lvl == 2:
            e_ref, _, _ = solve_lb(CubedSphereMesh2D(n), 2)
            assert abs(e - e_ref) <= 1e-6 * e_ref   # different error rule (degree 31 instead of 36) and an iterative solve
    slope = np.polyfit(np.log10(dxs), np.log10(errs), 1)[0]
    assert slope >= 2


def test_surface_gradient_and_divergence_match_oracle(gg):
    # the fields of test/Autodiff/seq/SurfDiffOperatorTests.jl:25-33
    from oracle import forms
    from oracle.mesh import CubedSphereMesh2D
    from test_oracle_forms import _ambient_vecX, _grad_ambient_fX, _grad_ambient_vecX
    r = 1.7
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(r), n=3)
    dΩ = gg.Measure(gg.Triangulation(model), 5)
    mesh = CubedSphereMesh2D(3, r)
    q = forms.CellQuadrature(mesh, 5)
    sg = gg.surface_gradient(dΩ, _grad_ambient_fX).get().reshape(model.num_cells, -1, 2)
    ref = forms.surface_gradient(_grad_ambient_fX, mesh, q.x, mesh.cell_to_chart)
    assert np.abs(sg - ref).max() <= 1e-12 * np.abs(ref).max()
    sd = gg.surface_divergence(dΩ, _ambient_vecX, _grad_ambient_vecX).get().reshape(model.num_cells, -1)
    ref = forms.surface_divergence(_ambient_vecX, _grad_ambient_vecX, mesh, q.x, mesh.cell_to_chart)
    assert np.abs(sd - ref).max() <= 1e-12 * np.abs(ref).max()


def test_helmholtz_driver_convergence(gg):
    """test/Laplacian/Helmholtz.jl:19-86 with the blocks assembled on the device; the indefinite system goes to the caller's LU
    (LUSolver() in the reference), the manufactured right-hand side and the error norm stay on the device."""
    import scipy.sparse.linalg as spla
    from test_oracle_forms import solve_helmholtz
    from oracle.mesh import CubedSphereMesh2D
    errs = []
    for lvl in (1, 2, 3, 4):
        model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=2**lvl)
        Ω = gg.Triangulation(model)
        dΩ, dΩe = gg.Measure(Ω, 18), gg.Measure(Ω, 31)
        V = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, 2), conformity="H1")
        A = gg.assemble_matrix(gg.Helmholtz(dΩ), V, V).to_scipy()
        _, xyz = dΩ.quadrature_points(chart=False)
        rhs = fX(xyz).reshape(-1) + gg.surface_laplacian(dΩ, grad_fX, hess_fX).get()
        b = gg.assemble_vector(gg.Source(dΩ, rhs), V).get()
        x = spla.spsolve(A.tocsc(), b)
        _, xe = dΩe.quadrature_points(chart=False)
        errs.append(gg.l2_error(x, V, dΩe, fq=fX(xe)))
        if lvl == 2:
            e_ref = solve_helmholtz(CubedSphereMesh2D(4), 2)
            assert abs(errs[-1] - e_ref) <= 1e-6 * e_ref
    slope = np.polyfit(np.log10([1 / 2.0**lvl for lvl in (1, 2, 3, 4)]), np.log10(errs), 1)[0]
    assert slope >= 2


def test_hodge_laplacian_scalar_driver_convergence(gg):
    """test/Laplacian/HodgeLaplacian_scalar.jl:23-121: mixed RT_2 x DG_2 system from device-assembled blocks (RT mass, div,
    manufactured source), caller's LU, device error norms for the pressure and for the contravariant velocity."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from test_oracle_forms import grad_sinlat, hess_sinlat, sinlat, solve_hodge_scalar
    from oracle.mesh import CubedSphereMesh2D
    p, degree = 2, 12
    eu, ep = [], []
    for lvl in (1, 2, 3):
        model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=2**lvl)
        Ω = gg.Triangulation(model)
        dΩ, dΩe = gg.Measure(Ω, degree), gg.Measure(Ω, 2 * degree)
        V = gg.TestFESpace(model, gg.ReferenceFE(gg.raviart_thomas, float, p), conformity="HDiv")
        Q = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="L2", constraint="zeromean")
        M = gg.assemble_matrix(gg.RTMass(dΩ), V, V).to_scipy()
        B = gg.assemble_matrix(gg.Div(dΩ), V, Q).to_scipy()                  # rows: test Q, columns: trial V
        bq = gg.assemble_vector(gg.Source(dΩ, gg.surface_laplacian(dΩ, grad_sinlat, hess_sinlat), scale=-1.0), Q).get()
        A = sp.bmat([[M, -B.T], [B, None]]).tocsc()
        x = spla.spsolve(A, np.concatenate([np.zeros(V.num_free_dofs), bq]))
        u, ph = x[: V.num_free_dofs], x[V.num_free_dofs:]
        mean = gg.chart_mean(ph, Q)
        _, xe = dΩe.quadrature_points(chart=False)
        ep.append(gg.l2_error(ph, Q, dΩe, fq=sinlat(xe) + mean))
        minus_sigma = -gg.surface_gradient(dΩe, grad_sinlat).get()
        eu.append(gg.l2_error(u, V, dΩe, fq=minus_sigma))
        if lvl == 2:
            eu_ref, ep_ref = solve_hodge_scalar(CubedSphereMesh2D(4), p)
            assert abs(eu[-1] - eu_ref) <= 1e-8 * eu_ref and abs(ep[-1] - ep_ref) <= 1e-8 * ep_ref
    lv = np.log10([1 / 2.0**lvl for lvl in (1, 2, 3)])
    assert np.polyfit(lv, np.log10(eu), 1)[0] >= 2 and np.polyfit(lv, np.log10(ep), 1)[0] >= 2


def test_error_behaviour_of_the_driver_hooks(gg):
    """Wrong sizes / unsupported requests return error codes (GGError), never garbage."""
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=2)
    dΩ = gg.Measure(gg.Triangulation(model), 6)
    V = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, 1), conformity="H1")
    W = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, gg.VectorValue, 1), conformity="H1")
    Q = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, 1), conformity="L2")
    u = np.zeros(V.num_free_dofs)
    with pytest.raises(gg.GGError):
        gg.l2_error(u, V, dΩ, fq=np.zeros(7))                               # f at the wrong number of points
    with pytest.raises(gg.GGError):
        gg.l2_error(np.zeros(3), V, dΩ)                                     # DOF vector of the wrong length
    with pytest.raises(gg.GGError):
        gg.assemble_vector(gg.VectorSource(dΩ, np.zeros(model.num_cells * dΩ.num_points)), W)   # one component only
    with pytest.raises(gg.GGError):
        gg.assemble_matrix(gg.VectorMass(dΩ), V, V)                         # scalar space under the vector form
    with pytest.raises(gg.GGError):
        gg.change_of_basis(V)                                               # only vector H1 spaces carry one
    with pytest.raises(ValueError):
        gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, gg.VectorValue, 1), conformity="L2")
    with pytest.raises(gg.GGError):
        gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, gg.VectorValue, 3), conformity="H1")  # order 3 not built
    with pytest.raises(gg.GGError):
        gg.CSRMatrix(V, V, skeleton=True)                                   # skeleton patterns are for DG spaces
    A = gg.CSRMatrix(Q, Q, skeleton=True)
    with pytest.raises(gg.GGError):
        form = gg.ScalarMass(dΩ)                                            # generic numeric phase on a skeleton pattern
        args = form.args(Q, Q)
        import ctypes
        gg._lib.check(Q.lib.gg_assemble_matrix(form.form_id, ctypes.byref(args), A.h, 0))
    shell = gg.AtlasDiscreteModel(gg.CubedSphereWithThicknessMesh(1.0, 0.1), n=2)
    with pytest.raises(gg.GGError):
        gg.surface_laplacian(gg.Measure(gg.Triangulation(shell), 4), grad_fX, hess_fX)      # 2-D surfaces only
    with pytest.raises(gg.GGError):
        gg.RKStepper(gg.RungeKutta(None, None, 1e-3, "EXRK_SSP_3_3"), gg.ThermalShallowWaterOperator(model, 3, lambda X: X[..., 2]))
