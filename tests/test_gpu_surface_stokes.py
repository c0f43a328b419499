"""GPU parity: the vector Laplacian and Taylor-Hood Stokes blocks on the vector H1 space (GG_FORM_VECTOR_LAPLACIAN, GG_FORM_VECTOR_DIV,
gg_h1_seminorm_sq) against oracle/surface_stokes.py and the committed fixture, and the two drivers of test/SurfaceStokes/
(VectorLaplacian2D.jl, SurfaceStokes_TaylorHood.jl) with device-assembled blocks, the caller's LU (LUSolver() in the reference) and
device error norms.  Tolerance: 1e-12 relative to the largest entry (north_star), written where it is used."""
import os

import numpy as np
import pytest

from __graft_entry__ import load_package
from oracle import forms
from oracle import surface_stokes as ss
from oracle import vector_h1 as vh
from oracle.mesh import CubedSphereMesh2D
from oracle.spaces import CGSpace
from test_oracle_surface_stokes import grad_pX, grad_uX, pX, uX, uX_sym

pytestmark = pytest.mark.gpu
# the reference measures the errors with Measure(Ω, 2 degree) = 36 (VectorLaplacian2D.jl:30-33) and Measure(Ω, 3 degree) = 54
# (SurfaceStokes_TaylorHood.jl:45-48): the same rules on both sides
ERR_DEGREE_VL, ERR_DEGREE_ST = 36, 54
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gg():
    return load_package()


def wX(X):
    return np.stack([X[..., 2] * X[..., 1], -X[..., 0] ** 2, X[..., 0] + 0.3], axis=-1)


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def spaces(gg, model, p):
    V = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, gg.VectorValue, p), conformity="H1")
    Q = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, max(p - 1, 1)), conformity="H1", constraint="zeromean")
    return V, Q


@pytest.mark.parametrize("n,p,degree", [(2, 1, 12), (3, 2, 18), (4, 2, 7)])
def test_element_and_assembled_matrices_match_oracle(gg, n, p, degree):
    r = 1.3
    mesh = CubedSphereMesh2D(n, r)
    Vo, Qo = vh.VectorCGSpace(mesh, p), CGSpace(mesh, max(p - 1, 1), zeromean=True)
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(r), n=n)
    V, Q = spaces(gg, model, p)
    dΩ = gg.Measure(gg.Triangulation(model), degree)
    q = forms.CellQuadrature(mesh, degree)
    Ko = ss.vector_laplacian_element_matrices(Vo, q)
    Ke = gg.element_matrices(gg.VectorLaplacian(dΩ), V, V)
    assert rel(Ke, Ko) <= 1e-12
    Bo = ss.div_element_matrices(Qo, Vo, q)
    Be = gg.element_matrices(gg.VectorDiv(dΩ), V, Q)
    assert Be.shape == Bo.shape and rel(Be, Bo) <= 1e-12
    Ao = forms.assemble_matrix(Vo, Vo, 0.7 * Ko)
    A = gg.assemble_matrix(gg.VectorLaplacian(dΩ, scale=0.7), V, V).to_scipy()
    assert (A.indptr == Ao.indptr).all() and (A.indices == Ao.indices).all() and rel(A.data, Ao.data) <= 1e-12
    Bao = forms.assemble_matrix(Qo, Vo, Bo)
    Ba = gg.assemble_matrix(gg.VectorDiv(dΩ), V, Q).to_scipy()
    assert Ba.shape == Bao.shape and (Ba.indptr == Bao.indptr).all() and (Ba.indices == Bao.indices).all()
    assert rel(Ba.data, Bao.data) <= 1e-12
    U = vh.interpolate(Vo, wX)
    h1o = np.sqrt(ss.h1_error(Vo, q, U, 0 * U) ** 2 - ss.fe_l2_error(Vo, q, U, 0 * U) ** 2)
    assert abs(gg.h1_seminorm(U, V, degree) - h1o) <= 1e-11 * h1o


def test_against_the_committed_fixture(gg):
    g = np.load(os.path.join(G, "surface_stokes_n3.npz"))
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=3)
    V, Q = spaces(gg, model, 2)
    dΩ = gg.Measure(gg.Triangulation(model), 18)
    assert rel(gg.element_matrices(gg.VectorLaplacian(dΩ), V, V, 0, 4), g["K_first4"]) <= 1e-12
    assert rel(gg.element_matrices(gg.VectorDiv(dΩ), V, Q, 0, 4), g["B_first4"]) <= 1e-12
    A = gg.assemble_matrix(gg.VectorLaplacian(dΩ), V, V).to_scipy()
    assert (A.indptr == g["A_indptr"]).all() and (A.indices == g["A_indices"]).all() and rel(A.data, g["A_data"]) <= 1e-12
    B = gg.assemble_matrix(gg.VectorDiv(dΩ), V, Q).to_scipy()
    assert (B.indptr == g["B_indptr"]).all() and (B.indices == g["B_indices"]).all() and rel(B.data, g["B_data"]) <= 1e-12
    U = gg.interpolate(wX, V)
    assert abs(gg.h1_seminorm(U, V, 18) ** 2 - float(g["h1_of_interp"])) <= 1e-11 * float(g["h1_of_interp"])


def test_form_argument_errors(gg):
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=2)
    V, Q = spaces(gg, model, 2)
    dΩ = gg.Measure(gg.Triangulation(model), 6)
    with pytest.raises(gg.GGError):
        gg.element_matrices(gg.VectorLaplacian(dΩ), Q, Q)                   # scalar spaces
    with pytest.raises(gg.GGError):
        gg.element_matrices(gg.VectorDiv(dΩ), Q, V)                         # test / trial swapped
    with pytest.raises(gg.GGError):
        gg.h1_seminorm(np.zeros(Q.num_free_dofs), Q, 6)                      # not a vector H1 space


def test_vector_laplacian_driver_converges(gg):
    """test/SurfaceStokes/VectorLaplacian2D.jl: p = 2, degree 6 (p + 1), errors with 2 degree; rhs = -vecΔs_2D(uX) = 2 u in closed
    form (oracle test: equal to the reference's chain of definitions on all panels)."""
    import scipy.sparse.linalg as spla
    p, degree = 2, 18
    levels = (1, 2, 3, 4)
    errs = []
    for lvl in levels:
        model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=2**lvl)
        V, _ = spaces(gg, model, p)
        Ω = gg.Triangulation(model)
        dΩ = gg.Measure(Ω, degree)
        rhs = 2.0 * gg.contravariant_components(dΩ, uX).get()
        op = gg.AffineFEOperator(gg.VectorLaplacian(dΩ), gg.VectorSource(dΩ, rhs), V, V)
        uh = spla.spsolve(op.get_matrix().to_scipy().tocsc(), op.get_vector().get())
        e = gg.interpolate(uX, V).get() - uh
        errs.append(gg.norm(e, V, ERR_DEGREE_VL))
        if lvl == 2:
            el2, eh1, _, _ = ss.vector_laplacian_2d(CubedSphereMesh2D(4), p, uX, rhs=lambda q: 2.0 * vh.contravariant(q.mesh, q, uX),
                                                        error_degree=ERR_DEGREE_VL)
            assert abs(errs[-1] - el2) <= 1e-8 * el2
            assert abs(np.hypot(errs[-1], gg.h1_seminorm(e, V, ERR_DEGREE_VL)) - eh1) <= 1e-8 * eh1
    assert np.polyfit(np.log10([1 / 2.0**lvl for lvl in levels]), np.log10(errs), 1)[0] >= p


def test_surface_stokes_driver_converges(gg):
    """test/SurfaceStokes/SurfaceStokes_TaylorHood.jl: Q2 vector x Q1 zero-mean, alpha = nu = 1; the tested quantity is eu_h1."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    p, degree, alpha, nu = 2, 18, 1.0, 1.0
    levels = (1, 2, 3, 4)
    eh1, epl2 = [], []
    for lvl in levels:
        model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=2**lvl)
        V, Q = spaces(gg, model, p)
        Ω = gg.Triangulation(model)
        dΩ, dΩe = gg.Measure(Ω, degree), gg.Measure(Ω, ERR_DEGREE_ST)
        u_c = gg.contravariant_components(dΩ, uX).get()
        rhs = (2.0 * nu + alpha) * u_c + gg.surface_gradient(dΩ, grad_pX).get()
        sigma = gg.surface_divergence(dΩ, uX, grad_uX)
        Auu = (gg.assemble_matrix(gg.VectorLaplacian(dΩ, scale=nu), V, V).to_scipy()
               + gg.assemble_matrix(gg.VectorMass(dΩ, scale=alpha), V, V).to_scipy())
        B = gg.assemble_matrix(gg.VectorDiv(dΩ), V, Q).to_scipy()
        bu = gg.assemble_vector(gg.VectorSource(dΩ, rhs), V).get()
        bq = gg.assemble_vector(gg.Source(dΩ, sigma), Q).get()
        x = spla.spsolve(sp.bmat([[Auu, -B.T], [B, None]]).tocsc(), np.concatenate([bu, bq]))
        uh, ph = x[: V.num_free_dofs], x[V.num_free_dofs:]
        e = gg.interpolate(uX, V).get() - uh
        eh1.append(np.hypot(gg.norm(e, V, ERR_DEGREE_ST), gg.h1_seminorm(e, V, ERR_DEGREE_ST)))
        # ZeroMeanFESpace: solution and interpolant are both shifted to zero chart mean
        p_int = gg.interpolate(pX, Q)
        pi = p_int.get()
        d_int = float(pX(_fixed_node_xyz(gg, Q)))
        ep = (pi - gg.chart_mean(pi, Q, dirichlet=d_int)) - (ph - gg.chart_mean(ph, Q))
        ep_fixed = (d_int - gg.chart_mean(pi, Q, dirichlet=d_int)) - (0.0 - gg.chart_mean(ph, Q))
        epl2.append(gg.l2_error(ep, Q, dΩe, dirichlet=ep_fixed))
        if lvl == 2:
            r_h1, r_p, _ = ss.surface_stokes(CubedSphereMesh2D(4), p, uX, pX, uX_sym, grad_pX, grad_uX, error_degree=ERR_DEGREE_ST)
            assert abs(eh1[-1] - r_h1) <= 1e-8 * r_h1 and abs(epl2[-1] - r_p) <= 1e-8 * r_p
    h = np.log10([1 / 2.0**lvl for lvl in levels])
    assert np.polyfit(h, np.log10(eh1), 1)[0] >= p
    assert np.polyfit(h, np.log10(epl2), 1)[0] >= 1.8


def _fixed_node_xyz(gg, Q):
    """ambient position of the fixed (last) node of a zero-mean CG space: the one local DOF with id <= 0."""
    ids = Q.get_cell_dof_ids()
    c, a = np.argwhere(ids <= 0)[0]
    return gg.interpolation_points(Q)[1][c, a]
