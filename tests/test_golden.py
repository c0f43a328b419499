"""The committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py) against the oracle (CPU) -- guards the
fixtures against drift of the restatement -- and, under -m gpu, against the CUDA path through the C ABI."""
import os

import numpy as np
import pytest

from __graft_entry__ import load_package
from oracle import advection, forms, shell, surface_stokes, swe, tswe, vector_h1
from oracle import rk as ork
from oracle.mesh import CubedSphereMesh2D
from oracle.spaces import CGSpace

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_reproduces_the_golden_fixtures():
    g = np.load(os.path.join(G, "laplace_beltrami_n4.npz"))
    mesh = CubedSphereMesh2D(4, 1.0)
    assert (mesh.cell_nodes == g["cell_nodes"]).all()
    V = CGSpace(mesh, 2)
    A = forms.assemble_matrix(V, V, forms.lb_element_matrices(V, forms.CellQuadrature(mesh, 18)))
    assert (A.indptr == g["q2_indptr"]).all() and (A.indices == g["q2_indices"]).all()
    assert np.abs(A.data - g["q2_data"]).max() < 1e-13 * np.abs(g["q2_data"]).max()
    t = np.load(os.path.join(G, "transient_n3.npz"))
    W = ork.WaveProblem(CubedSphereMesh2D(3, 1.0), 1)
    assert np.abs(W.residual(t["wave_x"]) - t["wave_res"]).max() < 1e-13 * np.abs(t["wave_res"]).max()
    P = swe.ShallowWaterProblem(CubedSphereMesh2D(3, 1.0), 1)
    assert np.abs(P.diagnostics(t["swe_x0"]) - t["swe_y0"]).max() < 1e-11 * np.abs(t["swe_y0"]).max()


def test_oracle_reproduces_the_next_row_fixtures():
    g = np.load(os.path.join(G, "next_rows_n3.npz"))
    mesh = CubedSphereMesh2D(3, 1.0)
    T = tswe.ThermalShallowWaterProblem(mesh, 1)
    T.dt, T.tau = float(g["tswe_dt"]), float(g["tswe_dt"]) / 2
    assert np.abs(T.diagnostics(g["tswe_x0"]) - g["tswe_y0"]).max() < 1e-11 * np.abs(g["tswe_y0"]).max()
    r = T.residual_x(g["tswe_x0"], g["tswe_y0"], g["tswe_y0"])
    assert np.abs(r - g["tswe_res0"]).max() < 1e-12 * np.abs(g["tswe_res0"]).max()
    Vv = vector_h1.VectorCGSpace(mesh, 2)
    assert (Vv.cell_dofs == g["vh1_cell_dofs"]).all() and np.abs(Vv.cob - g["vh1_cob"]).max() < 1e-14


def test_oracle_reproduces_the_surface_stokes_fixture():
    g = np.load(os.path.join(G, "surface_stokes_n3.npz"))
    mesh = CubedSphereMesh2D(3, 1.0)
    Vv, Qs = vector_h1.VectorCGSpace(mesh, 2), CGSpace(mesh, 1, zeromean=True)
    q = forms.CellQuadrature(mesh, 18)
    K = surface_stokes.vector_laplacian_element_matrices(Vv, q)
    B = surface_stokes.div_element_matrices(Qs, Vv, q)
    assert np.abs(K[:4] - g["K_first4"]).max() < 1e-13 * np.abs(g["K_first4"]).max()
    assert np.abs(B[:4] - g["B_first4"]).max() < 1e-13 * np.abs(g["B_first4"]).max()
    A = forms.assemble_matrix(Vv, Vv, K)
    assert (A.indptr == g["A_indptr"]).all() and (A.indices == g["A_indices"]).all()
    assert np.abs(A.data - g["A_data"]).max() < 1e-13 * np.abs(g["A_data"]).max()


_wX = lambda X: np.stack([X[..., 2] * X[..., 1], -X[..., 0] ** 2, X[..., 0] + 0.3], -1)   # noqa: E731
_gfX = lambda X: np.stack([2 * X[..., 0], 2 * X[..., 1] * X[..., 2], X[..., 1] ** 2], -1)    # noqa: E731


def _hfX(X):
    H = np.zeros(X.shape + (3,))
    H[..., 0, 0] = 2.0
    H[..., 1, 1] = 2 * X[..., 2]
    H[..., 1, 2] = H[..., 2, 1] = 2 * X[..., 1]
    return H


@pytest.mark.gpu
def test_gpu_against_the_next_row_fixtures():
    gg = load_package()
    g = np.load(os.path.join(G, "next_rows_n3.npz"))
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()   # noqa: E731
    m3 = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=3)
    st = gg.RKStepper(gg.RungeKutta(gg.CGSolver(rtol=1e-14), None, float(g["tswe_dt"]), "EXRK_SSP_3_3"),
                      gg.ThermalShallowWaterOperator(m3, 1, swe.coriolis()))
    assert rel(st.diagnostics(g["tswe_x0"]), g["tswe_y0"]) < 1e-10
    assert rel(st.stage_residual(g["tswe_x0"], g["tswe_y0"]), g["tswe_res0"]) < 1e-11
    st.set_state(g["tswe_x0"])
    st.step(1)
    assert rel(st.get_state(), g["tswe_x_after_1_step"]) < 1e-9
    V = gg.TestFESpace(m3, gg.ReferenceFE(gg.lagrangian, gg.VectorValue, 2), conformity="H1")
    assert (V.get_cell_dof_ids() - 1 == g["vh1_cell_dofs"]).all() and np.abs(gg.change_of_basis(V) - g["vh1_cob"]).max() < 1e-13
    A = gg.assemble_matrix(gg.VectorMass(gg.Measure(gg.Triangulation(m3), 12)), V, V).to_scipy()
    assert (A.indptr == g["vh1_mass_indptr"]).all() and (A.indices == g["vh1_mass_indices"]).all()
    assert rel(A.data, g["vh1_mass_data"]) < 1e-12
    assert rel(gg.interpolate(_wX, V).get(), g["vh1_interp"]) < 1e-12
    d6 = gg.Measure(gg.Triangulation(m3), 6)
    assert rel(gg.surface_laplacian(d6, _gfX, _hfX).get(), g["surf_lap"].reshape(-1)) < 1e-12
    assert rel(gg.surface_gradient(d6, _gfX).get(), g["surf_grad"].reshape(-1)) < 1e-12
    assert rel(gg.surface_divergence(d6, _gfX, _hfX).get(), g["surf_div_of_grad"].reshape(-1)) < 1e-12
    wind = lambda X: np.stack([-X[..., 1], X[..., 0], 0 * X[..., 0]], -1)   # noqa: E731
    Q = gg.TestFESpace(m3, gg.ReferenceFE(gg.lagrangian, float, 1), conformity="L2")
    Aa = gg.assemble_advection_matrix(gg.AdvectionOperator(m3, 1, wind, degree=4), Q).to_scipy().toarray()
    assert np.abs(Aa - g["adv_matrix_dense_n3_p1"]).max() < 1e-12 * np.abs(g["adv_matrix_dense_n3_p1"]).max()


@pytest.mark.gpu
def test_gpu_against_the_golden_fixtures():
    gg = load_package()
    g = np.load(os.path.join(G, "laplace_beltrami_n4.npz"))
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=4)
    assert (model.get_cell_node_ids() - 1 == g["cell_nodes"]).all()
    for p, deg in ((1, 12), (2, 18)):
        V = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="H1")
        dΩ = gg.Measure(gg.Triangulation(model), deg)
        assert (V.get_cell_dof_ids() - 1 == g[f"q{p}_cell_dofs"]).all()
        Ke = gg.element_matrices(gg.LaplaceBeltrami(dΩ), V, V, first=0, n=8)
        assert np.abs(Ke - g[f"q{p}_Ke_first8"]).max() < 1e-12 * np.abs(g[f"q{p}_Ke_first8"]).max()
        A = gg.assemble_matrix(gg.LaplaceBeltrami(dΩ), V, V).to_scipy()
        assert (A.indptr == g[f"q{p}_indptr"]).all() and (A.indices == g[f"q{p}_indices"]).all()   # bit-exact pattern
        assert np.abs(A.data - g[f"q{p}_data"]).max() < 1e-12 * np.abs(g[f"q{p}_data"]).max()
    t = np.load(os.path.join(G, "transient_n3.npz"))
    m3 = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=3)
    Vr = gg.TestFESpace(m3, gg.ReferenceFE(gg.raviart_thomas, float, 1), conformity="HDiv")
    assert (Vr.get_cell_dof_ids() - 1 == t["rt1_cell_dofs"]).all() and (Vr.get_cell_dof_signs() == t["rt1_signs"]).all()
    rk = gg.RungeKutta(gg.CGSolver(rtol=1e-14), None, 0.01, "EXRK_SSP_3_3")
    st = gg.RKStepper(rk, gg.WaveOperator(m3, 1))
    assert np.abs(st.residual(t["wave_x"]) - t["wave_res"]).max() < 1e-12 * np.abs(t["wave_res"]).max()
    st.set_state(t["wave_x"])
    st.step(3)
    ref = t["wave_x_after_3_steps_dt0.01_ssp33"]
    assert np.abs(st.get_state() - ref).max() < 1e-11 * np.abs(ref).max()
    op = gg.ShallowWaterOperator(m3, 1, swe.coriolis(), gravity=swe._g)
    ss = gg.RKStepper(gg.RungeKutta(gg.CGSolver(rtol=1e-14), None, float(t["swe_dt"]), "EXRK_SSP_3_3"), op)
    y = ss.diagnostics(t["swe_x0"])
    assert np.abs(y - t["swe_y0"]).max() < 1e-10 * np.abs(t["swe_y0"]).max()
    r = ss.stage_residual(t["swe_x0"], t["swe_y0"])
    assert np.abs(r - t["swe_res0"]).max() < 1e-12 * np.abs(t["swe_res0"]).max()
    ss.set_state(t["swe_x0"])
    ss.step(2)
    assert np.abs(ss.get_state() - t["swe_x_after_2_steps"]).max() < 1e-9 * np.abs(t["swe_x_after_2_steps"]).max()
    s = np.load(os.path.join(G, "shell_advection.npz"))
    sh = gg.AtlasDiscreteModel(gg.CubedSphereWithThicknessMesh(1.0, 0.19), n=2)
    V3 = gg.TestFESpace(sh, gg.ReferenceFE(gg.lagrangian, float, 2), conformity="H1")
    assert (sh.get_cell_node_ids() - 1 == s["shell_cell_nodes"]).all() and (V3.get_cell_dof_ids() - 1 == s["shell_cell_dofs"]).all()
    A = gg.assemble_matrix(gg.LaplaceBeltrami(gg.Measure(gg.Triangulation(sh), 8)), V3, V3).to_scipy()
    assert (A.indptr == s["shell_indptr"]).all() and (A.indices == s["shell_indices"]).all()
    assert np.abs(A.data - s["shell_data"]).max() < 1e-12 * np.abs(s["shell_data"]).max()
    m4 = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=4)
    wind = lambda X: np.stack([-X[..., 1], X[..., 0], 0 * X[..., 0]], -1)   # noqa: E731
    sa = gg.RKStepper(gg.RungeKutta(None, None, 0.01, "EXRK_SSP_3_3"), gg.AdvectionOperator(m4, 2, wind))
    assert np.abs(sa.residual(s["adv_u"]) - s["adv_res"]).max() < 1e-12 * np.abs(s["adv_res"]).max()
