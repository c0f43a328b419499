"""The C restatement (CPU baseline) agrees with the numpy oracle (CPU only)."""
import numpy as np
import pytest

from oracle import c_oracle, forms
from oracle.mesh import CubedSphereMesh2D
from oracle.spaces import CGSpace


@pytest.mark.parametrize("n,p,degree", [(2, 1, 12), (4, 2, 18), (3, 2, 8)])
def test_c_oracle_matches_numpy_oracle(n, p, degree):
    mesh = CubedSphereMesh2D(n, 1.4)
    V = CGSpace(mesh, p, zeromean=True)
    q = forms.CellQuadrature(mesh, degree)
    Ko = forms.lb_element_matrices(V, q)
    Kc = c_oracle.lb_element_matrices(mesh.cell_chart_coords, 1.4, p, degree)
    assert np.abs(Kc - Ko).max() < 1e-13 * np.abs(Ko).max()
    A = forms.assemble_matrix(V, V, Ko)
    vals = c_oracle.scatter_csr(V.cell_dofs, Kc, A.indptr, A.indices)
    assert np.abs(vals - A.data).max() < 1e-13 * np.abs(A.data).max()
    u = np.random.default_rng(0).standard_normal(V.num_free)
    ue = forms.gather(V, u, 0.25)
    fe = c_oracle.lb_element_vectors(mesh.cell_chart_coords, 1.4, p, degree, ue)
    feo = np.einsum("cab,cb->ca", Ko, ue)
    assert np.abs(fe - feo).max() < 1e-13 * np.abs(feo).max()
    r = c_oracle.scatter_vector(V.cell_dofs, fe, V.num_free)
    assert np.abs(r - forms.assemble_vector(V, feo)).max() < 1e-13 * np.abs(r).max()
    # threaded == serial (static cell ranges, no reduction across threads)
    K1 = c_oracle.lb_element_matrices(mesh.cell_chart_coords, 1.4, p, degree, nthreads=1)
    assert (K1 == Kc).all()


# ---- explicit-RK half (oracle/c/gg_oracle_rk.c) against the numpy oracle ------------------------------------------------------------

@pytest.mark.parametrize("n,p", [(3, 0), (3, 1), (2, 2)])
def test_c_wave_matches_numpy_oracle(n, p):
    from oracle import rk as ork
    mesh = CubedSphereMesh2D(n)
    W, Wc = ork.WaveProblem(mesh, p), c_oracle.CWaveProblem(mesh, p, nthreads=2)
    assert abs(Wc.Mu - W.Mu).max() < 1e-13 * abs(W.Mu).max()
    x0 = np.random.default_rng(11).standard_normal(W.nu + W.np_)
    ro = W.residual(x0)
    ru, rpe = Wc.residual(x0)
    rp = np.zeros(W.np_)
    rp[Wc.T.Q.cell_dofs] = rpe
    assert np.abs(np.concatenate([ru, rp]) - ro).max() < 1e-12 * np.abs(ro).max()
    x, xc = W.step(x0, 0.01), Wc.step(x0, 0.01)
    assert np.abs(x - xc).max() < 1e-9 * np.abs(x).max()


@pytest.mark.parametrize("n,p,q0", [(3, 1, "alias"), (2, 2, "alias"), (3, 1, "start_of_step")])
def test_c_shallow_water_matches_numpy_oracle(n, p, q0):
    from oracle import swe
    mesh = CubedSphereMesh2D(n)
    S = swe.ShallowWaterProblem(mesh, p, q0_mode=q0)
    Sc = c_oracle.CShallowWaterProblem(mesh, p, swe.coriolis(), swe._g, q0_mode=q0, nthreads=2)
    x0 = S.initial_state()
    x0 = x0 + 1e-3 * np.abs(x0).max() * np.random.default_rng(5).standard_normal(x0.size)
    y, yc = S.diagnostics(x0), Sc.diagnostics(x0)
    for a, b in zip(S.split_y(y), S.split_y(yc)):
        assert np.abs(a - b).max() < 1e-9 * np.abs(a).max()
    dt = swe.reference_dt(mesh, max(p, 1))
    S.dt = Sc.dt = dt
    S.tau = Sc.tau = dt / 2
    y0 = 0.9 * y
    ro = S.residual_x(x0, y, y0)
    ru, rpe = Sc.residual_x(x0, y, y0)
    rp = np.zeros(S.np_)
    rp[Sc.T.Q.cell_dofs] = rpe
    assert np.abs(np.concatenate([ru, rp]) - ro).max() < 1e-11 * np.abs(ro).max()
    xF, yF = S.step(x0, dt)
    xFc, yFc = Sc.step(x0, dt)
    assert np.abs(xF - xFc).max() < 1e-9 * np.abs(xF).max()
    assert np.abs(yF - yFc).max() < 1e-7 * np.abs(yF).max()
